"""GPU parity of the headline kernel: the checkerboard Metropolis kernel must reproduce the CPU model of the
same per-task random-bit procedure (oracle/rrrmc_oracle.c:orc_checkerboard_sweeps) bit for bit, and its
observables must agree with the reference's random-site Metropolis within 3σ."""
import ctypes as C

import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from rrrmc_b200._ffi import check, lib, ptr
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu


def _fwd(A, J, L, D):
    """forward-bond couplings [N][D] from the reference (A,J) layout."""
    N = A.shape[0]
    out = np.zeros((N, D), np.int8)
    for i in range(N):
        stride = 1
        for d in range(D):
            c = (i // stride) % L
            up = i + (((c + 1) % L) - c) * stride
            ks = np.flatnonzero(A[i] == up + 1)
            # L=2: two bonds to the same neighbour; the lower site's forward bond sits first
            k = ks[0] if (len(ks) == 1 or i < up) else ks[1]
            out[i, d] = J[i, k]
            stride *= L
    return out


def _multispin(Cfg):
    """(R,N) bits -> multispin words [N][W]"""
    b = Cfg.s.astype(np.uint8)
    R, N = b.shape
    W = (R + 31) // 32
    pad = np.zeros((W * 32, N), np.uint8); pad[:R] = b
    packed = np.ascontiguousarray(np.packbits(np.ascontiguousarray(pad.T).reshape(N, W, 32), axis=2, bitorder="little"))
    return np.ascontiguousarray(packed.reshape(N, W * 4).view(np.uint32).reshape(N, W))


def _from_multispin(sp, R):
    N, W = sp.shape
    bits = np.unpackbits(sp.view(np.uint8).reshape(N, W * 4), axis=1, bitorder="little")[:, :R]
    return rb.Config.from_bits(bits.T)


@pytest.mark.parametrize("L,D,R,K,M,beta", [(4, 2, 32, 6, 0, 0.5), (6, 2, 96, 3, 8, 1.0), (4, 3, 128, 6, 8, 0.7), (6, 3, 160, 0, 0, 1.2),
                                            (8, 3, 256, 8, 4, 0.3), (2, 3, 64, 6, 8, 0.9), (4, 1, 32, 4, 4, 0.6), (8, 3, 100, 6, 8, 2.0),
                                            (8, 3, 512, 0, 4, 0.2), (8, 3, 384, 1, 12, 0.4), (6, 3, 128, 4, 28, 1.0), (8, 2, 1024, 7, 8, 1.0)])
def test_checkerboard_bit_exact_vs_cpu_model(L, D, R, K, M, beta):
    A, J = ea_instance(L, D, seed=L * 10 + D)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(7))
    thr = ffi.thresholds_fixed64(beta, D)
    seed, nsw = 0xC0FFEE1234, 5
    X._upload(C0)
    check(lib().rrrmc_checkerboard_sweeps(X._state, ptr(thr), D, K, M, seed, 3, nsw))
    got = X._download()
    Rp = ((R + 31) // 32) * 32
    sp = _multispin(C0)
    acc = np.zeros(Rp, np.int64)
    ffi.checkerboard_sweeps(L, D, Rp, sp, _fwd(A, J, L, D), thr, K, seed, 3, nsw, acc, M=M)
    want = _from_multispin(sp, R)
    assert got == want
    assert not (got == C0)


def _sparse_tbl(beta, D):
    thr = ffi.thresholds_fixed64(beta, D)
    n = ffi.CB_T1 + (D - 1) * ffi.CB_TC
    tbl = np.zeros(n, np.uint32)
    check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), D, ptr(tbl), n))
    return tbl


@pytest.mark.parametrize("L,D,R,beta", [(4, 2, 32, 0.5), (6, 2, 96, 1.0), (4, 3, 128, 0.7), (6, 3, 160, 1.2), (8, 3, 256, 0.3),
                                        (2, 3, 64, 0.9), (4, 1, 32, 0.6), (8, 3, 100, 2.0), (8, 3, 512, 0.05), (8, 3, 384, 1.0),
                                        (6, 3, 128, 0.0), (8, 2, 1024, 1.0), (16, 3, 1024, 1.0)])
def test_checkerboard_sparse_bit_exact_vs_cpu_model(L, D, R, beta):
    """Sparse acceptance procedure (binomial counts + positions) against orc_checkerboard_sweeps_sparse. Small β makes
    almost every lane pass, which drives the duplicate-redraw and window-refill paths hard."""
    A, J = ea_instance(L, D, seed=L * 10 + D)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(7))
    tbl = _sparse_tbl(beta, D)
    seed, nsw = 0xC0FFEE1234, 5
    X._upload(C0)
    check(lib().rrrmc_checkerboard_sweeps_sparse(X._state, ptr(tbl), len(tbl), seed, (1 << 33) + 3, nsw))
    got = X._download()
    Rp = ((R + 31) // 32) * 32
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_sparse(L, D, Rp, sp, _fwd(A, J, L, D), tbl, seed, (1 << 33) + 3, nsw)
    want = _from_multispin(sp, R)
    assert got == want
    assert not (got == C0)


def _poisson_tbl(beta, D):
    thr = ffi.thresholds_fixed64(beta, D)
    tbl = np.zeros(ffi.CBP_LEN, np.uint32)
    check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), D, ptr(tbl), len(tbl)))
    assert np.array_equal(tbl, ffi.cb_poisson_tables(thr))
    return tbl


@pytest.mark.parametrize("L,D,R,beta,NW", [(4, 2, 32, 0.5, 1), (6, 2, 96, 1.0, 2), (4, 3, 128, 0.7, 4), (6, 3, 160, 1.2, 2), (8, 3, 256, 0.5, 1),
                                           (2, 3, 64, 0.9, 6), (4, 1, 32, 0.6, 2), (8, 3, 100, 2.0, 1), (8, 3, 512, 0.45, 2), (8, 3, 384, 1.0, 2),
                                           (6, 3, 128, 0.6, 6), (8, 2, 1024, 1.0, 4), (16, 3, 1024, 1.0, 2), (8, 3, 256, 0.8, 4),
                                           (8, 1, 128, 0.3, 1), (8, 2, 256, 0.4, 1), (32, 3, 1024, 1.0, 2), (64, 3, 128, 1.1, 2),
                                           (12, 3, 256, 0.9, 2), (20, 3, 1024, 1.0, 2)])
def test_checkerboard_poisson_bit_exact_vs_cpu_model(L, D, R, beta, NW):
    """Poisson acceptance procedure (hit counts + positions with replacement) against orc_checkerboard_sweeps_poisson.
    Warm β with few static slots drives the overflow stream, the multi-hit level-2/3 paths and the ambiguous lookup
    buckets hard; cold β is the branch-free path the benchmark runs."""
    A, J = ea_instance(L, D, seed=L * 10 + D)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(7))
    tbl = _poisson_tbl(beta, D)
    seed, nsw = 0xC0FFEE1234, (5 if L < 32 else 2)   # the big lattices walk several bricks per persistent block
    X._upload(C0)
    check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbl), len(tbl), NW, seed, (1 << 33) + 3, nsw))
    got = X._download()
    Rp = ((R + 31) // 32) * 32
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_poisson(L, D, Rp, sp, _fwd(A, J, L, D), tbl, NW, seed, (1 << 33) + 3, nsw)
    want = _from_multispin(sp, R)
    assert got == want
    assert not (got == C0)


@pytest.mark.parametrize("variant", ["128", "8", "16", "32", "512", "640", "1024", "1152"])
def test_checkerboard_poisson_kernel_variants_agree(variant, monkeypatch):
    """The persistent kernels (one task per thread; two tasks = two replica groups of a site per thread) with two blocks
    (every block walks 32 bricks), the one-task-per-thread kernel, the row mapping and the launch without programmatic
    serialization all produce the oracle's trajectory (warm β with one static word: most tasks run the second tier)."""
    L, D, R, beta, NW = 16, 3, 1024, 0.8, 1
    A, J = ea_instance(L, D, seed=11)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(17))
    tbl = _poisson_tbl(beta, D)
    X._upload(C0)
    monkeypatch.setenv("RRRMC_CB_VARIANT", variant)
    check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbl), len(tbl), NW, 5, 0, 2))
    got = X._download()
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_poisson(L, D, R, sp, _fwd(A, J, L, D), tbl, NW, 5, 0, 2)
    assert got == _from_multispin(sp, R)


def _poisson_run_vs_oracle(L, R, beta, NW, nsw, seed, sweep0, A=None, J=None, Jfwd=None, cfg_seed=7):
    """-> (device configuration, oracle configuration) after nsw poisson sweeps from the same start."""
    D = 3
    if A is None:
        A, J = ea_instance(L, D, seed=L * 10 + D)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(cfg_seed))
    tbl = _poisson_tbl(beta, D)
    X._upload(C0)
    l0 = X.ctx.launch_count()
    check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbl), len(tbl), NW, seed, sweep0, nsw))
    launches = X.ctx.launch_count() - l0
    got = X._download()
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_poisson(L, D, R, sp, _fwd(A, J, L, D) if Jfwd is None else Jfwd, tbl, NW, seed, sweep0, nsw)
    return got, _from_multispin(sp, R), launches


# the TMA-staged brick kernel (ea_tma.cu) is what AUTO runs for 3D lattices with L % 8 == 0 and whole 1024-replica slabs.
# L = 8: one brick along x (both x faces wrap onto the brick itself), two along y and z (every halo wraps); L = 16, 24:
# interior bricks; R = 2048, 3072: several slabs; NW = 4, 6: the second static Philox call; warm β: second tier.
@pytest.mark.parametrize("L,R,beta,NW", [(8, 1024, 1.0, 2), (8, 2048, 0.8, 1), (16, 1024, 0.6, 4), (16, 3072, 1.0, 2),
                                         (24, 1024, 1.2, 1), (16, 1024, 0.5, 6), (32, 1024, 1.0, 2)])
def test_checkerboard_tma_bit_exact_vs_cpu_model(L, R, beta, NW, monkeypatch):
    monkeypatch.delenv("RRRMC_CB_VARIANT", raising=False)
    nsw = 4 if L <= 16 else 2
    got, want, launches = _poisson_run_vs_oracle(L, R, beta, NW, nsw, 0xC0FFEE1234, (1 << 33) + 3)
    assert got == want
    assert launches in (1, 2)                        # ONE launch of the multi-sweep kernel (+ the one-off bond-mask reorder)
    # the per-colour TMA launches (bit 12 forbids the multi-sweep kernel): two launches per sweep, same trajectory
    monkeypatch.setenv("RRRMC_CB_VARIANT", "4096")
    got1, _, launches1 = _poisson_run_vs_oracle(L, R, beta, NW, nsw, 0xC0FFEE1234, (1 << 33) + 3)
    assert got1 == want
    assert launches1 in (2 * nsw, 2 * nsw + 1)
    # the cp.async kernels (TMA forbidden) give the same trajectory
    monkeypatch.setenv("RRRMC_CB_VARIANT", "2048")
    got2, _, _ = _poisson_run_vs_oracle(L, R, beta, NW, nsw, 0xC0FFEE1234, (1 << 33) + 3)
    assert got2 == want


@pytest.mark.parametrize("variant", ["128", "1", "129", "4224", "4097", "4128", "4225"])
def test_checkerboard_tma_kernel_variants_agree(variant, monkeypatch):
    """TMA kernel with two blocks only (every block walks 16 bricks through its two-stage ring: stage reuse, barrier
    phases), one block per SM, and without programmatic dependent launch."""
    monkeypatch.setenv("RRRMC_CB_VARIANT", variant)
    got, want, _ = _poisson_run_vs_oracle(16, 1024, 0.8, 1, 3, 5, 0)
    assert got == want


@pytest.mark.parametrize("variant,NW", [("2176", 2), ("2688", 2), ("2048", 2), ("3200", 1)])
def test_checkerboard_persist2_multibrick_bit_exact(variant, NW, monkeypatch):
    """The cp.async persistent kernels in the regime round 1 benched them in: NW = 2 (two tasks per thread) with many
    bricks per block. L = 32, R = 1024 has 512 bricks; bit 7 (128) launches two blocks (256 bricks each through the
    software pipeline), bit 9 (512) forces the two-task kernel, bit 10 (1024) the one-task kernel; bit 11 keeps TMA off."""
    monkeypatch.setenv("RRRMC_CB_VARIANT", variant)
    got, want, _ = _poisson_run_vs_oracle(32, 1024, 1.0, NW, 2, 77, 5)
    assert got == want


@pytest.mark.parametrize("variant", [None, "4096", "2048"])
def test_checkerboard_poisson_full_size_bit_exact(variant, monkeypatch):
    """BASELINE configs[1] at size: L = 64, D = 3, R = 1024, β = 1, NW = 2 (what bench.py runs), two sweeps against
    orc_checkerboard_sweeps_poisson — the TMA kernel (2048 bricks over 296 blocks: ~7 bricks per block) and the
    round-1 kernel k_checkerboard_poisson_persist2<2,4> (4096 bricks over 740 blocks)."""
    if variant is None:
        monkeypatch.delenv("RRRMC_CB_VARIANT", raising=False)
    else:
        monkeypatch.setenv("RRRMC_CB_VARIANT", variant)
    L, D, R = 64, 3, 1024
    rng = np.random.default_rng(64)
    A = ffi.gen_EA(L, D)
    idx = np.arange(L ** D)
    Jf = rng.choice(np.array([-1, 1], np.int8), (L ** D, D))
    # reference (A, J) layout from the forward bonds: slot of i+e_d in A[i] and of i in A[i+e_d]
    J = np.zeros(A.shape, np.int64)
    for d in range(D):
        c = (idx // L ** d) % L
        up = idx + (((c + 1) % L) - c) * L ** d
        J[idx, (A == (up + 1)[:, None]).argmax(axis=1)] = Jf[:, d]
        J[up, (A[up] == (idx + 1)[:, None]).argmax(axis=1)] = Jf[:, d]
    got, want, _ = _poisson_run_vs_oracle(L, R, 1.0, 2, 2, 0x5EEDEA64, 0, A=A, J=J, Jfwd=Jf, cfg_seed=1)
    assert got == want


def test_checkerboard_poisson_argument_checks():
    A, J = ea_instance(4, 3, seed=1)
    X = rb.GraphEA(4, 3, replicas=32, A=A, J=J)
    X._upload(rb.Config(X.N, 32, rng=np.random.default_rng(0)))
    tbl = _poisson_tbl(1.0, 3)
    for nw in (0, 3, 5, 7):
        with pytest.raises(ValueError):
            check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbl), len(tbl), nw, 1, 0, 1))
    with pytest.raises(ValueError):
        check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbl), len(tbl) - 1, 2, 1, 0, 1))
    bad = tbl.copy(); bad[ffi.CBP_KA - 1] = 5
    with pytest.raises(ValueError):
        check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(bad), len(bad), 2, 1, 0, 1))
    # NW selection: colder needs fewer static slots; too warm has none
    assert lib().rrrmc_checkerboard_poisson_nw(ptr(_poisson_tbl(2.0, 3)), 1.5e-3) == 1
    assert lib().rrrmc_checkerboard_poisson_nw(ptr(tbl), 1.5e-3) == ffi.cb_poisson_nw(tbl, 1.5e-3) == 4
    assert lib().rrrmc_checkerboard_poisson_nw(ptr(tbl), 0.0) == ffi.cb_poisson_nw(tbl) == 2
    assert lib().rrrmc_checkerboard_poisson_nw(ptr(_poisson_tbl(0.3, 3)), 0.0) == 0
    with pytest.raises(NotImplementedError):
        rb.standardMC(X, 0.3, 10 * X.N, quiet=True, cb_method="poisson")


def test_standardMC_poisson_energies_and_accepted():
    """standardMC through the public API with the default (auto -> poisson at β=1.1) procedure: energies, accepted
    counters and configurations at every hook against the CPU model."""
    L, D, R, beta = 6, 3, 64, 1.1
    A, J = ea_instance(L, D, seed=3)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_int(A, J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(8))
    seen = []

    def hook(it, X_, C, acc, E):
        seen.append((it, np.array(acc), np.array(E), C.chunks.copy()))
        return True
    N = X.N
    Es, Cf = rb.standardMC(X, beta, 6 * N, step=2 * N, seed=77, C0=C0, hook=hook, quiet=True, schedule="checkerboard")
    sp = _multispin(C0); acc = np.zeros(R, np.int64)
    tbl = ffi.cb_poisson_tables(ffi.thresholds_fixed64(beta, D))
    NW = ffi.cb_poisson_nw(tbl)
    for k in range(3):
        ffi.checkerboard_sweeps_poisson(L, D, R, sp, _fwd(A, J, L, D), tbl, NW, 77, 2 * k, 2, acc)
        cfg = _from_multispin(sp, R)
        assert np.array_equal(seen[k][3], cfg.chunks)
        assert np.array_equal(seen[k][1], acc)
        e = np.array([g.energy(cfg.chunks[r]) for r in range(R)])
        assert np.array_equal(seen[k][2], e.astype(np.int64)) and np.array_equal(Es[k], e.astype(np.int64))
    assert Cf == _from_multispin(sp, R)
    Es2, _ = rb.standardMC(X, beta, 6 * N, step=2 * N, seed=77, C0=C0, quiet=True, cb_method="poisson")
    assert np.array_equal(Es, Es2)


def test_standardMC_sparse_energies_and_accepted():
    L, D, R, beta = 6, 3, 64, 1.1
    A, J = ea_instance(L, D, seed=3)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_int(A, J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(8))
    seen = []

    def hook(it, X_, C, acc, E):
        seen.append((it, np.array(acc), np.array(E), C.chunks.copy()))
        return True
    N = X.N
    Es, Cf = rb.standardMC(X, beta, 6 * N, step=2 * N, seed=77, C0=C0, hook=hook, quiet=True, cb_method="sparse")
    sp = _multispin(C0); acc = np.zeros(R, np.int64)
    tbl = ffi.cb_sparse_tables(ffi.thresholds_fixed64(beta, D))
    for k in range(3):
        ffi.checkerboard_sweeps_sparse(L, D, R, sp, _fwd(A, J, L, D), tbl, 77, 2 * k, 2, acc)
        cfg = _from_multispin(sp, R)
        assert np.array_equal(seen[k][3], cfg.chunks)
        assert np.array_equal(seen[k][1], acc)
        e = np.array([g.energy(cfg.chunks[r]) for r in range(R)])
        assert np.array_equal(seen[k][2], e.astype(np.int64)) and np.array_equal(Es[k], e.astype(np.int64))
    assert Cf == _from_multispin(sp, R)
    # the two procedures are different consumers of the same counter stream: trajectories differ, both are valid
    Es2, _ = rb.standardMC(X, beta, 6 * N, step=2 * N, seed=77, C0=C0, quiet=True, cb_method="planes")
    assert not np.array_equal(Es, Es2)


@pytest.mark.parametrize("method,beta", [("sparse", 1.0), ("sparse", 0.6), ("planes", 1.0), ("poisson", 1.0), ("poisson", 0.65),
                                         ("auto", 0.5)])
def test_checkerboard_methods_statistics_vs_reference_sampler(method, beta):
    """⟨E⟩ after equilibration: each acceptance procedure vs the oracle's random-site standardMC within 3σ."""
    L, D = 6, 3
    A, J = ea_instance(L, D, seed=22)
    R = 512
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    N = X.N
    Es, _ = rb.standardMC(X, beta, 600 * N, step=600 * N, seed=6, quiet=True, cb_method=method)
    e_gpu = Es[-1] / N
    g = ffi.Graph.ea_int(A, J)
    e_cpu = []
    for r in range(96):
        src = ffi.PhiloxDraws(4321, chain=r)
        s = src.config(N)
        E, _ = ffi.standardMC(g, beta, 600 * N, s, src, step=600 * N)
        e_cpu.append(E[-1] / N)
    e_cpu = np.array(e_cpu)
    sigma = np.sqrt(e_gpu.var(ddof=1) / R + e_cpu.var(ddof=1) / len(e_cpu))
    assert abs(e_gpu.mean() - e_cpu.mean()) < 3 * sigma, (method, e_gpu.mean(), e_cpu.mean(), sigma)


def test_standardMC_checkerboard_energies_and_accepted():
    L, D, R, beta = 6, 3, 64, 0.8
    A, J = ea_instance(L, D, seed=3)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_int(A, J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(8))
    seen = []

    def hook(it, X_, C, acc, E):
        seen.append((it, np.array(acc), np.array(E), C.chunks.copy()))
        return True
    N = X.N
    Es, Cf = rb.standardMC(X, beta, 6 * N, step=2 * N, seed=77, C0=C0, hook=hook, quiet=True, planes_K=6, planes_M=4,
                           cb_method="planes")
    assert Es.shape == (3, R) and [s[0] for s in seen] == [2 * N, 4 * N, 6 * N]
    # CPU model with the same seed / thresholds
    sp = _multispin(C0); acc = np.zeros(R, np.int64)
    thr = ffi.thresholds_fixed64(beta, D)
    for k in range(3):
        ffi.checkerboard_sweeps(L, D, R, sp, _fwd(A, J, L, D), thr, 6, 77, 2 * k, 2, acc, M=4)
        cfg = _from_multispin(sp, R)
        assert np.array_equal(seen[k][3], cfg.chunks)
        assert np.array_equal(seen[k][1], acc)                       # exact accepted counters
        e = np.array([g.energy(cfg.chunks[r]) for r in range(R)])
        assert np.array_equal(seen[k][2], e.astype(np.int64)) and np.array_equal(Es[k], e.astype(np.int64))
    assert Cf == _from_multispin(sp, R)


def test_checkerboard_statistics_vs_reference_sampler():
    """⟨E⟩ from checkerboard sweeps vs the oracle's random-site standardMC (RRRMC.jl:81-127) within 3σ (L=6, 3D)."""
    L, D, beta = 6, 3, 0.6
    A, J = ea_instance(L, D, seed=21)
    R = 256
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    N = X.N
    Es, _ = rb.standardMC(X, beta, 400 * N, step=400 * N, seed=5, quiet=True, schedule="checkerboard")
    e_gpu = Es[-1] / N
    g = ffi.Graph.ea_int(A, J)
    e_cpu = []
    for r in range(64):
        src = ffi.PhiloxDraws(1234, chain=r)
        s = src.config(N)
        E, _ = ffi.standardMC(g, beta, 400 * N, s, src, step=400 * N)
        e_cpu.append(E[-1] / N)
    e_cpu = np.array(e_cpu)
    sigma = np.sqrt(e_gpu.var(ddof=1) / R + e_cpu.var(ddof=1) / len(e_cpu))
    assert abs(e_gpu.mean() - e_cpu.mean()) < 3 * sigma, (e_gpu.mean(), e_cpu.mean(), sigma)
    m = np.zeros(R); check(lib().rrrmc_magnetization(X._state, ptr(m)))
    assert abs(m.mean() / N) < 0.1


def test_full_size_properties_L64_R1024():
    """BASELINE config (3D L=64 ±J × 1024): energy kernel vs oracle on sampled replicas after device sweeps;
    flipping every site twice is the identity; sweeps at β=0 flip every spin every sweep (ΔE<=0 or U<1)."""
    L, D, R = 64, 3, 1024
    X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(0))
    st = X._ensure_state()
    check(lib().rrrmc_state_randomize(st, 99))
    thr = ffi.thresholds_fixed64(1.0, D)
    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), D, 6, 8, 1, 0, 4))
    E = np.zeros(R); check(lib().rrrmc_energy(st, ptr(E)))
    Cf = X._download()
    g = ffi.Graph.ea_int(X.A, X.J)
    for r in (0, 511, 1023):
        assert g.energy(Cf.chunks[r]) == E[r]
    assert (E < -0.5 * X.N).all()  # four sweeps at β=1 are already far below the random-state energy 0
    # β=0: thresholds saturate (p=1-2^-64): every attempt is accepted, two sweeps restore the state
    thr0 = np.full(D, 2 ** 64 - 1, np.uint64)
    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr0), D, 6, 8, 2, 0, 1))
    mid = X._download()
    assert np.array_equal(mid.chunks, ~Cf.chunks)
    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr0), D, 6, 8, 2, 1, 1))
    assert X._download() == Cf
    # sparse procedure at full size: energies keep decreasing towards equilibrium and match the oracle's energy()
    tbl = _sparse_tbl(1.0, D)
    check(lib().rrrmc_checkerboard_sweeps_sparse(st, ptr(tbl), len(tbl), 3, 0, 6))
    E2 = np.zeros(R); check(lib().rrrmc_energy(st, ptr(E2)))
    C2 = X._download()
    for r in (1, 700):
        assert g.energy(C2.chunks[r]) == E2[r]
    assert E2.mean() < E.mean()
    # poisson procedure at full size
    ptbl = _poisson_tbl(1.0, D)
    check(lib().rrrmc_checkerboard_sweeps_poisson(st, ptr(ptbl), len(ptbl), 2, 5, 0, 6))
    E3 = np.zeros(R); check(lib().rrrmc_energy(st, ptr(E3)))
    C3 = X._download()
    for r in (2, 900):
        assert g.energy(C3.chunks[r]) == E3[r]
    assert E3.mean() < E2.mean()


# ---- β ladder on the checkerboard schedule: one β per 128-replica group (multi-sweep brick kernel) ----
def _ladder_tbls(betas, D=3):
    return np.stack([_poisson_tbl(float(b), D) for b in betas])


@pytest.mark.parametrize("L,R,betas,NW", [(8, 1024, [0.6, 0.7, 0.8, 0.9, 1.0, 1.2, 1.5, 2.0], 4),
                                          (16, 1024, [2.0, 0.9, 1.0, 1.1, 0.8, 1.3, 1.4, 0.75], 2),
                                          (16, 2048, list(np.linspace(0.7, 2.2, 16)), 2)])
def test_checkerboard_ladder_bit_exact_vs_cpu_model(L, R, betas, NW, monkeypatch):
    """Every 128-replica group runs at its own β (eight or sixteen distinct values in one batch): bit-exact against the
    oracle's ladder restatement, and each group equals the one-β run of a batch that holds that β everywhere."""
    monkeypatch.delenv("RRRMC_CB_VARIANT", raising=False)
    D, nsw, seed, sweep0 = 3, 3, 0xBEEF77, 11
    A, J = ea_instance(L, D, seed=L + 5)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(23))
    tbls = _ladder_tbls(betas)
    X._upload(C0)
    check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(X._state, ptr(tbls), len(betas), NW, seed, sweep0, nsw))
    got = X._download()
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_poisson_ladder(L, D, R, sp, _fwd(A, J, L, D), tbls, NW, seed, sweep0, nsw)
    assert got == _from_multispin(sp, R)
    # group g of the ladder == group g of a one-β batch at betas[g] (same counters, same tables)
    g = 3
    X._upload(C0)
    check(lib().rrrmc_checkerboard_sweeps_poisson(X._state, ptr(tbls[g]), ffi.CBP_LEN, NW, seed, sweep0, nsw))
    one = X._download()
    assert np.array_equal(np.asarray(got.chunks)[128 * g:128 * (g + 1)], np.asarray(one.chunks)[128 * g:128 * (g + 1)])


def test_standard_mc_checkerboard_ladder_and_errors():
    """standardMC(schedule=checkerboard) with a per-replica β vector: constant inside 128-replica groups runs the ladder
    (colder groups end at lower energy); a β that changes inside a group and a lattice the brick kernel cannot take are
    rejected with a message."""
    L, D, R = 8, 3, 1024
    X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(3))
    betas = np.repeat(np.array([0.6, 0.7, 0.8, 0.9, 1.0, 1.2, 1.5, 2.5]), 128)
    Es, C = rb.standardMC(X, betas, 60 * X.N, step=60 * X.N, seed=4, schedule="checkerboard", C0=rb.Config(X.N, R, rng=np.random.default_rng(5)))
    e = np.asarray(Es)[-1].reshape(8, 128).mean(axis=1)
    assert np.array_equal(rb.energy(X, C), np.asarray(Es)[-1])
    assert e[0] > e[3] > e[7]
    bad = betas.copy(); bad[5] = 0.61
    with pytest.raises(Exception, match="constant inside each group"):
        rb.standardMC(X, bad, X.N, schedule="checkerboard")
    X2 = rb.GraphEA(6, 3, replicas=256, rng=np.random.default_rng(3))
    with pytest.raises(Exception, match="brick kernel"):
        rb.standardMC(X2, np.repeat([1.0, 1.1], 128), X2.N, schedule="checkerboard")


def test_tempering_exchange_vs_cpu_model():
    """rrrmc_tempering_exchange: the device's decisions equal orc_tempering_decide on the same energies, and accepted
    pairs exchange exactly their configurations (both parities; counters accumulate until read)."""
    L, D, R = 8, 3, 1024
    X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(9))
    bg = np.array([0.5, 0.6, 0.75, 0.9, 1.1, 1.3, 1.6, 2.0])
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(10))
    X._upload(C0)
    # a few ladder sweeps so that the energies of neighbouring groups differ in both directions
    tbls = _ladder_tbls(bg)
    check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(X._state, ptr(tbls), 8, 4, 3, 0, 5))
    total = np.zeros(7, np.int64)
    for rd in (0, 1, 2):
        before = X._download()
        E = rb.energy(X, before)
        want = ffi.tempering_decide(bg, E, 77, rd).astype(np.int64)
        acc = np.zeros(7, np.int64)
        check(lib().rrrmc_tempering_exchange(X._state, ptr(bg), 8, 77, rd, ptr(acc) if rd != 1 else None))
        after = X._download()
        perm = np.arange(R)
        for g in range(7):
            for l in np.nonzero(want[g])[0]:
                perm[128 * g + l], perm[128 * (g + 1) + l] = perm[128 * (g + 1) + l], perm[128 * g + l]
        assert np.array_equal(np.asarray(after.chunks), np.asarray(before.chunks)[perm])
        total += want.sum(axis=1)
        if rd == 0:
            assert np.array_equal(acc, want.sum(axis=1))
            total[:] = 0
        if rd == 2:
            assert np.array_equal(acc, total)      # rounds 1 and 2 together: round 1 did not read the counters
    assert want.sum() > 0 and (1 - want[::2]).sum() > 0


def test_tempered_checkerboard_orders_the_ladder():
    """sharding.tempered_checkerboard: after tempering, colder rungs sit at lower energy and neighbouring rungs exchange
    at a sensible rate."""
    from rrrmc_b200 import sharding
    L, D, R = 8, 3, 1024
    X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(2))
    bg = np.linspace(0.6, 1.3, 8)
    acc, att = sharding.tempered_checkerboard(X, bg, rounds=40, sweeps_per_round=5, seed=5, C0=rb.Config(X.N, R, rng=np.random.default_rng(6)))
    E = rb.energy(X, X._download()).reshape(8, 128).mean(axis=1)
    assert np.all(np.diff(E) < 0)
    rate = acc / att
    assert np.all(rate > 0.05) and np.all(rate <= 1.0)


def test_checkerboard_ladder_full_size_bit_exact():
    """BASELINE configs[1] at size with a β ladder: L = 64, R = 1024, eight distinct β (0.8 … 1.6, NW = 4), one sweep of the
    multi-sweep brick kernel with per-group tables against orc_checkerboard_sweeps_poisson_ladder, then a tempering
    exchange whose decisions are checked against orc_tempering_decide on the device's own energies."""
    L, D, R = 64, 3, 1024
    A, J = ea_instance(L, D, seed=64)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    bg = np.geomspace(0.8, 1.6, 8)
    tbls = _ladder_tbls(bg)
    X._upload(C0)
    check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(X._state, ptr(tbls), 8, 4, 0xABCDEF, 5, 1))
    got = X._download()
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_poisson_ladder(L, D, R, sp, _fwd(A, J, L, D), tbls, 4, 0xABCDEF, 5, 1)
    assert got == _from_multispin(sp, R)
    E = rb.energy(X, got)
    want = ffi.tempering_decide(bg, E, 3, 0).astype(np.int64)
    acc = np.zeros(7, np.int64)
    check(lib().rrrmc_tempering_exchange(X._state, ptr(bg), 8, 3, 0, ptr(acc)))
    assert np.array_equal(acc, want.sum(axis=1))
    after = X._download()
    perm = np.arange(R)
    for g in range(7):
        for l in np.nonzero(want[g])[0]:
            perm[128 * g + l], perm[128 * (g + 1) + l] = perm[128 * (g + 1) + l], perm[128 * g + l]
    assert np.array_equal(np.asarray(after.chunks), np.asarray(got.chunks)[perm])


def test_checkerboard_ladder_full_size_repeated_runs_agree():
    """Regression for a shared-memory stage race of the brick kernels (the consumers' "stage is free" arrive was issued
    while their ld.shared were still outstanding, and the refill through the async proxy could overtake them: about one
    run in ten of the ladder kernel at this size computed one warp's task from the next brick's bytes). Forty runs from
    the same start must all reproduce the oracle's trajectory (scripts/stress_flow.py is the long version)."""
    L, D, R = 64, 3, 1024
    A, J = ea_instance(L, D, seed=64)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    tbls = _ladder_tbls(np.geomspace(1.0, 1.6, 8))
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_poisson_ladder(L, D, R, sp, _fwd(A, J, L, D), tbls, 2, 0xABCDEF, 5, 1)
    want = np.asarray(_from_multispin(sp, R).chunks)
    bad = 0
    for _ in range(40):
        X._upload(C0)
        check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(X._state, ptr(tbls), 8, 2, 0xABCDEF, 5, 1))
        bad += not np.array_equal(np.asarray(X._download().chunks), want)
    assert bad == 0, "%d of 40 runs differ from the oracle" % bad


def test_two_contexts_run_the_multi_sweep_kernel_from_two_threads():
    """Two batches on two contexts (two streams) driven by two host threads at once, as bench.py's end-to-end arm does:
    the persistent multi-sweep kernels are chained device-wide (two half-resident cooperative grids would wait for each
    other forever), and every batch still reproduces the oracle's trajectory."""
    import threading
    L, D, R, beta, NW, nsw = 16, 3, 1024, 1.0, 2, 6
    A, J = ea_instance(L, D, seed=3)
    tbl = _poisson_tbl(beta, D)
    Xs = [rb.GraphEA(L, D, replicas=R, A=A, J=J, ctx=rb.Context(0)) for _ in range(2)]
    C0s = [rb.Config(Xs[0].N, R, rng=np.random.default_rng(40 + k)) for k in range(2)]
    out, errs = [None, None], []

    def work(k):
        try:
            Xs[k]._upload(C0s[k])
            for rep in range(5):
                check(lib().rrrmc_checkerboard_sweeps_poisson(Xs[k]._state, ptr(tbl), len(tbl), NW, 100 + k, rep * nsw, nsw))
            out[k] = Xs[k]._download()
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not errs and all(not t.is_alive() for t in th)
    for k in range(2):
        sp = _multispin(C0s[k])
        ffi.checkerboard_sweeps_poisson(L, D, R, sp, _fwd(A, J, L, D), tbl, NW, 100 + k, 0, 5 * nsw)
        assert out[k] == _from_multispin(sp, R)
