"""The warp-cooperative rrrMC / bklMC kernel (csrc/chain_warp.cu, site_pick="rank"): bit-exact against its CPU model
(oracle orc_rank_rrrMC / orc_rank_bklMC: the reference's chains with the member of a ΔE class picked by rank in site
order), statistically equal to the reference-order samplers, resumable through the hook protocol."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu


def _oracle(fn, g, beta, iters, step, C0, seed, R):
    Es, Cs, res = [], [], []
    for r in range(R):
        s = C0.chunks[r].copy()
        E, info = fn(g, beta, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step)
        Es.append(E); Cs.append(s); res.append(info)
    return np.array(Es).T, np.array(Cs), res


@pytest.mark.parametrize("L,D", [(4, 3), (3, 3), (6, 2), (5, 2), (8, 1), (8, 3), (12, 3)])
@pytest.mark.parametrize("sampler", ["rrr", "bkl"])
def test_rank_kernel_bit_exact_vs_cpu_model(L, D, sampler):
    R, beta, iters, step = 5, 1.7, 4000, 100
    A, J = ea_instance(L, D, (-1, 1), seed=L + D)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_int(A, J, (-1, 1))
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    if sampler == "rrr":
        Es, Cf = rb.rrrMC(X, beta, iters, step=step, seed=99, C0=C0, quiet=True, site_pick="rank")
        wantE, wantC, res = _oracle(ffi.rank_rrrMC, g, beta, iters, step, C0, 99, R)
    else:
        Es, Cf = rb.bklMC(X, beta, iters, step=step, seed=99, C0=C0, quiet=True, site_pick="rank")
        wantE, wantC, res = _oracle(ffi.rank_bklMC, g, beta, iters, step, C0, 99, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)
    assert X.last_run.accepted_total == sum(r.accepted for r in res)


@pytest.mark.parametrize("N,K", [(40, 3), (30, 4), (50, 5), (24, 6), (64, 2)])
@pytest.mark.parametrize("sampler", ["rrr", "bkl"])
def test_rank_kernel_on_rrg_bit_exact_vs_cpu_model(N, K, sampler):
    """The adjacency-table instantiation: ±J GraphRRG of degree 2..6 (odd degrees: allΔE = 2, 6, ..: no zero class)."""
    R, beta, iters, step = 4, 1.5, 3000, 100
    rng = np.random.default_rng(100 * N + K)
    A = rb.gen_RRG(N, K, rng)
    J = rb.gen_J_graph(lambda n: rng.choice([-1.0, 1.0], n), A).astype(np.int64)
    X = rb.GraphRRG(N, K, replicas=R, A=A, J=J)
    g = ffi.Graph.rrg_int(A, J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(5))
    if sampler == "rrr":
        Es, Cf = rb.rrrMC(X, beta, iters, step=step, seed=17, C0=C0, quiet=True, site_pick="rank")
        wantE, wantC, _ = _oracle(ffi.rank_rrrMC, g, beta, iters, step, C0, 17, R)
    else:
        Es, Cf = rb.bklMC(X, beta, iters, step=step, seed=17, C0=C0, quiet=True, site_pick="rank")
        wantE, wantC, _ = _oracle(ffi.rank_bklMC, g, beta, iters, step, C0, 17, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)


def test_rank_kernel_per_replica_beta_and_hook():
    """Per-replica β, a hook after every sample (the kernel pauses and rebuilds its shared-memory state from the
    configuration at every launch) and an early stop."""
    L, D, R = 6, 3, 4
    A, J = ea_instance(L, D, (-1, 1), seed=8)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_int(A, J, (-1, 1))
    betas = np.array([0.8, 1.3, 2.0, 3.0])
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(4))
    seen = []

    def hook(it, X_, C, acc, E):
        seen.append((it, np.array(E, np.float64).copy()))
        return it < 1500
    Es, Cf = rb.rrrMC(X, betas, 3000, step=250, seed=5, C0=C0, quiet=True, site_pick="rank", hook=hook)
    assert [it for it, _ in seen] == [250, 500, 750, 1000, 1250, 1500]
    for r in range(R):
        s = C0.chunks[r].copy()
        E, _ = ffi.rank_rrrMC(g, betas[r], 3000, s, ffi.PhiloxDraws(5, chain=r), step=250)
        assert np.array_equal(np.array([e[r] for _, e in seen]), E[:6])


def test_rank_kernel_statistics_vs_reference_order():
    """Same chain law as the reference-order kernel: mean energies of 96 chains agree within 3σ (rrrMC and bklMC)."""
    L, D, R, beta = 6, 3, 96, 1.2
    A, J = ea_instance(L, D, (-1, 1), seed=21)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    for fn, iters in ((rb.rrrMC, 40000), (rb.bklMC, 400000)):
        C0 = rb.Config(X.N, R, rng=np.random.default_rng(7))
        Ea, _ = fn(X, beta, iters, step=iters // 4, seed=11, C0=C0, quiet=True)
        Eb, _ = fn(X, beta, iters, step=iters // 4, seed=12, C0=C0, quiet=True, site_pick="rank")
        a, b = np.asarray(Ea)[-1], np.asarray(Eb)[-1]
        sigma = np.sqrt(a.var(ddof=1) / R + b.var(ddof=1) / R)
        assert abs(a.mean() - b.mean()) < 3 * sigma, (fn.__name__, a.mean(), b.mean(), sigma)


def test_rank_kernel_baseline_config3_size():
    """BASELINE configs[2] at size (L = 32, D = 3, β = 3): three chains, rrrMC and bklMC, bit-exact against the CPU model."""
    L, D, R, beta = 32, 3, 3, 3.0
    A, J = ea_instance(L, D, (-1, 1), seed=32)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_int(A, J, (-1, 1))
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    Es, Cf = rb.rrrMC(X, beta, 3000, step=500, seed=1, C0=C0, quiet=True, site_pick="rank")
    wantE, wantC, _ = _oracle(ffi.rank_rrrMC, g, beta, 3000, 500, C0, 1, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE) and np.array_equal(Cf.chunks, wantC)
    Es, Cf = rb.bklMC(X, beta, 6000, step=1000, seed=2, C0=C0, quiet=True, site_pick="rank")
    wantE, wantC, _ = _oracle(ffi.rank_bklMC, g, beta, 6000, 1000, C0, 2, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE) and np.array_equal(Cf.chunks, wantC)


def test_rank_kernel_rejects_what_it_cannot_take():
    X = rb.GraphEA(2, 3, replicas=2, rng=np.random.default_rng(1))       # L = 2: double bonds (repeated neighbours)
    with pytest.raises(Exception, match="site_pick = RANK"):
        rb.rrrMC(X, 1.0, 10, site_pick="rank", quiet=True)
    Xn = rb.GraphEANormal(4, 2, replicas=2, rng=np.random.default_rng(1))
    with pytest.raises(Exception, match="site_pick = RANK"):
        rb.bklMC(Xn, 1.0, 10, site_pick="rank", quiet=True)
