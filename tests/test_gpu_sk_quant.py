"""GPU parity for the fully connected and Suzuki-Trotter families: GraphSK, GraphSKNormal,
GraphQT, GraphQuant over {Empty, SK, SKNormal} — the instances of test/runtests.jl:66-67, 78-81 plus larger ones.
Interface queries must equal the oracle's (integers bit-exact, Float64 to 1e-6 relative per north_star; the chain
path is in fact bit-exact because it keeps the reference's summation order), and every sampler must reproduce the
oracle trajectory bit for bit when both consume the same Philox draw stream."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance, sk_binary, sk_gauss

pytestmark = pytest.mark.gpu


def _mk(name, R, seed=0):
    """-> (engine graph, oracle graph factory)"""
    if name.startswith("SKNormal"):
        N = int(name.split("(")[1][:-1]); J = sk_gauss(N, seed)
        return rb.GraphSKNormal(N, replicas=R, J=J), (lambda: ffi.Graph.sk_f64(J))
    if name.startswith("SK("):
        N = int(name.split("(")[1][:-1]); J = sk_binary(N, seed)
        return rb.GraphSK(N, replicas=R, J=J), (lambda: ffi.Graph.sk_bin(J))
    if name.startswith("QT"):
        N, M = [int(v) for v in name.split("(")[1][:-1].split(",")]
        return rb.GraphQT(N, M, 0.73, replicas=R), (lambda: ffi.Graph.qt(N, M, 0.73))
    if name.startswith("QEAT"):   # GraphQEAT(L, D, M): GraphQuant over GraphEANormal (QAliases.jl:51-81)
        L, D, M = [int(v) for v in name.split("(")[1][:-1].split(",")]
        A, J = ea_instance(L, D, seed=seed + 6, gaussian=True)
        return (rb.GraphQEAT(L, D, M, 0.5, 2.0, replicas=R, A=A, J=J),
                (lambda: ffi.Graph.quant(L ** D, M, 0.5, 2.0, ffi.EA_F64, J, A)))
    # Quant(Nk,M,inner)
    a = name.split("(")[1][:-1].split(",")
    Nk, M, inner = int(a[0]), int(a[1]), a[2]
    G, b = 0.5, 2.0
    if inner == "Empty":
        return rb.GraphQ0T(Nk, M, G, b, replicas=R), (lambda: ffi.Graph.quant(Nk, M, G, b, ffi.EMPTY))
    if inner == "SK":
        J = sk_binary(Nk, seed + 3)
        return rb.GraphQSKT(Nk, M, G, b, replicas=R, J=J), (lambda: ffi.Graph.quant(Nk, M, G, b, ffi.SK_BIN, J))
    J = sk_gauss(Nk, seed + 4)
    return rb.GraphQSKNormalT(Nk, M, G, b, replicas=R, J=J), (lambda: ffi.Graph.quant(Nk, M, G, b, ffi.SK_F64, J))


GRAPHS = ["SK(10)", "SKNormal(10)", "SK(37)", "SKNormal(33)", "QT(24,4)", "Quant(10,8,Empty)", "Quant(10,8,SK)",
          "Quant(10,8,SKNormal)", "Quant(17,5,SK)", "QEAT(3,2,5)", "QEAT(2,3,4)", "QEAT(4,3,6)"]


@pytest.mark.parametrize("name", GRAPHS)
def test_interface_queries_match_oracle(name):
    R = 6
    X, mk = _mk(name, R)
    g = mk()
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    E = np.atleast_1d(rb.energy(X, C0))
    for r in range(R):
        assert E[r] == g.energy(C0.chunks[r])
    for i in (1, X.N // 2, X.N):
        dE = np.atleast_1d(rb.delta_energy(X, C0, i))
        res = np.atleast_1d(rb.delta_energy_residual(X, C0, i))
        for r in range(R):
            g.energy(C0.chunks[r])
            assert dE[r] == g.delta_energy(C0.chunks[r], i)
            assert res[r] == g.delta_energy_residual(C0.chunks[r], i)
        assert tuple(g.neighbors(i)) == rb.neighbors(X, i)
    g.energy(C0.chunks[2])
    all_dE = rb.all_delta_energy(X, C0, 2)
    assert all(all_dE[i] == g.delta_energy(C0.chunks[2], i + 1) for i in range(X.N))
    if name.startswith(("QT", "Quant", "QEAT")):
        assert tuple(g.allDE()) == rb.allDeltaE(X)
    # ΔE ≡ energy(flipped) − energy (generic fallback Interface.jl:130-138), to rounding
    i = 3
    Cf = C0.copy(); Cf.chunks[:, 0] ^= np.uint64(1 << (i - 1))
    d = np.atleast_1d(rb.energy(X, Cf)) - E
    assert np.allclose(d, np.atleast_1d(rb.delta_energy(X, C0, i)), rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("name", ["Quant(10,8,Empty)", "Quant(10,8,SK)", "Quant(10,8,SKNormal)", "Quant(9,7,SK)", "QEAT(3,2,6)"])
def test_quant_observables_match_oracle(name):
    R = 4
    X, mk = _mk(name, R)
    g = mk()
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(5))
    tm = rb.transverse_mag(X, C0, 2.0); qe = rb.Qenergy(X, C0); re = rb.Renergies(X, C0); ov = rb.overlaps(X, C0)
    for r in range(R):
        s = C0.chunks[r]
        g.energy(s)
        assert np.isclose(tm[r], ffi.lib().orc_transverse_mag(g.h, s, 2.0), rtol=1e-13)
        assert np.isclose(qe[r], ffi.lib().orc_Qenergy(g.h, s), rtol=1e-12, atol=1e-12)
        want = np.zeros(X.M); ffi.lib().orc_Renergies(g.h, want)
        assert np.array_equal(re[r], want)
        wo = np.zeros(X.M // 2); ffi.lib().orc_overlaps(g.h, wo)
        assert np.array_equal(ov[r], wo)


def _oracle_run(fn, mk, beta, iters, step, C0, seed, R, **kw):
    Es, Cs = [], []
    for r in range(R):
        g = mk()
        s = C0.chunks[r].copy()
        E, _ = fn(g, beta, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step, **kw)
        Es.append(E); Cs.append(s)
    return np.array(Es).T, np.array(Cs)


@pytest.mark.parametrize("name", GRAPHS)
def test_standardMC_bit_exact(name):
    R, beta, iters, step = 4, 1.1, 3000, 100
    X, mk = _mk(name, R, seed=1)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(1))
    Es, Cf = rb.standardMC(X, beta, iters, step=step, seed=4242, C0=C0, quiet=True)
    wantE, wantC = _oracle_run(ffi.standardMC, mk, beta, iters, step, C0, 4242, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)


@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("thr", [float("nan"), 0.0, 1.0])
def test_rrrMC_bit_exact(name, thr):
    R, beta, iters, step = 3, 2.0, 1500, 50
    X, mk = _mk(name, R, seed=2)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    Es, Cf = rb.rrrMC(X, beta, iters, step=step, seed=99, C0=C0, staged_thr=thr, quiet=True)
    wantE, wantC = _oracle_run(ffi.rrrMC, mk, beta, iters, step, C0, 99, R, staged_thr=thr)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)


@pytest.mark.parametrize("name", GRAPHS)
def test_bklMC_bit_exact(name):
    R, beta, iters, step = 3, 2.0, 4000, 100
    X, mk = _mk(name, R, seed=3)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(4))
    Es, Cf = rb.bklMC(X, beta, iters, step=step, seed=7, C0=C0, quiet=True)
    wantE, wantC = _oracle_run(ffi.bklMC, mk, beta, iters, step, C0, 7, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)


@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("sampler", ["standardMC", "rrrMC", "bklMC"])
def test_checkenergy_hook(name, sampler):
    """checkenergy_hook of test/runtests.jl:12-15: the tracked energy equals a from-scratch energy(X, C) at every
    sample (taken before the move of that iteration), for every sampler."""
    R = 3
    X, _ = _mk(name, R, seed=5)
    n = [0]

    def hook(it, X_, C, acc, E):
        n[0] += 1
        fresh = np.atleast_1d(rb.energy(X_, rb.Config(C.N, C.R, chunks=C.chunks)))
        assert np.allclose(fresh, E, atol=1e-11 * max(1.0, X_.N)), (it, fresh, E)
        return True
    getattr(rb, sampler)(X, 2.0, 2000, step=250, seed=11, hook=hook, quiet=True)
    assert n[0] == 8


def test_hook_and_restart_on_quant():
    """restart from C0=C with the energy check hook (test/runtests.jl:141-147) on GraphQSKT."""
    R = 3
    X, mk = _mk("Quant(10,8,SK)", R)
    Es, C1 = rb.rrrMC(X, 2.0, 2000, step=100, seed=5, quiet=True)
    seen = []

    def hook(it, X_, C, acc, E):
        seen.append(it)
        assert np.allclose(np.atleast_1d(rb.energy(X_, rb.Config(C.N, C.R, chunks=C.chunks))), E, atol=1e-9)
        return it < 600
    Es2, C2 = rb.rrrMC(X, 2.0, 2000, step=100, seed=6, C0=C1, hook=hook, quiet=True)
    assert seen == [100, 200, 300, 400, 500, 600] and len(Es2) == 6


def test_sk_larger_batch_statistics():
    """SKNormal N=256 × 64 replicas: Metropolis lowers the energy per spin towards the SK value; every replica's
    tracked energy equals a fresh recompute at the sample (relative 1e-9)."""
    N, R = 256, 64
    X = rb.GraphSKNormal(N, replicas=R, rng=np.random.default_rng(0))

    def hook(it, X_, C, acc, E):
        assert np.allclose(rb.energy(X_, rb.Config(C.N, C.R, chunks=C.chunks)), E, rtol=1e-9)
        return True
    Es, Cf = rb.standardMC(X, 1.5, 40 * N, step=40 * N, seed=3, hook=hook, quiet=True)
    e = Es[-1] / N
    assert -0.80 < e.mean() < -0.55
