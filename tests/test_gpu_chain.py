"""GPU parity of the sequential samplers (reference order): with the shared Philox draw source each GPU chain must
reproduce the oracle's standardMC / rrrMC / bklMC trajectory bit for bit (energies at every step, final Config),
and a dumped typed-draw trace must replay to the same trajectory (north_star 'replay mode')."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu

GRAPHS = [(4, 2, (-1, 1)), (4, 3, (-1, 1)), (3, 2, (-1, 1)), (2, 3, (-1, 1)), (3, 3, (-1, 0, 1)), (6, 2, "normal"), (3, 3, "normal")]


def _mk(L, D, lev, R, seed=0):
    if lev == "normal":
        A, J = ea_instance(L, D, seed=seed, gaussian=True)
        return rb.GraphEANormal(L, D, replicas=R, A=A, J=J), ffi.Graph.ea_f64(A, J)
    A, J = ea_instance(L, D, lev, seed)
    return rb.GraphEA(L, D, lev, replicas=R, A=A, J=J), ffi.Graph.ea_int(A, J, lev)


def _oracle_run(fn, g, beta, iters, step, C0, seed, R, **kw):
    Es, Cs, res = [], [], []
    for r in range(R):
        s = C0.chunks[r].copy()
        E, info = fn(g, beta, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step, **kw)
        Es.append(E); Cs.append(s); res.append(info)
    return np.array(Es).T, np.array(Cs), res


@pytest.mark.parametrize("L,D,lev", GRAPHS)
def test_standardMC_random_site_bit_exact(L, D, lev):
    R, beta, iters, step = 5, 1.3, 3000, 100
    X, g = _mk(L, D, lev, R)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(1))
    Es, Cf = rb.standardMC(X, beta, iters, step=step, seed=4242, C0=C0, schedule="random", quiet=True)
    wantE, wantC, _ = _oracle_run(ffi.standardMC, g, beta, iters, step, C0, 4242, R)
    assert Es.shape == (iters // step, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)


@pytest.mark.parametrize("L,D,lev", GRAPHS)
@pytest.mark.parametrize("thr", [float("nan"), 0.0, 1.0])
def test_rrrMC_bit_exact(L, D, lev, thr):
    R, beta, iters, step = 4, 2.0, 2000, 50
    X, g = _mk(L, D, lev, R, seed=1)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    Es, Cf = rb.rrrMC(X, beta, iters, step=step, seed=99, C0=C0, staged_thr=thr, quiet=True)
    wantE, wantC, _ = _oracle_run(ffi.rrrMC, g, beta, iters, step, C0, 99, R, staged_thr=thr)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)


@pytest.mark.parametrize("L,D,lev", GRAPHS)
def test_bklMC_bit_exact(L, D, lev):
    R, beta, iters, step = 4, 2.0, 5000, 100
    X, g = _mk(L, D, lev, R, seed=2)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    Es, Cf = rb.bklMC(X, beta, iters, step=step, seed=7, C0=C0, quiet=True)
    wantE, wantC, _ = _oracle_run(ffi.bklMC, g, beta, iters, step, C0, 7, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)


@pytest.mark.parametrize("sampler", ["standardMC", "rrrMC", "bklMC"])
def test_hook_energy_consistency_and_stop(sampler):
    """The reference's own test (runtests.jl:12-15): tracked E equals a from-scratch energy at every hook; a hook
    returning false stops the run (RRRMC.jl:108)."""
    X, g = _mk(4, 3, (-1, 1), 6, seed=5)
    fn = {"standardMC": lambda *a, **k: rb.standardMC(*a, schedule="random", **k), "rrrMC": rb.rrrMC, "bklMC": rb.bklMC}[sampler]
    calls = []

    def hook(it, X_, C, acc, E):
        fresh = np.array([g.energy(C.chunks[r]) for r in range(6)])
        assert np.array_equal(np.asarray(E, np.float64), fresh)
        calls.append(it)
        return len(calls) < 7
    Es, Cf = fn(X, 2.0, 10_000, step=100, seed=11, hook=hook, quiet=True)
    assert calls == [100 * k for k in range(1, 8)] and len(Es) == 7


@pytest.mark.parametrize("sampler,ofn", [("standardMC", ffi.standardMC), ("rrrMC", ffi.rrrMC), ("bklMC", ffi.bklMC)])
@pytest.mark.parametrize("L,D,lev", [(4, 3, (-1, 1)), (6, 2, "normal")])
def test_replay_reference_trace(sampler, ofn, L, D, lev):
    """Replay mode: a typed draw stream recorded from the oracle (stand-in for a dump of the reference)
    fed to one GPU chain reproduces the spin trajectory bit for bit."""
    X, g = _mk(L, D, lev, 3, seed=9)
    C0 = rb.Config(X.N, 3, rng=np.random.default_rng(4))
    rec = ffi.Recorder(ffi.PhiloxDraws(31337, chain=17, tag=5))
    s = C0.chunks[1].copy()
    wantE, _ = ofn(g, 1.7, 4000, s, rec, step=40)
    kind, ival, fval = rec.arrays()
    Es, Cf = rb.replay(X, C0, sampler, 1.7, 4000, kind, ival, fval, step=40, replica=1)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks[1], s)
    assert np.array_equal(Cf.chunks[0], C0.chunks[0]) and np.array_equal(Cf.chunks[2], C0.chunks[2])


@pytest.mark.parametrize("L,D,lev", [(4, 3, (-1, 1)), (5, 2, (-1, 1)), (3, 3, (-1, 0, 1))])
def test_replay_wtm_and_extremal_opt(L, D, lev):
    """Replay entries of wtmMC (RRRMC.jl:376-430) and extremal_opt (RRRMC.jl:468-521): a typed draw stream recorded from
    the oracle drives one GPU chain to the same energies, final configuration and (extremal_opt) Emin / Cmin / itmin."""
    X, g = _mk(L, D, lev, 3, seed=5)
    C0 = rb.Config(X.N, 3, rng=np.random.default_rng(8))
    rec = ffi.Recorder(ffi.PhiloxDraws(4711, chain=3, tag=2))
    s = C0.chunks[2].copy()
    wantE, _ = ffi.wtmMC(g, 1.4, 60, s, rec, step=2.5)
    kind, ival, fval = rec.arrays()
    Es, Cf = rb.replay_wtm(X, C0, 1.4, 60, kind, ival, fval, step=2.5, replica=2)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks[2], s) and np.array_equal(Cf.chunks[0], C0.chunks[0])
    tau = 1.3
    ftau = np.cumsum(np.arange(1, X.N + 1, dtype=np.float64) ** (-tau))
    rec = ffi.Recorder(ffi.PhiloxDraws(99, chain=1, tag=7))
    s = C0.chunks[1].copy()
    wantE, wantCmin, res = ffi.extremal_opt(g, ftau, 3000, s, rec, step=100)
    kind, ival, fval = rec.arrays()
    Es, Cf, Emin, Cmin, itmin = rb.replay_extremal_opt(X, C0, tau, 3000, kind, ival, fval, step=100, replica=1)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks[1], s) and np.array_equal(Cmin, wantCmin)
    assert Emin == res.Emin and itmin == res.itmin


def test_replay_rejects_wrong_trace():
    X, g = _mk(4, 2, (-1, 1), 1)
    C0 = rb.Config(X.N, 1, rng=np.random.default_rng(4))
    with pytest.raises(rb.RRRMCError):
        rb.replay(X, C0, "standardMC", 1.0, 100, np.ones(5, np.uint8), np.zeros(5, np.int64), np.zeros(5), step=10)


def test_per_replica_beta_on_chains():
    """β is per replica on the sequential samplers (parallel-tempering ladders)."""
    X, g = _mk(4, 3, (-1, 1), 3, seed=3)
    C0 = rb.Config(X.N, 3, rng=np.random.default_rng(6))
    betas = np.array([0.5, 1.0, 2.0])
    Es, Cf = rb.standardMC(X, betas, 2000, step=500, seed=5, C0=C0, schedule="random", quiet=True)
    for r in range(3):
        s = C0.chunks[r].copy()
        E, _ = ffi.standardMC(g, betas[r], 2000, s, ffi.PhiloxDraws(5, chain=r), step=500)
        assert np.array_equal(np.asarray(Es[:, r], np.float64), E) and np.array_equal(Cf.chunks[r], s)


@pytest.mark.parametrize("sampler,ofn,iters", [("rrrMC", ffi.rrrMC, 6000), ("bklMC", ffi.bklMC, 60000)])
@pytest.mark.parametrize("L,D", [(8, 3), (6, 2), (4, 1)])
@pytest.mark.parametrize("generic", [False, True])
def test_ea_fast_path_and_generic_kernel_agree_with_oracle(sampler, ofn, iters, L, D, generic, monkeypatch):
    """GraphEA ±J runs rrrMC/bklMC on the compact-state kernel (chain_ea.cu); RRRMC_CHAIN_GENERIC forces the generic
    one. Both must reproduce the oracle bit for bit (energies at every step, final Config, accepted counters), at low
    temperature where the rejection-free samplers are meant to be used, with a hook pausing the kernels."""
    if generic:
        monkeypatch.setenv("RRRMC_CHAIN_GENERIC", "1")
    R, beta, step = 6, 2.5, iters // 12
    X, g = _mk(L, D, (-1, 1), R, seed=L + D)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(12))
    fn = {"rrrMC": rb.rrrMC, "bklMC": rb.bklMC}[sampler]
    accs = []

    def hook(it, X_, C, acc, E):
        accs.append(np.array(acc))
        return True
    Es, Cf = fn(X, beta, iters, step=step, seed=2024, C0=C0, hook=hook, quiet=True)
    wantE, wantC, res = _oracle_run(ofn, g, beta, iters, step, C0, 2024, R)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks, wantC)
    assert len(accs) == 12
    # run without a hook (single launch) must give the same trajectory
    Es2, Cf2 = fn(X, beta, iters, step=step, seed=2024, C0=C0, quiet=True)
    assert np.array_equal(Es2, Es) and Cf2 == Cf
    assert X.last_run.accepted_total == sum(r.accepted for r in res)


def test_ea_fast_path_forced_modes_and_restart():
    """Forced staged / eager modes (runtests.jl:140-191) on the fast path, then a second run continuing from the
    final configuration of the first (fresh caches, same graph object)."""
    R, beta = 3, 1.2
    X, g = _mk(6, 3, (-1, 1), R, seed=17)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    for thr in (0.0, 1.0):
        Es, Cf = rb.rrrMC(X, beta, 3000, step=100, seed=5, C0=C0, staged_thr=thr, quiet=True)
        wantE, wantC, _ = _oracle_run(ffi.rrrMC, g, beta, 3000, 100, C0, 5, R, staged_thr=thr)
        assert np.array_equal(np.asarray(Es, np.float64), wantE) and np.array_equal(Cf.chunks, wantC)
    Es2, Cf2 = rb.bklMC(X, beta, 20000, step=1000, seed=6, C0=Cf, quiet=True)
    wantE2, wantC2, _ = _oracle_run(ffi.bklMC, g, beta, 20000, 1000, Cf, 6, R)
    assert np.array_equal(np.asarray(Es2, np.float64), wantE2) and np.array_equal(Cf2.chunks, wantC2)
