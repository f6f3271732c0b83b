"""Reference test strategy (test/runtests.jl:125-191) applied to the oracle: every sampler, fresh / restart /
forced-eager / forced-staged, with the energy-consistency hook (runtests.jl:12-15); plus exact Boltzmann
stationarity on a tiny instance (the idea of RRRMC.jl:528-543,593-676) and draw-trace record/replay."""
import itertools

import numpy as np
import pytest

from oracle import ffi
from tests.helpers import bits, ea_instance, random_config, reference_graphs

BETA, ITERS, STEP = 2.0, 10_000, 100


def _check_hook(g, s):
    bad = []

    def hook(it, E, acc):
        e = g.energy(s)  # also resets the cache, exactly like the reference's checkenergy_hook
        if not abs(E - e) <= 1e-11 * max(1.0, abs(e)) + 1e-11:
            bad.append((it, E, e))
        return True
    return hook, bad


@pytest.mark.parametrize("name", list(reference_graphs().keys()))
def test_energy_consistency_all_samplers(name):
    g = reference_graphs()[name]
    src = ffi.PhiloxDraws(seed=8426732438942, chain=0)
    s = src.config(g.N)
    Es, r = ffi.standardMC(g, BETA, ITERS, s, src, step=STEP)
    assert len(Es) == ITERS // STEP and r.iters_done == ITERS
    hook, bad = _check_hook(g, s)
    ffi.standardMC(g, BETA, ITERS, s, src, step=STEP, hook=hook)
    assert not bad, bad[:3]

    Es, r = ffi.bklMC(g, BETA, ITERS, s, src, step=STEP)
    hook, bad = _check_hook(g, s)
    ffi.bklMC(g, BETA, ITERS, s, src, step=STEP, hook=hook)
    assert not bad, bad[:3]

    for thr in (float("nan"), 0.0, 1.0):  # default, always eager, always staged (runtests.jl:153-163)
        hook, bad = _check_hook(g, s)
        Es, r = ffi.rrrMC(g, BETA, ITERS, s, src, step=STEP, hook=hook, staged_thr=thr)
        assert not bad, (thr, bad[:3])
        if thr == 0.0:
            assert r.staged_its == 0
        if thr == 1.0:  # acc_rate can round to exactly 1.0 for tiny N (λ=5/N), where `<` fails — reference behaviour
            assert r.staged_its > 0

    if g.kind in (ffi.QUANT, ffi.EA_DISCR):  # runtests.jl:165-190: samplers on inner_graph(X) as well
        g0 = g.inner()
        hook, bad = _check_hook(g0, s)
        ffi.bklMC(g0, BETA, ITERS, s, src, step=STEP, hook=hook)
        ffi.rrrMC(g0, BETA, ITERS, s, src, step=STEP, hook=hook)
        assert not bad


def test_hook_can_stop_and_sampling_is_before_move():
    A, J = ea_instance(3, 2)
    g = ffi.Graph.ea_int(A, J)
    src = ffi.PhiloxDraws(1)
    s = src.config(g.N)
    e0 = g.energy(s)
    seen = []
    Es, r = ffi.standardMC(g, 1.0, 1000, s, src, step=1, hook=lambda it, E, acc: (seen.append(it), it < 5)[1])
    assert seen == [1, 2, 3, 4, 5] and r.iters_done == 5
    assert Es[0] == e0  # sample at it=1 is taken before the first move (RRRMC.jl:101-109)


def _boltzmann(g, N, beta):
    E = np.zeros(2 ** N)
    for c in range(2 ** N):
        E[c] = g.energy(np.array([c], np.uint64))
    p = np.exp(-beta * (E - E.min()))
    return p / p.sum()


@pytest.mark.parametrize("sampler", ["standard", "rrr", "rrr_staged", "bkl"])
def test_boltzmann_stationarity_tiny_EA(sampler):
    """All samplers target the Boltzmann distribution (truep, RRRMC.jl:528-543): 3x3 EA ±J, N=9, χ² on 512 states."""
    A, J = ea_instance(3, 2, seed=11)
    g = ffi.Graph.ea_int(A, J)
    N, beta = 9, 0.7
    p = _boltzmann(g, N, beta)
    src = ffi.PhiloxDraws(12345, chain=7)
    s = src.config(N)
    counts = np.zeros(2 ** N)
    weights = np.zeros(2 ** N)

    def hook(it, E, acc):
        counts[int(s[0])] += 1
        return True
    iters, step = 400_000, 4
    if sampler == "standard":
        ffi.standardMC(g, beta, iters, s, src, step=step, hook=hook)
    elif sampler == "rrr":
        ffi.rrrMC(g, beta, iters, s, src, step=step, hook=hook)
    elif sampler == "rrr_staged":
        ffi.rrrMC(g, beta, iters, s, src, step=step, hook=hook, staged_thr=1.0)
    else:
        ffi.bklMC(g, beta, iters, s, src, step=step, hook=hook)
    n = counts.sum()
    # samples are correlated: compare with a loose relative tolerance on well-populated states
    big = p > 2e-3
    rel = np.abs(counts[big] / n - p[big]) / p[big]
    assert rel.max() < 0.15, rel.max()
    assert abs((counts / n) @ np.arange(2 ** N) - p @ np.arange(2 ** N)) < 6.0


def test_boltzmann_stationarity_quant():
    """rrrMC(::DoubleGraph) (RRRMC.jl:221-290) samples exp(-βE) of the full GraphQuant energy: Nk=3, M=3."""
    Nk, M, beta = 3, 3, 1.1
    Jsk = np.array([[0, 1, 0], [1, 0, 1], [0, 1, 0]], np.uint8)
    g = ffi.Graph.quant(Nk, M, 0.8, beta, ffi.SK_BIN, Jsk)
    N = Nk * M
    p = _boltzmann(g, N, beta)
    src = ffi.PhiloxDraws(99, chain=1)
    s = src.config(N)
    counts = np.zeros(2 ** N)

    def hook(it, E, acc):
        counts[int(s[0])] += 1
        return True
    ffi.rrrMC(g, beta, 600_000, s, src, step=3, hook=hook)
    n = counts.sum()
    big = p > 2e-3
    rel = np.abs(counts[big] / n - p[big]) / p[big]
    assert rel.max() < 0.15, rel.max()


@pytest.mark.parametrize("sampler", ["standard", "rrr", "bkl"])
def test_trace_record_replay_bit_exact(sampler):
    """A recorded typed-draw trace replays to the same Es and final Config."""
    A, J = ea_instance(4, 3, seed=5)
    g = ffi.Graph.ea_int(A, J)
    fn = {"standard": ffi.standardMC, "rrr": ffi.rrrMC, "bkl": ffi.bklMC}[sampler]
    src = ffi.PhiloxDraws(2024, chain=3)
    s0 = src.config(g.N)
    rec = ffi.Recorder(src)
    s = s0.copy()
    Es, r = fn(g, 1.5, 5000, s, rec, step=50)
    kind, ival, fval = rec.arrays()
    assert len(kind) > 0
    rep = ffi.Replayer(kind, ival, fval)
    s2 = s0.copy()
    Es2, r2 = fn(g, 1.5, 5000, s2, rep, step=50)
    assert rep.error == 0 and rep.consumed == len(kind)
    assert np.array_equal(Es, Es2) and np.array_equal(s, s2) and r.accepted == r2.accepted
    # draw-order contract for standardMC (Appendix A.8): RANGE, then FLOAT only if ΔE>0
    if sampler == "standard":
        assert kind[0] == 0
        assert (np.diff(np.flatnonzero(kind == 0)) <= 2).all()


def test_discretized_splits_couplings_and_energy():
    """GraphEANormalDiscretized (EA.jl:311-344): levels + residuals add up to the continuous couplings (discretize,
    Common.jl:38-49: nearest level, first wins ties), energy = inner + residual = the GraphEANormal energy of cJ, and
    delta_energy = energy(flipped) - energy (the generic definition, Interface.jl:130-138)."""
    A, cJ = ea_instance(3, 3, seed=31, gaussian=True)
    g = ffi.Graph.ea_discretized(A, cJ, (-1, 0, 1))
    gn = ffi.Graph.ea_f64(A, cJ)
    assert np.array_equal(g.allDE(), ffi.Graph.ea_int(A, np.zeros_like(cJ, dtype=np.int64), (-1, 0, 1)).allDE())
    s = random_config(g.N, seed=2)
    E = g.energy(s)
    assert abs(E - gn.energy(s)) < 1e-12 * g.N
    assert abs(E - (g.inner().energy(s) + (E - g.inner().energy(s)))) == 0
    g.energy(s)
    for i in range(1, g.N + 1):
        d = g.delta_energy(s, i)
        t = s.copy(); t[(i - 1) >> 6] ^= np.uint64(1 << ((i - 1) & 63))
        assert abs(d - (gn.energy(t) - gn.energy(s))) < 1e-11
    assert tuple(g.neighbors(1)) == tuple(gn.neighbors(1))


def test_boltzmann_stationarity_discretized():
    """rrrMC(::DoubleGraph) on GraphEANormalDiscretized samples exp(-βE) of the FULL energy (levels + residuals)."""
    A, cJ = ea_instance(3, 2, seed=12, gaussian=True)
    g = ffi.Graph.ea_discretized(A, cJ, (-1, 0, 1))
    N, beta = 9, 0.8
    p = _boltzmann(g, N, beta)
    for thr in (float("nan"), 1.0):
        src = ffi.PhiloxDraws(77, chain=2)
        s = src.config(N)
        counts = np.zeros(2 ** N)

        def hook(it, E, acc):
            counts[int(s[0])] += 1
            return True
        ffi.rrrMC(g, beta, 3_000_000, s, src, step=4, hook=hook, staged_thr=thr)
        n = counts.sum()
        big = p > 2e-3
        rel = np.abs(counts[big] / n - p[big]) / p[big]
        assert rel.max() < 0.1, (thr, rel.max())


@pytest.mark.parametrize("name", list(reference_graphs().keys()))
def test_wtmMC_energy_consistency(name):
    """wtmMC (RRRMC.jl:376-430) under the reference's check-energy hook (test/runtests.jl:140-150)."""
    g = reference_graphs()[name]
    src = ffi.PhiloxDraws(seed=77, chain=3)
    s = src.config(g.N)
    hook, bad = _check_hook(g, s)
    Es, r = ffi.wtmMC(g, BETA, 80, s, src, step=0.8 * g.N, hook=hook)
    assert not bad, bad[:3]
    assert len(Es) == 80 and r.iters_done > 0


def test_wtmMC_boltzmann_stationarity():
    """Samples taken at equal intervals of the global time of the waiting-time method follow exp(-βE)."""
    A, J = ea_instance(3, 2, seed=11)
    g = ffi.Graph.ea_int(A, J)
    N, beta = 9, 0.7
    p = _boltzmann(g, N, beta)
    src = ffi.PhiloxDraws(12345, chain=7)
    s = src.config(N)
    counts = np.zeros(2 ** N)

    def hook(it, E, acc):
        counts[int(s[0])] += 1
        return True
    ffi.wtmMC(g, beta, 200_000, s, src, step=float(N), hook=hook)
    n = counts.sum()
    big = p > 2e-3
    rel = np.abs(counts[big] / n - p[big]) / p[big]
    assert rel.max() < 0.12, rel.max()


# ---- extremal_opt (RRRMC.jl:468-521, EOCache DeltaE.jl:413-543) -------------------------------------------------
@pytest.mark.parametrize("name", ["EA(2,3)", "EA(3,2)", "EA(3,2,(-1,0,1))", "RRG(10,3)", "RRG(10,3,(-1,0,1))"])
def test_extremal_opt_energy_consistency(name):
    """E tracked through the ranked moves ≡ energy(X, C) at every hook (the invariant of test/runtests.jl:12-20, which
    the reference applies to its Monte Carlo samplers); Emin/Cmin/itmin are what the run actually visited."""
    g = reference_graphs()[name]
    N = g.N
    s = random_config(N, 5)
    ftau = np.cumsum(np.arange(1, N + 1, dtype=np.float64) ** -1.3)
    g2 = reference_graphs()[name]
    seen = []

    def hook(it, E, Emin):
        assert E == g2.energy(s.copy()) and Emin <= E
        seen.append((it, E))
        return True
    Es, Cmin, res = ffi.extremal_opt(g, ftau, 400, s, ffi.PhiloxDraws(3, 0), step=4, hook=hook)
    assert len(seen) == 100 and [e for _, e in seen] == list(Es)
    assert res.Efinal == g2.energy(s.copy()) and g2.energy(Cmin.copy()) == res.Emin
    assert res.Emin <= Es.min() and 0 <= res.itmin <= 400


def test_extremal_opt_rank_distribution():
    """On a graph with no bonds to break the ranking (GraphQT with fourK = 0 has every ΔE = 0: one class), τ-EO picks
    sites uniformly; with distinct classes the lowest-ΔE class must be picked with probability fτ[n₁]/z."""
    A, J = ea_instance(4, 2, seed=2)
    g = ffi.Graph.ea_int(A, J)
    N = g.N
    tau = 1.5
    ftau = np.cumsum(np.arange(1, N + 1, dtype=np.float64) ** -tau)
    hits, tot = 0, 4000
    for k in range(tot):
        s = random_config(N, 77)      # same start each time, one move, different draws
        dE = np.array([g.delta_energy(s, i) for i in range(1, N + 1)]) if g.energy(s) is not None else None
        n1 = int((dE == dE.min()).sum())
        Es, _, res = ffi.extremal_opt(g, ftau, 2, s, ffi.PhiloxDraws(1000 + k, 0), step=1)
        hits += (Es[1] - Es[0]) == dE.min()
    p = ftau[n1 - 1] / ftau[-1]
    assert abs(hits / tot - p) < 4 * np.sqrt(p * (1 - p) / tot)
