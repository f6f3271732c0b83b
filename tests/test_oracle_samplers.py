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

    if g.kind == ffi.QUANT:  # runtests.jl:165-190: samplers on inner_graph(X) as well
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
