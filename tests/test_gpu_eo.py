"""GPU parity of extremal_opt (RRRMC.jl:468-521 on the EOCache of DeltaE.jl:413-543) on the DiscrGraph families: the
chain kernel and the oracle consume the same Philox draw stream, so the energies at every hook instant, the final
configuration, Emin, Cmin and itmin must agree bit for bit; plus the hook contract and the argument errors."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu


def _mk(name, R):
    if name == "EA(4,3)":
        A, J = ea_instance(4, 3, seed=3)
        return rb.GraphEA(4, 3, replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J))
    if name == "EA(2,3)":   # L = 2: doubled bonds (EA.jl:24-43)
        A, J = ea_instance(2, 3, seed=8)
        return rb.GraphEA(2, 3, replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J))
    if name == "EA(6,2,(-1,0,1))":
        A, J = ea_instance(6, 2, (-1, 0, 1), seed=4)
        return rb.GraphEA(6, 2, (-1, 0, 1), replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J, (-1, 0, 1)))
    if name == "EA(5,2,(-2,-1,1,2))":
        A, J = ea_instance(5, 2, (-2, -1, 1, 2), seed=5)
        return rb.GraphEA(5, 2, (-2, -1, 1, 2), replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J, (-2, -1, 1, 2)))
    if name == "QT(12,4)":
        return rb.GraphQT(12, 4, 0.73, replicas=R), (lambda: ffi.Graph.qt(12, 4, 0.73))
    if name.startswith("RRG"):   # odd degree: allΔE has no zero (K = 2L classes); zero level: neighbors() skips J = 0
        _, N, K, lev = name.split(":")
        N, K = int(N), int(K)
        lev = {"pm1": (-1, 1), "z": (-1, 0, 1)}[lev]
        rng = np.random.default_rng(100 * N + K)
        A = rb.gen_RRG(N, K, rng)
        J = rb.gen_J_graph(lambda n: rng.choice(np.asarray(lev, np.float64), n), A).astype(np.int64)
        return rb.GraphRRG(N, K, lev, replicas=R, A=A, J=J), (lambda: ffi.Graph.rrg_int(A, J, lev))
    raise KeyError(name)


GRAPHS = ["EA(4,3)", "EA(2,3)", "EA(6,2,(-1,0,1))", "EA(5,2,(-2,-1,1,2))", "QT(12,4)", "RRG:40:3:pm1", "RRG:30:4:z", "RRG:50:5:pm1"]


@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("tau,step", [(1.3, 7), (2.2, 1)])
def test_extremal_opt_bit_exact_vs_oracle(name, tau, step):
    R, iters, seed = 5, 900, 4242
    X, mk = _mk(name, R)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    ftau = rb.eo_ftau(X.N, tau)
    Cf, Emin, Cmin, itmin, Es = rb.extremal_opt(X, tau, iters, step=step, seed=seed, C0=C0, quiet=True, return_Es=True)
    Es = np.asarray(Es, np.float64).reshape(-1, R)
    assert Es.shape[0] == iters // step
    for r in range(R):
        g = mk()
        s = C0.chunks[r].copy()
        want, cmin, res = ffi.extremal_opt(g, ftau, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step)
        assert np.array_equal(Es[:, r], want), (name, r)
        assert np.array_equal(Cf.chunks[r], s), (name, r)
        assert float(np.atleast_1d(Emin)[r]) == res.Emin and int(np.atleast_1d(itmin)[r]) == res.itmin, (name, r)
        assert np.array_equal(Cmin.chunks[r], cmin), (name, r)
        assert g.energy(cmin) == res.Emin
    assert X.last_run.iters_done == iters


def test_extremal_opt_per_chain_tau_and_larger_lattice():
    R, iters = 8, 20000
    X, mk = _mk("EA(4,3)", R)
    A, J = ea_instance(8, 3, seed=11)
    X = rb.GraphEA(8, 3, replicas=R, A=A, J=J)
    taus = np.linspace(1.1, 2.5, R)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(5))
    Cf, Emin, Cmin, itmin = rb.extremal_opt(X, taus, iters, step=iters, seed=77, C0=C0, quiet=True)
    for r in (0, 3, 7):
        g = ffi.Graph.ea_int(A, J)
        s = C0.chunks[r].copy()
        _, cmin, res = ffi.extremal_opt(g, rb.eo_ftau(X.N, taus[r]), iters, s, ffi.PhiloxDraws(77, chain=r), step=iters)
        assert Emin[r] == res.Emin and itmin[r] == res.itmin
        assert np.array_equal(Cmin.chunks[r], cmin) and np.array_equal(Cf.chunks[r], s)
    # τ-EO finds states far below a random configuration's energy (≈ 0): a sanity bound, not a parity statement
    assert (np.asarray(Emin) < -1.4 * X.N).all()


def test_extremal_opt_hook_contract():
    """hook(it, X, C, E, Emin) before the move of iteration it (RRRMC.jl:497-501); false stops the run; E tracks energy(X, C)."""
    R = 3
    X, mk = _mk("EA(6,2,(-1,0,1))", R)
    g = mk()
    seen, bad = [], []

    def hook(it, X_, C, E, Emin):
        seen.append(it)
        e = np.array([g.energy(C.chunks[r]) for r in range(R)])
        if not np.array_equal(np.atleast_1d(E), e) or (np.atleast_1d(Emin) > np.atleast_1d(E)).any():
            bad.append((it, E, e, Emin))
        return len(seen) < 6
    Cf, Emin, Cmin, itmin = rb.extremal_opt(X, 1.4, 1000, step=50, seed=9, hook=hook, quiet=True)
    assert not bad, bad[:2]
    assert seen == [50, 100, 150, 200, 250, 300]
    assert (np.asarray(itmin) < 300).all()
    for r in range(R):
        assert g.energy(Cmin.chunks[r]) == np.atleast_1d(Emin)[r]


def test_extremal_opt_argument_errors():
    X, _ = _mk("EA(4,3)", 2)
    with pytest.raises(ValueError):
        rb.extremal_opt(X, 1.3, 10, step=0, quiet=True)
    with pytest.raises(ValueError):
        rb.extremal_opt(X, 1.3, 10, C0=rb.Config(X.N + 1, 2), quiet=True)
    with pytest.raises(ValueError):
        rb.extremal_opt(X, 1.3, 10, ftau=np.ones(X.N + 3), quiet=True)
    with pytest.raises((ValueError, rb.RRRMCError)):   # decreasing table
        rb.extremal_opt(X, 1.3, 10, ftau=np.linspace(2, 1, X.N), quiet=True)
    Y = rb.GraphQSKT(8, 4, 0.3, 1.0, replicas=2, rng=np.random.default_rng(1))   # a DoubleGraph: not on this path
    with pytest.raises(NotImplementedError):
        rb.extremal_opt(Y, 1.3, 10, quiet=True)


@pytest.mark.parametrize("name", ["EANormal(4,3)", "EANormal(6,2)", "SKNormal(48)"])
@pytest.mark.parametrize("tau,step", [(1.3, 10), (2.0, 1)])
def test_extremal_opt_cont_bit_exact_vs_oracle(name, tau, step):
    """EOCacheCont (DeltaE.jl:555-635): extremal_opt on the Float64 SimpleGraphs — ΔEs of every spin kept sorted, one
    draw per move. Bit-exact against the oracle's restatement (energies at the sampling instants, final configuration,
    Emin, itmin, Cmin)."""
    R, iters, seed = 4, 600, 777
    if name.startswith("EANormal"):
        L, D = (4, 3) if "4,3" in name else (6, 2)
        A, J = ea_instance(L, D, seed=L, gaussian=True)
        X, mk = rb.GraphEANormal(L, D, replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_f64(A, J))
    else:
        Jm = rb.gen_J_gauss(48, np.random.default_rng(5))
        X, mk = rb.GraphSKNormal(48, replicas=R, J=Jm), (lambda: ffi.Graph.sk_f64(Jm))
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    ftau = rb.eo_ftau(X.N, tau)
    Cf, Emin, Cmin, itmin, Es = rb.extremal_opt(X, tau, iters, step=step, seed=seed, C0=C0, quiet=True, return_Es=True)
    Es = np.asarray(Es, np.float64).reshape(-1, R)
    for r in range(R):
        g = mk()
        s = C0.chunks[r].copy()
        want, cmin, res = ffi.extremal_opt(g, ftau, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step)
        assert np.array_equal(Es[:, r], want), (name, r)
        assert np.array_equal(Cf.chunks[r], s), (name, r)
        assert float(np.atleast_1d(Emin)[r]) == res.Emin and int(np.atleast_1d(itmin)[r]) == res.itmin, (name, r)
        assert np.array_equal(Cmin.chunks[r], cmin), (name, r)
