"""GPU parity of GraphEANormalDiscretized (EA.jl:311-529, integer levels): interface queries and the three samplers
against the oracle on the same Philox draw stream, bit for bit (Float64 sums in the reference's order)."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu


def _pair(L, D, lev, R, seed):
    A, cJ = ea_instance(L, D, seed=seed, gaussian=True)
    return rb.GraphEANormalDiscretized(L, D, lev, replicas=R, A=A, cJ=cJ), ffi.Graph.ea_discretized(A, cJ, lev), A, cJ


@pytest.mark.parametrize("L,D,lev", [(2, 3, (-1, 0, 1)), (3, 2, (-1, 0, 1)), (4, 3, (-2, -1, 0, 1, 2)), (6, 2, (-1, 1))])
def test_interface_matches_oracle(L, D, lev):
    R = 5
    X, g, A, cJ = _pair(L, D, lev, R, seed=40 + L)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    E = np.atleast_1d(rb.energy(X, C0))
    for r in range(R):
        assert E[r] == g.energy(C0.chunks[r])
    assert np.array_equal(np.asarray(rb.allDeltaE(X), np.float64), g.allDE())
    g.energy(C0.chunks[2])
    dE = np.asarray(rb.all_delta_energy(X, C0, 2), np.float64)
    assert np.array_equal(dE, np.array([g.delta_energy(C0.chunks[2], i) for i in range(1, X.N + 1)]))
    for i in (1, X.N):
        assert tuple(rb.neighbors(X, i)) == tuple(g.neighbors(i))
        res = np.atleast_1d(rb.delta_energy_residual(X, C0, i))
        for r in range(R):
            g.energy(C0.chunks[r])
            assert res[r] == g.delta_energy_residual(C0.chunks[r], i)
    # same energy as GraphEANormal on the undiscretised couplings (up to summation order)
    Y = rb.GraphEANormal(L, D, replicas=R, A=A, J=cJ)
    assert np.allclose(np.atleast_1d(rb.energy(Y, C0)), E, rtol=0, atol=1e-10 * X.N)


@pytest.mark.parametrize("sampler", ["standardMC", "rrrMC", "rrrMC_staged", "rrrMC_eager", "bklMC"])
@pytest.mark.parametrize("L,D,lev,beta", [(3, 2, (-1, 0, 1), 2.0), (4, 3, (-1, 0, 1), 1.3), (2, 3, (-1, 0, 1), 0.7)])
def test_samplers_bit_exact_vs_oracle(sampler, L, D, lev, beta):
    R, iters, step, seed = 6, 3000, 250, 9091
    X, _, A, cJ = _pair(L, D, lev, R, seed=50 + L)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(5))
    kw = {}
    if sampler == "standardMC":
        fn, ofn, kw = rb.standardMC, ffi.standardMC, {"schedule": "random"}
    elif sampler == "bklMC":
        fn, ofn = rb.bklMC, ffi.bklMC
    else:
        fn, ofn = rb.rrrMC, ffi.rrrMC
        if sampler == "rrrMC_staged":
            kw = {"staged_thr": 1.0}
        if sampler == "rrrMC_eager":
            kw = {"staged_thr": 0.0}
    Es, Cf = fn(X, beta, iters, step=step, seed=seed, C0=C0, quiet=True, **kw)
    Es = np.asarray(Es, np.float64).reshape(-1, R)
    okw = {k: v for k, v in kw.items() if k != "schedule"}
    for r in range(R):
        g = ffi.Graph.ea_discretized(A, cJ, lev)
        s = C0.chunks[r].copy()
        want, res = ofn(g, beta, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step, **okw)
        assert np.array_equal(Es[:len(want), r], want), (sampler, r)
        assert np.array_equal(Cf.chunks[r], s), (sampler, r)


def test_energy_consistency_hook_like_the_reference_tests():
    """test/runtests.jl:12-20 on the engine: at every hook the tracked energy equals energy(X, C) recomputed."""
    R = 4
    X, g, A, cJ = _pair(4, 2, (-1, 0, 1), R, seed=77)
    bad = []

    def hook(it, X_, C, acc, E):
        e = np.array([g.energy(C.chunks[r]) for r in range(R)])
        if not np.allclose(np.atleast_1d(E), e, rtol=0, atol=1e-11 * X.N):
            bad.append((it, np.atleast_1d(E) - e))
        return True
    for fn, kw in ((rb.standardMC, {"schedule": "random"}), (rb.rrrMC, {}), (rb.rrrMC, {"staged_thr": 1.0}), (rb.bklMC, {})):
        fn(X, 2.0, 4000, step=200, seed=5, hook=hook, quiet=True, **kw)
        assert not bad, bad[:2]


def test_argument_errors():
    A, cJ = ea_instance(3, 2, seed=1, gaussian=True)
    Xf = rb.GraphEANormalDiscretized(3, 2, (-1.5, 0.0, 1.5), A=A, cJ=cJ)    # DFloat64 levels: runs in units of 1.5
    assert Xf.LEV == (-1, 0, 1) and Xf.dfloat_g == 150000
    with pytest.raises(NotImplementedError):
        rb.GraphEANormalDiscretized(3, 2, (0.00001, 1.0), A=A, cJ=cJ)       # would need couplings beyond int8
    with pytest.raises(ValueError):
        rb.GraphEANormalDiscretized(3, 2, (-1, -1, 1), A=A, cJ=cJ)
    bad = cJ.copy(); bad[0, 0] += 1.0   # breaks the symmetry of the discretised couplings
    with pytest.raises(ValueError):
        rb.GraphEANormalDiscretized(3, 2, (-1, 0, 1), A=A, cJ=bad)
