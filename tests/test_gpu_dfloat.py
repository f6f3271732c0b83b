"""GraphEA with fractional levels — the reference's DFloat64 path (EA.jl:191, src/DFloats.jl: five-digit fixed point, exact
integer arithmetic). The engine runs the equivalent integer-level graph in units of u = g/10^5 at β·u; energies come back as
Float64(DFloat64) = units·g/10^5. Checked against the oracle's integer-level graph, chain by chain."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu
LEV, ILEV, G = (-1.5, 0.5, 1.5), (-3, 1, 3), 50000


def _mk(R):
    A, Ji = ea_instance(4, 2, ILEV, seed=31)
    Jreal = Ji * (G / 1e5)
    return rb.GraphEA(4, 2, LEV, replicas=R, A=A, J=Jreal), A, Ji


def test_dfloat_interface():
    R = 3
    X, A, Ji = _mk(R)
    assert X.ET is float and X.LEV == ILEV
    g = ffi.Graph.ea_int(A, Ji, ILEV)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(1))
    E = np.atleast_1d(rb.energy(X, C0))
    for r in range(R):
        assert E[r] == g.energy(C0.chunks[r]) * G / 1e5
    g.energy(C0.chunks[1])
    want = np.array([g.delta_energy(C0.chunks[1], i) for i in range(1, X.N + 1)]) * G / 1e5
    assert np.array_equal(rb.all_delta_energy(X, C0, 1), want)
    assert rb.allDeltaE(X) == tuple(float(v) * G / 1e5 for v in g.allDE())
    assert rb.allDeltaE(X) == (0.0, 2.0, 4.0, 6.0, 8.0, 10.0, 12.0)   # 2·|Σ of four couplings from ±{0.5, 1.5}|


@pytest.mark.parametrize("sampler", ["standardMC", "rrrMC", "bklMC"])
def test_dfloat_samplers_match_integer_unit_oracle(sampler):
    R, beta, iters, step, seed = 3, 0.9, 1500, 50, 5
    X, A, Ji = _mk(R)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    kw = {"schedule": "random"} if sampler == "standardMC" else {}
    Es, Cf = getattr(rb, sampler)(X, beta, iters, step=step, seed=seed, C0=C0, quiet=True, **kw)
    Es = np.asarray(Es, np.float64).reshape(-1, R)
    for r in range(R):
        g = ffi.Graph.ea_int(A, Ji, ILEV)
        s = C0.chunks[r].copy()
        want, _ = getattr(ffi, sampler)(g, beta * (G / 1e5), iters, s, ffi.PhiloxDraws(seed, chain=r), step=step)
        assert np.array_equal(Es[:, r], want * G / 1e5), (sampler, r)
        assert np.array_equal(Cf.chunks[r], s), (sampler, r)


def test_dfloat_rejects_bad_levels():
    with pytest.raises(ValueError):
        rb.GraphEA(4, 2, (0.123456, 1.0))
    with pytest.raises(ValueError):
        A, Ji = ea_instance(4, 2, ILEV, seed=31)
        rb.GraphEA(4, 2, LEV, A=A, J=Ji * 0.7)           # couplings that are not levels (EA.jl:161)


# ---- against the oracle's restatement of src/DFloats.jl, and on the other graph families ----
def test_dfloat_arithmetic_restatement():
    """orc_dfloat_*: convert = round(x·10^5) to nearest even, Float64() = i/10^5, / Integer truncates (DFloats.jl:24-39)."""
    assert ffi.lib().orc_dfloat_from_f64(1.5) == 150000 and ffi.lib().orc_dfloat_from_f64(-0.000005) == 0
    assert ffi.lib().orc_dfloat_from_f64(0.000015) == 2 and ffi.lib().orc_dfloat_from_f64(0.12345) == 12345
    assert ffi.lib().orc_dfloat_to_f64(-350000) == -3.5
    assert ffi.lib().orc_dfloat_div_int(-7, 2) == -3 and ffi.lib().orc_dfloat_div_int(7, 2) == 3


def test_dfloat_energy_equals_dfloat_arithmetic():
    """GraphEA with levels (-1.5, 0.5, 1.5) and (0.25, -0.75): energies and ΔE equal the DFloat64 arithmetic of the
    reference (every coupling converted to round(J·10^5), integer sums, Float64 = sum/10^5), not just the engine's own
    integer-unit reformulation."""
    for LEVr, seed in (((-1.5, 0.5, 1.5), 31), ((0.25, -0.75), 7), ((-0.5, 0.25, 1.25), 9)):
        ilev, g = rb.interface.dfloat_levels(LEVr)
        A, Ji = ea_instance(4, 3, ilev, seed=seed)
        Jreal = Ji * (g / 1e5)
        X = rb.GraphEA(4, 3, LEVr, replicas=4, A=A, J=Jreal)
        C0 = rb.Config(X.N, 4, rng=np.random.default_rng(seed))
        E = np.atleast_1d(rb.energy(X, C0))
        for r in range(4):
            want, lf2 = ffi.dfloat_ea_energy(A, Jreal, C0.chunks[r], fields=True)
            assert E[r] == want
            assert np.array_equal(rb.all_delta_energy(X, C0, r), -lf2)      # ΔE_i = -lfields[i] (EA.jl:274)


def test_dfloat_rrg_and_discretized():
    """DFloat64 levels on GraphRRG (RRG.jl:162) and on the discretised DoubleGraphs (EA.jl:360, RRG.jl:330)."""
    LEVr = (-1.5, 0.5, 1.5)
    ilev, g = rb.interface.dfloat_levels(LEVr)
    u = g / 1e5
    rng = np.random.default_rng(3)
    X = rb.GraphRRG(40, 3, LEVr, replicas=3, rng=rng)
    assert X.ET is float and X.LEV == ilev
    C0 = rb.Config(X.N, 3, rng=np.random.default_rng(4))
    E = np.atleast_1d(rb.energy(X, C0))
    Jreal = X.J * u
    for r in range(3):
        assert E[r] == ffi.dfloat_ea_energy(X.A, Jreal, C0.chunks[r])
    Es, Cf = rb.rrrMC(X, 1.1, 2000, step=200, seed=3, C0=C0, quiet=True)
    for r in range(3):
        assert np.atleast_1d(rb.energy(X, Cf))[r] == ffi.dfloat_ea_energy(X.A, Jreal, Cf.chunks[r])
    # discretised graphs: the same energy as the continuous graph on cJ (the levels and residuals add up to cJ)
    for mk, mkn in ((lambda **k: rb.GraphEANormalDiscretized(4, 3, LEVr, replicas=3, **k), lambda A, J: rb.GraphEANormal(4, 3, replicas=3, A=A, J=J)),):
        Xd = mk(rng=np.random.default_rng(5))
        Xn = mkn(Xd.A, Xd.cJ)
        Cd = rb.Config(Xd.N, 3, rng=np.random.default_rng(6))
        Ed, En = np.atleast_1d(rb.energy(Xd, Cd)), np.atleast_1d(rb.energy(Xn, Cd))
        assert np.allclose(Ed, En, rtol=1e-12, atol=1e-10)
        Es, Cf = rb.rrrMC(Xd, 0.9, 3000, step=500, seed=8, C0=Cd, quiet=True)
        assert np.allclose(np.atleast_1d(rb.energy(Xd, Cf)), np.atleast_1d(rb.energy(Xn, Cf)), rtol=1e-12, atol=1e-10)
    Xr = rb.GraphRRGNormalDiscretized(60, 3, LEVr, replicas=2, rng=np.random.default_rng(9))
    Cr = rb.Config(Xr.N, 2, rng=np.random.default_rng(10))
    Er = np.atleast_1d(rb.energy(Xr, Cr))
    # direct sum over the edge list with the continuous couplings
    for r in range(2):
        sig = 2.0 * np.array([(int(Cr.chunks[r][i >> 6]) >> (i & 63)) & 1 for i in range(Xr.N)]) - 1
        tot = 0.0
        for x in range(Xr.N):
            for k in range(Xr.A.shape[1]):
                y = Xr.A[x, k] - 1
                if y > x:
                    tot -= Xr.cJ[x, k] * sig[x] * sig[y]
        assert abs(Er[r] - tot) < 1e-9
