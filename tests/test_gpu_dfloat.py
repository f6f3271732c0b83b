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
