"""GPU parity of wtmMC (RRRMC.jl:376-430, WaitingTimes.jl) on every hot-path graph family: the chain kernel and the
oracle consume the same Philox draw stream, so energies at every sample and the final configurations must agree
bit for bit; plus the reference's energy-consistency hook (test/runtests.jl:12-20,140-150)."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance, sk_binary, sk_gauss

pytestmark = pytest.mark.gpu


def _mk(name, R):
    if name == "EA(4,3)":
        A, J = ea_instance(4, 3, seed=3)
        return rb.GraphEA(4, 3, replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J))
    if name == "EA(3,2,(-1,0,1))":
        A, J = ea_instance(3, 2, (-1, 0, 1), seed=4)
        return rb.GraphEA(3, 2, (-1, 0, 1), replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J, (-1, 0, 1)))
    if name == "EANormal(4,2)":
        A, J = ea_instance(4, 2, seed=5, gaussian=True)
        return rb.GraphEANormal(4, 2, replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_f64(A, J))
    if name == "EANormalDiscretized(3,3)":
        A, cJ = ea_instance(3, 3, seed=6, gaussian=True)
        return rb.GraphEANormalDiscretized(3, 3, (-1, 0, 1), replicas=R, A=A, cJ=cJ), (lambda: ffi.Graph.ea_discretized(A, cJ, (-1, 0, 1)))
    if name == "SK(12)":
        J = sk_binary(12, 7)
        return rb.GraphSK(12, replicas=R, J=J), (lambda: ffi.Graph.sk_bin(J))
    if name == "SKNormal(11)":
        J = sk_gauss(11, 8)
        return rb.GraphSKNormal(11, replicas=R, J=J), (lambda: ffi.Graph.sk_f64(J))
    if name == "QT(12,4)":
        return rb.GraphQT(12, 4, 0.73, replicas=R), (lambda: ffi.Graph.qt(12, 4, 0.73))
    if name == "Quant(8,5,SK)":
        J = sk_binary(8, 9)
        return rb.GraphQSKT(8, 5, 0.5, 2.0, replicas=R, J=J), (lambda: ffi.Graph.quant(8, 5, 0.5, 2.0, ffi.SK_BIN, J))
    if name == "QEAT(3,2,4)":
        A, J = ea_instance(3, 2, seed=10, gaussian=True)
        return rb.GraphQEAT(3, 2, 4, 0.5, 2.0, replicas=R, A=A, J=J), (lambda: ffi.Graph.quant(9, 4, 0.5, 2.0, ffi.EA_F64, J, A))
    raise KeyError(name)


GRAPHS = ["EA(4,3)", "EA(3,2,(-1,0,1))", "EANormal(4,2)", "EANormalDiscretized(3,3)", "SK(12)", "SKNormal(11)", "QT(12,4)",
          "Quant(8,5,SK)", "QEAT(3,2,4)"]


@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("beta,step", [(1.0, 0.7), (2.5, 3.0)])
def test_wtmMC_bit_exact_vs_oracle(name, beta, step):
    R, samples, seed = 5, 120, 31337
    X, mk = _mk(name, R)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(2))
    Es, Cf = rb.wtmMC(X, beta, samples, step=step, seed=seed, C0=C0, quiet=True)
    Es = np.asarray(Es, np.float64).reshape(-1, R)
    assert Es.shape[0] == samples
    moves = 0
    for r in range(R):
        g = mk()
        s = C0.chunks[r].copy()
        want, res = ffi.wtmMC(g, beta, samples, s, ffi.PhiloxDraws(seed, chain=r), step=step)
        assert np.array_equal(Es[:, r], want), (name, r)
        assert np.array_equal(Cf.chunks[r], s), (name, r)
        moves += res.iters_done
    assert X.last_run.accepted_total == moves


def test_wtmMC_hook_energy_consistency_and_early_stop():
    R = 3
    X, mk = _mk("EANormal(4,2)", R)
    g = mk()
    seen, bad = [], []

    def hook(t, X_, C, num_moves, E):
        seen.append(t)
        e = np.array([g.energy(C.chunks[r]) for r in range(R)])
        if not np.allclose(np.atleast_1d(E), e, rtol=0, atol=1e-11 * X.N):
            bad.append((t, np.atleast_1d(E) - e))
        return len(seen) < 7
    Es, _ = rb.wtmMC(X, 1.5, 50, step=2.0, seed=9, hook=hook, quiet=True)
    assert not bad, bad[:2]
    assert len(seen) == 7 and Es.shape[0] == 7
    assert np.allclose(seen, [(k + 1) * 2.0 / X.N for k in range(7)])   # the reference passes the global time (RRRMC.jl:405)


def test_wtmMC_argument_errors():
    X, _ = _mk("EA(4,3)", 2)
    with pytest.raises(ValueError):
        rb.wtmMC(X, 1.0, 10, step=0.0, quiet=True)
    with pytest.raises(ValueError):
        rb.wtmMC(X, float("inf"), 10, quiet=True)
    with pytest.raises(ValueError):
        rb.wtmMC(X, 1.0, 10, C0=rb.Config(X.N + 1, 2), quiet=True)
