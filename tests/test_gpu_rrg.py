"""GPU parity of the random-regular-graph family (src/graphs/RRG.jl; the benchmark of the RRR paper): GraphRRG ±J and
GraphRRGNormal on explicit K-regular adjacencies. Interface queries and every sampler against the oracle's
general-adjacency GraphEA arithmetic (identical to RRG.jl:165-250 when couplings are non-zero), bit for bit."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi

pytestmark = pytest.mark.gpu


def _mk(name, R):
    kind, N, K = name.split(",")
    N, K = int(N), int(K)
    rng = np.random.default_rng(100 * N + K)
    A = rb.gen_RRG(N, K, rng)
    assert A.shape == (N, K) and (np.diff(A, axis=1) > 0).all() and not (A == np.arange(1, N + 1)[:, None]).any()
    if kind == "pm1":
        J = rb.gen_J_graph(lambda n: rng.choice([-1.0, 1.0], n), A).astype(np.int64)
        return rb.GraphRRG(N, K, replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J))
    if kind == "int":
        J = rb.gen_J_graph(lambda n: rng.choice([-2.0, -1.0, 1.0, 2.0], n), A).astype(np.int64)
        return rb.GraphRRG(N, K, (-2, -1, 1, 2), replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_int(A, J, (-2, -1, 1, 2)))
    if kind == "int0":   # zero level: neighbors() skips the zero couplings (RRG.jl:133)
        J = rb.gen_J_graph(lambda n: rng.choice([-1.0, 0.0, 1.0], n), A).astype(np.int64)
        assert (J == 0).any()
        return rb.GraphRRG(N, K, (-1, 0, 1), replicas=R, A=A, J=J), (lambda: ffi.Graph.rrg_int(A, J, (-1, 0, 1)))
    if kind == "disc":   # GraphRRGNormalDiscretized (RRG.jl:274-330)
        cJ = rb.gen_J_graph(lambda n: rng.standard_normal(n), A)
        return rb.GraphRRGNormalDiscretized(N, K, (-1, 0, 1), replicas=R, A=A, cJ=cJ), (lambda: ffi.Graph.rrg_discretized(A, cJ, (-1, 0, 1)))
    J = rb.gen_J_graph(lambda n: rng.standard_normal(n), A)
    return rb.GraphRRGNormal(N, K, replicas=R, A=A, J=J), (lambda: ffi.Graph.ea_f64(A, J))


GRAPHS = ["pm1,10,3", "pm1,40,4", "pm1,64,6", "pm1,30,5", "int,20,3", "normal,10,3", "normal,36,4", "int0,10,3", "int0,30,4",
          "disc,10,3", "disc,24,5"]


@pytest.mark.parametrize("name", GRAPHS)
def test_interface_matches_oracle(name):
    R = 4
    X, mk = _mk(name, R)
    g = mk()
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    E = np.atleast_1d(rb.energy(X, C0))
    for r in range(R):
        assert E[r] == g.energy(C0.chunks[r])
    g.energy(C0.chunks[1])
    dE = np.asarray(rb.all_delta_energy(X, C0, 1), np.float64)
    assert np.array_equal(dE, np.array([g.delta_energy(C0.chunks[1], i) for i in range(1, X.N + 1)]))
    for i in (1, X.N):
        assert tuple(rb.neighbors(X, i)) == tuple(g.neighbors(i))
        if name.startswith("int0"):
            assert tuple(rb.neighbors(X, i)) == tuple(X.A[i - 1][X.J[i - 1] != 0])
        else:
            assert tuple(rb.neighbors(X, i)) == tuple(X.A[i - 1])
    if not name.startswith("normal"):
        assert np.array_equal(np.asarray(rb.allDeltaE(X), np.float64), g.allDE())
        if name.startswith("pm1"):   # RRG.jl:252-255: (0,4,..) for even K, (2,6,..) for odd K
            K = X.K
            assert tuple(rb.allDeltaE(X)) == (tuple(4 * d for d in range(K // 2 + 1)) if K % 2 == 0 else tuple(2 * (2 * d + 1) for d in range((K + 1) // 2)))


@pytest.mark.parametrize("name", GRAPHS)
@pytest.mark.parametrize("sampler", ["standardMC", "rrrMC", "rrrMC_staged", "bklMC", "wtmMC"])
def test_samplers_bit_exact_vs_oracle(name, sampler):
    R, beta, iters, step, seed = 4, 1.5, 2400, 200, 777
    X, mk = _mk(name, R)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(5))
    if sampler == "wtmMC":
        Es, Cf = rb.wtmMC(X, beta, 60, step=1.3, seed=seed, C0=C0, quiet=True)
        run = lambda g, s, r: ffi.wtmMC(g, beta, 60, s, ffi.PhiloxDraws(seed, chain=r), step=1.3)
    elif sampler == "standardMC":
        Es, Cf = rb.standardMC(X, beta, iters, step=step, seed=seed, C0=C0, quiet=True, schedule="random")
        run = lambda g, s, r: ffi.standardMC(g, beta, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step)
    elif sampler == "bklMC":
        Es, Cf = rb.bklMC(X, beta, iters, step=step, seed=seed, C0=C0, quiet=True)
        run = lambda g, s, r: ffi.bklMC(g, beta, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step)
    else:
        thr = 1.0 if sampler == "rrrMC_staged" else float("nan")
        Es, Cf = rb.rrrMC(X, beta, iters, step=step, seed=seed, C0=C0, quiet=True, staged_thr=thr)
        run = lambda g, s, r: ffi.rrrMC(g, beta, iters, s, ffi.PhiloxDraws(seed, chain=r), step=step, staged_thr=thr)
    Es = np.asarray(Es, np.float64).reshape(-1, R)
    for r in range(R):
        s = C0.chunks[r].copy()
        want, _ = run(mk(), s, r)
        assert np.array_equal(Es[:len(want), r], want), (name, sampler, r)
        assert np.array_equal(Cf.chunks[r], s), (name, sampler, r)


def test_argument_errors():
    A = rb.gen_RRG(10, 3, np.random.default_rng(1))
    J = rb.gen_J_graph(lambda n: np.ones(n), A).astype(np.int64)
    with pytest.raises(ValueError):
        rb.gen_RRG(9, 3)                                   # N*K odd, RRG.jl:29
    bad = J.copy(); bad[0, 0] = -bad[0, 0]
    with pytest.raises(ValueError):
        rb.GraphRRG(10, 3, A=A, J=bad)                     # not symmetric
    with pytest.raises(ValueError):
        rb.GraphRRG(10, 3, (-1.5, 0.5), A=A, J=J)          # fractional levels (DFloat64): J = ±1 are not levels
    Ab = A.copy(); Ab[0, 0], Ab[0, 1] = Ab[0, 1], Ab[0, 0]
    with pytest.raises(ValueError):
        rb.GraphRRG(10, 3, A=Ab, J=J)                      # rows must ascend
    X = rb.GraphRRG(10, 3, A=A, J=J)
    with pytest.raises(NotImplementedError):
        rb.standardMC(X, 1.0, 100, schedule="checkerboard", quiet=True)
