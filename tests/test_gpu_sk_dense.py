"""GPU parity of the dense GraphSKNormal path (BASELINE config 4): tensor-core local-field initialisation against
the oracle's energy() cache (SK.jl:212-237; tolerance 1e-6 relative per north_star, measured far tighter), the
ordered CUDA-core path bit-exact, and the lock-step Metropolis kernel through invariants and 3σ statistics against
the reference-order sampler."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import sk_gauss

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,R", [(64, 32), (200, 70), (384, 129), (1024, 64)])
def test_fields_init_tensor_cores_vs_oracle(N, R):
    J = sk_gauss(N, seed=N)
    X = rb.GraphSKNormal(N, replicas=R, J=J)
    g = ffi.Graph.sk_f64(J)
    C0 = rb.Config(N, R, rng=np.random.default_rng(1))
    lf_tc, E_tc, _ = rb.sk_fields_init(X, C0, tensor_cores=True)
    lf_cc, E_cc, _ = rb.sk_fields_init(X, C0, tensor_cores=False)
    for r in (0, R // 2, R - 1):
        E = g.energy(C0.chunks[r])
        want = g.lfields()
        assert np.array_equal(lf_cc[r], want) and E_cc[r] == E          # reference summation order: bit-exact
        scale = np.abs(want).max()
        assert np.abs(lf_tc[r] - want).max() <= 1e-6 * scale            # north_star tolerance
        assert np.abs(lf_tc[r] - want).max() <= 4e-9 * max(1.0, N / 64)  # what the 40-bit fixed point actually gives
        assert abs(E_tc[r] - E) <= 1e-6 * abs(E) + 1e-9


def test_fields_init_is_exact_for_dyadic_couplings():
    """With couplings on a 2^-20 grid the fixed-point digit planes represent J exactly: the tensor-core result must
    then equal the ordered sum bit for bit (all partial sums are exact in both)."""
    N, R = 256, 64
    rng = np.random.default_rng(3)
    J = np.triu(rng.integers(-2 ** 14, 2 ** 14, (N, N)).astype(np.float64) / 2 ** 20, 1); J = J + J.T
    X = rb.GraphSKNormal(N, replicas=R, J=J)
    C0 = rb.Config(N, R, rng=rng)
    lf_tc, E_tc, _ = rb.sk_fields_init(X, C0, tensor_cores=True)
    lf_cc, E_cc, _ = rb.sk_fields_init(X, C0, tensor_cores=False)
    assert np.array_equal(lf_tc, lf_cc) and np.array_equal(E_tc, E_cc)


@pytest.mark.parametrize("N,R", [(96, 33), (512, 64)])
def test_lockstep_invariants(N, R):
    J = sk_gauss(N, seed=7)
    X = rb.GraphSKNormal(N, replicas=R, J=J)
    C0 = rb.Config(N, R, rng=np.random.default_rng(2))
    rb.sk_fields_init(X, C0, tensor_cores=True)
    E, acc, C1 = rb.sk_metropolis_sweeps(X, 1.2, 5, seed=9)
    assert not (C1 == C0) and (acc > 0).all() and (acc <= 5 * N).all()
    lf_dev = np.zeros((R, N)); rb._ffi.check(rb._ffi.lib().rrrmc_sk_get_fields(X._state, rb._ffi.ptr(lf_dev)))
    # tracked energy and incrementally updated fields equal a from-scratch recompute (checkenergy_hook invariant)
    lf, E_fresh, _ = rb.sk_fields_init(X, C1, tensor_cores=False)
    assert np.allclose(E, E_fresh, rtol=1e-9, atol=1e-9)
    assert np.allclose(lf_dev, lf, rtol=0, atol=1e-9 * np.abs(lf).max())
    g = ffi.Graph.sk_f64(J)
    assert np.isclose(g.energy(C1.chunks[R - 1]), E[R - 1], rtol=1e-9)
    # same seed, same start -> same trajectory; continuing = one longer run
    rb.sk_fields_init(X, C0, tensor_cores=False)
    E2, _, C2 = rb.sk_metropolis_sweeps(X, 1.2, 2, seed=9, sweep0=0)
    E3, _, C3 = rb.sk_metropolis_sweeps(X, 1.2, 3, seed=9, sweep0=2)
    assert C3 == C1 and np.allclose(E3, E, rtol=1e-12)


@pytest.mark.parametrize("variant", ["0", "2", "1"])   # fields in registers / TMA rows, fields in shared memory / plain loads
@pytest.mark.parametrize("N,R,nsw", [(2, 5, 9), (64, 300, 6), (96, 33, 5), (512, 64, 3), (1026, 149, 2), (4096, 8, 1), (4098, 6, 1)])
def test_lockstep_trajectory_vs_oracle(N, R, nsw, variant, monkeypatch):
    """Bit-exact trajectory parity of every lock-step kernel with the CPU restatement (orc_sk_lockstep_sweeps): same
    configurations, fields, tracked energies and acceptance counts, per-replica β; two calls continue one run."""
    monkeypatch.setenv("RRRMC_SK_VARIANT", variant)
    J = sk_gauss(N, seed=100 + N)
    X = rb.GraphSKNormal(N, replicas=R, J=J)
    C0 = rb.Config(N, R, rng=np.random.default_rng(N + R))
    lf0, E0, _ = rb.sk_fields_init(X, C0, tensor_cores=False)
    beta = np.linspace(0.3, 2.0, R)
    seed = 0xabcdef12345 + N
    n1 = nsw // 2
    rb.sk_metropolis_sweeps(X, beta, n1, seed=seed, sweep0=(1 << 32) - 1)
    E, acc, C1 = rb.sk_metropolis_sweeps(X, beta, nsw - n1, seed=seed, sweep0=(1 << 32) - 1 + n1)
    lf_dev = np.zeros((R, N)); rb._ffi.check(rb._ffi.lib().rrrmc_sk_get_fields(X._state, rb._ffi.ptr(lf_dev)))
    chunks = np.ascontiguousarray(C0.chunks, np.uint64).copy()
    lf = np.ascontiguousarray(lf0, np.float64).copy(); Eo = np.ascontiguousarray(E0, np.float64).copy()
    acco = np.zeros(R, np.int64)
    ffi.sk_lockstep_sweeps(J, chunks, lf, Eo, acco, beta, seed, (1 << 32) - 1, nsw)
    assert np.array_equal(np.asarray(C1.chunks, np.uint64), chunks)
    assert np.array_equal(lf_dev, lf) and np.array_equal(E, Eo) and np.array_equal(acc, acco)
    assert acco.sum() > 0


def test_lockstep_statistics_vs_reference_sampler():
    """⟨E⟩/N after equilibration: lock-step sweeps vs the reference-order standardMC chains, within 3σ."""
    N, R, beta = 64, 256, 0.8
    J = sk_gauss(N, seed=11)
    X = rb.GraphSKNormal(N, replicas=R, J=J)
    C0 = rb.Config(N, R, rng=np.random.default_rng(5))
    rb.sk_fields_init(X, C0)
    E, _, _ = rb.sk_metropolis_sweeps(X, beta, 300, seed=3)
    Y = rb.GraphSKNormal(N, replicas=R, J=J)
    Es, _ = rb.standardMC(Y, beta, 300 * N, step=300 * N, seed=4, C0=C0, quiet=True)
    a, b = E / N, Es[-1] / N
    sigma = np.sqrt(a.var(ddof=1) / R + b.var(ddof=1) / R)
    assert abs(a.mean() - b.mean()) < 3 * sigma, (a.mean(), b.mean(), sigma)


def test_config4_shape_smoke():
    """BASELINE config 4 (N=4096 × 512 replicas): tensor-core fields against the ordered path on sampled replicas."""
    N, R = 4096, 512
    X = rb.GraphSKNormal(N, replicas=R, rng=np.random.default_rng(0))
    C0 = rb.Config(N, R, rng=np.random.default_rng(1))
    lf_tc, E_tc, ms = rb.sk_fields_init(X, C0, tensor_cores=True)
    g = ffi.Graph.sk_f64(X.J)
    for r in (0, 511):
        E = g.energy(C0.chunks[r]); want = g.lfields()
        assert np.abs(lf_tc[r] - want).max() <= 1e-6 * np.abs(want).max()
        assert abs(E_tc[r] - E) <= 1e-6 * abs(E)
    E, acc, C1 = rb.sk_metropolis_sweeps(X, 1.0, 1, seed=1)
    assert (E < E_tc).all() and (acc > 0).all()
