"""Checkerboard Metropolis for continuous couplings (GraphEANormal, EA.jl:534-680) on the replica batch — VERDICT r1
row J2. The kernel (csrc/ea_normal.cu) must reproduce its CPU restatement (oracle: orc_checkerboard_sweeps_f64) bit
for bit: same Float64 ΔE (slot order of energy(), EA.jl:590-603), same accept(-βΔE) (RRRMC.jl:39), same Philox
uniforms; its observables must agree with the reference's random-site Metropolis within 3σ (north_star)."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance
from tests.test_gpu_checkerboard import _from_multispin, _multispin

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("L,D,R", [(4, 2, 32), (6, 2, 96), (4, 3, 128), (8, 3, 100), (2, 3, 64), (6, 3, 160), (4, 1, 40), (8, 2, 1024)])
def test_checkerboard_f64_bit_exact_vs_cpu_model(L, D, R):
    A, J = ea_instance(L, D, seed=L * 10 + D, gaussian=True)
    X = rb.GraphEANormal(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(7))
    betas = np.linspace(0.3, 2.5, R)                   # a different β on every replica
    seed, sweep0, nsw = 0xC0FFEE1234, (1 << 33) + 3, 5
    got = rb.checkerboard_sweeps_normal(X, betas, nsw, seed=seed, sweep0=sweep0, C0=C0)
    Rp = ((R + 31) // 32) * 32
    sp = _multispin(C0)
    ffi.checkerboard_sweeps_f64(L, D, R, sp, A, J, betas, seed, sweep0, nsw)
    assert sp.shape == (X.N, Rp // 32)
    assert got == _from_multispin(sp, R)
    assert not (got == C0)


def test_checkerboard_f64_delta_energy_is_the_reference_value():
    """The ΔE the kernel acts on is delta_energy() after energy() (EA.jl:665-672): at β = +inf-like (β = 1e6) a sweep
    flips exactly the lanes with ΔE <= 0 as the interface reports them for the configuration at the time of the update.
    Colour 0 is updated from the initial configuration, so its flips can be predicted from rb.all_delta_energy."""
    L, D, R = 6, 3, 64
    A, J = ea_instance(L, D, seed=3, gaussian=True)
    X = rb.GraphEANormal(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(1))
    dE = np.array([rb.all_delta_energy(X, C0, r) for r in (0, 31, 63)])       # Float64, bit-equal to the oracle (test_gpu_interface)
    g = ffi.Graph.ea_f64(A, J)
    for k, r in enumerate((0, 31, 63)):
        g.energy(C0.chunks[r])
        want = np.array([g.delta_energy(C0.chunks[r], i + 1) for i in range(X.N)])
        assert np.allclose(dE[k], want, rtol=1e-6, atol=0) and np.array_equal(dE[k], want)
    got = rb.checkerboard_sweeps_normal(X, 1e6, 1, seed=5, C0=C0)
    idx = np.arange(X.N)
    colour0 = ((idx % L) + (idx // L) % L + idx // (L * L)) % 2 == 0
    for k, r in enumerate((0, 31, 63)):
        flipped = got.s[r] != C0.s[r]
        assert np.array_equal(flipped[colour0], dE[k][colour0] <= 0)


def test_standardMC_checkerboard_on_EANormal_energies_and_accepted():
    L, D, R, beta = 4, 3, 70, 0.9
    A, J = ea_instance(L, D, seed=21, gaussian=True)
    X = rb.GraphEANormal(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_f64(A, J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(8))
    seen = []

    def hook(it, X_, C, acc, E):
        seen.append((it, np.array(acc), np.array(E), C.chunks.copy()))
        return True
    N = X.N
    Es, Cf = rb.standardMC(X, beta, 6 * N, step=2 * N, seed=77, C0=C0, hook=hook, quiet=True, schedule="checkerboard")
    sp = _multispin(C0); acc = np.zeros(R, np.int64)
    for k in range(3):
        ffi.checkerboard_sweeps_f64(L, D, R, sp, A, J, beta, 77, 2 * k, 2, acc)
        cfg = _from_multispin(sp, R)
        assert seen[k][0] == 2 * (k + 1) * N
        assert np.array_equal(seen[k][3], cfg.chunks)
        assert np.array_equal(seen[k][1], acc)
        e = np.array([g.energy(cfg.chunks[r]) for r in range(R)])
        assert np.array_equal(seen[k][2], e) and np.array_equal(Es[k], e)
    assert Cf == _from_multispin(sp, R)
    with pytest.raises(NotImplementedError):
        rb.standardMC(rb.GraphEANormal(3, 2), beta, 100, schedule="checkerboard", quiet=True)   # odd L: not two-colourable


def test_checkerboard_f64_statistics_within_3_sigma_of_random_site():
    """north_star: energy from the checkerboard kernel agrees with the reference sampler (random-site Metropolis in the
    reference order, the chain engine — itself bit-exact against the oracle) within 3σ over independent replicas."""
    L, D, R, beta = 6, 3, 256, 0.8
    A, J = ea_instance(L, D, seed=5, gaussian=True)
    X = rb.GraphEANormal(L, D, replicas=R, A=A, J=J)
    N = X.N
    C = rb.checkerboard_sweeps_normal(X, beta, 400, seed=6, C0=rb.Config(N, R, rng=np.random.default_rng(2)))
    e_cb = np.atleast_1d(rb.energy(X, C)) / N
    Y = rb.GraphEANormal(L, D, replicas=R, A=A, J=J)
    Es, _ = rb.standardMC(Y, beta, 400 * N, step=400 * N, seed=9, C0=rb.Config(N, R, rng=np.random.default_rng(3)), quiet=True)
    e_rs = Es[-1] / N
    sigma = np.sqrt(e_cb.var(ddof=1) / R + e_rs.var(ddof=1) / R)
    assert abs(e_cb.mean() - e_rs.mean()) < 3 * sigma, (e_cb.mean(), e_rs.mean(), sigma)
    m = (2.0 * C.s.astype(np.float64) - 1.0).mean(axis=1)
    assert abs(m.mean()) < 3 * m.std(ddof=1) / np.sqrt(R) + 0.02       # no net magnetisation in the symmetric-coupling glass
