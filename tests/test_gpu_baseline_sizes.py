"""Engine vs oracle AT THE SIZES BASELINE.json quotes (the parity suite only ran toy sizes).

C1  GraphEA 2D 32x32 ±J, standardMC β=1, 10^6 iterations, step 10^3, 1 replica (the reference's CPU-runnable case)
C3  GraphEA 3D L=32 ±J, rrrMC and bklMC at β=3 — a few of the 256 replicas against the oracle, all 256 for the invariant
C5  GraphQSKT Nk=1024, M=64 (N = 65 536), rrrMC(DoubleGraph), >= 10^4 iterations, + the hook-side observables
C2 (the headline) at size is tests/test_gpu_checkerboard.py::test_checkerboard_poisson_full_size_bit_exact.
Every comparison is bit for bit: both sides consume the same Philox draw stream."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance, sk_binary

pytestmark = pytest.mark.gpu


def test_C1_standardMC_2d_32x32_one_replica_1e6_iterations():
    L, D, beta, iters, step = 32, 2, 1.0, 1_000_000, 1000
    A, J = ea_instance(L, D, seed=321)
    X = rb.GraphEA(L, D, replicas=1, A=A, J=J)
    g = ffi.Graph.ea_int(A, J)
    C0 = rb.Config(X.N, 1, rng=np.random.default_rng(1))
    Es, Cf = rb.standardMC(X, beta, iters, step=step, seed=167432777111, C0=C0, quiet=True)   # default = the reference order
    s = C0.chunks[0].copy()
    wantE, info = ffi.standardMC(g, beta, iters, s, ffi.PhiloxDraws(167432777111, chain=0), step=step)
    assert np.asarray(Es).shape == (iters // step,)
    assert np.array_equal(np.asarray(Es, np.float64), wantE)
    assert np.array_equal(Cf.chunks[0], s)
    assert X.last_run.accepted_total == info.accepted
    # the run equilibrates: the 2D ±J energy per spin at β=1 is far below the random start and the tracked energy
    # equals a fresh recompute at the end (test/runtests.jl:12-15)
    assert wantE[-1] / X.N < -1.0
    assert g.energy(Cf.chunks[0]) >= wantE[-1] - 8 and g.energy(Cf.chunks[0]) <= wantE[-1] + 8   # one move after the last sample


@pytest.mark.parametrize("sampler", ["rrrMC", "bklMC"])
def test_C3_rejection_free_L32_beta3(sampler):
    """L=32 (N = 32 768) at β=3: the compact-state fast path (chain_ea.cu) on 256 chains; chains 0, 100 and 255 are
    compared with the oracle move by move (energies at ten samples and the final configuration); every chain
    keeps the energy-consistency invariant."""
    L, D, beta, R = 32, 3, 3.0, 256
    iters, step = (500_000, 50_000) if sampler == "rrrMC" else (10_000_000, 1_000_000)
    A, J = ea_instance(L, D, seed=33)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_int(A, J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(3))
    fn, ofn = (rb.rrrMC, ffi.rrrMC) if sampler == "rrrMC" else (rb.bklMC, ffi.bklMC)
    Es, Cf = fn(X, beta, iters, step=step, seed=2024, C0=C0, quiet=True)
    assert np.asarray(Es).shape == (iters // step, R)
    for r in (0, 100, 255):
        s = C0.chunks[r].copy()
        wantE, _ = ofn(g, beta, iters, s, ffi.PhiloxDraws(2024, chain=r), step=step)
        assert np.array_equal(np.asarray(Es, np.float64)[:, r], wantE), r
        assert np.array_equal(Cf.chunks[r], s), r
    Efin = np.atleast_1d(rb.energy(X, Cf))
    for r in (1, 17, 254):
        assert Efin[r] == g.energy(Cf.chunks[r])
    assert (np.asarray(Es)[-1] < np.asarray(Es)[0]).all()      # every chain is still relaxing downhill from a random start


def test_C5_quantum_rrrMC_Nk1024_M64_with_observables():
    """GraphQSKT(1024, 64, Γ=0.3, β=2): rrrMC on the DoubleGraph (QT inner graph + SK residual, RRRMC.jl:221-290), 50 000
    iterations on 4 replicas against the oracle; then Qenergy / transverse_mag / Renergies / overlaps of the final
    configurations against QT.jl:113-121, 201-268 as the oracle restates them."""
    Nk, M, G, beta, R, iters, step = 1024, 64, 0.3, 2.0, 4, 50_000, 5_000
    Jb = sk_binary(Nk, seed=5)
    X = rb.GraphQSKT(Nk, M, G, beta, replicas=R, J=Jb)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(6))
    Es, Cf = rb.rrrMC(X, beta, iters, step=step, seed=555, C0=C0, quiet=True)
    assert np.asarray(Es).shape == (iters // step, R)
    for r in range(R):
        g = ffi.Graph.quant(Nk, M, G, beta, ffi.SK_BIN, Jb)
        s = C0.chunks[r].copy()
        wantE, _ = ffi.rrrMC(g, beta, iters, s, ffi.PhiloxDraws(555, chain=r), step=step)
        assert np.array_equal(np.asarray(Es, np.float64)[:, r], wantE), r
        assert np.array_equal(Cf.chunks[r], s), r
    tm = rb.transverse_mag(X, Cf, beta); qe = rb.Qenergy(X, Cf); re = rb.Renergies(X, Cf); ov = rb.overlaps(X, Cf)
    g = ffi.Graph.quant(Nk, M, G, beta, ffi.SK_BIN, Jb)
    for r in range(R):
        s = Cf.chunks[r]
        g.energy(s)
        assert np.isclose(tm[r], ffi.lib().orc_transverse_mag(g.h, s, beta), rtol=1e-13)
        assert np.isclose(qe[r], ffi.lib().orc_Qenergy(g.h, s), rtol=1e-12, atol=1e-12)
        want = np.zeros(M); ffi.lib().orc_Renergies(g.h, want)
        assert np.array_equal(re[r], want)
        wo = np.zeros(M // 2); ffi.lib().orc_overlaps(g.h, wo)
        assert np.array_equal(ov[r], wo)
