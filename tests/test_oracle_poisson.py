"""CPU checks of the poisson acceptance procedure of the checkerboard sweeps: the count tables against
high-precision arithmetic, and the CPU model's per-class acceptance frequencies against exp(-βΔE)."""
import numpy as np
import pytest

from oracle import ffi
from tests.helpers import ea_instance
from tests.test_oracle_sparse import _fwd


@pytest.mark.parametrize("beta,D", [(1.0, 3), (0.8, 3), (2.0, 3), (0.5, 3), (0.6, 2), (1.3, 1), (0.0, 1)])
def test_tables_match_high_precision_poisson_cdf(beta, D):
    thr = ffi.thresholds_fixed64(beta, D)
    tbl = ffi.cb_poisson_tables(thr)
    exact = ffi.cb_poisson_tables_exact(thr)
    assert len(tbl) == ffi.CBP_LEN
    assert np.abs(tbl.astype(np.int64) - exact.astype(np.int64)).max() <= 1
    TA = tbl[:ffi.CBP_KA]
    TB0, TB, TC = (tbl[ffi.CBP_KA + k * ffi.CBP_KR: ffi.CBP_KA + (k + 1) * ffi.CBP_KR] for k in range(3))
    for T in (TA, TB, TC):
        assert T[-1] == 0xffffffff and (np.diff(T.astype(np.int64)) >= 0).all()
    assert TB0[-1] == TC[0] and (np.diff(TB0.astype(np.int64)) >= 0).all() and TB0.max() <= TC[0]
    if D < 3:
        assert TC[0] == 0xffffffff and np.array_equal(TB0, TB)
    if D < 2:
        assert TB[0] == 0xffffffff
    # a lane of class 1 is hit by some level with probability 1 - exp(-(mean_a + mean_b + mean_c)/128) = p1
    if beta > 0:
        mean = sum(((2.0 ** 32 - 1 - T[:-1].astype(np.float64)) / 2.0 ** 32).sum() for T in (TA, TB, TC))
        assert abs(-np.expm1(-mean / 128) - np.exp(-4 * beta)) < 1e-8


@pytest.mark.parametrize("beta,NW", [(0.5, 1), (0.5, 2), (0.7, 4), (0.9, 6)])
def test_cpu_model_acceptance_frequencies(beta, NW):
    """One sweep from a fixed state: the fraction of flipped lanes per ΔE class must match exp(-βΔE) within 4σ, and
    lanes with ΔE<=0 always flip (accept(), RRRMC.jl:39). Small NW at warm β exercises the overflow stream, the
    level-2 and level-3 hits and the fresh count uniform."""
    L, D, R = 4, 3, 128
    A, J = ea_instance(L, D, seed=5)
    N = L ** D
    g = ffi.Graph.ea_int(A, J)
    thr = ffi.thresholds_fixed64(beta, D)
    tbl = ffi.cb_poisson_tables(thr)
    assert tbl[ffi.CBP_KA - 2] == 0xffffffff
    Jf = _fwd(A, J, L, D)
    rng = np.random.default_rng(3)
    tot = np.zeros(4); acc = np.zeros(4)
    co = np.indices((L,) * D).reshape(D, -1)[::-1]
    col0 = np.flatnonzero(co.sum(axis=0) % 2 == 0)   # only colour-0 sites see the initial state on all neighbours
    for trial in range(60):
        sp = rng.integers(0, 2 ** 32, (N, R // 32), dtype=np.uint32)
        before = np.unpackbits(sp.view(np.uint8).reshape(N, R // 8), axis=1, bitorder="little").T.copy()  # [R][N]
        ffi.checkerboard_sweeps_poisson(L, D, R, sp, Jf, tbl, NW, 1000 + trial, 0, 1)
        after = np.unpackbits(sp.view(np.uint8).reshape(N, R // 8), axis=1, bitorder="little").T
        for r in range(0, R, 8):
            ch = np.packbits(before[r], bitorder="little").view(np.uint64).copy()
            g.energy(ch)
            for i in col0:
                dE = int(g.delta_energy(ch, int(i) + 1))
                c = 0 if dE <= 0 else dE // 4
                tot[c] += 1; acc[c] += before[r, i] != after[r, i]
    assert acc[0] == tot[0]
    for c in range(1, D + 1):
        p = np.exp(-beta * 4 * c)
        sigma = np.sqrt(p * (1 - p) / tot[c])
        assert abs(acc[c] / tot[c] - p) < 4 * sigma, (c, acc[c] / tot[c], p, sigma)


def test_cpu_model_lanes_are_independent():
    """Pair statistics inside one task: the joint flip frequency of two class-1 lanes of the same task must be p1²
    (hits are placed with replacement from a Poisson count, which is what makes lanes independent)."""
    L, D, R, beta = 2, 1, 128, 0.25          # 2-site ring with a double bond: ΔE ∈ {-4, 0, +4}... use aligned spins
    Jf = np.ones((2, 1), np.int8)
    thr = ffi.thresholds_fixed64(beta, D)
    tbl = ffi.cb_poisson_tables(thr)
    p1 = np.exp(-4 * beta)
    n, both, one = 0, 0, 0
    for trial in range(400):
        sp = np.zeros((2, 4), np.uint32)     # all spins equal, J=+1: every lane of site 0 has ΔE = +4 (class 1)
        ffi.checkerboard_sweeps_poisson(L, D, R, sp, Jf, tbl, 1, 77 + trial, 0, 1)
        # colour 0 = site 0 moved first from the all-aligned state
        bits = np.unpackbits(sp[0].view(np.uint8), bitorder="little")
        n += 64; one += bits.sum(); both += (bits[0::2] & bits[1::2]).sum()
    assert abs(one / (2 * n) - p1) < 4 * np.sqrt(p1 * (1 - p1) / (2 * n))
    assert abs(both / n - p1 * p1) < 4 * np.sqrt(p1 * p1 * (1 - p1 * p1) / n)


def test_cpu_model_replicas_are_independent_of_batch_composition():
    L, D, beta = 4, 2, 0.9
    A, J = ea_instance(L, D, seed=6)
    N = L ** D
    tbl = ffi.cb_poisson_tables(ffi.thresholds_fixed64(beta, D))
    Jf = _fwd(A, J, L, D)
    sp = np.random.default_rng(1).integers(0, 2 ** 32, (N, 8), dtype=np.uint32)
    a = sp.copy(); b = np.ascontiguousarray(sp[:, :4])
    ffi.checkerboard_sweeps_poisson(L, D, 256, a, Jf, tbl, 2, 9, 2, 3)
    ffi.checkerboard_sweeps_poisson(L, D, 128, b, Jf, tbl, 2, 9, 2, 3)
    assert np.array_equal(a[:, :4], b)


@pytest.mark.parametrize("beta,D", [(1.0, 3), (0.77, 3), (3.0, 3), (0.5, 2), (0.0, 1)])
def test_library_table_builder_matches_oracle(beta, D):
    """rrrmc_checkerboard_poisson_tables is host-only code of the C ABI (no device needed): it must produce the very
    tables the oracle builds, since parity tests feed one table to both sides."""
    from rrrmc_b200._ffi import check, lib, ptr
    thr = ffi.thresholds_fixed64(beta, D)
    tbl = np.zeros(ffi.CBP_LEN, np.uint32)
    check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), D, ptr(tbl), ffi.CBP_LEN))
    assert np.array_equal(tbl, ffi.cb_poisson_tables(thr))
    with pytest.raises(ValueError):
        check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), D, ptr(tbl), ffi.CBP_LEN - 1))
