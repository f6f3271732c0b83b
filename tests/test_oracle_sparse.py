"""CPU checks of the sparse acceptance procedure of the checkerboard sweeps: the count tables
against exact rational arithmetic, and the CPU model's per-class acceptance frequencies against exp(-βΔE)."""
import numpy as np
import pytest

from oracle import ffi
from tests.helpers import ea_instance


@pytest.mark.parametrize("beta,D", [(1.0, 3), (0.8, 3), (2.0, 3), (0.3, 2), (1.3, 1), (0.0, 3)])
def test_tables_match_exact_binomial_cdf(beta, D):
    thr = ffi.thresholds_fixed64(beta, D)
    tbl = ffi.cb_sparse_tables(thr)
    exact = ffi.cb_sparse_tables_exact(thr)
    assert len(tbl) == ffi.CB_T1 + (D - 1) * ffi.CB_TC
    assert np.abs(tbl.astype(np.int64) - exact.astype(np.int64)).max() <= 1
    assert tbl[ffi.CB_T1 - 1] == 0xffffffff
    for c in range(2, D + 1):
        T = tbl[ffi.CB_T1 + (c - 2) * ffi.CB_TC: ffi.CB_T1 + (c - 1) * ffi.CB_TC]
        assert T[-1] == 0xffffffff and (np.diff(T.astype(np.int64)) >= 0).all()
    # mean count implied by the table = n·p to 2^-32 resolution
    T1 = tbl[:ffi.CB_T1].astype(np.float64)
    mean = ((2.0 ** 32 - 1 - T1[:-1]) / 2.0 ** 32).sum()
    assert abs(mean - 32 * np.exp(-4 * beta)) < 1e-6


def _fwd(A, J, L, D):
    N = A.shape[0]
    out = np.zeros((N, D), np.int8)
    for i in range(N):
        stride = 1
        for d in range(D):
            c = (i // stride) % L
            up = i + (((c + 1) % L) - c) * stride
            out[i, d] = J[i, np.flatnonzero(A[i] == up + 1)[0]]
            stride *= L
    return out


def test_cpu_model_acceptance_frequencies():
    """One sweep from a fixed state: the fraction of flipped lanes per ΔE class must match exp(-βΔE) within 4σ, and
    lanes with ΔE<=0 always flip (accept(), RRRMC.jl:39)."""
    L, D, R, beta = 4, 3, 128, 0.35
    A, J = ea_instance(L, D, seed=5)
    N = L ** D
    g = ffi.Graph.ea_int(A, J)
    thr = ffi.thresholds_fixed64(beta, D)
    tbl = ffi.cb_sparse_tables(thr)
    Jf = _fwd(A, J, L, D)
    rng = np.random.default_rng(3)
    tot = np.zeros(4); acc = np.zeros(4)
    for trial in range(60):
        sp = rng.integers(0, 2 ** 32, (N, R // 32), dtype=np.uint32)
        before = np.unpackbits(sp.view(np.uint8).reshape(N, R // 8), axis=1, bitorder="little").T.copy()  # [R][N]
        ffi.checkerboard_sweeps_sparse(L, D, R, sp, Jf, tbl, 1000 + trial, 0, 1)
        after = np.unpackbits(sp.view(np.uint8).reshape(N, R // 8), axis=1, bitorder="little").T
        # only colour-0 sites see the initial state on all their neighbours
        co = np.indices((L,) * D).reshape(D, -1)[::-1]
        col0 = np.flatnonzero(co.sum(axis=0) % 2 == 0)
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


def test_cpu_model_replicas_are_independent_of_batch_composition():
    """Counter-based draws: a task's outcome depends only on (seed, sweep, site, group), so sweeping the first 128
    replicas alone reproduces their trajectory inside a 256-replica batch."""
    L, D, beta = 4, 2, 0.9
    A, J = ea_instance(L, D, seed=6)
    N = L ** D
    tbl = ffi.cb_sparse_tables(ffi.thresholds_fixed64(beta, D))
    Jf = _fwd(A, J, L, D)
    sp = np.random.default_rng(1).integers(0, 2 ** 32, (N, 8), dtype=np.uint32)
    a = sp.copy(); b = np.ascontiguousarray(sp[:, :4])
    ffi.checkerboard_sweeps_sparse(L, D, 256, a, Jf, tbl, 9, 2, 3)
    ffi.checkerboard_sweeps_sparse(L, D, 128, b, Jf, tbl, 9, 2, 3)
    assert np.array_equal(a[:, :4], b)


@pytest.mark.parametrize("beta,D", [(1.0, 3), (0.77, 3), (3.0, 3), (0.5, 2), (0.0, 1)])
def test_library_table_builder_matches_oracle(beta, D):
    """rrrmc_checkerboard_sparse_tables is host-only code of the C ABI (no device needed): it must produce the very
    tables the oracle builds, since parity tests feed one table to both sides."""
    from rrrmc_b200._ffi import check, lib, ptr
    thr = ffi.thresholds_fixed64(beta, D)
    n = ffi.CB_T1 + (D - 1) * ffi.CB_TC
    tbl = np.zeros(n, np.uint32)
    check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), D, ptr(tbl), n))
    assert np.array_equal(tbl, ffi.cb_sparse_tables(thr))
    with pytest.raises(ValueError):
        check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), D, ptr(tbl), n - 1))
