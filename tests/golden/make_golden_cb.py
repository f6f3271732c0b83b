#!/usr/bin/env python
"""Generates tests/golden/checkerboard_v2.npz from the CPU oracle: the engine's count-table acceptance procedures
(poisson: oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_poisson; sparse: orc_checkerboard_sweeps_sparse) on a small
3D ±J instance, with the tables themselves. Same caveat as make_golden.py: the reference (pure Julia, no checkerboard
schedule) cannot pin these; they freeze the oracle's and the engine's answers against silent drift.
Run:  python tests/golden/make_golden_cb.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ffi  # noqa: E402
from tests.helpers import ea_instance  # noqa: E402

L, D, R, SEED, SWEEP0, NSW = 8, 3, 128, 4242, (1 << 32) + 5, 3
CASES = {"poisson_b0.9_nw2": (0.9, 2), "poisson_b0.6_nw1": (0.6, 1), "poisson_b1.5_nw1": (1.5, 1)}


def forward_couplings(A, J):
    N = A.shape[0]
    Jf = np.zeros((N, D), np.int8)
    for i in range(N):
        stride = 1
        for d in range(D):
            c = (i // stride) % L
            up = i + (((c + 1) % L) - c) * stride
            Jf[i, d] = J[i, np.flatnonzero(A[i] == up + 1)[0]]
            stride *= L
    return Jf


def compute():
    out = {}
    A, J = ea_instance(L, D, seed=21)
    Jf = forward_couplings(A, J)
    sp0 = np.random.default_rng(23).integers(0, 2 ** 32, (L ** D, R // 32), dtype=np.uint32)
    out["initial"] = sp0
    for name, (beta, nw) in CASES.items():
        thr = ffi.thresholds_fixed64(beta, D)
        tbl = ffi.cb_poisson_tables(thr)
        sp = sp0.copy(); acc = np.zeros(R, np.int64)
        ffi.checkerboard_sweeps_poisson(L, D, R, sp, Jf, tbl, nw, SEED, SWEEP0, NSW, acc)
        out[f"{name}/thr"] = thr; out[f"{name}/tbl"] = tbl; out[f"{name}/final"] = sp; out[f"{name}/accepted"] = acc
    thr = ffi.thresholds_fixed64(0.9, D)
    tbl = ffi.cb_sparse_tables(thr)
    sp = sp0.copy(); acc = np.zeros(R, np.int64)
    ffi.checkerboard_sweeps_sparse(L, D, R, sp, Jf, tbl, SEED, SWEEP0, NSW, acc)
    out["sparse_b0.9/thr"] = thr; out["sparse_b0.9/tbl"] = tbl; out["sparse_b0.9/final"] = sp; out["sparse_b0.9/accepted"] = acc
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "checkerboard_v2.npz")
    np.savez_compressed(path, **compute())
    print("wrote", path, os.path.getsize(path), "bytes")
