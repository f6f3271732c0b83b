"""Workload for ncu captures of the checkerboard kernels: equilibrate, then run a few sweeps of one procedure.
usage: python scripts/prof_cb.py [beta] [sparse|planes|poisson] [nsweeps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb
from rrrmc_b200._ffi import check, lib, ptr

beta = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
method = sys.argv[2] if len(sys.argv) > 2 else "sparse"
nsw = int(sys.argv[3]) if len(sys.argv) > 3 else 10
X = rb.GraphEA(64, 3, replicas=1024, rng=np.random.default_rng(1))
st = X._ensure_state()
check(lib().rrrmc_state_randomize(st, 5))
thr = np.array([min(int(np.exp(-beta * 4 * c) * 2.0 ** 64), 2 ** 64 - 1) for c in range(1, 4)], dtype=np.uint64)
tbl = np.zeros(33 + 2 * 129, np.uint32)
check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), 3, ptr(tbl), len(tbl)))
check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), 3, 5, 4, 1, 0, 200))       # equilibrate (planes kernel)
if method == "poisson":
    ptbl = np.zeros(160, np.uint32)
    check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), 3, ptr(ptbl), 160))
    NW = lib().rrrmc_checkerboard_poisson_nw(ptr(ptbl), 0.0)
    check(lib().rrrmc_checkerboard_sweeps_poisson(st, ptr(ptbl), 160, NW, 1, 200, nsw))
elif method == "sparse":
    check(lib().rrrmc_checkerboard_sweeps_sparse(st, ptr(tbl), len(tbl), 1, 200, nsw))
else:
    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), 3, 5, 4, 1, 200, nsw))
X.ctx.sync()
