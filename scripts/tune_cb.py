"""Device-only timing sweep of the checkerboard kernel over planes_K / planes_M / occupancy variants / beta (tuning aid)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb
from rrrmc_b200._ffi import check, lib, ptr

L, D, R = 64, 3, int(os.environ.get("TUNE_R", "1024"))
X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(1))
st = X._ensure_state()
check(lib().rrrmc_state_randomize(st, 5))
ctx = X.ctx
NSW = 100
for beta in [float(b) for b in os.environ.get("TUNE_BETAS", "1.0").split(",")]:
    thr = np.array([min(int(np.exp(-beta * 4 * c) * 2.0 ** 64), 2 ** 64 - 1) for c in range(1, D + 1)], dtype=np.uint64)
    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), D, 6, 8, 1, 0, 300))  # equilibrate a bit at this beta
    for variant in [int(v) for v in os.environ.get("TUNE_VARIANTS", "0").split(",")]:
        os.environ["RRRMC_CB_VARIANT"] = str(variant)
        if os.environ.get("TUNE_SPARSE", "1") == "1":   # sparse acceptance procedure (no K/M knobs)
            n = 33 + (D - 1) * 129
            tbl = np.zeros(n, np.uint32)
            check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), D, ptr(tbl), n))
            check(lib().rrrmc_checkerboard_sweeps_sparse(st, ptr(tbl), n, 2, 0, 20))
            ctx.sync()
            best = 1e9
            for rep in range(3):
                ctx.timer_start()
                check(lib().rrrmc_checkerboard_sweeps_sparse(st, ptr(tbl), n, 3, 1000 * rep, NSW))
                best = min(best, ctx.timer_stop())
            rate = NSW * X.N * R / (best * 1e-3)
            print(f"beta={beta} variant={variant} sparse      {best / NSW * 1e3:8.2f} us/sweep  {rate:.3e} attempts/s", flush=True)
        for M in [int(m) for m in os.environ.get("TUNE_MS", "0,4,8,12").split(",")]:
            for K in [int(k) for k in os.environ.get("TUNE_KS", "4,5,6,7,8,10").split(",")]:
                check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), D, K, M, 2, 0, 20))
                ctx.sync()
                best = 1e9
                for rep in range(3):
                    ctx.timer_start()
                    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), D, K, M, 3, 1000 * rep, NSW))
                    best = min(best, ctx.timer_stop())
                rate = NSW * X.N * R / (best * 1e-3)
                print(f"beta={beta} variant={variant} K={K:2d} M={M:2d}  {best / NSW * 1e3:8.2f} us/sweep  {rate:.3e} attempts/s", flush=True)
