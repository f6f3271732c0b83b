"""Device-only timing of the checkerboard acceptance procedures at the headline size (tuning aid):
sparse vs poisson with NW static position words, over β. TUNE_BETAS, TUNE_NWS, TUNE_VARIANTS, TUNE_R."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb
from rrrmc_b200._ffi import check, lib, ptr

L, D, R = int(os.environ.get("TUNE_L", "64")), 3, int(os.environ.get("TUNE_R", "1024"))
X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(1))
st = X._ensure_state()
check(lib().rrrmc_state_randomize(st, 5))
ctx = X.ctx
NSW = 100


def timed(fn):
    fn(2, 0, 20)
    ctx.sync()
    best = 1e9
    for rep in range(3):
        ctx.timer_start()
        fn(3, 1000 * rep, NSW)
        best = min(best, ctx.timer_stop())
    return best


for beta in [float(b) for b in os.environ.get("TUNE_BETAS", "1.0").split(",")]:
    thr = np.array([min(int(np.exp(-beta * 4 * c) * 2.0 ** 64), 2 ** 64 - 1) for c in range(1, D + 1)], dtype=np.uint64)
    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), D, 6, 8, 1, 0, 300))  # equilibrate a bit at this beta
    n = 33 + (D - 1) * 129
    tbl = np.zeros(n, np.uint32)
    check(lib().rrrmc_checkerboard_sparse_tables(ptr(thr), D, ptr(tbl), n))
    ptbl = np.zeros(160, np.uint32)
    check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), D, ptr(ptbl), 160))
    for variant in [int(v) for v in os.environ.get("TUNE_VARIANTS", "0").split(",")]:
        os.environ["RRRMC_CB_VARIANT"] = str(variant)
        if os.environ.get("TUNE_SPARSE", "1") == "1":
            best = timed(lambda seed, t0, k: check(lib().rrrmc_checkerboard_sweeps_sparse(st, ptr(tbl), n, seed, t0, k)))
            print(f"beta={beta} variant={variant} sparse        {best / NSW * 1e3:8.2f} us/sweep  {NSW * X.N * R / (best * 1e-3):.3e} attempts/s", flush=True)
        for NW in [int(v) for v in os.environ.get("TUNE_NWS", "1,2,4,6").split(",")]:
            pover = 1.0 - (float(ptbl[4 * NW - 1]) + 1.0) / 2.0 ** 32
            best = timed(lambda seed, t0, k: check(lib().rrrmc_checkerboard_sweeps_poisson(st, ptr(ptbl), 160, NW, seed, t0, k)))
            print(f"beta={beta} variant={variant} poisson NW={NW}  {best / NSW * 1e3:8.2f} us/sweep  {NSW * X.N * R / (best * 1e-3):.3e} attempts/s  P(overflow)={pover:.2e}", flush=True)
