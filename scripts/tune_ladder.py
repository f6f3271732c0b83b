"""Device timing of the β-ladder checkerboard sweeps (per-group count tables) against the one-β kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb
from rrrmc_b200 import _ffi
from rrrmc_b200._ffi import check, lib, ptr

L, D, R = 64, 3, 1024
X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(1))
st = X._ensure_state()
check(lib().rrrmc_state_randomize(st, 5))
ctx = X.ctx
def tables(b):
    thr = np.array([min(int(np.exp(-b * 4 * c) * 2.0 ** 64), 2 ** 64 - 1) for c in range(1, D + 1)], dtype=np.uint64)
    t = np.zeros(_ffi.CBP_LEN, np.uint32)
    check(lib().rrrmc_checkerboard_poisson_tables(ptr(thr), D, ptr(t), len(t)))
    return t
for name, bg, NW in [("all 1.0", [1.0] * 8, 2), ("1.00..1.035", list(np.linspace(1.0, 1.035, 8)), 2), ("0.8..1.6", list(np.geomspace(0.8, 1.6, 8)), 4), ("1.5..2.2", list(np.linspace(1.5, 2.2, 8)), 1)]:
    tb = np.stack([tables(b) for b in bg])
    check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(st, ptr(tb), 8, NW, 3, 0, 50))
    ctx.sync()
    best = 1e9
    for rep in range(3):
        ctx.timer_start()
        check(lib().rrrmc_checkerboard_sweeps_poisson_ladder(st, ptr(tb), 8, NW, 3, 100 * rep, 100))
        best = min(best, ctx.timer_stop())
    t1 = tables(bg[0])
    check(lib().rrrmc_checkerboard_sweeps_poisson(st, ptr(t1), len(t1), NW, 3, 0, 20))
    ctx.timer_start()
    check(lib().rrrmc_checkerboard_sweeps_poisson(st, ptr(t1), len(t1), NW, 3, 0, 100))
    one = ctx.timer_stop()
    print(f"ladder {name} NW={NW}: {best * 10:.2f} us/sweep ({100 * X.N * R / (best * 1e-3):.3e} attempts/s); one-β kernel at β={bg[0]:.2f}: {one * 10:.2f} us/sweep", flush=True)
