"""Probe: is the checkerboard launch time bound by something other than instruction count? (tuning aid)"""
import os, sys, subprocess, threading
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb
from rrrmc_b200._ffi import check, lib, ptr

def run(L, R, beta, K, M, nsw=100):
    X = rb.GraphEA(L, 3, replicas=R, rng=np.random.default_rng(1))
    st = X._ensure_state()
    check(lib().rrrmc_state_randomize(st, 5))
    thr = np.array([min(int(np.exp(-beta * 4 * c) * 2.0 ** 64), 2 ** 64 - 1) for c in range(1, 4)], dtype=np.uint64)
    check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), 3, 6, 8, 1, 0, 100))
    X.ctx.sync()
    best = 1e9
    for rep in range(3):
        X.ctx.timer_start()
        check(lib().rrrmc_checkerboard_sweeps(st, ptr(thr), 3, K, M, 3, 1000 * rep, nsw))
        best = min(best, X.ctx.timer_stop())
    print(f"L={L} R={R} beta={beta} K={K} M={M}: {best / nsw / 2 * 1e3:7.2f} us/launch  {nsw * X.N * R / (best * 1e-3):.3e} attempts/s", flush=True)

p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-lms", "200"], stdout=subprocess.PIPE, text=True)
lines = []
threading.Thread(target=lambda: [lines.append(l.strip()) for l in p.stdout], daemon=True).start()
for (L, R) in ((64, 1024), (64, 2048), (64, 512), (48, 1024), (32, 1024)):
    for beta in (1.0, 0.01):
        for (K, M) in ((0, 0), (4, 0), (6, 8), (10, 8), (16, 8)):
            run(L, R, beta, K, M)
p.terminate()
print("clock samples:", sorted(set(lines))[:12])
