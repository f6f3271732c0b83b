"""Workload for ncu captures of the chain kernels (C3: GraphEA 3D L=32 ±J, β=3, 256 replicas).
usage: python scripts/prof_chain.py [rrrMC|bklMC] [iters] [replicas]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb

name = sys.argv[1] if len(sys.argv) > 1 else "rrrMC"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
R = int(sys.argv[3]) if len(sys.argv) > 3 else 256
L, D, beta = 32, 3, 3.0
X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(1))
_, C = rb.standardMC(X, beta, 200 * X.N, step=200 * X.N, seed=1, quiet=True, schedule="checkerboard")
fn = {"rrrMC": rb.rrrMC, "bklMC": rb.bklMC}[name]
Es, C2 = fn(X, beta, iters, step=iters, seed=3, C0=C, quiet=True)
info = X.last_run
print(name, "R", R, "moves/s %.4g" % (info.accepted_total / (info.device_ms * 1e-3)), "device_ms %.1f" % info.device_ms)
