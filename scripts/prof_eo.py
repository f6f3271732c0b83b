"""Workload for an ncu capture of the chain kernel running extremal_opt (RRG N=10^4, K=3, τ=1.3).
usage: python scripts/prof_eo.py [iters] [replicas]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N, K, tau = 10_000, 3, 1.3
rng = np.random.default_rng(8370000274 % 2 ** 32)
A = rb.gen_RRG(N, K, rng)
J = rb.gen_J_graph(lambda n: rng.choice([-1.0, 1.0], n), A).astype(np.int64)
X = rb.GraphRRG(N, K, replicas=R, A=A, J=J)
C0 = rb.Config(N, R, rng=np.random.default_rng(4))
rb.extremal_opt(X, tau, iters, step=iters, seed=3, C0=C0, quiet=True)
info = X.last_run
print("extremal_opt R", R, "moves/s %.4g" % (R * iters / (info.device_ms * 1e-3)), "device_ms %.1f" % info.device_ms)
