"""Workload for ncu captures of the chain kernel on BASELINE config 5 (GraphQSKT Nk=1024, M=64, rrrMC)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb
R = int(sys.argv[2]) if len(sys.argv) > 2 else 64
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
X = rb.GraphQSKT(1024, 64, 0.3, 2.0, replicas=R, rng=np.random.default_rng(4))
Es, C = rb.rrrMC(X, 2.0, iters, step=iters, seed=3, quiet=True)
print("it/s %.4g" % (R * iters / (X.last_run.device_ms * 1e-3)))
