"""Workload for ncu captures of k_chain_warp: C3 (L=32, 3D, beta=3, 256 chains), a short rrrMC run."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rrrmc_b200 as rb
L, D, R, beta = 32, 3, 256, 3.0
X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(1))
_, C = rb.standardMC(X, beta, 100 * X.N, step=100 * X.N, seed=1, quiet=True, schedule="checkerboard")
fn = rb.bklMC if (len(sys.argv) > 1 and sys.argv[1] == "bkl") else rb.rrrMC
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
fn(X, beta, iters, step=iters, seed=3, C0=C, quiet=True, site_pick="rank")
