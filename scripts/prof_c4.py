#!/usr/bin/env python
"""One lock-step sweep of BASELINE config 4 (GraphSKNormal N=4096 x 512 replicas) for an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:k_sk_lockstep -c 1 -o gpurun_out/c4 python scripts/prof_c4.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rrrmc_b200 as rb

N, R = 4096, 512
X = rb.GraphSKNormal(N, replicas=R, rng=np.random.default_rng(2))
C0 = rb.Config(N, R, rng=np.random.default_rng(3))
rb.sk_fields_init(X, C0, tensor_cores=True)
E, acc, _ = rb.sk_metropolis_sweeps(X, 1.0, int(os.environ.get("NSW", "2")), seed=1)
print("acceptance", acc.sum() / (R * N * int(os.environ.get("NSW", "2"))))
