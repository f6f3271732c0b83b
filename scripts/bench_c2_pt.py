#!/usr/bin/env python
"""BASELINE config 2 with parallel tempering: GraphEA 3D L=64 ±J, 1024 replicas per GPU on the checkerboard schedule, a
β ladder of eight rungs laid over the 128-replica groups (128 independent ladders per GPU). A round = `C2_SWEEPS` sweeps
(one launch of the multi-sweep brick kernel with per-group count tables) + rrrmc_tempering_exchange (energies, decisions
and the exchange of configurations are device kernels). Replica shards hold whole ladders: no collective inside the loop;
NCCL carries only the barrier, the max-over-ranks time and the final reduction of the observables.

  python scripts/bench_c2_pt.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_c2_pt.py
env: C2_ROUNDS (default 40), C2_SWEEPS (sweeps per round, default 10), C2_BLO / C2_BHI (ladder ends, default 1.0 / 1.035: at N = 262144 neighbouring rungs must be ~0.005 apart to exchange)"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import rrrmc_b200 as rb

from rrrmc_b200 import sharding as sh

L, D, R = 64, 3, 1024
rounds = int(os.environ.get("C2_ROUNDS", "40")); nsw = int(os.environ.get("C2_SWEEPS", "10"))
bg = np.geomspace(float(os.environ.get("C2_BLO", "1.0")), float(os.environ.get("C2_BHI", "1.035")), R // 128)
ctx = rb.Context(device=local)
X = rb.GraphEA(L, D, replicas=R, rng=np.random.default_rng(4), ctx=ctx)           # same instance on every rank
C0 = rb.Config(X.N, R, rng=np.random.default_rng(100 + rank))                      # different configurations per rank
sh.tempered_checkerboard(X, bg, 2, nsw, seed=1 + 1000 * rank, C0=C0)              # warm-up
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
ctx.timer_start()
acc, att = sh.tempered_checkerboard(X, bg, rounds, nsw, seed=7 + 1000 * rank)
dev_ms = ctx.timer_stop()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
E = rb.energy(X, X._download()).reshape(R // 128, 128).mean(axis=1) / X.N
t = torch.tensor([dt, dev_ms * 1e-3], device="cuda", dtype=torch.float64)
obs = torch.tensor(np.concatenate([acc / np.maximum(att, 1), E]), device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(obs, op=dist.ReduceOp.SUM)
obs = (obs / world).cpu().numpy()
if rank == 0:
    wall, dev = float(t[0]), float(t[1])
    attempts = world * X.N * R * rounds * nsw
    print(json.dumps({"config": "C2+PT", "sampler": "standardMC(checkerboard, β ladder) + device exchange", "L": L, "D": D, "n_gpus": world,
                      "replicas_per_gpu": R, "ladders_per_gpu": 128, "beta_ladder": [float(b) for b in bg],
                      "rounds": rounds, "sweeps_per_round": nsw,
                      "attempts_per_s_wall": attempts / wall, "attempts_per_s_device": attempts / dev,
                      "wall_s": wall, "device_s_max_over_ranks": dev,
                      "swap_accept_rate_per_pair": [float(x) for x in obs[:R // 128 - 1]],
                      "mean_E_per_N_per_rung": [float(x) for x in obs[R // 128 - 1:]],
                      "exchange": "device kernels (k_energy_pm1, k_pt_decide, k_pt_exchange); no PCIe traffic and no collective in the loop"}), flush=True)
if world > 1:
    dist.destroy_process_group()
