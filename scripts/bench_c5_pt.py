#!/usr/bin/env python
"""BASELINE config 5 with its multi-GPU part: GraphQSKT (Suzuki-Trotter, Nk=1024, M=64, Γ=0.3) under rrrMC, replicas
sharded across the GPUs of one node (one process per GPU), parallel-tempering swaps of β labels after every round.
The only exchange is an NCCL all-gather of three scalars per replica (E, e0, E_cl); every rank takes identical swap
decisions from a shared counter RNG, and a replica's fourK follows the β it holds (QT.jl:165). No spin data moves.

  python scripts/bench_c5_pt.py                                   # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/bench_c5_pt.py
env: C5_R (replicas per GPU, default 128), C5_ROUNDS (default 6), C5_ITERS (iterations per round, default 20000)"""
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
os.environ.setdefault("RRRMC_DEVICE", str(local))
import rrrmc_b200 as rb

from rrrmc_b200 import sharding as sh

Nk, M, G = 1024, 64, 0.3
Rl = int(os.environ.get("C5_R", "128")); R = Rl * world
rounds = int(os.environ.get("C5_ROUNDS", "6")); iters = int(os.environ.get("C5_ITERS", "20000"))
ctx = rb.Context(device=local)
X = rb.GraphQSKT(Nk, M, G, 2.0, replicas=Rl, rng=np.random.default_rng(4), ctx=ctx)   # same instance on every rank
betas = np.geomspace(0.5, 4.0, R)
ladder = sh.TemperingLadder(betas, seed=17, action=sh.quantum_action(M, G))
shard = sh.ReplicaShard(R, rank=rank, world=world)
dev_ms = []


def sampler(X_, b, it, **kw):
    out = rb.rrrMC(X_, b, it, **kw)
    dev_ms.append(X_.last_run.device_ms)
    return out


sh.tempered_run(X, ladder, shard, 1, iters // 10, sampler, seed=1, terms_fn=sh.quant_terms, on_device=True)      # warm-up round
dev_ms.clear()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
hist, C = sh.tempered_run(X, ladder, shard, rounds, iters, sampler, seed=100, terms_fn=sh.quant_terms, on_device=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
t = torch.tensor([dt, sum(dev_ms) * 1e-3], device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    wall, dev = float(t[0]), float(t[1])
    print(json.dumps({"config": "C5+PT", "sampler": "rrrMC(DoubleGraph)", "Nk": Nk, "M": M, "Gamma": G, "n_gpus": world,
                      "replicas_total": R, "replicas_per_gpu": Rl, "beta_ladder": [float(betas[0]), float(betas[-1])],
                      "rounds": rounds, "iters_per_round": iters,
                      "iterations_per_s_wall": R * rounds * iters / wall, "iterations_per_s_device": R * rounds * iters / dev,
                      "swap_accept_rate": float(ladder.accepts.sum() / max(1, ladder.attempts.sum())),
                      "wall_s": wall, "device_s_max_over_ranks": dev,
                      "exchange": "all-gather of (E, e0, E_cl) per replica per round; labels only, no spin data",
                      "mean_E_per_N_coldest": float(hist[-1][ladder.order[-1]] / (Nk * M))}), flush=True)
if world > 1:
    dist.destroy_process_group()
