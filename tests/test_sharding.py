"""Host logic of the multi-GPU path (replica sharding, observable all-gather, parallel-tempering label swaps),
including world-size-2 gloo runs on CPU. No compute kernels are involved."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("rrrmc_sharding", os.path.join(ROOT, "rrrmc.jl_b200", "sharding.py"))
sh = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sh)


def test_replica_ranges_partition_and_align():
    for total, world in ((1024, 1), (1024, 2), (1024, 8), (1152, 4), (8192, 8), (384, 8)):
        r = [sh.replica_range(k, world, total) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % 128 == 0 and hi % 128 == 0 for lo, hi in r)
    with pytest.raises(ValueError):
        sh.replica_range(0, 2, 1000)


def test_tempering_is_deterministic_and_detailed_balance():
    betas = np.linspace(0.5, 2.0, 8)
    a = sh.TemperingLadder(betas, seed=3); b = sh.TemperingLadder(betas, seed=3)
    rng = np.random.default_rng(0)
    for sweep in range(50):
        E = rng.normal(-100, 10, 8)
        assert np.array_equal(a.swap(E, sweep), b.swap(E, sweep))
    assert sorted(a.order) == list(range(8)) and np.array_equal(np.sort(a.beta_of_replica()), betas)
    # a swap that lowers the action is always accepted; one that raises it by a lot never
    l = sh.TemperingLadder([1.0, 2.0], seed=1)
    l.swap(np.array([-50.0, -10.0]), 0)   # the colder label (β=2) moves to the lower energy: ΔS = (β0-β1)(E1-E0) = -40
    assert list(l.order) == [1, 0]
    l.swap(np.array([-5000.0, -10.0]), 0)  # moving it back would cost e^{-4990}
    assert list(l.order) == [1, 0]
    # exact two-level check: acceptance frequency of an uphill swap ~ exp(-ΔS)
    acc = 0
    for sweep in range(0, 4000, 2):
        l2 = sh.TemperingLadder([1.0, 2.0], seed=9)
        l2.swap(np.array([-0.5, -1.0]), sweep)  # ΔS = (β0-β1)(E1-E0) = 0.5
        acc += list(l2.order) == [1, 0]
    assert abs(acc / 2000 - np.exp(-0.5)) < 0.04


def test_quantum_action_reduces_to_classical_when_trotter_term_vanishes():
    act = sh.quantum_action(M=8, Gamma=0.5)
    assert np.isclose(act(2.0, (0.0, 16.0)), 2.0 * 16.0 / 8)
    l = sh.TemperingLadder([1.0, 2.0], seed=0, action=act)
    l.swap((np.array([0.0, 0.0]), np.array([-80.0, -8.0])), 0)
    assert list(l.order) == [1, 0]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 384  # 3 blocks of 128 over 2 ranks: 128 + 256
        shard = sh.ReplicaShard(total)
        rng = np.random.default_rng(100 + rank)
        E_local = rng.normal(size=shard.count)
        E_all = sh.all_gather(E_local)
        m_all = sh.all_gather(np.stack([E_local, -E_local], axis=1))
        ladder = sh.TemperingLadder(np.linspace(0.5, 2.0, total), seed=5)
        b = None
        for sweep in range(4):
            b = ladder.swap(E_all, sweep)
        q.put((rank, shard.lo, shard.hi, E_all, m_all, b, shard.local(b)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_gather_and_identical_swaps():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, lo0, hi0, E0, m0, b0, bl0), (r1, lo1, hi1, E1, m1, b1, bl1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 128, 128, 384)
    want = np.concatenate([np.random.default_rng(100).normal(size=128), np.random.default_rng(101).normal(size=256)])
    assert np.array_equal(E0, want) and np.array_equal(E1, want)          # gather in rank order, ragged shards
    assert np.array_equal(m0[:, 1], -want) and np.array_equal(m1, m0)
    assert np.array_equal(b0, b1)                                          # every rank took the same swap decisions
    assert np.array_equal(np.concatenate([bl0, bl1]), b0)
    assert not np.array_equal(b0, np.linspace(0.5, 2.0, 384))


class _ToyEngine:
    """Stand-in for a GraphQuant batch in the tempered driver: the 'configuration' is one number per replica that
    relaxes towards −β (so colder labels end lower), observable terms are simple functions of it."""

    def __init__(self, n):
        self.replicas, self.M, self.betas_seen, self.seeds_seen = n, 4, [], []

    def set_betas(self, b):
        self.betas_seen.append(np.array(b, np.float64))


def _toy_sampler(X, b, iters, *, step, seed, C0, quiet):
    x = np.zeros(X.replicas) if C0 is None else C0
    X.seeds_seen.append(seed)
    rng = np.random.default_rng(seed)
    x = 0.5 * x - 0.5 * np.asarray(b) * 10 + rng.normal(size=X.replicas) * 0.01
    return x[None, :], x


def _pt_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        total = 256
        shard = sh.ReplicaShard(total)
        X = _ToyEngine(shard.count)
        ladder = sh.TemperingLadder(np.geomspace(0.5, 4.0, total), seed=11, action=sh.quantum_action(X.M, 0.3))
        hist, C = sh.tempered_run(X, ladder, shard, 5, 100, _toy_sampler, seed=3, energy_fn=lambda X_, c: c,
                                  terms_fn=lambda X_, c: (np.rint(c), 2.0 * c))
        q.put((rank, hist, ladder.order.copy(), ladder.accepts.copy(), [b.copy() for b in X.betas_seen], shard.lo, shard.hi, list(X.seeds_seen)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_tempered_run_quantum_ladder():
    """The tempered driver over two ranks: identical swap decisions everywhere, each rank's engine is told the β its
    replicas hold before every round (GraphQuant's fourK follows the label), histories are the all-gathered energies."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_pt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda t: t[0])
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, h0, o0, a0, seen0, lo0, hi0, sd0), (_, h1, o1, a1, seen1, lo1, hi1, sd1) = res
    # the shards' sampler seeds differ in every round (independent chains and initial configurations across ranks),
    # and are reproducible functions of (seed, round, first global replica of the shard)
    assert len(sd0) == 5 and len(sd1) == 5 and not set(sd0) & set(sd1)
    assert sd0 == [3 + 7919 * rd for rd in range(5)] and sd1 == [3 + sh.RANK_SEED_STRIDE * 128 + 7919 * rd for rd in range(5)]
    assert np.array_equal(h0, h1) and h0.shape == (5, 256)
    assert np.array_equal(o0, o1) and np.array_equal(a0, a1) and a0.sum() > 0
    assert (lo0, hi0, lo1, hi1) == (0, 128, 128, 256)
    assert len(seen0) == 5 and len(seen1) == 5
    assert np.array_equal(np.concatenate([seen0[0], seen1[0]]), np.geomspace(0.5, 4.0, 256))   # round 0: the initial ladder
    moved = [not np.array_equal(np.concatenate([a, b]), np.geomspace(0.5, 4.0, 256)) for a, b in zip(seen0[1:], seen1[1:])]
    assert any(moved)
    for a, b in zip(seen0, seen1):      # every round the two shards together hold each rung exactly once
        assert np.array_equal(np.sort(np.concatenate([a, b])), np.geomspace(0.5, 4.0, 256))
