"""Parallel tempering by β-label exchange on the GPU engine (single rank; the multi-rank plumbing is covered on CPU
by tests/test_sharding.py): per-replica β in the reference-order samplers, swap bookkeeping, and the expected
monotone ⟨E⟩(β)."""
import importlib.util
import os

import numpy as np
import pytest

import rrrmc_b200 as rb
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("rrrmc_sharding", os.path.join(ROOT, "rrrmc.jl_b200", "sharding.py"))
sh = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(sh)


def test_tempered_run_on_ea():
    L, D, R = 4, 3, 128
    A, J = ea_instance(L, D, seed=5)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    ladder = sh.TemperingLadder(np.linspace(0.2, 2.0, R), seed=1)
    shard = sh.ReplicaShard(R, rank=0, world=1)
    hist, C = sh.tempered_run(X, ladder, shard, rounds=30, iters_per_round=20 * X.N,
                              sampler=lambda X_, b, it, **kw: rb.standardMC(X_, b, it, schedule="random", **kw), seed=3)
    assert hist.shape == (30, R) and sorted(ladder.order) == list(range(R))
    assert ladder.accepts.sum() > 0 and (ladder.accepts <= ladder.attempts).all()
    # energies sorted by the β each replica holds at the end: colder labels sit at lower energy on average
    b = ladder.beta_of_replica()
    E = hist[-1]
    assert E[np.argsort(b)][: R // 4].mean() > E[np.argsort(b)][-R // 4:].mean()


def test_per_replica_beta_in_chain_samplers():
    """β[R] is honoured per replica: a batch with two temperatures equals two separate runs, chain by chain."""
    L, D, R = 4, 2, 4
    A, J = ea_instance(L, D, seed=6)
    X = rb.GraphEA(L, D, replicas=R, A=A, J=J)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(0))
    betas = np.array([0.5, 0.5, 2.0, 2.0])
    Es, Cf = rb.rrrMC(X, betas, 1000, step=100, seed=9, C0=C0, quiet=True)
    for b in (0.5, 2.0):
        Es1, C1 = rb.rrrMC(X, b, 1000, step=100, seed=9, C0=C0, quiet=True)
        sel = betas == b
        assert np.array_equal(Es[:, sel], Es1[:, sel]) and np.array_equal(Cf.chunks[sel], C1.chunks[sel])
