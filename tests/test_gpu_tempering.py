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


def _quant(R, Nk=8, M=5, G=0.5, beta0=2.0, seed=9):
    from tests.helpers import sk_binary
    J = sk_binary(Nk, seed)
    return rb.GraphQSKT(Nk, M, G, beta0, replicas=R, J=J), J


@pytest.mark.parametrize("sampler", ["standardMC", "rrrMC", "bklMC", "wtmMC"])
def test_quant_beta_ladder_bit_exact_vs_oracle(sampler):
    """A GraphQuant batch as a β ladder: replica r must behave exactly like the reference's GraphQuant(Nk, M, Γ, β_r, …)
    — a different fourK type parameter per β (QT.jl:126,165) — chain by chain against the oracle."""
    from oracle import ffi
    R, Nk, M, G = 4, 8, 5, 0.5
    X, J = _quant(R, Nk, M, G)
    betas = np.array([0.7, 1.3, 2.0, 3.1])
    fk = X.set_betas(betas)
    C0 = rb.Config(X.N, R, rng=np.random.default_rng(4))
    gs = [ffi.Graph.quant(Nk, M, G, b, ffi.SK_BIN, J) for b in betas]
    assert np.array_equal(fk, [g.fourK() for g in gs]) and len(set(fk)) == R
    E = np.atleast_1d(rb.energy(X, C0))
    for r in range(R):
        assert E[r] == gs[r].energy(C0.chunks[r])
        want = np.array([gs[r].delta_energy(C0.chunks[r], i) for i in range(1, X.N + 1)])
        assert np.array_equal(np.asarray(rb.all_delta_energy(X, C0, r), np.float64), want)
    if sampler == "wtmMC":
        Es, Cf = rb.wtmMC(X, betas, 60, step=1.5, seed=11, C0=C0, quiet=True)
    elif sampler == "standardMC":
        Es, Cf = rb.standardMC(X, betas, 3000, step=50, seed=11, C0=C0, quiet=True, schedule="random")
    else:
        Es, Cf = getattr(rb, sampler)(X, betas, 3000, step=50, seed=11, C0=C0, quiet=True)
    Es = np.asarray(Es, np.float64).reshape(-1, R)
    for r in range(R):
        s = C0.chunks[r].copy()
        if sampler == "wtmMC":
            want, _ = ffi.wtmMC(gs[r], betas[r], 60, s, ffi.PhiloxDraws(11, chain=r), step=1.5)
        else:
            want, _ = getattr(ffi, sampler)(gs[r], betas[r], 3000, s, ffi.PhiloxDraws(11, chain=r), step=50)
        assert np.array_equal(Es[:, r], want), (sampler, r)
        assert np.array_equal(Cf.chunks[r], s), (sampler, r)
    # observables follow the replica's own (β, fourK): Qenergy QT.jl:253-268, transverse_mag QT.jl:113-121
    Q = np.atleast_1d(rb.Qenergy(X, Cf))
    for r in range(R):
        want = ffi.lib().orc_Qenergy(gs[r].h, Cf.chunks[r].copy())
        assert np.isclose(Q[r], want, rtol=1e-12, atol=1e-12), (r, Q[r], want)
    # back to the graph's own β
    X.set_betas(None)
    g0 = ffi.Graph.quant(Nk, M, G, 2.0, ffi.SK_BIN, J)
    assert np.atleast_1d(rb.energy(X, C0))[2] == g0.energy(C0.chunks[2])


def test_quant_tempered_run_swaps_follow_quantum_action():
    """Parallel tempering of a GraphQuant batch (BASELINE config 5): β labels move between replicas, each replica's fourK
    follows its label, energies stay consistent with energy(X, C) under the current labels."""
    from oracle import ffi
    R, Nk, M, G = 128, 8, 4, 0.6
    X, J = _quant(R, Nk, M, G)
    betas = np.geomspace(0.5, 4.0, R)
    ladder = sh.TemperingLadder(betas, seed=3, action=sh.quantum_action(M, G))
    shard = sh.ReplicaShard(R, rank=0, world=1)
    hist, C = sh.tempered_run(X, ladder, shard, rounds=12, iters_per_round=40 * X.N, sampler=rb.rrrMC, seed=5,
                              terms_fn=sh.quant_terms)
    assert hist.shape == (12, R) and sorted(ladder.order) == list(range(R))
    assert ladder.accepts.sum() > 0 and (ladder.accepts <= ladder.attempts).all()
    assert not np.array_equal(ladder.beta_of_replica(), betas)            # labels moved
    # the last round ran under the labels before the final swap: recompute them and check two replicas on the oracle
    e0, ecl = sh.quant_terms(X, C)
    for r in (0, 77):
        g = ffi.Graph.quant(Nk, M, G, X.betas[r], ffi.SK_BIN, J)
        E = g.energy(C.chunks[r])
        assert abs(E - (g.fourK() / 4 * e0[r] + ecl[r] / M)) <= 1e-9 * max(1.0, abs(E))
        assert abs(hist[-1, r] - E) <= 1e-9 * max(1.0, abs(E))


def test_tempered_run_on_device_matches_host_round_trips():
    """tempered_run(on_device=True) keeps the batch on the device between rounds (ON_DEVICE sentinel): same energies,
    swaps and final configuration as the default path that round-trips the configuration through the host."""
    from rrrmc_b200 import sharding as sh
    X = rb.GraphEA(4, 3, replicas=128, rng=np.random.default_rng(3))
    C0 = rb.Config(X.N, 128, rng=np.random.default_rng(4))
    out = []
    for on_dev in (False, True):
        ladder = sh.TemperingLadder(np.geomspace(0.5, 3.0, 128), seed=5)
        shard = sh.ReplicaShard(128, rank=0, world=1)
        hist, C = sh.tempered_run(X, ladder, shard, 4, 500, rb.rrrMC, seed=9, C0=C0, on_device=on_dev)
        out.append((hist, np.asarray(C.chunks).copy(), ladder.order.copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
