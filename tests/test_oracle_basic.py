"""Pins for the CPU oracle that do not need the reference to run (SURVEY §8c 'additional pins')."""
import itertools

import numpy as np
import pytest

from oracle import ffi
from tests.helpers import bits, ea_instance, random_config, reference_graphs, sk_binary, sk_gauss


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
        ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
        ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
         [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
    ]
    for ctr, key, want in kats:
        out = np.zeros(4, np.uint32)
        ffi.lib().orc_philox4x32_10(np.array(ctr, np.uint32), np.array(key, np.uint32), out)
        assert out.tolist() == want


def test_gen_EA_conventions():
    # SURVEY Appendix A.2 (verified against EA.jl:24-43): L=4,D=3 site 1
    assert ffi.gen_EA(4, 3)[0].tolist() == [2, 4, 5, 13, 17, 49]
    # L=2: every neighbour twice (double bonds)
    assert ffi.gen_EA(2, 3)[0].tolist() == [2, 2, 3, 3, 5, 5]
    A = ffi.gen_EA(3, 2)
    assert A[0].tolist() == [2, 3, 4, 7] and A[4].tolist() == [2, 4, 6, 8]
    for L, D in ((5, 1), (4, 2), (3, 3), (6, 3)):
        A = ffi.gen_EA(L, D)
        assert (np.diff(A, axis=1) >= 0).all()
        # symmetric adjacency with equal multiplicities
        for x in range(len(A)):
            for y in A[x]:
                assert (A[y - 1] == x + 1).sum() == (A[x] == y).sum()


def test_gen_J_symmetric_and_draw_order():
    A = ffi.gen_EA(4, 2)
    nb = int((A > np.arange(1, 17)[:, None]).sum())
    draws = np.arange(1, nb + 1, dtype=np.float64)
    J = ffi.gen_J(A, draws)
    # first bond drawn belongs to site 1's first neighbour with y>x
    assert J[0, 0] == 1.0
    for x in range(16):
        for k, y in enumerate(A[x]):
            l = list(A[y - 1]).index(x + 1)
            assert J[x, k] == J[y - 1, l]


def test_allDE():
    A, J = ea_instance(3, 2)
    assert ffi.Graph.ea_int(A, J).allDE().tolist() == [0, 4, 8]           # EA.jl:293
    A, J = ea_instance(4, 3)
    assert ffi.Graph.ea_int(A, J).allDE().tolist() == [0, 4, 8, 12]
    A, J = ea_instance(3, 2, (-1, 0, 1))
    assert ffi.Graph.ea_int(A, J, (-1, 0, 1)).allDE().tolist() == [0, 2, 4, 6, 8]  # EA.jl:295-309
    g = ffi.Graph.quant(10, 8, 0.5, 2.0, ffi.EMPTY)
    fourK = round(2 / 2.0 * np.log(1 / np.tanh(2.0 * 0.5 / 8)), 8)      # QT.jl:165
    assert g.fourK() == fourK and g.allDE().tolist() == [0.0, fourK]     # QT.jl:111


@pytest.mark.parametrize("name", list(reference_graphs().keys()))
def test_delta_energy_matches_energy_difference(name):
    """delta_energy ≡ energy(flipped) − energy (generic fallback Interface.jl:130-138)."""
    g = reference_graphs()[name]
    s = random_config(g.N, 5)
    for i in range(1, g.N + 1):
        e0 = g.energy(s)
        d = g.delta_energy(s, i)
        s2 = s.copy(); s2[(i - 1) >> 6] ^= np.uint64(1 << ((i - 1) & 63))
        e1 = g.energy(s2)
        g.energy(s)
        assert d == pytest.approx(e1 - e0, abs=1e-11), (name, i)


@pytest.mark.parametrize("name", list(reference_graphs().keys()))
def test_cache_matches_fresh_recompute_after_moves(name):
    """lfields after update_cache! ≡ lfields from energy() (SK.jl:125-130 commented check), incl. the undo fast path."""
    g = reference_graphs()[name]
    rng = np.random.default_rng(3)
    s = random_config(g.N, 6)
    g.energy(s)
    seq = list(rng.integers(1, g.N + 1, 200))
    seq += [seq[-1], seq[-1], 3, 3, 3]  # exercise move_last == move
    for i in seq:
        g.spinflip(s, int(i))
        ds = np.array([g.delta_energy(s, j) for j in range(1, g.N + 1)])
        g.energy(s)
        ds2 = np.array([g.delta_energy(s, j) for j in range(1, g.N + 1)])
        assert np.allclose(ds, ds2, atol=1e-10), name


def test_closed_forms_tiny():
    # Ising ring of 3 ferro bonds as EA(3,1): energy −3 when aligned (cf. ThreeSpin.jl:26-47)
    A = ffi.gen_EA(3, 1)
    J = np.ones_like(A)
    g = ffi.Graph.ea_int(A, J)
    s = np.array([0b111], np.uint64)
    assert g.energy(s) == -3
    assert [g.delta_energy(s, i) for i in (1, 2, 3)] == [4, 4, 4]
    s = np.array([0b011], np.uint64)
    assert g.energy(s) == 1
    assert [g.delta_energy(s, i) for i in (1, 2, 3)] == [0, 0, -4]
    # GraphEmpty: energy ≡ 0 (Empty.jl:28-31)
    e = ffi.Graph.empty(7)
    assert e.energy(random_config(7)) == 0 and e.delta_energy(random_config(7), 3) == 0


def test_sk_binary_energy_definition():
    """GraphSK energy equals −Σ_{i<j} J_ij σ_i σ_j /√N with J=±1 (SK.jl:82-91 commented naive form)."""
    N = 10
    Jb = sk_binary(N, 2)
    g = ffi.Graph.sk_bin(Jb)
    s = random_config(N, 9)
    sig = 2 * bits(s, N) - 1
    Jpm = 2 * Jb.astype(np.int64) - 1
    np.fill_diagonal(Jpm, 0)
    want = -0.5 * sig @ Jpm @ sig / np.sqrt(N)
    assert g.energy(s) == pytest.approx(want, abs=1e-12)


def test_sk_normal_energy_definition():
    N = 10
    J = sk_gauss(N, 2)
    g = ffi.Graph.sk_f64(J)
    s = random_config(N, 9)
    sig = 2 * bits(s, N) - 1
    assert g.energy(s) == pytest.approx(-0.5 * sig @ J @ sig, abs=1e-12)


def test_quant_energy_definition():
    """E = E_QT + (1/M) Σ_k E_classical(slice k) (QT.jl:185-199); Trotter ring of QT.jl:68-84."""
    Nk, M = 10, 8
    J = sk_gauss(Nk, 7)
    g = ffi.Graph.quant(Nk, M, 0.5, 2.0, ffi.SK_F64, J)
    s = random_config(Nk * M, 4)
    sig = (2 * bits(s, Nk * M) - 1).reshape(M, Nk)
    e_cl = sum(-0.5 * sig[k] @ J @ sig[k] for k in range(M)) / M
    e_qt = -(sig * np.roll(sig, -1, axis=0)).sum() * g.fourK() / 4
    assert g.energy(s) == pytest.approx(e_cl + e_qt, abs=1e-10)
    nb = g.neighbors(13)  # slice 2, inner site 3: Trotter neighbours 3 and 23, then slice-2 sites except 13
    assert nb[:2].tolist() == [3, 23] and nb[2:].tolist() == [j for j in range(11, 21) if j != 13]


def test_discrete_cache_consistency():
    """check_consistency (DeltaE.jl:120-136) after every eager apply_move! on EA and Quant graphs."""
    rng = np.random.default_rng(0)
    for name in ("EA(3,2)", "EA(2,3)", "EA(3,2,(-1,0,1))", "Quant(10,8,SK)"):
        g = reference_graphs()[name]
        s = random_config(g.N, 2)
        sites = rng.integers(1, g.N + 1, 300).astype(np.int64)
        assert ffi.lib().orc_check_discrete_cache(g.h, s, 1.3, sites, len(sites)) == 0, name


def test_sk_lockstep_sweeps_restatement():
    """orc_sk_lockstep_sweeps against a line-by-line numpy version of the same schedule (accept of RRRMC.jl:39 on
    ΔE = lfields[i], update_cache! of SK.jl:252-265), and the checkenergy invariant: tracked E and incrementally updated
    fields equal a from-scratch energy() (SK.jl:212-237)."""
    import math
    N, R, nsw, seed, sweep0 = 12, 3, 4, 0x1234567890ab, (1 << 32) - 2
    J = sk_gauss(N, 5)
    g = ffi.Graph.sk_f64(J)
    beta = np.array([0.4, 1.0, 2.5])
    chunks = np.stack([random_config(N, 20 + r) for r in range(R)]).astype(np.uint64)
    lf = np.zeros((R, N)); E = np.zeros(R)
    for r in range(R):
        E[r] = g.energy(chunks[r]); lf[r] = g.lfields()
    want_s = np.stack([bits(chunks[r], N) for r in range(R)]).astype(np.int64)
    want_lf, want_E, want_acc = lf.copy(), E.copy(), np.zeros(R, np.int64)
    key = np.array([seed & 0xffffffff, seed >> 32], np.uint32)
    for r in range(R):
        for sw in range(nsw):
            t = sweep0 + sw
            for i in range(N):
                dE = want_lf[r, i]; x = -beta[r] * dE
                ok = x >= 0
                if not ok:
                    ctr = np.array([i, r, t & 0xffffffff, (t >> 32) ^ 0x534b4c53], np.uint32); o = np.zeros(4, np.uint32)
                    ffi.lib().orc_philox4x32_10(ctr, key, o)
                    u = float(((int(o[1]) << 32) | int(o[0])) >> 11) * 2.0 ** -53
                    ok = u < math.exp(x)
                if not ok:
                    continue
                want_E[r] += dE; want_acc[r] += 1
                want_s[r, i] ^= 1
                for j in range(N):
                    if j != i:
                        want_lf[r, j] = want_lf[r, j] + 4 * ((1 - 2 * (want_s[r, i] ^ want_s[r, j])) * J[i, j])
                want_lf[r, i] = -want_lf[r, i]
    acc = np.zeros(R, np.int64)
    ffi.sk_lockstep_sweeps(J, chunks, lf, E, acc, beta, seed, sweep0, nsw)
    got_s = np.stack([bits(chunks[r], N) for r in range(R)])
    assert np.array_equal(got_s, want_s) and np.array_equal(lf, want_lf) and np.array_equal(E, want_E)
    assert np.array_equal(acc, want_acc) and acc.sum() > 0
    for r in range(R):
        assert g.energy(chunks[r]) == pytest.approx(E[r], abs=1e-12)
        assert np.allclose(g.lfields(), lf[r], atol=1e-12)
