"""GPU parity (through the C ABI): Interface queries vs the oracle — bit-exact for integer couplings,
1e-6 relative (in practice exact) for Float64 couplings. Covers ragged replica counts, odd L, L=2 double bonds."""
import numpy as np
import pytest

import rrrmc_b200 as rb
from oracle import ffi
from tests.helpers import ea_instance

pytestmark = pytest.mark.gpu

CASES = [(4, 2, (-1, 1), 1), (6, 2, (-1, 1), 33), (4, 3, (-1, 1), 100), (8, 3, (-1, 1), 128), (3, 2, (-1, 1), 5),
         (2, 3, (-1, 1), 40), (5, 3, (-1, 1), 32), (4, 3, (-1, 0, 1), 7), (3, 2, (-1, 0, 1), 64), (6, 1, (-1, 1), 3)]


def _graph(L, D, lev, R, seed=0):
    A, J = ea_instance(L, D, lev, seed)
    return rb.GraphEA(L, D, lev, replicas=R, A=A, J=J), ffi.Graph.ea_int(A, J, lev)


@pytest.mark.parametrize("L,D,lev,R", CASES)
def test_energy_and_delta_energy_bit_exact(L, D, lev, R):
    X, g = _graph(L, D, lev, R)
    C = rb.Config(X.N, R, rng=np.random.default_rng(1))
    E = np.atleast_1d(rb.energy(X, C))
    want = np.array([g.energy(C.chunks[r]) for r in range(R)])
    assert np.array_equal(E, want.astype(np.int64))
    for site in sorted({1, 2, X.N // 2 + 1, X.N}):
        dE = np.atleast_1d(rb.delta_energy(X, C, site))
        want = []
        for r in range(R):
            g.energy(C.chunks[r])
            want.append(g.delta_energy(C.chunks[r], site))
        assert np.array_equal(dE, np.array(want).astype(np.int64)), site
    for r in sorted({0, R // 2, R - 1}):
        g.energy(C.chunks[r])
        want = np.array([g.delta_energy(C.chunks[r], i) for i in range(1, X.N + 1)]).astype(np.int64)
        assert np.array_equal(rb.all_delta_energy(X, C, r), want)
        assert set(np.abs(want).tolist()) <= set(rb.allΔE(X))  # Interface.jl:123-125


@pytest.mark.parametrize("L,D,lev,R", CASES)
def test_neighbors_allDE_roundtrip(L, D, lev, R):
    X, g = _graph(L, D, lev, R)
    assert rb.allΔE(X) == tuple(int(v) for v in g.allDE())
    for i in (1, X.N // 2, X.N):
        assert rb.neighbors(X, i) == tuple(g.neighbors(i).tolist())
    C = rb.Config(X.N, R, rng=np.random.default_rng(2))
    X._upload(C)
    assert X._download() == C
    # partial download of a ragged range
    if R > 3:
        part = np.zeros((2, C.chunks.shape[1]), np.uint64)
        rb._ffi.check(rb._ffi.lib().rrrmc_state_download(X._state, R - 3, 2, rb._ffi.ptr(part)))
        assert np.array_equal(part, C.chunks[R - 3:R - 1])


def test_spinflip_and_update():
    X, g = _graph(4, 3, (-1, 1), 70)
    C = rb.Config(X.N, 70, rng=np.random.default_rng(3))
    ref = C.copy()
    mask = np.zeros(70, np.uint8); mask[[0, 5, 33, 69]] = 1
    rb.spinflip(X, C, 17, mask)
    for r in np.flatnonzero(mask):
        ref.chunks[r, 0] ^= np.uint64(1 << 16)
    assert C == ref
    rb.spinflip(X, C, 64)
    ref.chunks[:, 0] ^= np.uint64(1 << 63)
    assert C == ref


def test_update_cache_does_not_flip_again():
    """update_cache!(X, C, move) runs AFTER the caller flipped the spin (Interface.jl:69-92): it refreshes the device
    copy and must leave `C` as the caller made it; ΔE of the moved site changes sign, as after spinflip!."""
    X, g = _graph(4, 3, (-1, 1), 40)
    C = rb.Config(X.N, 40, rng=np.random.default_rng(8))
    before = np.atleast_1d(rb.delta_energy(X, C, 17))
    C.chunks[:, 0] ^= np.uint64(1 << 16)          # spinflip!(C, 17) on the host copy
    flipped = C.copy()
    rb.update_cache(X, C, 17)
    assert C == flipped
    assert X._download() == flipped
    assert np.array_equal(np.atleast_1d(rb.delta_energy(X, C, 17)), -before)
    with pytest.raises(ValueError):
        rb.update_cache(X, C, X.N + 1)


def test_randomize_after_chain_run_is_not_lost():
    """C ABI: a chain sampler leaves the chain layout current; rrrmc_state_randomize then writes the multispin copy and
    must mark it current, or the next query re-transposes the stale chain copy over it."""
    from rrrmc_b200._ffi import check, lib, ptr
    X, g = _graph(4, 3, (-1, 1), 64)
    C = rb.Config(X.N, 64, rng=np.random.default_rng(9))
    X._upload(C)
    st = X._state
    betas = np.full(64, 1.0)
    check(lib().rrrmc_rrr_mc(st, ptr(betas), 50, 50, 11, rb._ffi.C.cast(None, rb._ffi.HOOK), None, None, None, 0, None))
    check(lib().rrrmc_state_randomize(st, 12345))
    got1 = X._download()
    check(lib().rrrmc_state_randomize(st, 12345))
    got2 = X._download()
    assert got1 == got2                               # the randomisation is what the state holds ...
    want = np.array([g.energy(got1.chunks[r]) for r in range(64)])
    Eo = np.zeros(64); check(lib().rrrmc_energy(st, ptr(Eo)))
    assert np.array_equal(Eo, want.astype(np.float64))  # ... and what energy() sees


@pytest.mark.parametrize("L,D,R", [(4, 2, 3), (4, 3, 40), (3, 3, 8), (2, 3, 4)])
def test_float_couplings_within_1e6(L, D, R):
    A, J = ea_instance(L, D, seed=4, gaussian=True)
    X = rb.GraphEANormal(L, D, replicas=R, A=A, J=J)
    g = ffi.Graph.ea_f64(A, J)
    C = rb.Config(X.N, R, rng=np.random.default_rng(5))
    E = np.atleast_1d(rb.energy(X, C))
    want = np.array([g.energy(C.chunks[r]) for r in range(R)])
    assert np.allclose(E, want, rtol=1e-6, atol=0)
    assert np.array_equal(E, want)  # same summation order => same bits
    g.energy(C.chunks[R - 1])
    want = np.array([g.delta_energy(C.chunks[R - 1], i) for i in range(1, X.N + 1)])
    got = rb.all_delta_energy(X, C, R - 1)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-12) and np.array_equal(got, want)
    dE = np.atleast_1d(rb.delta_energy(X, C, 3))
    want = []
    for r in range(R):
        g.energy(C.chunks[r]); want.append(g.delta_energy(C.chunks[r], 3))
    assert np.array_equal(dE, np.array(want))


def test_argument_errors_match_reference():
    A, J = ea_instance(4, 2)
    with pytest.raises(ValueError):  # EA.jl:156 "does not look like an EA graph"
        rb.GraphEA(4, 2, A=A[:, ::-1].copy(), J=J)
    with pytest.raises(ValueError):  # EA.jl:161 incompatible levels
        rb.GraphEA(4, 2, A=A, J=2 * J)
    with pytest.raises(ValueError):  # EA.jl:25
        rb.gen_EA(1, 2)
    X = rb.GraphEA(4, 2, A=A, J=J)
    with pytest.raises(ValueError):  # RRRMC.jl:94 wrong N
        rb.standardMC(X, 1.0, 100, C0=rb.Config(15), quiet=True)
    Xodd = rb.GraphEA(3, 2)
    with pytest.raises(NotImplementedError):  # checkerboard needs a two-colourable lattice
        rb.standardMC(Xodd, 1.0, 100, schedule="checkerboard", quiet=True)
