"""Host-side mirror of RRRMC.jl's graph Interface and samplers for the B200 engine.

Same names, argument meaning and error behaviour as the reference (src/Interface.jl, src/RRRMC.jl), with one
extension: a graph carries a *replica batch* of R independent chains, so energies are arrays of length R and a
`Config` holds R bit-vectors.  With replicas=1 the calls read exactly like the reference's:

    X = GraphEA(32, 2)                                   # src/graphs/EA.jl:181-191
    Es, C = standardMC(X, 1.0, 10**6, step=10**3)        # src/RRRMC.jl:81-127

All compute goes through the C ABI (include/rrrmc_b200.h) into CUDA kernels; there is no CPU path here.
"""
import ctypes as C
import math

import numpy as np

from . import _ffi
from ._ffi import check, lib, ptr

DEFAULT_SEED = 167432777111  # src/RRRMC.jl:82


# ----------------------------------------------------------------------------------------------------
class Context:
    """One device + one stream (rrrmc_ctx_t)."""
    _default = {}

    def __init__(self, device=0, stream=None):
        h = C.c_void_p()
        check(lib().rrrmc_ctx_create(device, stream, C.byref(h)))
        self.h, self.device = h, device

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = cls(device)
        return cls._default[device]

    def sync(self):
        check(lib().rrrmc_ctx_sync(self.h))

    def timer_start(self):
        check(lib().rrrmc_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().rrrmc_ctx_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_uint64()
        check(lib().rrrmc_ctx_launch_count(self.h, C.byref(n)))
        return n.value

    def flush_l2(self):
        check(lib().rrrmc_ctx_flush_l2(self.h))


# ----------------------------------------------------------------------------------------------------
class Config:
    """Config (src/Interface.jl:21-54) for a batch: `chunks[r]` is replica r's BitVector chunk array
    (site i = bit (i-1)&63 of chunk (i-1)>>6). `s` gives the bits as a (R, N) bool array."""

    def __init__(self, N, replicas=1, chunks=None, init=True, rng=None):
        self.N, self.R = int(N), int(replicas)
        nch = (self.N + 63) // 64
        if chunks is not None:
            self.chunks = np.ascontiguousarray(chunks, dtype=np.uint64).reshape(self.R, nch)
        elif init:
            rng = rng or np.random.default_rng()
            self.chunks = rng.integers(0, 2 ** 64, (self.R, nch), dtype=np.uint64)
            if self.N % 64:
                self.chunks[:, -1] &= np.uint64((1 << (self.N % 64)) - 1)
        else:
            self.chunks = np.zeros((self.R, nch), np.uint64)

    def __len__(self):
        return self.N

    @property
    def s(self):
        b = np.unpackbits(self.chunks.view(np.uint8), axis=1, bitorder="little")[:, :self.N]
        return b.astype(bool)

    @classmethod
    def from_bits(cls, bits):
        bits = np.atleast_2d(np.asarray(bits, dtype=np.uint8))
        R, N = bits.shape
        pad = (-N) % 64
        packed = np.ascontiguousarray(np.packbits(np.pad(np.ascontiguousarray(bits), ((0, 0), (0, pad))), axis=1, bitorder="little"))
        return cls(N, R, chunks=packed.view(np.uint64))

    def copy(self):
        return Config(self.N, self.R, chunks=self.chunks.copy())

    def __eq__(self, other):
        return isinstance(other, Config) and self.N == other.N and np.array_equal(self.chunks, other.chunks)


# ----------------------------------------------------------------------------------------------------
def gen_EA(L, D):
    """gen_EA (src/graphs/EA.jl:24-43): (N, 2D) int64, 1-based, rows sorted ascending."""
    if L < 2:
        raise ValueError(f"L must be ≥ 2, given: {L}")
    if D < 1:
        raise ValueError(f"D must be ≥ 0, given: {D}")
    A = np.zeros((L ** D, 2 * D), np.int64)
    check(lib().rrrmc_gen_ea_adjacency(L, D, ptr(A)))
    return A


def gen_J(f, A):
    """gen_J (src/graphs/EA.jl:45-71): one draw f() per bond x<y in (x, slot) order, mirrored into the first
    still-empty slot of J[y]. `f(n)` must return n draws in consumption order."""
    N, twoD = A.shape
    x = np.arange(1, N + 1)[:, None]
    fwd = A > x
    J = np.zeros((N, twoD), np.float64)
    draws = np.asarray(f(int(fwd.sum())), dtype=np.float64)
    J[fwd] = draws
    if (A[:, 1:] == A[:, :-1]).any():  # L=2: duplicated neighbours, follow the fill order literally
        filled = fwd.copy()
        for xi in range(N):
            for k in range(twoD):
                y = A[xi, k] - 1
                if xi < y:
                    l = int(np.flatnonzero(~filled[y])[0])
                    J[y, l] = J[xi, k]; filled[y, l] = True
        return J
    xs, ks = np.nonzero(fwd)
    ys = A[xs, ks] - 1
    ls = np.array([np.searchsorted(A[y], xx + 1) for y, xx in zip(ys, xs)]) if len(xs) < 4096 else \
        (A[ys] == (xs + 1)[:, None]).argmax(axis=1)
    J[ys, ls] = J[xs, ks]
    return J


class AbstractGraph:
    """AbstractGraph{ET} (src/Interface.jl:66): a model plus the device-resident replica batch it samples."""
    _state = None
    ET = float

    def getN(self):
        return self.N

    # -- state plumbing
    def _ensure_state(self):
        if self._state is None:
            h = C.c_void_p()
            check(lib().rrrmc_state_create(self._h, self.replicas, C.byref(h)))
            self._state = h
        return self._state

    def _upload(self, Cfg):
        if Cfg is ON_DEVICE:
            self._ensure_state()
            return
        if Cfg.N != self.N:
            raise ValueError(f"Invalid C0, wrong N, expected {self.N}, given: {Cfg.N}")  # RRRMC.jl:94
        if Cfg.R != self.replicas:
            raise ValueError(f"Config holds {Cfg.R} replicas, graph batch has {self.replicas}")
        check(lib().rrrmc_state_upload(self._ensure_state(), 0, self.replicas, ptr(Cfg.chunks)))

    def _download(self):
        out = Config(self.N, self.replicas, init=False)
        check(lib().rrrmc_state_download(self._ensure_state(), 0, self.replicas, ptr(out.chunks)))
        return out

    def __del__(self):
        try:
            if self._state is not None:
                lib().rrrmc_state_destroy(self._state)
            if getattr(self, "_h", None) is not None:
                lib().rrrmc_graph_destroy(self._h)
        except Exception:
            pass


class _OnDevice:
    """Sentinel for `C0` / `Cfg`: the configuration the device batch currently holds (no host round trip). Samplers
    called with C0=ON_DEVICE continue from it and return it as their Config without downloading; energy(X, ON_DEVICE)
    evaluates it in place. Used by the tempering drivers, which only need energies between rounds."""

    def __repr__(self):
        return "ON_DEVICE"


ON_DEVICE = _OnDevice()

MAXDIGITS = 5  # src/DFloats.jl:12: a DFloat64 is the integer round(x·10^5)


def dfloat_levels(LEV):
    """Levels with a fractional part are DFloat64 in the reference (GraphEA(L, D, LEV::Tuple{Float64,…}), EA.jl:191;
    src/DFloats.jl): fixed-point integers with five decimal digits, all arithmetic exact. Returns (integer levels,
    g) with levels = round(LEV·10^5) / g, g the gcd of the fixed-point values — so the model is an integer-level graph whose
    energies are counted in units of u = g / 10^5."""
    import math
    ints = []
    for l in LEV:
        v = round(float(l) * 10 ** MAXDIGITS)
        if abs(float(l) * 10 ** MAXDIGITS - v) > 1e-6:
            raise ValueError(f"up to {MAXDIGITS} decimal digits supported in levels, given: {LEV}")  # EA.jl:133
        ints.append(int(v))
    g = 0
    for v in ints:
        g = math.gcd(g, abs(v))
    if g == 0:
        raise ValueError(f"all levels are zero: {LEV}")
    lev = tuple(v // g for v in ints)
    if max(abs(v) for v in lev) > 127:
        raise NotImplementedError(f"levels {LEV} need integer couplings beyond int8 after reduction ({lev})")
    return lev, g


def _out(X, a):
    """Engine energies -> the graph's energy type: Int for integer levels, Float64(DFloat64) = units·u for DFloat64
    levels (the division by 10^5 of DFloats.jl:28 folded into u), Float64 as is."""
    g = getattr(X, "dfloat_g", None)
    if g is not None and getattr(X, "dfloat_mixed", False):
        return np.asarray(a, np.float64) * (g / 10 ** MAXDIGITS)   # DoubleGraph: DFloat64 levels + Float64 residuals, in units of u
    if g is not None:
        return np.rint(a) * g / 10 ** MAXDIGITS   # Float64(x::DFloat64) = d2i(x) / dfact, one rounding (DFloats.jl:28)
    return np.rint(a).astype(np.int64) if X.ET is int else a


def _beta_in(X, beta):
    """β as the engine sees it: integer-unit graphs of DFloat64 levels run at β·u (β·ΔE = (β·u)·ΔE_units; the reference
    evaluates β·(ΔE_units·u), equal up to the last ulp of the exponent — the chain law is the same)."""
    g = getattr(X, "dfloat_g", None)
    return beta if g is None else np.asarray(beta, np.float64) * (g / 10 ** MAXDIGITS)


def _scalarize(X, a):
    a = _out(X, np.asarray(a))
    return a[0] if X.replicas == 1 else a


def energy(X, Cfg):
    """energy(X, C) (Interface.jl:105): array of R energies (a scalar when replicas == 1). Resets caches."""
    X._upload(Cfg)
    E = np.zeros(X.replicas, np.float64)
    check(lib().rrrmc_energy(X._state, ptr(E)))
    return _scalarize(X, E)


def delta_energy(X, Cfg, move):
    """delta_energy(X, C, move) (Interface.jl:130): ΔE of flipping 1-based spin `move`, per replica."""
    if not 1 <= move <= X.N:
        raise ValueError(f"move out of range 1..{X.N}: {move}")
    X._upload(Cfg)
    dE = np.zeros(X.replicas, np.float64)
    check(lib().rrrmc_delta_energy(X._state, move, ptr(dE)))
    return _scalarize(X, dE)


def all_delta_energy(X, Cfg, replica=0):
    """[delta_energy(X, C, i) for i in 1:N] for one replica (what gen_ΔEcache evaluates, DeltaE.jl:79-80)."""
    X._upload(Cfg)
    dE = np.zeros(X.N, np.float64)
    check(lib().rrrmc_all_delta_energy(X._state, replica, ptr(dE)))
    return _out(X, dE)


def neighbors(X, i):
    """neighbors(X, i) (Interface.jl:158), in the reference's iteration order."""
    m = C.c_int64()
    check(lib().rrrmc_max_neighbors(X._h, C.byref(m)))
    out = np.zeros(max(64, m.value), np.int64); n = C.c_int()
    check(lib().rrrmc_neighbors(X._h, i, ptr(out), C.byref(n)))
    return tuple(int(v) for v in out[:n.value])


def allDeltaE(X):
    """allΔE(X) (Interface.jl:200-201)."""
    out = np.zeros(64, np.float64); n = C.c_int()
    check(lib().rrrmc_allDE(X._h, ptr(out), C.byref(n)))
    vals = out[:n.value]
    vals = _out(X, vals)
    return tuple(int(v) for v in vals) if X.ET is int else tuple(float(v) for v in vals)


allΔE = allDeltaE


def spinflip(X, Cfg, move, replica_mask=None):
    """spinflip!(X, C, move) (Interface.jl:89-92): flips the spin on the device batch and in `Cfg`."""
    X._upload(Cfg)
    m = None
    if replica_mask is not None:
        bits = np.zeros(((X.replicas + 31) // 32) * 32, np.uint8); bits[:X.replicas] = replica_mask
        m = np.packbits(bits, bitorder="little").view(np.uint32).copy()
    check(lib().rrrmc_spinflip(X._state, move, ptr(m)))
    Cfg.chunks[...] = X._download().chunks


def update_cache(X, Cfg, move):
    """update_cache!(X, C, move) (Interface.jl:69-86): invoked AFTER the caller flipped spin `move` in `Cfg`; it must
    not flip again. The device batch holds no cache the caller can see beside the configuration itself, so the update
    is to make the device copy equal to the (already flipped) `Cfg`; the local-field caches of the sequential samplers
    are rebuilt from it on their next use, like the reference's caches after `energy`."""
    if not (1 <= int(move) <= X.N):
        raise ValueError(f"move out of range: {move}")
    X._upload(Cfg)


class GraphEA(AbstractGraph):
    """GraphEA(L, D, LEV=(-1,1)) <: DiscrGraph (src/graphs/EA.jl:138-191) on a replica batch.
    Pass `A`, `J` (reference layout) to wrap an existing instance: GraphEA{ET,LEV,twoD}(A, J), EA.jl:145."""
    ET = int

    def __init__(self, L, D, LEV=(-1, 1), replicas=1, A=None, J=None, rng=None, ctx=None):
        if len(set(LEV)) != len(LEV):
            raise ValueError(f"repeated levels in LEV: {LEV}")  # EA.jl:122
        if not all(float(l).is_integer() for l in LEV):
            # DFloat64 levels (EA.jl:191, src/DFloats.jl): an integer-level graph in units of u; J is given/drawn in real units
            ilev, self.dfloat_g = dfloat_levels(LEV)
            self.ET, self.LEV_real = float, tuple(float(l) for l in LEV)
            if J is not None:
                J = np.rint(np.asarray(J, np.float64) * 10 ** MAXDIGITS / self.dfloat_g)
            LEV = ilev
        self.L, self.D, self.LEV, self.replicas = L, D, tuple(int(l) for l in LEV), int(replicas)
        self.ctx = ctx or Context.default()
        self.A = gen_EA(L, D) if A is None else np.ascontiguousarray(A, np.int64)
        self.N = self.A.shape[0]
        if J is None:
            rng = rng or np.random.default_rng()
            lev = np.asarray(self.LEV, np.float64)
            J = gen_J(lambda n: rng.choice(lev, n), self.A)  # rand(vLEV), EA.jl:185-187
        self.J = np.ascontiguousarray(np.rint(J), np.int64)
        if not np.isin(self.J, self.LEV).all():
            raise ValueError(f"the given J is incompatible with levels {LEV}")  # EA.jl:161
        kind = _ffi.EA_PM1 if set(self.LEV) == {-1, 1} else _ffi.EA_INT
        h = C.c_void_p()
        check(lib().rrrmc_graph_ea_create(self.ctx.h, L, D, kind, ptr(self.A), ptr(self.J), C.byref(h)))
        self._h = h


class GraphEANormal(AbstractGraph):
    """GraphEANormal(L, D) <: SimpleGraph{Float64} (src/graphs/EA.jl:534-574)."""
    ET = float

    def __init__(self, L, D, replicas=1, A=None, J=None, rng=None, ctx=None):
        self.L, self.D, self.replicas = L, D, int(replicas)
        self.ctx = ctx or Context.default()
        self.A = gen_EA(L, D) if A is None else np.ascontiguousarray(A, np.int64)
        self.N = self.A.shape[0]
        if J is None:
            rng = rng or np.random.default_rng()
            J = gen_J(lambda n: rng.standard_normal(n), self.A)
        self.J = np.ascontiguousarray(J, np.float64)
        h = C.c_void_p()
        check(lib().rrrmc_graph_ea_create(self.ctx.h, L, D, _ffi.EA_F64, ptr(self.A), ptr(self.J), C.byref(h)))
        self._h = h

    @classmethod
    def from_file(cls, fname, replicas=1, ctx=None):
        """GraphEANormal(fname::AbstractString) (src/graphs/EA.jl:576-580): a 2D instance in gen_AJ's file format."""
        L, D, A, J = gen_AJ(fname)
        return cls(L, D, replicas=replicas, A=A, J=J, ctx=ctx)


def gen_AJ(fname):
    """gen_AJ(fname) (src/graphs/EA.jl:73-118): the reference's on-disk instance format, 2D lattices only — three header
    lines `type: …`, `size: L`, `name: …`, then one `x y Jxy` line per bond (1-based sites of gen_EA(L, 2)); every bond
    must appear exactly once and be a lattice bond. Returns (L, D, A, J) with J slot-aligned with A (float64)."""
    D = 2
    with open(fname) as f:
        if not f.readline().strip().startswith("type:"):
            raise ValueError(f"{fname}: first line must start with 'type:'")
        ls = f.readline().split()
        if len(ls) != 2 or ls[0] != "size:":
            raise ValueError(f"{fname}: second line must be 'size: L'")
        L = int(ls[1])
        if not f.readline().strip().startswith("name:"):
            raise ValueError(f"{fname}: third line must start with 'name:'")
        A = gen_EA(L, D)
        N = A.shape[0]
        J = np.full((N, 2 * D), np.nan)            # NaN plays the reference's sentinel (EA.jl:88-90)
        for ln, l in enumerate(f, 4):
            ls = l.split()
            if len(ls) != 3:
                raise ValueError(f"{fname}:{ln}: expected 'x y Jxy'")
            x, y, Jxy = int(ls[0]), int(ls[1]), float(ls[2])
            if not (1 <= x <= N and 1 <= y <= N):
                raise ValueError(f"{fname}:{ln}: site out of range 1..{N}")
            for a, b in ((x, y), (y, x)):
                k = np.flatnonzero(A[a - 1] == b)  # findfirst(Ax, y), EA.jl:98-107
                if len(k) == 0:
                    raise ValueError(f"{fname}:{ln}: {a} and {b} are not neighbours on the {L}x{L} lattice")
                if not np.isnan(J[a - 1, k[0]]):
                    raise ValueError(f"{fname}:{ln}: bond {x}-{y} given twice")
                J[a - 1, k[0]] = Jxy
        if np.isnan(J).any():
            raise ValueError(f"{fname}: {int(np.isnan(J).sum()) // 2} bond(s) missing")  # EA.jl:110
    return L, D, A, J


def write_AJ(fname, L, A, J, name="instance", kind="EA2D"):
    """Writes an (A, J) 2D instance in the format gen_AJ reads (the reference ships a reader only)."""
    with open(fname, "w") as f:
        f.write(f"type: {kind}\nsize: {L}\nname: {name}\n")
        for x in range(A.shape[0]):
            for k in range(A.shape[1]):
                if A[x, k] > x + 1:
                    f.write(f"{x + 1} {A[x, k]} {float(J[x, k])!r}\n")


def gen_RRG(N, K, rng=None, max_attempts=100_000):
    """gen_RRG (src/graphs/RRG.jl:27-68): a random K-regular simple graph by the Bollobás pairing model, as an (N, K) array
    of 1-based neighbours, rows ascending. Restarts until there is neither a self-loop nor a double edge."""
    if (N * K) % 2:
        raise ValueError(f"N * K must be even, given N={N}, K={K}")
    rng = rng or np.random.default_rng()
    for _ in range(max_attempts):
        stubs = rng.permutation(np.repeat(np.arange(N), K))
        a, b = stubs[0::2], stubs[1::2]
        if (a == b).any():
            continue
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        if len(np.unique(lo * N + hi)) != len(lo):
            continue
        nbrs = [[] for _ in range(N)]
        for x, y in zip(a, b):
            nbrs[x].append(y + 1); nbrs[y].append(x + 1)
        return np.array([sorted(r) for r in nbrs], dtype=np.int64)
    raise RuntimeError("gen_RRG failed (K too large?)")


def gen_J_graph(f, A):
    """gen_J for a general adjacency (src/graphs/RRG.jl:70-96): one draw per bond (x < y) in row order, mirrored."""
    A = np.asarray(A); N, K = A.shape
    nb = int((A > np.arange(1, N + 1)[:, None]).sum())
    draws = np.asarray(f(nb), np.float64)
    J = np.zeros((N, K)); t = 0
    for x in range(N):
        for k in range(K):
            y = A[x, k] - 1
            if x < y:
                J[x, k] = draws[t]; t += 1
                J[y, int(np.flatnonzero(A[y] == x + 1)[0])] = J[x, k]
    return J


class GraphRRG(AbstractGraph):
    """GraphRRG(N, K, LEV=(-1,1)) <: DiscrGraph (src/graphs/RRG.jl:112-160), integer levels; neighbors() skips zero couplings."""
    ET = int

    def __init__(self, N, K, LEV=(-1, 1), replicas=1, A=None, J=None, rng=None, ctx=None):
        if not all(float(l).is_integer() for l in LEV):
            # DFloat64 levels (RRG.jl:162, src/DFloats.jl): the integer-level graph in units of u = gcd/10^5, as in GraphEA
            ilev, self.dfloat_g = dfloat_levels(LEV)
            self.ET, self.LEV_real = float, tuple(float(l) for l in LEV)
            if J is not None:
                J = np.rint(np.asarray(J, np.float64) * 10 ** MAXDIGITS / self.dfloat_g)
            LEV = ilev
        rng = rng or np.random.default_rng()
        self.N, self.K, self.LEV, self.replicas = int(N), int(K), tuple(int(l) for l in LEV), int(replicas)
        self.ctx = ctx or Context.default()
        self.A = gen_RRG(N, K, rng) if A is None else np.ascontiguousarray(A, np.int64)
        if J is None:
            lev = np.asarray(self.LEV, np.float64)
            J = gen_J_graph(lambda n: rng.choice(lev, n), self.A)
        self.J = np.ascontiguousarray(np.rint(J), np.int64)
        if not np.isin(self.J[self.J != 0], self.LEV).all():
            raise ValueError(f"the given J is incompatible with levels {LEV}")
        kind = _ffi.EA_PM1 if set(self.LEV) == {-1, 1} else _ffi.EA_INT
        h = C.c_void_p()
        check(lib().rrrmc_graph_rrg_create(self.ctx.h, self.N, self.K, kind, ptr(self.A), ptr(self.J), C.byref(h)))
        self._h = h


class GraphRRGNormalDiscretized(AbstractGraph):
    """GraphRRGNormalDiscretized(N, K, LEV) <: DoubleGraph{DiscrGraph{Int},Float64} (src/graphs/RRG.jl:274-330), integer levels."""
    ET = float

    def __init__(self, N, K, LEV=(-1, 0, 1), replicas=1, A=None, cJ=None, rng=None, ctx=None):
        if len(set(LEV)) != len(LEV):
            raise ValueError(f"repeated levels in LEV: {LEV}")
        unit = 1.0
        if not all(float(l).is_integer() for l in LEV):
            # DFloat64 levels (RRG.jl:330): the whole DoubleGraph in units of u = gcd/10^5 — integer levels LEV/u, couplings
            # cJ/u (the nearest level and the residual scale with it), β·u; energies come back multiplied by u
            ilev, self.dfloat_g = dfloat_levels(LEV)
            self.dfloat_mixed, self.LEV_real = True, tuple(float(l) for l in LEV)
            unit, LEV = self.dfloat_g / 10 ** MAXDIGITS, ilev
        rng = rng or np.random.default_rng()
        self.N, self.K, self.LEV, self.replicas = int(N), int(K), tuple(int(l) for l in LEV), int(replicas)
        self.ctx = ctx or Context.default()
        self.A = gen_RRG(N, K, rng) if A is None else np.ascontiguousarray(A, np.int64)
        self.cJ = np.ascontiguousarray(gen_J_graph(lambda n: rng.standard_normal(n), self.A) if cJ is None else cJ, np.float64)
        lev = np.ascontiguousarray(self.LEV, np.int64)
        cJu = np.ascontiguousarray(self.cJ / unit)
        h = C.c_void_p()
        check(lib().rrrmc_graph_rrg_discretized_create(self.ctx.h, self.N, self.K, ptr(self.A), ptr(cJu), ptr(lev), len(lev), C.byref(h)))
        self._h = h


class GraphRRGNormal(AbstractGraph):
    """GraphRRGNormal(N, K) <: SimpleGraph{Float64} (src/graphs/RRG.jl): unit-variance Gaussian couplings."""
    ET = float

    def __init__(self, N, K, replicas=1, A=None, J=None, rng=None, ctx=None):
        rng = rng or np.random.default_rng()
        self.N, self.K, self.replicas = int(N), int(K), int(replicas)
        self.ctx = ctx or Context.default()
        self.A = gen_RRG(N, K, rng) if A is None else np.ascontiguousarray(A, np.int64)
        self.J = np.ascontiguousarray(gen_J_graph(lambda n: rng.standard_normal(n), self.A) if J is None else J, np.float64)
        h = C.c_void_p()
        check(lib().rrrmc_graph_rrg_create(self.ctx.h, self.N, self.K, _ffi.EA_F64, ptr(self.A), ptr(self.J), C.byref(h)))
        self._h = h


class GraphEANormalDiscretized(AbstractGraph):
    """GraphEANormalDiscretized(L, D, LEV) <: DoubleGraph{DiscrGraph{Int},Float64} (src/graphs/EA.jl:311-360) with integer
    levels: unit-variance Gaussian couplings `cJ`, discretised to the nearest of LEV (inner GraphEA{Int,LEV}) plus Float64
    residuals. Same energy as GraphEANormal on `cJ`; rrrMC samples the levels reduced-rejection and filters the residual
    by accept(c, -βΔE1) (RRRMC.jl:221-290). Pass `A`, `cJ` to wrap an existing instance."""
    ET = float

    def __init__(self, L, D, LEV=(-1, 0, 1), replicas=1, A=None, cJ=None, rng=None, ctx=None):
        if len(set(LEV)) != len(LEV):
            raise ValueError(f"repeated levels in LEV: {LEV}")
        unit = 1.0
        if not all(float(l).is_integer() for l in LEV):
            # DFloat64 levels (EA.jl:360): the whole DoubleGraph in units of u = gcd/10^5 (see GraphRRGNormalDiscretized)
            ilev, self.dfloat_g = dfloat_levels(LEV)
            self.dfloat_mixed, self.LEV_real = True, tuple(float(l) for l in LEV)
            unit, LEV = self.dfloat_g / 10 ** MAXDIGITS, ilev
        self.L, self.D, self.LEV, self.replicas = L, D, tuple(int(l) for l in LEV), int(replicas)
        self.ctx = ctx or Context.default()
        self.A = gen_EA(L, D) if A is None else np.ascontiguousarray(A, np.int64)
        self.N = self.A.shape[0]
        if cJ is None:
            rng = rng or np.random.default_rng()
            cJ = gen_J(lambda n: rng.standard_normal(n), self.A)   # gen_J(Float64, N, A) do randn() end, EA.jl:323-325
        self.cJ = np.ascontiguousarray(cJ, np.float64)
        lev = np.ascontiguousarray(self.LEV, np.int64)
        cJu = np.ascontiguousarray(self.cJ / unit)
        h = C.c_void_p()
        check(lib().rrrmc_graph_ea_discretized_create(self.ctx.h, L, D, ptr(self.A), ptr(cJu), ptr(lev), len(lev), C.byref(h)))
        self._h = h


def gen_J_gauss(N, rng=None):
    """gen_J_gauss (src/graphs/SK.jl:170-179): symmetric N(0, 1/N) couplings, zero diagonal, as an (N, N) array."""
    rng = rng or np.random.default_rng()
    J = np.triu(rng.standard_normal((N, N)) / np.sqrt(N), 1)
    return J + J.T


def gen_J_bits(N, rng=None):
    """gen_J (src/graphs/SK.jl:17-26): symmetric random bit matrix, zero diagonal, as an (N, N) uint8 array."""
    rng = rng or np.random.default_rng()
    J = np.triu(rng.integers(0, 2, (N, N)), 1)
    return (J + J.T).astype(np.uint8)


class GraphSKNormal(AbstractGraph):
    """GraphSKNormal(N) <: SimpleGraph{Float64} (src/graphs/SK.jl:181-210): J_ij ~ N(0, 1/N)."""
    ET = float

    def __init__(self, N, replicas=1, J=None, rng=None, ctx=None):
        self.N, self.replicas = int(N), int(replicas)
        self.ctx = ctx or Context.default()
        self.J = np.ascontiguousarray(gen_J_gauss(N, rng) if J is None else J, np.float64)
        if self.J.shape != (self.N, self.N):
            raise ValueError(f"invalid J inner length, expected {self.N}")  # SK.jl:188
        h = C.c_void_p()
        check(lib().rrrmc_graph_sk_create(self.ctx.h, self.N, _ffi.SK_F64, ptr(self.J), C.byref(h)))
        self._h = h


class GraphSK(AbstractGraph):
    """GraphSK(N) <: SimpleGraph{Float64} (src/graphs/SK.jl:28-60): J_ij = ±1/√N stored as bits."""
    ET = float

    def __init__(self, N, replicas=1, J=None, rng=None, ctx=None):
        self.N, self.replicas = int(N), int(replicas)
        self.ctx = ctx or Context.default()
        self.J = np.ascontiguousarray(gen_J_bits(N, rng) if J is None else J, np.uint8)
        if self.J.shape != (self.N, self.N):
            raise ValueError(f"invalid J inner length, expected {self.N}")  # SK.jl:35
        h = C.c_void_p()
        check(lib().rrrmc_graph_sk_create(self.ctx.h, self.N, _ffi.SK_BIN, ptr(self.J), C.byref(h)))
        self._h = h


class GraphQT(AbstractGraph):
    """GraphQT{fourK}(N, M) <: DiscrGraph{Float64} (src/graphs/QT.jl:42-54): the Trotter-direction couplings."""
    ET = float

    def __init__(self, N, M, fourK, replicas=1, ctx=None):
        self.N, self.M, self.Nk, self.fourK, self.replicas = int(N), int(M), int(N) // int(M), float(fourK), int(replicas)
        self.ctx = ctx or Context.default()
        h = C.c_void_p()
        check(lib().rrrmc_graph_qt_create(self.ctx.h, self.N, self.M, self.fourK, C.byref(h)))
        self._h = h


class GraphQuant(AbstractGraph):
    """GraphQuant(Nk, M, Γ, β, inner, J) <: DoubleGraph{Float64} (src/graphs/QT.jl:126-170): M Suzuki-Trotter
    slices of a classical graph. inner: "SK" (GraphQSKT, QAliases.jl:34-43), "SKNormal" (GraphQSKNormalT, :46-47)
    or "Empty" (GraphQ0T, :19-31)."""
    ET = float

    def __init__(self, Nk, M, Γ, β, inner="SK", replicas=1, J=None, rng=None, ctx=None):
        self.Nk, self.M, self.N, self.Γ, self.β, self.replicas = int(Nk), int(M), int(Nk) * int(M), float(Γ), float(β), int(replicas)
        self.ctx = ctx or Context.default()
        self.inner = inner
        kind = {"SK": _ffi.SK_BIN, "SKNormal": _ffi.SK_F64, "Empty": _ffi.EMPTY}[inner]
        if inner == "SK":
            self.J = np.ascontiguousarray(gen_J_bits(Nk, rng) if J is None else J, np.uint8)
        elif inner == "SKNormal":
            self.J = np.ascontiguousarray(gen_J_gauss(Nk, rng) if J is None else J, np.float64)
        else:
            self.J = None
        h = C.c_void_p()
        check(lib().rrrmc_graph_quant_create(self.ctx.h, self.Nk, self.M, self.Γ, self.β, kind, ptr(self.J), C.byref(h)))
        self._h = h
        fk = C.c_double()
        check(lib().rrrmc_graph_fourK(self._h, C.byref(fk)))
        self.fourK = fk.value

    def set_betas(self, betas):
        """Turns the batch into a β ladder: replica r at betas[r] with its own fourK(β) (QT.jl:165; in the reference each
        β is a separate GraphQuant{fourK,G}). None restores the graph's β. Returns the per-replica fourK."""
        st = self._ensure_state()
        if betas is None:
            check(lib().rrrmc_state_set_quant_betas(st, None, None))
            self.betas = None
            return None
        b = np.ascontiguousarray(np.broadcast_to(np.asarray(betas, np.float64), (self.replicas,)))
        fk = np.zeros(self.replicas, np.float64)
        check(lib().rrrmc_state_set_quant_betas(st, ptr(b), ptr(fk)))
        self.betas = b.copy()
        return fk

    def inner_graph(self):
        """inner_graph(X) (Interface.jl:239-240; QT.jl:148): the GraphQT part, on its own replica batch."""
        return GraphQT(self.N, self.M, self.fourK, replicas=self.replicas, ctx=self.ctx)


class GraphQEAT(GraphQuant):
    """GraphQEAT(L, D, M, Γ, β) (src/QAliases.jl:51-81): GraphQuant over GraphEANormal{2D} — the transverse-field
    Edwards-Anderson model. Couplings uniform in [-2, 2) (4*rand() - 2, QAliases.jl:62-64) unless `A`, `J` are given."""

    def __init__(self, L, D, M, Γ, β, replicas=1, A=None, J=None, rng=None, ctx=None):
        self.L, self.D = int(L), int(D)
        self.A = gen_EA(L, D) if A is None else np.ascontiguousarray(A, np.int64)
        self.Nk, self.M = self.A.shape[0], int(M)
        self.N, self.Γ, self.β, self.replicas = self.Nk * self.M, float(Γ), float(β), int(replicas)
        self.ctx = ctx or Context.default()
        self.inner = "EANormal"
        if J is None:
            rng = rng or np.random.default_rng()
            J = gen_J(lambda n: 4 * rng.random(n) - 2, self.A)
        self.J = np.ascontiguousarray(J, np.float64)
        h = C.c_void_p()
        check(lib().rrrmc_graph_quant_ea_create(self.ctx.h, self.L, self.D, self.M, self.Γ, self.β, ptr(self.A), ptr(self.J), C.byref(h)))
        self._h = h
        fk = C.c_double()
        check(lib().rrrmc_graph_fourK(self._h, C.byref(fk)))
        self.fourK = fk.value


def GraphQSKT(N, M, Γ, β, **kw):
    """GraphQSKT(N, M, Γ, β) (src/QAliases.jl:34-43)."""
    return GraphQuant(N, M, Γ, β, "SK", **kw)


def GraphQSKNormalT(N, M, Γ, β, **kw):
    """GraphQSKNormalT(N, M, Γ, β) (src/QAliases.jl:46-47)."""
    return GraphQuant(N, M, Γ, β, "SKNormal", **kw)


def GraphQ0T(N, M, Γ, β, **kw):
    """GraphQ0T(N, M, Γ, β) (src/QAliases.jl:19-31)."""
    return GraphQuant(N, M, Γ, β, "Empty", **kw)


def delta_energy_residual(X, Cfg, move):
    """delta_energy_residual(X, C, move) (Interface.jl:254-261; QT.jl:270-281)."""
    X._upload(Cfg)
    dE = np.zeros(X.replicas, np.float64)
    check(lib().rrrmc_delta_energy_residual(X._state, move, ptr(dE)))
    return _scalarize(X, dE)


def _observable(fn, X, Cfg, n, *args):
    if Cfg is not None:
        X._upload(Cfg)
    out = np.zeros((X.replicas, n) if n > 1 else X.replicas, np.float64)
    check(fn(X._ensure_state(), *args, ptr(out)))
    return out[0] if X.replicas == 1 else out


def transverse_mag(X, Cfg, β):
    """transverse_mag(X, C, β) (QT.jl:113-121)."""
    return _observable(lib().rrrmc_transverse_mag, X, Cfg, 1, float(β))


def Qenergy(X, Cfg):
    """Qenergy(X, C) (QT.jl:253-268)."""
    return _observable(lib().rrrmc_Qenergy, X, Cfg, 1)


def Renergies(X, Cfg=None):
    """Renergies(X) (QT.jl:201-211): classical energy of every Trotter slice."""
    return _observable(lib().rrrmc_Renergies, X, Cfg, X.M)


def overlaps(X, Cfg=None):
    """overlaps(X) (QT.jl:213-251): mean overlap between slices at Trotter distance 1..M÷2."""
    return _observable(lib().rrrmc_overlaps, X, Cfg, X.M // 2)


def sk_fields_init(X, Cfg=None, tensor_cores=True):
    """Local fields of the whole batch, lfields[r][i] = 2σ_ri Σ_j J_ij σ_rj — the contraction inside energy(X, C) of
    SK.jl:212-237 — on the tensor cores (exact INT8 digit-plane GEMMs) or on CUDA cores in the reference's
    summation order. -> (lfields (R, N), E (R,), device_ms)"""
    if Cfg is not None:
        X._upload(Cfg)
    E = np.zeros(X.replicas, np.float64); ms = C.c_float()
    check(lib().rrrmc_sk_fields_init(X._ensure_state(), int(bool(tensor_cores)), ptr(E), C.byref(ms)))
    lf = np.zeros((X.replicas, X.N), np.float64)
    check(lib().rrrmc_sk_get_fields(X._state, ptr(lf)))
    return lf, E, ms.value


def sk_metropolis_sweeps(X, β, nsweeps, *, seed=DEFAULT_SEED, sweep0=0, C0=None):
    """Lock-step Metropolis sweeps on a GraphSKNormal batch (sites 1..N in order, all replicas together).
    -> (E (R,), accepted (R,), C)"""
    st = X._ensure_state()
    if C0 is not None:
        X._upload(C0)
    betas = np.ascontiguousarray(np.broadcast_to(np.asarray(β, np.float64), (X.replicas,)))
    E = np.zeros(X.replicas, np.float64); acc = np.zeros(X.replicas, np.int64)
    check(lib().rrrmc_sk_metropolis_sweeps(st, ptr(betas), int(seed), int(sweep0), int(nsweeps), ptr(E), ptr(acc)))
    return E, acc, X._download()


def checkerboard_sweeps_normal(X, β, nsweeps, *, seed=DEFAULT_SEED, sweep0=0, C0=None):
    """Checkerboard Metropolis sweeps on a GraphEANormal batch (continuous couplings, per-replica β): the kernel of
    csrc/ea_normal.cu behind rrrmc_checkerboard_sweeps_f64. -> C"""
    st = X._ensure_state()
    if C0 is not None:
        X._upload(C0)
    betas = np.ascontiguousarray(np.broadcast_to(np.asarray(β, np.float64), (X.replicas,)))
    check(lib().rrrmc_checkerboard_sweeps_f64(st, ptr(betas), int(seed), int(sweep0), int(nsweeps)))
    return X._download()


# ----------------------------------------------------------------------------------------------------
class _LazyConfig:
    """The `C` a hook sees: downloads the batch from the device on first access."""

    def __init__(self, X):
        self._X, self._c = X, None

    def _get(self):
        if self._c is None:
            self._c = self._X._download()
        return self._c

    def __getattr__(self, k):
        return getattr(self._get(), k)


def _run(fn, X, beta, iters, seed, step, hook, C0, quiet, opts, name):
    if step < 1:
        raise ValueError("step must be ≥ 1")
    st = X._ensure_state()
    if C0 is None:
        check(lib().rrrmc_state_randomize(st, seed if seed > 0 else np.random.SeedSequence().entropy & (2 ** 63 - 1)))
    else:
        X._upload(C0)
    R = X.replicas
    betas = np.ascontiguousarray(np.broadcast_to(np.asarray(_beta_in(X, beta), np.float64), (R,)))
    cap = min(10 ** 8, iters // step)  # RRRMC.jl:90
    Es = np.zeros((max(cap, 1), R), np.float64)
    info = _ffi.RunInfo()
    last = {}

    def _hook(user, it, E, acc, n):
        Ev = np.ctypeslib.as_array(E, (n,)).copy()
        av = np.ctypeslib.as_array(acc, (n,)).copy()
        last["acc"], last["it"] = av, it
        try:
            ok = hook(it, X, _LazyConfig(X), av if R > 1 else int(av[0]), _scalarize(X, Ev))
        except Exception as e:  # propagate after the C call returns
            last["exc"] = e
            return 0
        return 1 if ok else 0
    cb = _ffi.HOOK(_hook) if hook is not None else C.cast(None, _ffi.HOOK)
    check(fn(st, ptr(betas), int(iters), int(step), int(seed) if seed > 0 else 0, cb, None, C.byref(opts),
             ptr(Es), cap, C.byref(info)))
    if "exc" in last:
        raise last["exc"]
    Cout = ON_DEVICE if C0 is ON_DEVICE else X._download()
    Es = _out(X, Es[:info.nsamples])
    if not quiet:
        print("samples =", info.nsamples)
        print("iters =", info.iters_done)
        if "acc" in last and (last["acc"] >= 0).all():
            print("accept rate =", float(np.mean(last["acc"])) / max(1, last["it"]))
    X.last_run = info
    return (Es[:, 0] if R == 1 else Es), Cout


def _opts(schedule=None, planes_K=None, count_accepted=None, staged_thr=None, staged_thr_fact=None, planes_M=None,
          cb_method=None, site_pick=None):
    o = _ffi.Opts()
    check(lib().rrrmc_opts_default(C.byref(o)))
    if schedule is not None:
        o.schedule = {"checkerboard": _ffi.SCHED_CHECKERBOARD, "random": _ffi.SCHED_RANDOM_SITE}[schedule]
    if planes_K is not None:
        o.planes_K = planes_K
    if planes_M is not None:
        o.planes_M = planes_M
    if cb_method is not None:
        o.cb_method = {"auto": _ffi.CB_AUTO, "planes": _ffi.CB_PLANES, "sparse": _ffi.CB_SPARSE,
                       "poisson": _ffi.CB_POISSON}[cb_method]
    if count_accepted is not None:
        o.count_accepted = int(count_accepted)
    if staged_thr is not None:
        o.staged_thr = staged_thr
    if staged_thr_fact is not None:
        o.staged_thr_fact = staged_thr_fact
    if site_pick is not None:
        o.site_pick = {"reference": _ffi.PICK_REFERENCE, "rank": _ffi.PICK_RANK}[site_pick]
    return o


def standardMC(X, β, iters, *, seed=DEFAULT_SEED, step=1, hook=None, C0=None, quiet=False,
               schedule=None, planes_K=None, planes_M=None, count_accepted=None, cb_method=None):
    """standardMC(X, β, iters; seed, step, hook, C0, quiet) (src/RRRMC.jl:81-127) -> (Es, C).

    schedule="random" (default) is the reference's order and sampling contract: i = rand(1:N) per attempt, `Es` has
    iters÷step rows, sample `it` is taken before the move of iteration `it` (RRRMC.jl:100-119).
    schedule="checkerboard" (opt-in; implied by giving cb_method / planes_K / planes_M) updates all replicas of a
    two-colourable ±J lattice in lock step, a whole sweep (N attempts) at a time — `iters`/`step` are rounded up to
    whole sweeps and samples are post-sweep energies. cb_method selects how a task turns Philox bits into accept()
    decisions: "planes", "sparse", "poisson" or "auto" (include/rrrmc_b200.h)."""
    if schedule is None:
        schedule = "checkerboard" if (cb_method is not None or planes_K is not None or planes_M is not None) else "random"
    return _run(lib().rrrmc_standard_mc, X, β, iters, seed, step, hook, C0, quiet,
                _opts(schedule, planes_K, count_accepted, planes_M=planes_M, cb_method=cb_method), "standardMC")


def rrrMC(X, β, iters, *, seed=DEFAULT_SEED, step=1, hook=None, C0=None, staged_thr=float("nan"),
          staged_thr_fact=5.0, quiet=False, site_pick=None):
    """rrrMC(X, β, iters; ...) (src/RRRMC.jl:149-219). site_pick="rank" runs the warp-cooperative kernel (±J GraphEA
    lattices; the same chain law, the member of a ΔE class is picked by rank in site order instead of ArraySet order)."""
    if not np.all(np.isfinite(β)):
        raise ValueError(f"β must be finite, given: {β}")  # RRRMC.jl:159
    return _run(lib().rrrmc_rrr_mc, X, β, iters, seed, step, hook, C0, quiet,
                _opts(staged_thr=staged_thr, staged_thr_fact=staged_thr_fact, site_pick=site_pick), "rrrMC")


def bklMC(X, β, iters, *, seed=DEFAULT_SEED, step=1, hook=None, C0=None, quiet=False, site_pick=None):
    """bklMC(X, β, iters; ...) (src/RRRMC.jl:311-359). site_pick: see rrrMC."""
    return _run(lib().rrrmc_bkl_mc, X, β, iters, seed, step, hook, C0, quiet, _opts(site_pick=site_pick), "bklMC")


def wtmMC(X, β, samples, *, seed=DEFAULT_SEED, step=1.0, hook=None, C0=None, quiet=False):
    """wtmMC(X, β, samples; seed, step::Float64, hook, C0, quiet) (src/RRRMC.jl:376-430): the rejection-free waiting-time
    method. `step` is measured in the sampler's global time (scaled by N inside); the hook receives the global time
    k·step/N of sample k as its first argument, like the reference's."""
    if not step > 0:
        raise ValueError("step must be > 0")
    st = X._ensure_state()
    if C0 is None:
        check(lib().rrrmc_state_randomize(st, seed if seed > 0 else np.random.SeedSequence().entropy & (2 ** 63 - 1)))
    else:
        X._upload(C0)
    R = X.replicas
    betas = np.ascontiguousarray(np.broadcast_to(np.asarray(_beta_in(X, β), np.float64), (R,)))
    cap = min(10 ** 8, int(samples))
    Es = np.zeros((max(cap, 1), R), np.float64)
    info = _ffi.RunInfo()
    last = {}

    def _hook(user, k, E, acc, n):
        Ev = np.ctypeslib.as_array(E, (n,)).copy()
        av = np.ctypeslib.as_array(acc, (n,)).copy()
        try:
            ok = hook(k * (float(step) / X.N), X, _LazyConfig(X), av if R > 1 else int(av[0]), _scalarize(X, Ev))
        except Exception as e:  # propagate after the C call returns
            last["exc"] = e
            return 0
        return 1 if ok else 0
    cb = _ffi.HOOK(_hook) if hook is not None else C.cast(None, _ffi.HOOK)
    check(lib().rrrmc_wtm_mc(st, ptr(betas), int(samples), float(step), int(seed) if seed > 0 else 0, cb, None,
                             ptr(Es), cap, C.byref(info)))
    if "exc" in last:
        raise last["exc"]
    Cout = ON_DEVICE if C0 is ON_DEVICE else X._download()
    Es = _out(X, Es[:info.nsamples])
    if not quiet:
        print("samples =", info.nsamples)
        print("num_moves =", info.iters_done)
    X.last_run = info
    return (Es[:, 0] if R == 1 else Es), Cout


def eo_ftau(N, τ):
    """fτ = cumsum([j^(-τ) for j = 1:N]) (src/DeltaE.jl:443), summed the way Julia's `cumsum` sums a Float64 vector:
    Base.accumulate_pairwise! (blocks of < 128 elements accumulated left to right, block sums combined pairwise).
    The powers come from this host's libm `pow`; a Julia host passes its own table (julia/RRRMCB200.jl)."""
    v = np.arange(1, int(N) + 1, dtype=np.float64) ** (-float(τ))
    out = np.empty_like(v)
    if len(v) == 0:
        return out
    out[0] = v[0]

    def rec(s, i1, n):
        if n < 128:
            s_ = v[i1]
            out[i1] = s + s_
            for i in range(i1 + 1, i1 + n):
                s_ = s_ + v[i]
                out[i] = s + s_
            return s_
        n2 = n >> 1
        s1 = rec(s, i1, n2)
        s2 = rec(s + s1, i1 + n2, n - n2)
        return s1 + s2
    if len(v) > 1:
        rec(v[0], 1, len(v) - 1)
    return out


def extremal_opt(X, τ, iters, *, seed=DEFAULT_SEED, step=1, hook=None, C0=None, quiet=False, ftau=None, return_Es=False):
    """extremal_opt(X, τ, iters; seed, step, hook, C0, quiet) (src/RRRMC.jl:468-521) -> (C, Emin, Cmin, itmin), batched
    over the replicas (Emin, itmin are arrays when the batch holds more than one chain; τ may be one value per chain).
    hook(it, X, C, E, Emin)::Bool. DiscrGraph models (EOCache) and the Float64 SimpleGraphs (EOCacheCont). `ftau` overrides the table built by `eo_ftau`;
    `return_Es=True` appends the energies at the hook instants (a test aid, not in the reference)."""
    if step < 1:
        raise ValueError("step must be ≥ 1")
    st = X._ensure_state()
    if C0 is None:
        check(lib().rrrmc_state_randomize(st, seed if seed > 0 else np.random.SeedSequence().entropy & (2 ** 63 - 1)))
    else:
        X._upload(C0)
    R = X.replicas
    if ftau is None:
        taus = np.atleast_1d(np.asarray(τ, np.float64))
        ftau = eo_ftau(X.N, taus[0]) if len(taus) == 1 else np.stack([eo_ftau(X.N, t) for t in taus])
    ftau = np.ascontiguousarray(ftau, np.float64)
    if ftau.shape not in ((X.N,), (R, X.N)):
        raise ValueError(f"ftau must have shape ({X.N},) or ({R}, {X.N})")
    stride = 0 if ftau.ndim == 1 else X.N
    cap = min(10 ** 8, int(iters) // int(step))
    Es = np.zeros((max(cap, 1), R), np.float64)
    Emin = np.zeros(R, np.float64); itmin = np.zeros(R, np.int64)
    Cmin = Config(X.N, R, init=False)
    info = _ffi.RunInfo()
    last = {}

    def _hook(user, it, E, Em, n):
        Ev = np.ctypeslib.as_array(E, (n,)).copy()
        Mv = np.ctypeslib.as_array(Em, (n,)).copy()
        try:
            ok = hook(it, X, _LazyConfig(X), _scalarize(X, Ev), _scalarize(X, Mv))
        except Exception as e:  # propagate after the C call returns
            last["exc"] = e
            return 0
        return 1 if ok else 0
    cb = _ffi.EOHOOK(_hook) if hook is not None else C.cast(None, _ffi.EOHOOK)
    check(lib().rrrmc_extremal_opt(st, ptr(ftau), stride, int(iters), int(step), int(seed) if seed > 0 else 0, cb, None,
                                   ptr(Emin), ptr(itmin), ptr(Cmin.chunks), ptr(Es), cap, C.byref(info)))
    if "exc" in last:
        raise last["exc"]
    Cout = X._download()
    X.last_run = info
    if not quiet:
        print("iters =", info.iters_done)
        print(f"min [it = {itmin if R > 1 else int(itmin[0])}] = {_scalarize(X, Emin)}")
    out = (Cout, _scalarize(X, Emin), Cmin, itmin if R > 1 else int(itmin[0]))
    if return_Es:
        Es = _out(X, Es[:info.nsamples])
        out = out + ((Es[:, 0] if R == 1 else Es),)
    return out


def replay(X, C0, sampler, β, iters, kind, ival, fval, *, step=1, replica=0, staged_thr=float("nan"), staged_thr_fact=5.0):
    """Feed chain `replica` the typed draw stream the reference consumed -> (Es, C)."""
    X._upload(C0)
    kind = np.ascontiguousarray(kind, np.uint8); ival = np.ascontiguousarray(ival, np.int64); fval = np.ascontiguousarray(fval, np.float64)
    cap = iters // step
    Es = np.zeros(max(cap, 1), np.float64)
    info = _ffi.RunInfo()
    o = _opts(staged_thr=staged_thr, staged_thr_fact=staged_thr_fact)
    code = {"standardMC": 0, "rrrMC": 1, "bklMC": 2}[sampler]
    check(lib().rrrmc_replay(X._state, replica, code, float(_beta_in(X, β)), int(iters), int(step), ptr(kind), ptr(ival), ptr(fval),
                             len(kind), C.byref(o), ptr(Es), cap, C.byref(info)))
    Es = Es[:info.nsamples]
    X.last_run = info
    return _out(X, Es), X._download()


def replay_wtm(X, C0, β, samples, kind, ival, fval, *, step=1.0, replica=0):
    """wtmMC (RRRMC.jl:376-430) of chain `replica` fed a dumped draw stream -> (Es, C)."""
    X._upload(C0)
    kind = np.ascontiguousarray(kind, np.uint8); ival = np.ascontiguousarray(ival, np.int64); fval = np.ascontiguousarray(fval, np.float64)
    cap = int(samples)
    Es = np.zeros(max(cap, 1), np.float64)
    info = _ffi.RunInfo()
    check(lib().rrrmc_replay_wtm(X._state, replica, float(_beta_in(X, β)), int(samples), float(step), ptr(kind), ptr(ival), ptr(fval),
                                 len(kind), ptr(Es), cap, C.byref(info)))
    X.last_run = info
    return Es[:info.nsamples], X._download()


def replay_extremal_opt(X, C0, τ, iters, kind, ival, fval, *, step=1, replica=0):
    """extremal_opt (RRRMC.jl:468-521) of chain `replica` fed a dumped draw stream -> (Es, C, Emin, Cmin chunks, itmin)."""
    X._upload(C0)
    kind = np.ascontiguousarray(kind, np.uint8); ival = np.ascontiguousarray(ival, np.int64); fval = np.ascontiguousarray(fval, np.float64)
    ftau = np.cumsum(np.arange(1, X.N + 1, dtype=np.float64) ** (-float(τ)))      # fτ = cumsum(j^-τ), RRRMC.jl:483
    cap = int(iters) // int(step)
    Es = np.zeros(max(cap, 1), np.float64)
    nch = (X.N + 63) // 64
    emin = np.zeros(1, np.float64); itmin = np.zeros(1, np.int64); cmin = np.zeros(nch, np.uint64)
    info = _ffi.RunInfo()
    check(lib().rrrmc_replay_extremal_opt(X._state, replica, ptr(ftau), int(iters), int(step), ptr(kind), ptr(ival), ptr(fval), len(kind),
                                          ptr(emin), ptr(itmin), ptr(cmin), ptr(Es), cap, C.byref(info)))
    X.last_run = info
    return Es[:info.nsamples], X._download(), float(emin[0]), cmin, int(itmin[0])
