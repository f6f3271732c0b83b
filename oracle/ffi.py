"""ctypes binding of the CPU oracle (oracle/rrrmc_oracle.c).

TEST INFRASTRUCTURE ONLY: import from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librrrmc_oracle.so")

EA_INT, EA_F64, SK_BIN, SK_F64, QT, QUANT, EMPTY, EA_DISCR = 1, 2, 3, 4, 5, 6, 7, 8


def build(force=False):
    src = os.path.join(_HERE, "rrrmc_oracle.c")
    hdr = os.path.join(_HERE, "rrrmc_oracle.h")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return _SO


class Draws(C.Structure):
    _fields_ = [("f64", C.c_void_p), ("range", C.c_void_p), ("user", C.c_void_p)]


class PhiloxSrc(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("chain", C.c_uint64), ("n", C.c_uint64), ("tag", C.c_uint32)]


class XoshiroSrc(C.Structure):
    _fields_ = [("s", C.c_uint64 * 4)]


class Trace(C.Structure):
    pass


Trace._fields_ = [("len", C.c_int64), ("cap", C.c_int64), ("pos", C.c_int64),
                  ("kind", C.POINTER(C.c_uint8)), ("ival", C.POINTER(C.c_int64)), ("fval", C.POINTER(C.c_double)),
                  ("inner", Draws), ("error", C.c_int)]


class Result(C.Structure):
    _fields_ = [("nsamples", C.c_int64), ("iters_done", C.c_int64), ("accepted", C.c_int64),
                ("staged_its", C.c_int64), ("status", C.c_int)]


HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_int64)


class EOResult(C.Structure):
    _fields_ = [("nsamples", C.c_int64), ("iters_done", C.c_int64), ("itmin", C.c_int64),
                ("Emin", C.c_double), ("status", C.c_int), ("Efinal", C.c_double)]


EOHOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_double)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, i64, f64, i32 = C.c_void_p, C.c_int64, C.c_double, C.c_int
    p = np.ctypeslib.ndpointer
    sig = {
        "orc_philox4x32_10": (None, [p(np.uint32), p(np.uint32), p(np.uint32)]),
        "orc_philox_f64": (f64, [vp]), "orc_philox_range": (i64, [vp, i64]),
        "orc_philox_u64": (C.c_uint64, [C.POINTER(PhiloxSrc)]),
        "orc_xoshiro_seed": (None, [C.POINTER(XoshiroSrc), C.c_uint64]),
        "orc_xoshiro_f64": (f64, [vp]), "orc_xoshiro_range": (i64, [vp, i64]),
        "orc_trace_new": (C.POINTER(Trace), []), "orc_trace_free": (None, [C.POINTER(Trace)]),
        "orc_trace_load": (None, [C.POINTER(Trace), i64, p(np.uint8), p(np.int64), p(np.float64)]),
        "orc_gen_EA": (i64, [i64, i32, p(np.int64)]),
        "orc_gen_J_f64": (i32, [i64, i32, p(np.int64), p(np.float64), i64, p(np.float64)]),
        "orc_ea_int_create": (vp, [i64, i32, p(np.int64), p(np.int64), p(np.int64), i32]),
        "orc_ea_f64_create": (vp, [i64, i32, p(np.int64), p(np.float64)]),
        "orc_ea_discretized_create": (vp, [i64, i32, p(np.int64), p(np.float64), p(np.int64), i32]),
        "orc_rrg_int_create": (vp, [i64, i32, p(np.int64), p(np.int64), p(np.int64), i32]),
        "orc_rrg_discretized_create": (vp, [i64, i32, p(np.int64), p(np.float64), p(np.int64), i32]),
        "orc_sk_f64_create": (vp, [i64, p(np.float64)]),
        "orc_sk_bin_create": (vp, [i64, p(np.uint8)]),
        "orc_qt_create": (vp, [i64, i64, f64]),
        "orc_empty_create": (vp, [i64]),
        "orc_quant_create": (vp, [i64, i64, f64, f64, i32, vp, i32, vp]),
        "orc_graph_free": (None, [vp]),
        "orc_kind": (i32, [vp]), "orc_getN": (i64, [vp]),
        "orc_energy": (f64, [vp, p(np.uint64)]),
        "orc_delta_energy": (f64, [vp, p(np.uint64), i64]),
        "orc_delta_energy_residual": (f64, [vp, p(np.uint64), i64]),
        "orc_spinflip": (None, [vp, p(np.uint64), i64]),
        "orc_neighbors": (i32, [vp, i64, p(np.int64)]),
        "orc_allDE": (i32, [vp, p(np.float64)]),
        "orc_get_lfields": (i64, [vp, p(np.float64)]),
        "orc_quant_fourK": (f64, [vp]), "orc_inner_graph": (vp, [vp]),
        "orc_transverse_mag": (f64, [vp, p(np.uint64), f64]),
        "orc_Qenergy": (f64, [vp, p(np.uint64)]),
        "orc_Renergies": (None, [vp, p(np.float64)]), "orc_overlaps": (None, [vp, p(np.float64)]),
        "orc_standardMC": (Result, [vp, f64, i64, i64, p(np.uint64), Draws, HOOK, vp, vp, i64]),
        "orc_rrrMC": (Result, [vp, f64, i64, i64, p(np.uint64), Draws, f64, f64, HOOK, vp, vp, i64]),
        "orc_bklMC": (Result, [vp, f64, i64, i64, p(np.uint64), Draws, HOOK, vp, vp, i64]),
        "orc_wtmMC": (Result, [vp, f64, i64, f64, p(np.uint64), Draws, HOOK, vp, vp, i64]),
        "orc_rank_rrrMC": (Result, [vp, f64, i64, i64, p(np.uint64), Draws, HOOK, vp, vp, i64]),
        "orc_rank_bklMC": (Result, [vp, f64, i64, i64, p(np.uint64), Draws, HOOK, vp, vp, i64]),
        "orc_extremal_opt": (EOResult, [vp, p(np.float64), i64, i64, p(np.uint64), vp, Draws, EOHOOK, vp, vp, i64]),
        "orc_check_discrete_cache": (i32, [vp, p(np.uint64), f64, p(np.int64), i64]),
        "orc_ds_probe": (i32, [i64, p(np.float64), i64, p(np.int64), p(np.float64), p(np.float64), p(np.float64), i64, p(np.float64), p(np.int64)]),
        "orc_arrayset_probe": (i64, [i64, i64, p(np.int64), p(np.int64)]),
        "orc_checkerboard_sweeps": (None, [i32, i32, i64, p(np.uint32), p(np.int8), p(np.uint64), i32, i32,
                                           C.c_uint64, C.c_uint64, i64, vp]),
        "orc_checkerboard_sweeps_sparse": (None, [i32, i32, i64, p(np.uint32), p(np.int8), p(np.uint32),
                                                  C.c_uint64, C.c_uint64, i64, vp]),
        "orc_cb_sparse_tables": (None, [p(np.uint64), i32, p(np.uint32)]),
        "orc_checkerboard_sweeps_poisson": (None, [i32, i32, i64, p(np.uint32), p(np.int8), p(np.uint32), i32,
                                                   C.c_uint64, C.c_uint64, i64, vp]),
        "orc_checkerboard_sweeps_poisson_ladder": (None, [i32, i32, i64, p(np.uint32), p(np.int8), p(np.uint32), i32,
                                                          C.c_uint64, C.c_uint64, i64, vp]),
        "orc_cb_poisson_tables": (None, [p(np.uint64), i32, p(np.uint32)]),
        "orc_dfloat_from_f64": (i64, [f64]),
        "orc_dfloat_to_f64": (f64, [i64]),
        "orc_dfloat_div_int": (i64, [i64, i64]),
        "orc_dfloat_ea_energy": (f64, [i64, i32, p(np.int64), p(np.float64), p(np.uint64), vp]),
        "orc_sk_lockstep_sweeps": (None, [i32, i64, p(np.float64), p(np.uint64), i64, p(np.float64), p(np.float64), p(np.int64),
                                          p(np.float64), C.c_uint64, C.c_uint64, i64]),
        "orc_tempering_decide": (None, [i64, p(np.float64), p(np.float64), C.c_uint64, C.c_uint64, p(np.uint8)]),
        "orc_checkerboard_sweeps_f64": (None, [i32, i32, i64, p(np.uint32), p(np.int64), p(np.float64), p(np.float64),
                                               C.c_uint64, C.c_uint64, i64, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


def _fnptr(name):
    return C.cast(getattr(lib(), name), C.c_void_p).value


# ------------------------------------------------------------------ draw sources
class PhiloxDraws:
    """Philox4x32-10 chain source: key=seed, counter=(n, chain, tag) — same definition as the CUDA chain kernels."""

    def __init__(self, seed, chain=0, tag=0):
        self.src = PhiloxSrc(seed, chain, 0, tag)
        self.draws = Draws(_fnptr("orc_philox_f64"), _fnptr("orc_philox_range"), C.addressof(self.src))

    def u64(self):
        return lib().orc_philox_u64(C.byref(self.src))

    def config(self, N):
        """Config(N): ⌈N/64⌉ raw 64-bit words, unused high bits zero (Interface.jl:24-28)."""
        nch = (N + 63) // 64
        ch = np.array([self.u64() for _ in range(nch)], dtype=np.uint64)
        if N % 64:
            ch[-1] &= np.uint64((1 << (N % 64)) - 1)
        return ch


class XoshiroDraws:
    """xoshiro256++ source (CPU-baseline timing only)."""

    def __init__(self, seed):
        self.src = XoshiroSrc()
        lib().orc_xoshiro_seed(C.byref(self.src), seed)
        self.draws = Draws(_fnptr("orc_xoshiro_f64"), _fnptr("orc_xoshiro_range"), C.addressof(self.src))

    def config(self, N):
        rng = np.random.default_rng(int(self.src.s[0]) & 0xffffffff)
        nch = (N + 63) // 64
        ch = rng.integers(0, 2 ** 64, nch, dtype=np.uint64)
        if N % 64:
            ch[-1] &= np.uint64((1 << (N % 64)) - 1)
        return ch


class Recorder:
    """Wraps a source and records every typed draw (SURVEY Appendix B trace)."""

    def __init__(self, inner):
        self.inner = inner
        self.t = lib().orc_trace_new()
        self.t.contents.inner = inner.draws
        self.draws = Draws(_fnptr("orc_trace_rec_f64"), _fnptr("orc_trace_rec_range"), C.cast(self.t, C.c_void_p).value)

    def arrays(self):
        t = self.t.contents
        n = t.len
        kind = np.ctypeslib.as_array(t.kind, (n,)).copy() if n else np.zeros(0, np.uint8)
        ival = np.ctypeslib.as_array(t.ival, (n,)).copy() if n else np.zeros(0, np.int64)
        fval = np.ctypeslib.as_array(t.fval, (n,)).copy() if n else np.zeros(0, np.float64)
        return kind, ival, fval

    def __del__(self):
        try:
            lib().orc_trace_free(self.t)
        except Exception:
            pass


class Replayer:
    def __init__(self, kind, ival, fval):
        self.t = lib().orc_trace_new()
        lib().orc_trace_load(self.t, len(kind), np.ascontiguousarray(kind, np.uint8),
                             np.ascontiguousarray(ival, np.int64), np.ascontiguousarray(fval, np.float64))
        self.draws = Draws(_fnptr("orc_trace_play_f64"), _fnptr("orc_trace_play_range"), C.cast(self.t, C.c_void_p).value)

    @property
    def error(self):
        return self.t.contents.error

    @property
    def consumed(self):
        return self.t.contents.pos

    def __del__(self):
        try:
            lib().orc_trace_free(self.t)
        except Exception:
            pass


# ------------------------------------------------------------------ graphs
def gen_EA(L, D):
    N = L ** D
    A = np.zeros((N, 2 * D), dtype=np.int64)
    assert lib().orc_gen_EA(L, D, A) == N
    return A


def gen_J(A, draws):
    """gen_J(f, ET, N, A) with f() values supplied in consumption order (EA.jl:45-71)."""
    N, twoD = A.shape
    J = np.zeros((N, twoD), dtype=np.float64)
    draws = np.ascontiguousarray(draws, np.float64)
    rc = lib().orc_gen_J_f64(N, twoD, A, draws, len(draws), J)
    assert rc >= 0, rc
    return J


class Graph:
    def __init__(self, handle, keep=()):
        assert handle, "oracle graph creation failed"
        self.h = handle
        self._keep = keep
        self.N = lib().orc_getN(handle)
        self.kind = lib().orc_kind(handle)
        self._own = True

    @classmethod
    def ea_int(cls, A, J, lev=(-1, 1)):
        A = np.ascontiguousarray(A, np.int64); J = np.ascontiguousarray(J, np.int64)
        lev = np.ascontiguousarray(lev, np.int64)
        return cls(lib().orc_ea_int_create(A.shape[0], A.shape[1], A, J, lev, len(lev)))

    @classmethod
    def ea_f64(cls, A, J):
        A = np.ascontiguousarray(A, np.int64); J = np.ascontiguousarray(J, np.float64)
        return cls(lib().orc_ea_f64_create(A.shape[0], A.shape[1], A, J))

    @classmethod
    def ea_discretized(cls, A, cJ, lev=(-1, 0, 1)):
        """GraphEANormalDiscretized{Int,LEV,2D} from the continuous couplings cJ (EA.jl:311-344)."""
        A = np.ascontiguousarray(A, np.int64); cJ = np.ascontiguousarray(cJ, np.float64)
        lev = np.ascontiguousarray(lev, np.int64)
        return cls(lib().orc_ea_discretized_create(A.shape[0], A.shape[1], A, cJ, lev, len(lev)))

    @classmethod
    def rrg_int(cls, A, J, lev=(-1, 1)):
        """GraphRRG{Int,LEV,K}(A, J) (RRG.jl:112-137): neighbors() skips zero couplings."""
        A = np.ascontiguousarray(A, np.int64); J = np.ascontiguousarray(J, np.int64); lev = np.ascontiguousarray(lev, np.int64)
        return cls(lib().orc_rrg_int_create(A.shape[0], A.shape[1], A, J, lev, len(lev)))

    @classmethod
    def rrg_discretized(cls, A, cJ, lev=(-1, 0, 1)):
        """GraphRRGNormalDiscretized{Int,LEV,K} from the continuous couplings cJ (RRG.jl:274-310)."""
        A = np.ascontiguousarray(A, np.int64); cJ = np.ascontiguousarray(cJ, np.float64); lev = np.ascontiguousarray(lev, np.int64)
        return cls(lib().orc_rrg_discretized_create(A.shape[0], A.shape[1], A, cJ, lev, len(lev)))

    @classmethod
    def sk_f64(cls, J):
        J = np.ascontiguousarray(J, np.float64)
        return cls(lib().orc_sk_f64_create(J.shape[0], J))

    @classmethod
    def sk_bin(cls, J):
        J = np.ascontiguousarray(J, np.uint8)
        return cls(lib().orc_sk_bin_create(J.shape[0], J))

    @classmethod
    def qt(cls, N, M, fourK):
        return cls(lib().orc_qt_create(N, M, fourK))

    @classmethod
    def empty(cls, N):
        return cls(lib().orc_empty_create(N))

    @classmethod
    def quant(cls, Nk, M, Gamma, beta, inner_kind, J=None, A=None):
        twoD = 0
        Ap = None
        if inner_kind == SK_BIN:
            J = np.ascontiguousarray(J, np.uint8)
        elif inner_kind in (SK_F64, EA_F64):
            J = np.ascontiguousarray(J, np.float64)
        if inner_kind == EA_F64:
            A = np.ascontiguousarray(A, np.int64); twoD = A.shape[1]; Ap = A.ctypes.data
        Jp = J.ctypes.data if J is not None else None
        return cls(lib().orc_quant_create(Nk, M, Gamma, beta, inner_kind, Jp, twoD, Ap), keep=(J, A))

    def inner(self):
        g = Graph.__new__(Graph)
        g.h = lib().orc_inner_graph(self.h); g._keep = (self,); g._own = False
        g.N = lib().orc_getN(g.h); g.kind = lib().orc_kind(g.h)
        return g

    def energy(self, s): return lib().orc_energy(self.h, s)
    def delta_energy(self, s, i): return lib().orc_delta_energy(self.h, s, i)
    def delta_energy_residual(self, s, i): return lib().orc_delta_energy_residual(self.h, s, i)
    def spinflip(self, s, i): lib().orc_spinflip(self.h, s, i)

    def neighbors(self, i):
        out = np.zeros(self.N + 2, np.int64)
        n = lib().orc_neighbors(self.h, i, out)
        return out[:n].copy()

    def allDE(self):
        out = np.zeros(64, np.float64)
        n = lib().orc_allDE(self.h, out)
        assert n >= 0
        return out[:n].copy()

    def lfields(self):
        out = np.zeros(self.N, np.float64)
        n = lib().orc_get_lfields(self.h, out)
        return out[:n]

    def fourK(self): return lib().orc_quant_fourK(self.h)

    def __del__(self):
        try:
            if self._own:
                lib().orc_graph_free(self.h)
        except Exception:
            pass


def _mk_hook(hook):
    if hook is None:
        return HOOK(0), None
    def _h(user, it, E, acc):
        return 1 if hook(it, E, acc) else 0
    return HOOK(_h), _h


def _run(fn, g, beta, iters, step, s, src, hook, extra=()):
    cap = min(10 ** 8, iters // step)
    Es = np.zeros(max(cap, 1), np.float64)
    h, keep = _mk_hook(hook)
    res = fn(g.h, float(beta), int(iters), int(step), s, src.draws, *extra, h, None, Es.ctypes.data, cap)
    assert res.status == 0, res.status
    return Es[:min(res.nsamples, cap)].copy(), res


def standardMC(g, beta, iters, s, src, step=1, hook=None):
    return _run(lib().orc_standardMC, g, beta, iters, step, s, src, hook)


def rrrMC(g, beta, iters, s, src, step=1, hook=None, staged_thr=float("nan"), staged_thr_fact=5.0):
    return _run(lib().orc_rrrMC, g, beta, iters, step, s, src, hook, extra=(float(staged_thr), float(staged_thr_fact)))


def bklMC(g, beta, iters, s, src, step=1, hook=None):
    return _run(lib().orc_bklMC, g, beta, iters, step, s, src, hook)


def rank_rrrMC(g, beta, iters, s, src, step=1, hook=None):
    """rrrMC with the rank-select member pick (CPU model of csrc/chain_warp.cu)."""
    return _run(lib().orc_rank_rrrMC, g, beta, iters, step, s, src, hook)


def rank_bklMC(g, beta, iters, s, src, step=1, hook=None):
    return _run(lib().orc_rank_bklMC, g, beta, iters, step, s, src, hook)


def wtmMC(g, beta, samples, s, src, step=1.0, hook=None):
    """wtmMC(X, β, samples; step::Float64) (RRRMC.jl:376-430); the hook's first argument is the sample index."""
    cap = min(10 ** 8, samples)
    Es = np.zeros(max(cap, 1), np.float64)
    h, keep = _mk_hook(hook)
    res = lib().orc_wtmMC(g.h, float(beta), int(samples), float(step), s, src.draws, h, None, Es.ctypes.data, cap)
    assert res.status == 0, res.status
    return Es[:min(res.nsamples, cap)].copy(), res


def extremal_opt(g, ftau, iters, s, src, step=1, hook=None):
    """extremal_opt(X, τ, iters; step, hook) (RRRMC.jl:468-521) given fτ = cumsum(j^-τ); s is updated in place.
    Returns (Es at the hook instants, Cmin chunks, result struct with Emin / itmin / Efinal)."""
    cap = min(10 ** 8, iters // step)
    Es = np.zeros(max(cap, 1), np.float64)
    Cmin = np.zeros_like(s)
    if hook is None:
        h = EOHOOK(0)
    else:
        def _h(user, it, E, Emin):
            return 1 if hook(it, E, Emin) else 0
        h = EOHOOK(_h)
    ftau = np.ascontiguousarray(ftau, np.float64)
    assert ftau.shape == (g.N,)
    res = lib().orc_extremal_opt(g.h, ftau, int(iters), int(step), s, Cmin.ctypes.data, src.draws, h, None, Es.ctypes.data, cap)
    assert res.status == 0, res.status
    return Es[:min(res.nsamples, cap)].copy(), Cmin, res


def thresholds_fixed64(beta, D):
    """floor(exp(-β·4c)·2^64), c=1..D, as uint64 (host-supplied acceptance table)."""
    from fractions import Fraction
    import math
    out = []
    for c in range(1, D + 1):
        p = math.exp(-beta * 4 * c)
        v = int(Fraction(p) * (1 << 64))  # exact floor of the double p scaled by 2^64
        out.append(min(v, (1 << 64) - 1))
    return np.array(out, dtype=np.uint64)


def checkerboard_sweeps(L, D, R, spins, Jfwd, thr, K, seed, sweep0, nsweeps, accepted=None, M=0):
    """CPU model of the engine's checkerboard sweeps: K full bit planes, M merged planes, 32-bit tail."""
    acc_p = accepted.ctypes.data if accepted is not None else None
    lib().orc_checkerboard_sweeps(L, D, R, spins, np.ascontiguousarray(Jfwd, np.int8),
                                  np.ascontiguousarray(thr, np.uint64), K, M, seed, sweep0, nsweeps, acc_p)


CB_T1, CB_TC = 33, 129


def cb_sparse_tables(thr):
    """Count tables of the sparse procedure from the 64-bit fixed-point probabilities (oracle's long-double build)."""
    thr = np.ascontiguousarray(thr, np.uint64)
    tbl = np.zeros(CB_T1 + (len(thr) - 1) * CB_TC, np.uint32)
    lib().orc_cb_sparse_tables(thr, len(thr), tbl)
    return tbl


def cb_sparse_tables_exact(thr):
    """The same tables in exact rational arithmetic (pins the long-double builds to within one unit of 2^-32)."""
    from fractions import Fraction
    from math import comb
    out = []
    for c, t in enumerate(thr, start=1):
        n = 32 if c == 1 else 128
        p = Fraction(int(t), 1 << 64); q = 1 - p
        cdf = Fraction(0)
        for k in range(n + 1):
            cdf += comb(n, k) * p ** k * q ** (n - k)
            v = int(cdf * (1 << 32) + Fraction(1, 2))   # round half up
            out.append(0xffffffff if (k == n or v >= 1 << 32) else max(v, 1) - 1)
    return np.array(out, dtype=np.uint32)


def checkerboard_sweeps_sparse(L, D, R, spins, Jfwd, tbl, seed, sweep0, nsweeps, accepted=None):
    """CPU model of the engine's checkerboard sweeps with the sparse acceptance procedure."""
    acc_p = accepted.ctypes.data if accepted is not None else None
    assert len(tbl) == CB_T1 + (D - 1) * CB_TC
    lib().orc_checkerboard_sweeps_sparse(L, D, R, spins, np.ascontiguousarray(Jfwd, np.int8),
                                         np.ascontiguousarray(tbl, np.uint32), seed, sweep0, nsweeps, acc_p)


CBP_KA, CBP_KR = 64, 32
CBP_LEN = CBP_KA + 3 * CBP_KR


def cb_poisson_tables(thr):
    """Count tables TA | TB0 | TB | TC of the poisson procedure (oracle's long-double build)."""
    thr = np.ascontiguousarray(thr, np.uint64)
    tbl = np.zeros(CBP_LEN, np.uint32)
    lib().orc_cb_poisson_tables(thr, len(thr), tbl)
    return tbl


def cb_poisson_tables_exact(thr, digits=60):
    """The same tables in high-precision decimal arithmetic (pins the long-double builds to within one unit)."""
    from decimal import Decimal, getcontext, ROUND_HALF_EVEN
    getcontext().prec = digits
    lam = [Decimal(0)] * 5
    for c, t in enumerate(thr, start=1):
        lam[c] = -(1 - Decimal(int(t)) / Decimal(1 << 64)).ln()

    def table(mu, scale, last, n):
        out, pk, cdf = [], (-mu).exp(), Decimal(0)
        for k in range(n):
            cdf += pk
            v = int((cdf * scale).to_integral_value(rounding=ROUND_HALF_EVEN))
            out.append(last if (k == n - 1 or v > last) else max(v, 1) - 1)
            pk = pk * mu / (k + 1)
        return out
    TA = table(128 * (lam[1] - lam[2]), Decimal(1 << 32), 0xffffffff, CBP_KA)
    TB = table(128 * (lam[2] - lam[3]), Decimal(1 << 32), 0xffffffff, CBP_KR)
    TC = table(128 * lam[3], Decimal(1 << 32), 0xffffffff, CBP_KR)
    TB0 = table(128 * (lam[2] - lam[3]), Decimal(TC[0] + 1), TC[0], CBP_KR)
    return np.array(TA + TB0 + TB + TC, dtype=np.uint32)


def cb_poisson_nw(tbl, tol=None):
    """Smallest number of static position words NW in (1, 2, 4, 6) with P(level-1 count > 4·NW-1) <= tol, else 0;
    tol=None: the engine's per-NW defaults."""
    if tbl[CBP_KA - 2] != 0xffffffff:
        return 0
    for nw, d in ((1, 0.03), (2, 0.2), (4, 0.06), (6, 0.012)):
        if 1.0 - (float(tbl[4 * nw - 1]) + 1.0) / 2.0 ** 32 <= (d if tol is None else tol):
            return nw
    return 0


def checkerboard_sweeps_poisson(L, D, R, spins, Jfwd, tbl, NW, seed, sweep0, nsweeps, accepted=None):
    """CPU model of the engine's checkerboard sweeps with the poisson acceptance procedure."""
    acc_p = accepted.ctypes.data if accepted is not None else None
    assert len(tbl) == CBP_LEN and NW in (1, 2, 4, 6)
    lib().orc_checkerboard_sweeps_poisson(L, D, R, spins, np.ascontiguousarray(Jfwd, np.int8),
                                          np.ascontiguousarray(tbl, np.uint32), NW, seed, sweep0, nsweeps, acc_p)


def checkerboard_sweeps_poisson_ladder(L, D, R, spins, Jfwd, tbls, NW, seed, sweep0, nsweeps, accepted=None):
    """The same procedure on a β ladder: tbls[G][CBP_LEN], one table set per 128-replica group."""
    acc_p = accepted.ctypes.data if accepted is not None else None
    tbls = np.ascontiguousarray(tbls, np.uint32)
    assert tbls.shape == ((R + 127) // 128, CBP_LEN) and NW in (1, 2, 4, 6)
    lib().orc_checkerboard_sweeps_poisson_ladder(L, D, R, spins, np.ascontiguousarray(Jfwd, np.int8), tbls, NW, seed, sweep0,
                                                 nsweeps, acc_p)


def dfloat_ea_energy(A, J, s, fields=False):
    """energy(::GraphEA{DFloat64}) (EA.jl:195-222 on src/DFloats.jl arithmetic) for real couplings J in the reference
    (A, J) layout (zero entries of A = no neighbour) -> Float64 energy [, local fields 2·lf as Float64]."""
    A = np.ascontiguousarray(A, np.int64); J = np.ascontiguousarray(J, np.float64)
    N, twoD = A.shape
    lf2 = np.zeros(N, np.int64)
    E = lib().orc_dfloat_ea_energy(N, twoD, A, J, s, lf2.ctypes.data if fields else None)
    return (E, lf2 / 1e5) if fields else E


def sk_lockstep_sweeps(J, chunks, lf, E, acc, beta, seed, sweep0, nsweeps):
    """CPU model of the engine's lock-step Metropolis sweeps on GraphSKNormal; chunks [R, nchunks] uint64, lf [R, N], E and
    acc [R] are updated in place."""
    J = np.ascontiguousarray(J, np.float64); N = J.shape[0]; R = chunks.shape[0]
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(beta, np.float64), (R,)))
    assert chunks.dtype == np.uint64 and lf.shape == (R, N) and E.dtype == np.float64 and acc.dtype == np.int64
    lib().orc_sk_lockstep_sweeps(N, R, J, chunks, chunks.shape[1], lf, E, acc, b, seed, sweep0, nsweeps)


def tempering_decide(beta_group, E, seed, round_):
    """Exchange decisions of one tempering round -> uint8 [(G-1), 128] (CPU model of tempering.cu)."""
    bg = np.ascontiguousarray(beta_group, np.float64)
    E = np.ascontiguousarray(E, np.float64)
    assert len(E) == 128 * len(bg)
    swap = np.zeros((len(bg) - 1) * 128, np.uint8)
    lib().orc_tempering_decide(len(bg), bg, E, seed, round_, swap)
    return swap.reshape(len(bg) - 1, 128)


def checkerboard_sweeps_f64(L, D, R, spins, A, J, beta, seed, sweep0, nsweeps, accepted=None):
    """CPU model of the engine's checkerboard sweeps for continuous couplings (GraphEANormal)."""
    acc_p = accepted.ctypes.data if accepted is not None else None
    b = np.ascontiguousarray(np.broadcast_to(np.asarray(beta, np.float64), (R,)))
    lib().orc_checkerboard_sweeps_f64(L, D, R, spins, np.ascontiguousarray(A, np.int64), np.ascontiguousarray(J, np.float64),
                                      b, seed, sweep0, nsweeps, acc_p)
