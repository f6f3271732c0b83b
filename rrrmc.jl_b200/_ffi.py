"""ctypes binding of librrrmc_b200.so — the same C ABI (include/rrrmc_b200.h) a Julia host `ccall`s.
There is no CPU fallback: if the library is missing this module raises at import of the symbols."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "lib", "librrrmc_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE = 0, -1, -2, -3, -4
EA_PM1, EA_INT, EA_F64, SK_F64, SK_BIN, QT, QUANT, EMPTY, EA_DISCR = 1, 2, 3, 4, 5, 6, 7, 8, 9
SCHED_CHECKERBOARD, SCHED_RANDOM_SITE = 0, 1
CB_AUTO, CB_PLANES, CB_SPARSE, CB_POISSON = 0, 1, 2, 3
PICK_REFERENCE, PICK_RANK = 0, 1
CBP_LEN = 64 + 3 * 32   # count tables of the poisson procedure: TA[64] | TB0[32] | TB[32] | TC[32]
CBS_T1, CBS_TC = 33, 129

HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int64)
EOHOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int64)


class Opts(C.Structure):
    _fields_ = [("schedule", C.c_int), ("planes_K", C.c_int), ("count_accepted", C.c_int),
                ("staged_thr", C.c_double), ("staged_thr_fact", C.c_double), ("planes_M", C.c_int), ("cb_method", C.c_int),
                ("site_pick", C.c_int), ("reserved", C.c_int * 5)]


class RunInfo(C.Structure):
    _fields_ = [("nsamples", C.c_int64), ("iters_done", C.c_int64), ("launches", C.c_int64), ("device_ms", C.c_float),
                ("accepted_total", C.c_int64)]


# every symbol include/rrrmc_b200.h declares: name -> (restype, argtypes)
_vp, _i64, _u64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_uint64, C.c_int, C.c_double
_pp = C.POINTER(C.c_void_p)
_SAMPLER = [_vp, _vp, _i64, _i64, _u64, HOOK, _vp, C.POINTER(Opts), _vp, _i64, C.POINTER(RunInfo)]
SIGNATURES = {
    "rrrmc_last_error": (C.c_char_p, []),
    "rrrmc_version": (C.c_char_p, []),
    "rrrmc_ctx_create": (_i32, [_i32, _vp, _pp]),
    "rrrmc_ctx_destroy": (_i32, [_vp]),
    "rrrmc_ctx_sync": (_i32, [_vp]),
    "rrrmc_ctx_timer_start": (_i32, [_vp]),
    "rrrmc_ctx_timer_stop": (_i32, [_vp, C.POINTER(C.c_float)]),
    "rrrmc_ctx_launch_count": (_i32, [_vp, C.POINTER(C.c_uint64)]),
    "rrrmc_ctx_flush_l2": (_i32, [_vp]),
    "rrrmc_graph_ea_create": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _pp]),
    "rrrmc_graph_rrg_create": (_i32, [_vp, _i64, _i32, _i32, _vp, _vp, _pp]),
    "rrrmc_graph_rrg_discretized_create": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _i32, _pp]),
    "rrrmc_graph_ea_discretized_create": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _i32, _pp]),
    "rrrmc_graph_quant_ea_create": (_i32, [_vp, _i32, _i32, _i64, _f64, _f64, _vp, _vp, _pp]),
    "rrrmc_graph_sk_create": (_i32, [_vp, _i64, _i32, _vp, _pp]),
    "rrrmc_graph_quant_create": (_i32, [_vp, _i64, _i64, _f64, _f64, _i32, _vp, _pp]),
    "rrrmc_graph_qt_create": (_i32, [_vp, _i64, _i64, _f64, _pp]),
    "rrrmc_graph_fourK": (_i32, [_vp, C.POINTER(C.c_double)]),
    "rrrmc_gen_ea_adjacency": (_i32, [_i32, _i32, _vp]),
    "rrrmc_graph_destroy": (_i32, [_vp]),
    "rrrmc_getN": (_i32, [_vp, C.POINTER(C.c_int64)]),
    "rrrmc_neighbors": (_i32, [_vp, _i64, _vp, C.POINTER(C.c_int)]),
    "rrrmc_max_neighbors": (_i32, [_vp, C.POINTER(C.c_int64)]),
    "rrrmc_allDE": (_i32, [_vp, _vp, C.POINTER(C.c_int)]),
    "rrrmc_state_create": (_i32, [_vp, _i64, _pp]),
    "rrrmc_state_destroy": (_i32, [_vp]),
    "rrrmc_state_randomize": (_i32, [_vp, _u64]),
    "rrrmc_state_upload": (_i32, [_vp, _i64, _i64, _vp]),
    "rrrmc_state_download": (_i32, [_vp, _i64, _i64, _vp]),
    "rrrmc_energy": (_i32, [_vp, _vp]),
    "rrrmc_delta_energy": (_i32, [_vp, _i64, _vp]),
    "rrrmc_all_delta_energy": (_i32, [_vp, _i64, _vp]),
    "rrrmc_spinflip": (_i32, [_vp, _i64, _vp]),
    "rrrmc_magnetization": (_i32, [_vp, _vp]),
    "rrrmc_delta_energy_residual": (_i32, [_vp, _i64, _vp]),
    "rrrmc_transverse_mag": (_i32, [_vp, _f64, _vp]),
    "rrrmc_Qenergy": (_i32, [_vp, _vp]),
    "rrrmc_Renergies": (_i32, [_vp, _vp]),
    "rrrmc_overlaps": (_i32, [_vp, _vp]),
    "rrrmc_sk_fields_init": (_i32, [_vp, _i32, _vp, C.POINTER(C.c_float)]),
    "rrrmc_sk_get_fields": (_i32, [_vp, _vp]),
    "rrrmc_sk_metropolis_sweeps": (_i32, [_vp, _vp, _u64, _u64, _i64, _vp, _vp]),
    "rrrmc_opts_default": (_i32, [C.POINTER(Opts)]),
    "rrrmc_standard_mc": (_i32, _SAMPLER),
    "rrrmc_rrr_mc": (_i32, _SAMPLER),
    "rrrmc_bkl_mc": (_i32, _SAMPLER),
    "rrrmc_wtm_mc": (_i32, [_vp, _vp, _i64, _f64, _u64, HOOK, _vp, _vp, _i64, C.POINTER(RunInfo)]),
    "rrrmc_state_set_quant_betas": (_i32, [_vp, _vp, _vp]),
    "rrrmc_extremal_opt": (_i32, [_vp, _vp, _i64, _i64, _i64, _u64, EOHOOK, _vp, _vp, _vp, _vp, _vp, _i64, C.POINTER(RunInfo)]),
    "rrrmc_replay": (_i32, [_vp, _i64, _i32, _f64, _i64, _i64, _vp, _vp, _vp, _i64, C.POINTER(Opts), _vp, _i64, C.POINTER(RunInfo)]),
    "rrrmc_replay_wtm": (_i32, [_vp, _i64, _f64, _i64, _f64, _vp, _vp, _vp, _i64, _vp, _i64, C.POINTER(RunInfo)]),
    "rrrmc_replay_extremal_opt": (_i32, [_vp, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _i64, C.POINTER(RunInfo)]),
    "rrrmc_checkerboard_sweeps": (_i32, [_vp, _vp, _i32, _i32, _i32, _u64, _u64, _i64]),
    "rrrmc_checkerboard_sparse_tables": (_i32, [_vp, _i32, _vp, _i32]),
    "rrrmc_checkerboard_sweeps_sparse": (_i32, [_vp, _vp, _i32, _u64, _u64, _i64]),
    "rrrmc_checkerboard_poisson_tables": (_i32, [_vp, _i32, _vp, _i32]),
    "rrrmc_checkerboard_poisson_nw": (_i32, [_vp, C.c_double]),
    "rrrmc_checkerboard_sweeps_poisson": (_i32, [_vp, _vp, _i32, _i32, _u64, _u64, _i64]),
    "rrrmc_checkerboard_sweeps_poisson_ladder": (_i32, [_vp, _vp, _i32, _i32, _u64, _u64, _i64]),
    "rrrmc_checkerboard_sweeps_f64": (_i32, [_vp, _vp, _u64, _u64, _i64]),
    "rrrmc_tempering_exchange": (_i32, [_vp, _vp, _i32, _u64, _u64, _vp]),
}

_lib = None


class RRRMCError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(SO):
            raise RRRMCError(f"{SO} is missing: build it with `python rrrmc.jl_b200/build.py` (there is no CPU fallback)")
        L = C.CDLL(SO)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def check(status):
    """status codes -> exceptions, the way the Julia shim maps them (ERR_ARG -> ArgumentError)."""
    if status == OK:
        return
    msg = lib().rrrmc_last_error().decode()
    if status == ERR_ARG:
        raise ValueError(msg)
    if status == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RRRMCError(f"[{status}] {msg}")


def ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None
