"""rrrmc_b200 — B200-native engine for RRRMC.jl's single-spin-flip Monte Carlo hot path.

Layout: csrc/ (sm_100a CUDA kernels + the C ABI of include/rrrmc_b200.h), julia/ (the `ccall` shim a Julia host
uses), interface.py (this repository's executable mirror of the reference API, over ctypes)."""
from . import _ffi
from ._ffi import RRRMCError
from .interface import (DEFAULT_SEED, Config, Context, GraphEA, GraphEANormal, allDeltaE, all_delta_energy, bklMC,
                        delta_energy, energy, gen_EA, gen_J, neighbors, replay, rrrMC, spinflip, standardMC, update_cache)
from .interface import allΔE  # noqa: F401

__all__ = ["Config", "Context", "GraphEA", "GraphEANormal", "allDeltaE", "allΔE", "all_delta_energy", "bklMC",
           "delta_energy", "energy", "gen_EA", "gen_J", "neighbors", "replay", "rrrMC", "spinflip", "standardMC",
           "update_cache", "RRRMCError", "DEFAULT_SEED"]
