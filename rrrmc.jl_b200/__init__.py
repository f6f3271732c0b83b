"""rrrmc_b200 — B200-native engine for RRRMC.jl's single-spin-flip Monte Carlo hot path.

Layout: csrc/ (sm_100a CUDA kernels + the C ABI of include/rrrmc_b200.h), julia/ (the `ccall` shim a Julia host
uses), interface.py (this repository's executable mirror of the reference API, over ctypes)."""
from . import _ffi
from ._ffi import RRRMCError
from .interface import (DEFAULT_SEED, ON_DEVICE, Config, Context, GraphEA, GraphEANormal, GraphEANormalDiscretized, GraphQ0T, GraphQEAT, GraphRRG, GraphRRGNormal, GraphRRGNormalDiscretized, gen_RRG, gen_J_graph, GraphQSKNormalT, GraphQSKT, GraphQT,
                        GraphQuant, GraphSK, GraphSKNormal, Qenergy, Renergies, allDeltaE, all_delta_energy, bklMC, delta_energy,
                        delta_energy_residual, energy, gen_EA, gen_J, gen_J_bits, gen_J_gauss, neighbors, overlaps, replay, replay_extremal_opt, replay_wtm, rrrMC,
                        sk_fields_init, sk_metropolis_sweeps, checkerboard_sweeps_normal, spinflip, standardMC, transverse_mag, update_cache, wtmMC, extremal_opt, eo_ftau, gen_AJ, write_AJ)
from .interface import allΔE  # noqa: F401

__all__ = ["ON_DEVICE", "Config", "Context", "GraphEA", "GraphEANormal", "GraphEANormalDiscretized", "GraphSK", "GraphSKNormal", "GraphQT", "GraphQuant", "GraphQSKT",
           "GraphQSKNormalT", "GraphQ0T", "GraphQEAT", "GraphRRG", "GraphRRGNormal", "GraphRRGNormalDiscretized", "gen_RRG", "gen_J_graph", "allDeltaE", "allΔE", "all_delta_energy", "bklMC", "delta_energy",
           "delta_energy_residual", "energy", "gen_EA", "gen_J", "gen_J_gauss", "gen_J_bits", "neighbors", "replay", "replay_extremal_opt", "replay_wtm", "rrrMC",
           "spinflip", "standardMC", "wtmMC", "extremal_opt", "eo_ftau", "gen_AJ", "write_AJ", "update_cache", "sk_fields_init", "sk_metropolis_sweeps", "checkerboard_sweeps_normal", "transverse_mag", "Qenergy", "Renergies", "overlaps", "RRRMCError",
           "DEFAULT_SEED"]
