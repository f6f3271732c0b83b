// Launch parameters of the continuous-coupling checkerboard kernel (ea_normal.cu).
#pragma once
#include "common.cuh"

struct cbn_params {
    uint32_t *spins;        // [N][W]
    uint32_t *flips;        // [N][W] or nullptr: accept masks of the half-sweep (accepted counters)
    const int32_t *A;       // [N][2D] 0-based neighbours, reference slot order (EA.jl:24-43)
    const double *J;        // [N][2D] couplings aligned with A (EA.jl:45-71)
    const double *beta;     // [R] per-replica inverse temperature
    int L, D, twoD, W;
    int64_t R;
    int nwg;                // groups of four words per site
    int64_t ntasks;         // (N/2) * nwg warp tasks per colour
    uint32_t k0, k1;        // Philox key = seed
    uint32_t t_lo, t_hi16;  // sweep counter
};
rrrmc_status_t launch_checkerboard_f64(rrrmc_ctx *ctx, const cbn_params &p, int colour);
