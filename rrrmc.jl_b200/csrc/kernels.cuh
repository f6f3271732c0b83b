// Kernel launchers shared between translation units.
#pragma once
#include "common.cuh"

struct cb_params;
struct cbs_params;
struct cbp_params;
rrrmc_status_t launch_checkerboard(rrrmc_ctx *ctx, const cb_params &p, int D, int colour);
rrrmc_status_t launch_checkerboard_sparse(rrrmc_ctx *ctx, cbs_params &p, int D, int colour);
rrrmc_status_t launch_checkerboard_poisson(rrrmc_ctx *ctx, cbp_params &p, int D, int colour);
rrrmc_status_t launch_energy_pm1(rrrmc_state *s, int *d_unsat);
rrrmc_status_t launch_count_lanes(rrrmc_ctx *ctx, const uint32_t *masks, int64_t N, int W, long long *d_out);
rrrmc_status_t launch_delta_energy_site(rrrmc_state *s, int64_t site0, int *d_out);
rrrmc_status_t launch_delta_energy_replica(rrrmc_state *s, int64_t replica, int *d_out);
rrrmc_status_t launch_flip_site(rrrmc_state *s, int64_t site0, const uint32_t *d_mask);
rrrmc_status_t launch_randomize(rrrmc_state *s, uint64_t seed);
rrrmc_status_t launch_upload_transpose(rrrmc_state *s, int64_t first, int64_t count);
rrrmc_status_t launch_download_transpose(rrrmc_state *s, int64_t first, int64_t count);
rrrmc_status_t launch_flush(rrrmc_ctx *ctx);
rrrmc_status_t launch_tempering_exchange(rrrmc_state *s, const double *d_beta_group, uint32_t *d_masks, long long *d_accepted,
                                         uint64_t seed, uint64_t round);
