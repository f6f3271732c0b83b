// Sequential ("one Markov chain = one lane") samplers: the reference's standardMC / rrrMC / bklMC loops
// (src/RRRMC.jl:81-359) executed per replica with the reference's data structures in HBM.
#pragma once
#include "common.cuh"

enum { CHAIN_STANDARD = 0, CHAIN_RRR = 1, CHAIN_BKL = 2 };
static inline bool is_sk_kind(int k) { return k == RRRMC_SK_F64 || k == RRRMC_SK_BIN; }

void chain_free(rrrmc_state *s);
rrrmc_status_t chain_sync_to_multispin(rrrmc_state *s);   // make the multispin copy current
rrrmc_status_t chain_sync_from_multispin(rrrmc_state *s); // make the chain copy (d_chunks) current
rrrmc_status_t chain_energy(rrrmc_state *s, double *E_out);
rrrmc_status_t chain_delta_energy_site(rrrmc_state *s, int64_t site0, int what, double *out); // what: 0 ΔE, 1 residual
rrrmc_status_t chain_quant_observable(rrrmc_state *s, int what, double arg, double *out);
rrrmc_status_t chain_delta_energy_replica(rrrmc_state *s, int64_t replica, double *out);
rrrmc_status_t chain_run(rrrmc_state *s, int sampler, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                         rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info);
rrrmc_status_t chain_replay(rrrmc_state *s, int64_t replica, int sampler, double beta, int64_t iters, int64_t step,
                            const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                            const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info);

// dense GraphSKNormal kernels (sk_dense.cu): tensor-core local-field initialisation, lock-step Metropolis sweeps
void sk_dense_free(rrrmc_state *s);
void sk_dense_invalidate(rrrmc_state *s);
rrrmc_status_t sk_dense_fields_init(rrrmc_state *s, int use_tensor_cores, double *E_out, float *ms_out);
rrrmc_status_t sk_dense_get_fields(rrrmc_state *s, double *lf_out);
rrrmc_status_t sk_dense_sweeps(rrrmc_state *s, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                               double *E_out, int64_t *acc_out);
