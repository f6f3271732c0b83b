// Sequential ("one Markov chain = one lane") samplers: the reference's standardMC / rrrMC / bklMC loops
// (src/RRRMC.jl:81-359) executed per replica with the reference's data structures in HBM.
#pragma once
#include "common.cuh"

enum { CHAIN_STANDARD = 0, CHAIN_RRR = 1, CHAIN_BKL = 2, CHAIN_WTM = 3, CHAIN_EO = 4 };
static inline bool is_sk_kind(int k) { return k == RRRMC_SK_F64 || k == RRRMC_SK_BIN; }

// replay mode: the typed draw stream one chain consumes (kind 0 = rand(1:n) value, 1 = rand() value)
struct chain_trace_in { int64_t replica; const uint8_t *kind; const int64_t *ival; const double *fval; int64_t n; };

void chain_free(rrrmc_state *s);
rrrmc_status_t chain_sync_to_multispin(rrrmc_state *s);   // make the multispin copy current
rrrmc_status_t chain_sync_from_multispin(rrrmc_state *s); // make the chain copy (d_chunks) current
rrrmc_status_t chain_energy(rrrmc_state *s, double *E_out);
rrrmc_status_t chain_delta_energy_site(rrrmc_state *s, int64_t site0, int what, double *out); // what: 0 ΔE, 1 residual
rrrmc_status_t chain_quant_observable(rrrmc_state *s, int what, double arg, double *out);
rrrmc_status_t chain_delta_energy_replica(rrrmc_state *s, int64_t replica, double *out);
rrrmc_status_t chain_run(rrrmc_state *s, int sampler, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                         rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info);
// wtmMC(X, β, samples; step::Float64) (RRRMC.jl:376-430): `step` in units of the global time, before the division by N
rrrmc_status_t chain_run_wtm(rrrmc_state *s, const double *beta, int64_t samples, double step, uint64_t seed,
                             rrrmc_hook_fn hook, void *user, double *Es, int64_t Es_cap, rrrmc_run_info_t *info,
                             const chain_trace_in *tr = nullptr);
// extremal_opt(X, τ, iters; step, hook) (RRRMC.jl:468-521) on the EOCache of DeltaE.jl:413-543; DiscrGraph only
rrrmc_status_t chain_run_eo(rrrmc_state *s, const double *ftau, int64_t ftau_stride, int64_t iters, int64_t step, uint64_t seed,
                            rrrmc_eo_hook_fn hook, void *user, double *Emin_out, int64_t *itmin_out, uint64_t *Cmin_chunks,
                            double *Es, int64_t Es_cap, rrrmc_run_info_t *info, const chain_trace_in *tr = nullptr);
rrrmc_status_t chain_replay(rrrmc_state *s, int64_t replica, int sampler, double beta, int64_t iters, int64_t step,
                            const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                            const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info);

// ---- per-chain state shared by chain.cu and chain_ea.cu -------------------------------------------------------
constexpr int MAXL = 16;  // |allΔE| supported by the discrete cache
constexpr int MAXDEG = 8; // neighbours of a discrete graph on this path: 2D <= 8 (EA), 2 (QT)
constexpr unsigned FULLMASK = 0xffffffffu;

struct chain_hdr {
    double E, acc_rate, z, pdE;
    double T[2 * MAXL + 1];
    long long it, accepted, staged_its, nextstep, skip, rng_n;
    int t[2 * MAXL + 1];
    int pending, pmove, status, built, trefresh, done;
    double wt_next;               // wtmMC: global time of the next sample (RRRMC.jl:396)
    double Emin; long long itmin; // extremal_opt: minimum energy found and its iteration (RRRMC.jl:480-482)
};

struct chain_store {
    int64_t R = 0, N = 0, N2 = 0;
    int levs = 0, nDE = 0;
    bool f64 = false;             // fp64 local fields (EA F64, SK F64, QUANT over SK F64)
    bool cont_ready = false, disc_ready = false, wtm_ready = false;
    double *eo_ftau = nullptr; int64_t eo_ftau_len = 0; uint64_t *eo_cmin = nullptr; // extremal_opt: fτ table(s), Cmin [R][nchunks]
    double *wt_v = nullptr; int32_t *wt_node = nullptr, *wt_pos = nullptr; // wtmMC heap [R][N]: times / sites in heap order, heap position of a site
    int32_t *lfi = nullptr;       // [R][2][N] (EA: cur,last; SK family: per slice [2][Nk])
    double *lfd = nullptr;
    int32_t *ml = nullptr;        // [R][M] move_last per slice (0-based, -1 = none)
    uint8_t *sw = nullptr;        // [R][M] SK family: which half of the slice's field pair is current
    chain_hdr *hdr = nullptr;
    int32_t *av = nullptr, *apos = nullptr;
    uint8_t *cls = nullptr;
    double *dEs = nullptr, *dv = nullptr, *dps = nullptr;
    int32_t *csj = nullptr; double *csdE = nullptr, *csp = nullptr; // staged list of the continuous cache [R][N+1]
    double *d_Es = nullptr; int64_t Es_rows = 0;
    double *d_DE = nullptr, *d_beta = nullptr, *d_E = nullptr, *d_aux = nullptr; int64_t aux_len = 0;
    uint8_t *d_tkind = nullptr; int64_t *d_tival = nullptr; double *d_tfval = nullptr; int64_t tcap = 0;
    // compact, L2-resident state of the GraphEA ±J fast path (chain_ea.cu)
    int8_t *ea_lf = nullptr;      // [R][N] lfields (EA.jl:214), |value| <= 4D
    uint16_t *ea_apos = nullptr;  // [R][N] 0-based position of a site inside its class set
    uint16_t *ea_av = nullptr;    // [R][2L][N] class sets
};

struct chain_params {
    int kind, N, twoD, sampler, nDE, levs, cpw, coop;
    int Nk, M, inner;
    int nz;                       // 1: GraphRRG semantics — neighbors() of the integer graph skips zero couplings (RRG.jl:133)
    double fourK, sN;
    const double *fourK_r;        // [R] per-replica fourK of a GraphQuant β ladder (NULL: the graph's)
    int64_t R, N2, nchunks, chain0;
    const int32_t *A; const int8_t *J8; const double *Jd; const uint8_t *Jb;
    uint64_t *chunks;
    int32_t *lfi; double *lfd; int32_t *ml; uint8_t *sw;
    chain_hdr *hdr;
    int32_t *av, *apos; uint8_t *cls;
    double *dEs, *dv, *dps;
    int32_t *csj; double *csdE, *csp;
    const double *DE, *beta;
    double *Es; int64_t Es_rows, quota;
    long long iters, step;
    uint64_t seed;
    double staged_thr, staged_thr_fact;
    const uint8_t *tkind; const int64_t *tival; const double *tfval; int64_t tlen;
    double *wt_v; int32_t *wt_node, *wt_pos; double wt_step, wt_tmax;  // wtmMC: heap, step/N, step/N·samples
    const double *eo_ftau; int64_t eo_stride; uint64_t *eo_cmin;            // extremal_opt: fτ [N] (stride 0) or [R][N], Cmin
    int fast;                     // 1: GraphEA ±J fast path (chain_ea.cu); 2: warp-cooperative rank-select kernel (chain_warp.cu)
    const uint8_t *jcode; int latL; // chain_warp.cu: forward bond signs of the lattice, its side
    int8_t *ea_lf; uint16_t *ea_apos, *ea_av;
};

// GraphEA ±J fast path of rrrMC / bklMC (chain_ea.cu): same algorithm and draw stream as k_chain_run, compact state
bool chain_ea_eligible(const rrrmc_state *s, int sampler);
rrrmc_status_t chain_ea_prepare(rrrmc_state *s, chain_params &P);
rrrmc_status_t chain_ea_launch(rrrmc_state *s, const chain_params &P);

// warp-cooperative rrrMC / bklMC with the rank-select member pick (chain_warp.cu): GraphEA ±J lattices, L >= 3, D <= 3
bool chain_warp_eligible(const rrrmc_state *s, int sampler);
rrrmc_status_t chain_warp_launch(rrrmc_state *s, const chain_params &P);

// dense GraphSKNormal kernels (sk_dense.cu): tensor-core local-field initialisation, lock-step Metropolis sweeps
void sk_dense_free(rrrmc_state *s);
void sk_dense_invalidate(rrrmc_state *s);
rrrmc_status_t sk_dense_fields_init(rrrmc_state *s, int use_tensor_cores, double *E_out, float *ms_out);
rrrmc_status_t sk_dense_get_fields(rrrmc_state *s, double *lf_out);
rrrmc_status_t sk_dense_sweeps(rrrmc_state *s, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                               double *E_out, int64_t *acc_out);
