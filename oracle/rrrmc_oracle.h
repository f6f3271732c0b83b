/* rrrmc_oracle.h — CPU restatement of RRRMC.jl's single-spin-flip hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the CPU baseline.
 *
 * PARITY UNPINNED: the reference (pure Julia) cannot run in this image and ships
 * no golden vectors (test/runtests.jl asserts only the energy-consistency
 * invariant).  This restatement is pinned by (i) that invariant, (ii) closed
 * forms / hand-checked adjacency from the reference sources, (iii) exact
 * Boltzmann stationarity on tiny instances — see tests/test_oracle_*.py.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout, e.g. src/graphs/EA.jl:195-222).
 */
#ifndef RRRMC_ORACLE_H
#define RRRMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {void orc_sk_lockstep_sweeps(int N, int64_t R, const double *J, uint64_t *chunks, int64_t nchunks, double *lf, double *E,
                            int64_t *acc, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps);

#endif

enum {
    ORC_EA_INT = 1, /* GraphEA{Int,LEV,twoD}   <: DiscrGraph{Int}      src/graphs/EA.jl:138-169 */
    ORC_EA_F64 = 2, /* GraphEANormal{twoD}     <: SimpleGraph{Float64} src/graphs/EA.jl:534-553 */
    ORC_SK_BIN = 3, /* GraphSK                 <: SimpleGraph{Float64} src/graphs/SK.jl:28-49   */
    ORC_SK_F64 = 4, /* GraphSKNormal           <: SimpleGraph{Float64} src/graphs/SK.jl:181-199 */
    ORC_QT     = 5, /* GraphQT{fourK}          <: DiscrGraph{Float64}  src/graphs/QT.jl:42-54   */
    ORC_QUANT  = 6, /* GraphQuant{fourK,G}     <: DoubleGraph          src/graphs/QT.jl:126-147 */
    ORC_EMPTY  = 7, /* GraphEmpty              <: SimpleGraph{Int}     src/graphs/Empty.jl:14-31*/
    ORC_EA_DISCR = 8 /* GraphEANormalDiscretized{Int,LEV,twoD} <: DoubleGraph  src/graphs/EA.jl:311-344 */
};

typedef struct orc_graph orc_graph;

/* ---- draw source (injectable; Appendix A.8 of SURVEY.md lists the draw order) ---- */
typedef struct {
    double  (*f64)(void *user);              /* Julia rand()      in [0,1) */
    int64_t (*range)(void *user, int64_t n); /* Julia rand(1:n)   in 1..n  */
    void *user;
} orc_draws;

/* Counter-based source shared with the CUDA chain kernels: Philox4x32-10,
 * key=(seed_lo,seed_hi), counter=(n_lo,n_hi,chain,tag), one call per draw. */
typedef struct { uint64_t seed, chain, n; uint32_t tag; } orc_philox_src;
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
double   orc_philox_f64(void *src);
int64_t  orc_philox_range(void *src, int64_t n);
uint64_t orc_philox_u64(orc_philox_src *src);

/* xoshiro256++ source (Julia >= 1.7's default generator family) for CPU-baseline timing. */
typedef struct { uint64_t s[4]; } orc_xoshiro_src;
void     orc_xoshiro_seed(orc_xoshiro_src *g, uint64_t seed);
double   orc_xoshiro_f64(void *src);
int64_t  orc_xoshiro_range(void *src, int64_t n);

/* Trace of typed draws: kind 0 = RANGE (ival), 1 = FLOAT (fval). */
typedef struct {
    int64_t len, cap, pos;
    uint8_t *kind; int64_t *ival; double *fval;
    orc_draws inner; /* recorder: source being recorded; replayer: unused */
    int error;       /* replayer: set when kind mismatches / trace exhausted */
} orc_trace;
orc_trace *orc_trace_new(void);
void       orc_trace_free(orc_trace *t);
void       orc_trace_load(orc_trace *t, int64_t len, const uint8_t *kind, const int64_t *ival, const double *fval);
double     orc_trace_rec_f64(void *t);
int64_t    orc_trace_rec_range(void *t, int64_t n);
double     orc_trace_play_f64(void *t);
int64_t    orc_trace_play_range(void *t, int64_t n);

/* ---- lattice / coupling generators ---- */
int64_t orc_gen_EA(int64_t L, int D, int64_t *A_out /* [L^D * 2D], 1-based, rows sorted */);
int     orc_gen_J_f64(int64_t N, int twoD, const int64_t *A, const double *draws, int64_t ndraws, double *J_out);

/* ---- graph constructors (A, site indices are 1-based like the reference) ---- */
orc_graph *orc_ea_int_create(int64_t N, int twoD, const int64_t *A, const int64_t *J, const int64_t *lev, int nlev);
orc_graph *orc_ea_f64_create(int64_t N, int twoD, const int64_t *A, const double *J);
/* GraphEANormalDiscretized with integer levels: cJ = the continuous couplings (slot-aligned with A, symmetric) */
orc_graph *orc_ea_discretized_create(int64_t N, int twoD, const int64_t *A, const double *cJ, const int64_t *lev, int nlev);
/* GraphRRG{Int,LEV,K} (neighbors() = entries with non-zero coupling, RRG.jl:133) and GraphRRGNormalDiscretized (RRG.jl:274-310) */
orc_graph *orc_rrg_int_create(int64_t N, int K, const int64_t *A, const int64_t *J, const int64_t *lev, int nlev);
orc_graph *orc_rrg_discretized_create(int64_t N, int K, const int64_t *A, const double *cJ, const int64_t *lev, int nlev);
orc_graph *orc_sk_f64_create(int64_t N, const double *J /* [N*N] row-major, symmetric, zero diag */);
orc_graph *orc_sk_bin_create(int64_t N, const uint8_t *J /* [N*N] 0/1, symmetric, zero diag */);
orc_graph *orc_qt_create(int64_t N, int64_t M, double fourK);
orc_graph *orc_empty_create(int64_t N);
/* GraphQuant(Nk,M,Γ,β,Gconstr,args...) QT.jl:163-170: inner_kind ∈ {ORC_SK_BIN, ORC_SK_F64, ORC_EMPTY, ORC_EA_F64};
 * all M slices share the inner couplings (QAliases.jl:43), each slice owns its cache. For ORC_EA_F64 pass twoD and A. */
orc_graph *orc_quant_create(int64_t Nk, int64_t M, double Gamma, double beta, int inner_kind,
                            const void *J_inner, int twoD, const int64_t *A_inner);
void       orc_graph_free(orc_graph *g);

/* ---- Interface (src/Interface.jl:87-270) ---- */
int     orc_kind(const orc_graph *g);
int64_t orc_getN(const orc_graph *g);
double  orc_energy(orc_graph *g, const uint64_t *chunks);            /* (re)initialises caches */
double  orc_delta_energy(orc_graph *g, const uint64_t *chunks, int64_t i);
double  orc_delta_energy_residual(orc_graph *g, const uint64_t *chunks, int64_t i);
void    orc_spinflip(orc_graph *g, uint64_t *chunks, int64_t i);    /* flip + update_cache! */
int     orc_neighbors(const orc_graph *g, int64_t i, int64_t *out); /* returns count */
int     orc_allDE(const orc_graph *g, double *out);                  /* DiscrGraph / DoubleGraph{DiscrGraph} only */
int64_t orc_get_lfields(const orc_graph *g, double *out);            /* cached local fields (as double) */
double  orc_quant_fourK(const orc_graph *g);
orc_graph *orc_inner_graph(orc_graph *g);
/* observables QT.jl:113-121, 201-268 */
double  orc_transverse_mag(orc_graph *g, const uint64_t *chunks, double beta);
double  orc_Qenergy(orc_graph *g, const uint64_t *chunks);
void    orc_Renergies(orc_graph *g, double *out /* [M] */);
void    orc_overlaps(orc_graph *g, double *out /* [M/2] */);

/* ---- samplers (src/RRRMC.jl:81-359). hook returns 0 to stop. ---- */
typedef int (*orc_hook)(void *user, int64_t it, double E, int64_t accepted);
typedef struct {
    int64_t nsamples;   /* entries written to Es */
    int64_t iters_done; /* final `it` */
    int64_t accepted;
    int64_t staged_its; /* rrr only */
    int     status;     /* 0 ok, <0 error */
} orc_result;

orc_result orc_standardMC(orc_graph *g, double beta, int64_t iters, int64_t step, uint64_t *chunks,
                          orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap);
/* staged_thr = NaN selects the reference default (0.8 Simple / 0.5 Discr / 0.5 Double) */
orc_result orc_rrrMC(orc_graph *g, double beta, int64_t iters, int64_t step, uint64_t *chunks,
                     orc_draws d, double staged_thr, double staged_thr_fact,
                     orc_hook hook, void *user, double *Es, int64_t Es_cap);
orc_result orc_bklMC(orc_graph *g, double beta, int64_t iters, int64_t step, uint64_t *chunks,
                     orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap);

/* wtmMC(X, β, samples; step::Float64) — RRRMC.jl:376-430 + WaitingTimes.jl. The hook's `it` is the sample index. */
orc_result orc_wtmMC(orc_graph *g, double beta, int64_t samples, double step, uint64_t *chunks,
                     orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap);

/* extremal_opt(X, τ, iters; step, hook) — RRRMC.jl:468-521 on EOCache (DeltaE.jl:413-543); DiscrGraph only.
 * ftau[N] = cumsum(j^-τ, j = 1..N), computed by the caller. Cmin (may be NULL) receives the configuration of minimum
 * energy. Es (may be NULL) records E at every hook instant — a test aid, the reference returns no energy vector. */
typedef int (*orc_eo_hook)(void *user, int64_t it, double E, double Emin);
typedef struct { int64_t nsamples, iters_done, itmin; double Emin; int status; double Efinal; } orc_eo_result;
orc_eo_result orc_extremal_opt(orc_graph *g, const double *ftau, int64_t iters, int64_t step, uint64_t *chunks,
                               uint64_t *Cmin, orc_draws d, orc_eo_hook hook, void *user, double *Es, int64_t Es_cap);

/* ΔE-class cache consistency (DeltaE.jl:120-136, ArraySets.jl:27-42); exposed for tests:
 * builds a cache for (g,chunks,beta), applies `nmoves` eager apply_move! calls on sites[], checks
 * consistency after each, and returns 0 when consistent. */
/* test probes of the container types (DynamicSamplers.jl:18-176, ArraySets.jl:58-85) */
int orc_ds_probe(int64_t N, const double *v, int64_t nset, const int64_t *set_i, const double *set_x,
                 double *ps_out, double *z_out, int64_t nq, const double *xq, int64_t *el_out);
int64_t orc_arrayset_probe(int64_t N, int64_t nops, const int64_t *op, int64_t *v_out);
int orc_check_discrete_cache(orc_graph *g, uint64_t *chunks, double beta, const int64_t *sites, int64_t nmoves);

/* ---- CPU model of the engine's checkerboard Metropolis (NOT in the reference).
 * Restates, with scalar per-(site,replica) loops on top of orc-level ΔE, the exact per-task
 * random-bit procedure the CUDA kernel uses, so the two can be compared bit for bit.
 * spins: multispin words [N][R/32] (site-major, 0-based site = x + L*y + L*L*z), updated in place.
 * J: forward-bond couplings [N][D] (±1) for bonds to x+1,y+1,(z+1).
 * thr: per-class 64-bit fixed-point acceptance thresholds floor(exp(-β·ΔE_c)·2^64), c=1..D (ΔE=4c). */
void orc_checkerboard_sweeps(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                             const uint64_t *thr, int K, int M, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                             int64_t *accepted /* [R] += */);

/* Same sweeps with the engine's "sparse" acceptance procedure: instead of comparing a uniform per
 * lane, each task draws, per ΔE class, HOW MANY of its lanes pass (a binomial count by inverse CDF on one 32-bit
 * uniform) and then WHICH ones (uniform distinct positions, duplicates redrawn).
 * tbl: class 1 (ΔE=4): 33 entries T1[k] = round(P(Bin(32,p1) <= k)·2^32) - 1, one count per 32-lane word
 *      (more than k lanes pass iff the 32-bit uniform x > T1[k]; T1[32] = 2^32-1 ends the scan);
 *      classes c=2..D: 129 entries each, Tc[k] likewise for Bin(128,pc), one count per 128-lane task. */
#define ORC_CB_T1 33
#define ORC_CB_TC 129
void orc_checkerboard_sweeps_sparse(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                                    const uint32_t *tbl, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                                    int64_t *accepted /* [R] += */);
/* Reference construction of those tables from the 64-bit fixed-point acceptance probabilities (long double). */
void orc_cb_sparse_tables(const uint64_t *thr, int D, uint32_t *tbl);

/* Same sweeps with the engine's "poisson" acceptance procedure: per task and hit level a Poisson count
 * (inverse CDF on a 32-bit uniform) of uniformly placed hits, with replacement; a lane of class c flips iff it
 * received a hit of level >= c. tbl = TA[64] | TB0[32] | TB[32] | TC[32] (see rrrmc_oracle.c); NW = number of static
 * position words (1, 2, 4 or 6). */
#define ORC_CBP_KA 64
#define ORC_CBP_KR 32
#define ORC_CBP_LEN (ORC_CBP_KA + 3 * ORC_CBP_KR)
void orc_checkerboard_sweeps_poisson(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                                     const uint32_t *tbl, int NW, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                                     int64_t *accepted /* [R] += */);
/* β ladder: one table set per 128-replica group, tbls[G][ORC_CBP_LEN] */
void orc_checkerboard_sweeps_poisson_ladder(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                                            const uint32_t *tbls, int NW, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                                            int64_t *accepted /* [R] += */);
void orc_cb_poisson_tables(const uint64_t *thr, int D, uint32_t *tbl);

/* checkerboard Metropolis for continuous couplings (GraphEANormal): CPU model of csrc/ea_normal.cu */
void orc_checkerboard_sweeps_f64(int L, int D, int64_t R, uint32_t *spins, const int64_t *A, const double *J,
                                 const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps, int64_t *accepted);

#ifdef __cplusplus
}void orc_sk_lockstep_sweeps(int N, int64_t R, const double *J, uint64_t *chunks, int64_t nchunks, double *lf, double *E,
                            int64_t *acc, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps);

#endif
/* parallel-tempering exchange decisions: CPU model of csrc/tempering.cu */
void orc_tempering_decide(int64_t G, const double *beta_group, const double *E, uint64_t seed, uint64_t round, uint8_t *swap);

/* rank-select rrrMC / bklMC for GraphEA ±J: CPU model of csrc/chain_warp.cu (see rrrmc_oracle.c) */
orc_result orc_rank_rrrMC(orc_graph *g, double beta, int64_t iters, int64_t step, uint64_t *chunks,
                          orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap);
orc_result orc_rank_bklMC(orc_graph *g, double beta, int64_t iters, int64_t step, uint64_t *chunks,
                          orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap);

/* DFloat64 (src/DFloats.jl:11-62): five-digit fixed point as Int64; energy(::GraphEA{DFloat64}) (EA.jl:195-222) */
int64_t orc_dfloat_from_f64(double x);
double orc_dfloat_to_f64(int64_t d);
int64_t orc_dfloat_add(int64_t a, int64_t b);
int64_t orc_dfloat_sub(int64_t a, int64_t b);
int64_t orc_dfloat_mul_int(int64_t k, int64_t a);
int64_t orc_dfloat_div_int(int64_t a, int64_t k);
double orc_dfloat_ea_energy(int64_t N, int twoD, const int64_t *A, const double *J, const uint64_t *s, int64_t *lf2);
void orc_sk_lockstep_sweeps(int N, int64_t R, const double *J, uint64_t *chunks, int64_t nchunks, double *lf, double *E,
                            int64_t *acc, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps);

#endif
