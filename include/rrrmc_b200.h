/* rrrmc_b200.h — C ABI of the B200-native engine for RRRMC.jl's single-spin-flip Monte Carlo hot path.
 *
 * The reference (pure Julia) has no FFI; its "operator API" is multiple dispatch on
 * AbstractGraph (src/Interface.jl:87-270) plus the sampler functions (src/RRRMC.jl:81-359).
 * Each entry point below names the reference interface it replaces.  A Julia host binds these
 * with `ccall` (see INTEGRATION.md and rrrmc.jl_b200/julia/RRRMCB200.jl); this repository's
 * tests and bench drive the same symbols from Python through ctypes.
 *
 * Conventions kept from the reference: site indices are 1-based; a configuration is the
 * `BitVector` chunk array of src/Interface.jl:21-29 (site i = bit (i-1)&63 of UInt64 chunk
 * (i-1)>>6, σ = 2s-1); `energy` (re)initialises caches (Interface.jl:103); `delta_energy` is
 * evaluated before the flip, `update_cache`/`spinflip` after it (Interface.jl:84-85,127-128).
 * New: every call acts on a *replica batch* of R independent chains of the same graph.
 *
 * All functions return rrrmc_status_t (0 = ok). Negative codes map to Julia exceptions in the
 * shim (RRRMC_ERR_ARG -> ArgumentError). rrrmc_last_error() gives the thread-local message.
 * A context and everything created from it must be used from one host thread at a time.
 * There is no CPU fallback: every compute entry point fails with RRRMC_ERR_CUDA without a device.
 */
#ifndef RRRMC_B200_H
#define RRRMC_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int rrrmc_status_t;
#define RRRMC_OK               0
#define RRRMC_ERR_ARG         -1 /* invalid argument (Julia: ArgumentError)            */
#define RRRMC_ERR_CUDA        -2 /* CUDA runtime failure / no device                    */
#define RRRMC_ERR_UNSUPPORTED -3 /* valid request the engine does not implement         */
#define RRRMC_ERR_STATE       -4 /* object in the wrong state (e.g. energy not called)  */

typedef struct rrrmc_ctx   rrrmc_ctx_t;
typedef struct rrrmc_graph rrrmc_graph_t;
typedef struct rrrmc_state rrrmc_state_t;

const char *rrrmc_last_error(void);
const char *rrrmc_version(void);

/* ---- context: one device, one stream ------------------------------------------------------ */
/* cuda_stream: a cudaStream_t to launch on (e.g. torch's current stream) or NULL to create one. */
rrrmc_status_t rrrmc_ctx_create(int device, void *cuda_stream, rrrmc_ctx_t **out);
rrrmc_status_t rrrmc_ctx_destroy(rrrmc_ctx_t *ctx);
rrrmc_status_t rrrmc_ctx_sync(rrrmc_ctx_t *ctx);
/* CUDA-event stopwatch on the context's stream (bench.py times kernels with this). */
rrrmc_status_t rrrmc_ctx_timer_start(rrrmc_ctx_t *ctx);
rrrmc_status_t rrrmc_ctx_timer_stop(rrrmc_ctx_t *ctx, float *elapsed_ms);
/* number of kernels this context has launched so far (bench.py's gpu_launches). */
rrrmc_status_t rrrmc_ctx_launch_count(rrrmc_ctx_t *ctx, uint64_t *count);
/* writes >= `bytes` of device memory to evict L2 between timed steps. */
rrrmc_status_t rrrmc_ctx_flush_l2(rrrmc_ctx_t *ctx);

/* ---- graphs ------------------------------------------------------------------------------- */
#define RRRMC_EA_PM1 1 /* GraphEA{Int,(-1,1),2D}: J is int64, every entry ±1   (EA.jl:138-191) */
#define RRRMC_EA_INT 2 /* GraphEA{Int,LEV,2D}: J is int64, small integers      (EA.jl:181-190) */
#define RRRMC_EA_F64 3 /* GraphEANormal{2D}: J is double                       (EA.jl:534-574) */

/* Replaces GraphEA{ET,LEV,twoD}(A, J) (EA.jl:145-168) / GraphEANormal{twoD}(L, A, J) (EA.jl:540-552).
 * A: [N*2D] int64, 1-based, rows sorted ascending — must be the gen_EA(L,D) lattice (EA.jl:24-43);
 * J: [N*2D] aligned slot-for-slot with A (gen_J, EA.jl:45-71); int64 for PM1/INT, double for F64. */
rrrmc_status_t rrrmc_graph_ea_create(rrrmc_ctx_t *ctx, int L, int D, int coupling_kind,
                                     const int64_t *A, const void *J, rrrmc_graph_t **out);
#define RRRMC_SK_F64 4 /* GraphSKNormal: J is double [N*N], symmetric, zero diagonal (SK.jl:181-210) */
#define RRRMC_SK_BIN 5 /* GraphSK: J is uint8 [N*N] of 0/1 bits, J_ij = (2b-1)/sqrt(N) (SK.jl:28-60)  */
#define RRRMC_QT     6 /* GraphQT{fourK}: Trotter-direction ring couplings (QT.jl:42-121)              */
#define RRRMC_QUANT  7 /* GraphQuant{fourK,G}: M Trotter slices of a classical graph (QT.jl:126-321)   */
#define RRRMC_EMPTY  8 /* GraphEmpty as the inner graph of GraphQuant (GraphQ0T, QAliases.jl:19-31)    */
#define RRRMC_EA_DISCR 9 /* GraphEANormalDiscretized{Int,LEV,2D} <: DoubleGraph (EA.jl:311-344); see below    */

/* Replaces GraphRRG{Int,LEV,K}(A, J) (src/graphs/RRG.jl:112-137, ±J or small integer levels) and GraphRRGNormal
 * (continuous couplings) on an explicit K-regular adjacency A [N*K] (1-based, rows ascending; the reference draws it with
 * gen_RRG, RRG.jl:27-68) with slot-aligned symmetric couplings J (int64 / double by coupling_kind, as for
 * rrrmc_graph_ea_create). K <= 8. neighbors() lists the entries with a non-zero coupling (RRG.jl:133). Samplers run on
 * the chain engine (schedule RANDOM_SITE for standardMC). */
rrrmc_status_t rrrmc_graph_rrg_create(rrrmc_ctx_t *ctx, int64_t N, int K, int coupling_kind,
                                      const int64_t *A, const void *J, rrrmc_graph_t **out);
/* Replaces GraphRRGNormalDiscretized{Int,LEV,K} (RRG.jl:274-310), integer levels, continuous couplings cJ [N*K] passed in. */
rrrmc_status_t rrrmc_graph_rrg_discretized_create(rrrmc_ctx_t *ctx, int64_t N, int K, const int64_t *A, const double *cJ,
                                                  const int64_t *lev, int nlev, rrrmc_graph_t **out);

/* Replaces GraphEANormalDiscretized(L, D, LEV) with integer levels (EA.jl:311-344, e.g. (-1,0,1) as in test/runtests.jl:51)
 * and explicit continuous couplings cJ [N*2D] (double, slot-aligned with A, symmetric; the constructor draws them with
 * gen_J(randn), EA.jl:323-325). Every coupling is split by discretize (Common.jl:38-49) into a level (inner
 * GraphEA{Int,LEV}: what rrrMC's ΔE classes see) and a Float64 residual (accepted by accept(c, -βΔE1), RRRMC.jl:262). */
rrrmc_status_t rrrmc_graph_ea_discretized_create(rrrmc_ctx_t *ctx, int L, int D, const int64_t *A, const double *cJ,
                                                 const int64_t *lev, int nlev, rrrmc_graph_t **out);

/* Replaces GraphSKNormal(N) / GraphSK(N) with explicit couplings (SK.jl:181-199 / :28-49; the generators are
 * gen_J_gauss SK.jl:170-179 and gen_J SK.jl:17-26). J: [N*N] row-major, symmetric, zero diagonal; double for
 * RRRMC_SK_F64, uint8 0/1 for RRRMC_SK_BIN. */
rrrmc_status_t rrrmc_graph_sk_create(rrrmc_ctx_t *ctx, int64_t N, int coupling_kind, const void *J, rrrmc_graph_t **out);
/* Replaces GraphQuant(Nk, M, Γ, β, Gconstr, args...) (QT.jl:163-170) for the inner graphs on this path:
 * inner_kind RRRMC_SK_BIN (GraphQSKT, QAliases.jl:34-43), RRRMC_SK_F64 (GraphQSKNormalT, :46-47) or RRRMC_EMPTY
 * (GraphQ0T, :19-31; J_inner NULL). All M slices share J_inner [Nk*Nk]; fourK = round(2/β·log coth(βΓ/M), digits=8). */
rrrmc_status_t rrrmc_graph_quant_create(rrrmc_ctx_t *ctx, int64_t Nk, int64_t M, double Gamma, double beta,
                                        int inner_kind, const void *J_inner, rrrmc_graph_t **out);
/* Replaces GraphQEAT (QAliases.jl:51-81) = GraphQuant(N, M, Γ, β, GraphEANormal{2D}, L, A, J): the transverse-field
 * Edwards-Anderson model, M Trotter slices of one GraphEANormal instance (A, J in the reference layout, as for
 * rrrmc_graph_ea_create with RRRMC_EA_F64). */
rrrmc_status_t rrrmc_graph_quant_ea_create(rrrmc_ctx_t *ctx, int L, int D, int64_t M, double Gamma, double beta,
                                           const int64_t *A, const double *J, rrrmc_graph_t **out);
/* Replaces GraphQT{fourK}(N, M) (QT.jl:46-54) = inner_graph(X::GraphQuant) (Interface.jl:239-240). */
rrrmc_status_t rrrmc_graph_qt_create(rrrmc_ctx_t *ctx, int64_t N, int64_t M, double fourK, rrrmc_graph_t **out);
/* the fourK type parameter of a GraphQuant / GraphQT (QT.jl:42,165) */
rrrmc_status_t rrrmc_graph_fourK(const rrrmc_graph_t *g, double *fourK);
/* Convenience: the gen_EA adjacency itself (EA.jl:24-43) so hosts without the reference can build A. */
rrrmc_status_t rrrmc_gen_ea_adjacency(int L, int D, int64_t *A_out /* [L^D * 2D] */);
rrrmc_status_t rrrmc_graph_destroy(rrrmc_graph_t *g);

/* getN (Interface.jl:145) */
rrrmc_status_t rrrmc_getN(const rrrmc_graph_t *g, int64_t *N);
/* neighbors(X, i) (Interface.jl:158; EA.jl:292; Common.jl:78-92 AllButOne; QT.jl:105-108, :288-321), in the
 * reference's iteration order. `out` must have room for rrrmc_max_neighbors entries. */
rrrmc_status_t rrrmc_neighbors(const rrrmc_graph_t *g, int64_t site, int64_t *out, int *n);
rrrmc_status_t rrrmc_max_neighbors(const rrrmc_graph_t *g, int64_t *n);
/* allΔE(X) (Interface.jl:200-201; EA.jl:293-309): sorted non-negative |ΔE| values. out has room for 64. */
rrrmc_status_t rrrmc_allDE(const rrrmc_graph_t *g, double *out, int *n);

/* ---- replica-batch state (R chains; replaces `Config`, Interface.jl:21-54) ------------------ */
rrrmc_status_t rrrmc_state_create(rrrmc_graph_t *g, int64_t n_replicas, rrrmc_state_t **out);
rrrmc_status_t rrrmc_state_destroy(rrrmc_state_t *s);
/* Config(N) random init for every replica (Interface.jl:24-28), from Philox keyed by `seed`. */
rrrmc_status_t rrrmc_state_randomize(rrrmc_state_t *s, uint64_t seed);
/* chunks: [count][ceil(N/64)] uint64 in the reference BitVector layout; the library transposes
 * to/from its device layout. Unused high bits of the last chunk are written as zero on download. */
rrrmc_status_t rrrmc_state_upload(rrrmc_state_t *s, int64_t first_replica, int64_t count, const uint64_t *chunks);
rrrmc_status_t rrrmc_state_download(rrrmc_state_t *s, int64_t first_replica, int64_t count, uint64_t *chunks);

/* ---- Interface queries on the batch ------------------------------------------------------- */
/* energy(X, C) for every replica (Interface.jl:105; EA.jl:195-222): E_out[R]. Integer-valued
 * (exact) for PM1/INT graphs. Also (re)initialises the device-side local-field caches. */
rrrmc_status_t rrrmc_energy(rrrmc_state_t *s, double *E_out);
/* delta_energy(X, C, i) for every replica (Interface.jl:130; EA.jl:266-275): dE_out[R]. */
rrrmc_status_t rrrmc_delta_energy(rrrmc_state_t *s, int64_t site, double *dE_out);
/* delta_energy(X, C, i) for i = 1..N of one replica: dE_out[N]. */
rrrmc_status_t rrrmc_all_delta_energy(rrrmc_state_t *s, int64_t replica, double *dE_out);
/* spinflip!(X, C, i) = flip + update_cache! (Interface.jl:89-92; EA.jl:224-264) on the replicas
 * selected by replica_mask (bit r&31 of word r>>5; NULL = all replicas). */
rrrmc_status_t rrrmc_spinflip(rrrmc_state_t *s, int64_t site, const uint32_t *replica_mask);
/* magnetisation Σσ per replica (hook-side observable). */
rrrmc_status_t rrrmc_magnetization(rrrmc_state_t *s, double *m_out);
/* delta_energy_residual(X, C, i) (Interface.jl:254-261; QT.jl:270-281): dE_out[R]. 0 for single graphs.
 * Like the reference it reads the caches: call rrrmc_energy first. */
rrrmc_status_t rrrmc_delta_energy_residual(rrrmc_state_t *s, int64_t site, double *dE_out);
/* GraphQuant observables for hooks, per replica: transverse_mag(X, C, β) (QT.jl:113-121) -> out[R];
 * Qenergy(X, C) (QT.jl:253-268) -> out[R]; Renergies(X) (QT.jl:201-211) -> out[R*M]; overlaps(X) (QT.jl:213-251)
 * -> out[R*(M/2)]. They (re)compute the slice caches from the current configuration. */
rrrmc_status_t rrrmc_transverse_mag(rrrmc_state_t *s, double beta, double *out);
rrrmc_status_t rrrmc_Qenergy(rrrmc_state_t *s, double *out);
rrrmc_status_t rrrmc_Renergies(rrrmc_state_t *s, double *out);
rrrmc_status_t rrrmc_overlaps(rrrmc_state_t *s, double *out);

/* A GraphQuant batch as a parallel-tempering ladder: replica r sits at beta[r], hence at its own
 * fourK(β) = round(2/β·log coth(βΓ/M), digits=8) — a type parameter of GraphQuant{fourK,G} in the reference
 * (QT.jl:126,165), so there every β is a distinct graph; here it is per-replica state. Affects energy, delta_energy,
 * allΔE of the inner GraphQT inside the samplers, transverse_mag and Qenergy. beta = NULL restores the graph's β.
 * fourK_out (may be NULL) receives the R values. The β passed to a sampler must be the same vector. */
rrrmc_status_t rrrmc_state_set_quant_betas(rrrmc_state_t *s, const double *beta, double *fourK_out);

/* ---- samplers ------------------------------------------------------------------------------ */
/* hook(it, X, C, accepted, E)::Bool of RRRMC.jl:61-64,104-109, batched: E[R], accepted[R]
 * (accepted[r] = -1 when counting is disabled). Return 0 to stop the run. Called on the host
 * thread that invoked the sampler; the state may be queried/downloaded from inside the hook. */
typedef int (*rrrmc_hook_fn)(void *user, int64_t it, const double *E, const int64_t *accepted, int64_t R);

#define RRRMC_SCHED_CHECKERBOARD 0 /* two-colour lattice sweeps, all replicas in lock step (new engine) */
#define RRRMC_SCHED_RANDOM_SITE  1 /* the reference's order: i = rand(1:N) per attempt (RRRMC.jl:113)   */

/* How a checkerboard task turns Philox bits into the Metropolis filter accept() of RRRMC.jl:39.
 * Both are exact per-(site,replica) Bernoulli(exp(-βΔE)) decisions, independent across lanes; they consume the
 * counter stream differently, so trajectories differ between procedures (each has its own CPU restatement). */
#define RRRMC_CB_AUTO   0 /* poisson while its static slots cover the hit count (β >~ 0.5), else sparse/planes  */
#define RRRMC_CB_PLANES 1 /* bit-plane comparison of a (K+32)-bit uniform per lane against 64-bit thresholds */
#define RRRMC_CB_SPARSE 2 /* binomial count of passing lanes per ΔE class + uniform distinct positions       */
#define RRRMC_CB_POISSON 3 /* Poisson hit counts per task and level + uniform positions with replacement      */

#define RRRMC_PICK_REFERENCE 0
#define RRRMC_PICK_RANK 1
typedef struct {
    int    schedule;        /* RRRMC_SCHED_*; default RANDOM_SITE (the reference contract)           */
    int    planes_K;        /* checkerboard: full random bit planes, one Philox call each (default 5)  */
    int    count_accepted;  /* 1: exact per-replica accepted counters; 0: hook gets accepted = -1      */
    double staged_thr;      /* rrrMC: NaN = reference default (RRRMC.jl:163-165)                       */
    double staged_thr_fact; /* rrrMC: default 5.0 (RRRMC.jl:155)                                       */
    int    planes_M;        /* checkerboard: merged bit planes after the full ones, four per Philox
                               call (default 4, a multiple of 4); planes_K + planes_M <= 32. See DESIGN.md §5.          */
    int    cb_method;       /* checkerboard acceptance procedure: RRRMC_CB_AUTO (default), _PLANES, _SPARSE, _POISSON */
    int    site_pick;       /* rrrMC / bklMC: which member of the drawn ΔE class rand(1:t) picks. RRRMC_PICK_REFERENCE (default): the
                               reference's ArraySet order (ArraySets.jl:58-85) — bit-exact with the reference-order oracle and the
                               replay mode. RRRMC_PICK_RANK: the p-th member in site order, on the warp-cooperative kernel
                               (chain_warp.cu: one chain per warp in shared memory; ±J GraphEA lattices): the same chain law,
                               other trajectories; CPU model oracle/rrrmc_oracle.c:orc_rank_rrrMC / orc_rank_bklMC. */
    int    reserved[5];
} rrrmc_opts_t;
rrrmc_status_t rrrmc_opts_default(rrrmc_opts_t *o);

typedef struct {
    int64_t nsamples;   /* rows written to Es                                                  */
    int64_t iters_done; /* attempts per replica actually executed                              */
    int64_t launches;   /* kernels launched by this call                                       */
    float   device_ms;  /* CUDA-event time of the sampling kernels of this call (0 if unknown) */
    int64_t accepted_total; /* accepted moves summed over replicas (chain samplers; -1 if not counted) */
} rrrmc_run_info_t;

/* standardMC(X, β, iters; seed, step, hook, C0) (RRRMC.jl:81-127) on the batch. C0 is the state's
 * current contents (upload or randomize first). beta[R]: per-replica inverse temperature. Es: [Es_cap][R] row-major
 * or NULL. With the checkerboard schedule `iters` and `step` are rounded up to whole sweeps (N attempts per replica),
 * samples are post-sweep energies, and on ±J lattices β must be constant inside each group of 128 consecutive
 * replicas (the acceptance procedure draws one hit count per 128-lane task); GraphEANormal lattices take any β[r]. */
rrrmc_status_t rrrmc_standard_mc(rrrmc_state_t *s, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                                 rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *opts,
                                 double *Es, int64_t Es_cap, rrrmc_run_info_t *info);
/* rrrMC(X, β, iters; ...) (RRRMC.jl:149-290) and bklMC(X, β, iters; ...) (RRRMC.jl:311-359).
 * Es has iters ÷ step rows for every sampler. One documented deviation: with step > iters the reference's bklMC still
 * pushes ONE sample when a skip crosses `step` before the loop ends (RRRMC.jl:339-344 runs before the `it < iters`
 * test); whether that happens differs from chain to chain, a batch cannot return ragged rows, so no row is emitted. */
rrrmc_status_t rrrmc_rrr_mc(rrrmc_state_t *s, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                            rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *opts,
                            double *Es, int64_t Es_cap, rrrmc_run_info_t *info);
rrrmc_status_t rrrmc_bkl_mc(rrrmc_state_t *s, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                            rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *opts,
                            double *Es, int64_t Es_cap, rrrmc_run_info_t *info);

/* wtmMC(X, β, samples; step::Float64, seed, hook, C0) (RRRMC.jl:376-430): the rejection-free waiting-time method of
 * Dall and Sibani on src/WaitingTimes.jl. `step` is an interval of the sampler's global time (divided by N inside,
 * RRRMC.jl:394); at most `samples` rows of Es are produced, one per interval. The hook's `it` is the sample index k
 * (the reference passes the time k·step/N); its `accepted` is the number of moves so far. */
rrrmc_status_t rrrmc_wtm_mc(rrrmc_state_t *s, const double *beta, int64_t samples, double step, uint64_t seed,
                            rrrmc_hook_fn hook, void *user, double *Es, int64_t Es_cap, rrrmc_run_info_t *info);

/* extremal_opt(X, τ, iters; step, seed, hook, C0) (RRRMC.jl:468-521): τ-extremal optimisation on the EOCache of
 * DeltaE.jl:413-543 — spins ranked by ΔE class, rank drawn from the power law j^-τ, the chosen spin always flips.
 * DiscrGraph models (GraphEA / GraphRRG with integer levels, GraphQT) run on EOCache; the Float64 SimpleGraphs
 * (GraphEANormal, GraphRRGNormal, GraphSKNormal) on EOCacheCont (DeltaE.jl:555-635): ΔE of every spin kept in sorted
 * order (an insertion pass per move where the reference calls sortperm!), ONE draw per move, ties keep their order
 * (rankshuffle!, :608-633, only acts on equal ΔEs, which continuous couplings do not produce). DoubleGraphs return
 * RRRMC_ERR_UNSUPPORTED.
 * ftau: fτ = cumsum(j^-τ, j = 1..N) (DeltaE.jl:443), computed by the host language so that its own `^` and `cumsum`
 * bits are used; ftau_stride = 0: one table [N] for all chains, else chain r reads ftau + r·ftau_stride (per-chain τ).
 * The final configurations stay in the state (rrrmc_state_download). Outputs (any may be NULL): Emin_out[R],
 * itmin_out[R], Cmin_chunks[R][nchunks] (reference BitVector layout) — the reference's return values (C, Emin, Cmin,
 * itmin); Es[Es_cap][R]: energies at the hook instants (an aid, the reference returns no energy vector).
 * hook(it, X, C, E, Emin)::Bool of RRRMC.jl:499, batched over the chains. */
typedef int (*rrrmc_eo_hook_fn)(void *user, int64_t it, const double *E, const double *Emin, int64_t R);
rrrmc_status_t rrrmc_extremal_opt(rrrmc_state_t *s, const double *ftau, int64_t ftau_stride, int64_t iters, int64_t step,
                                  uint64_t seed, rrrmc_eo_hook_fn hook, void *user,
                                  double *Emin_out, int64_t *itmin_out, uint64_t *Cmin_chunks,
                                  double *Es, int64_t Es_cap, rrrmc_run_info_t *info);

/* Replay: feed one chain the typed draw stream the reference consumed
 * (kind 0 = rand(1:n) value, 1 = rand() value) and reproduce its trajectory.
 * sampler: 0 standardMC, 1 rrrMC, 2 bklMC. Es: [Es_cap] energies at every `step`. */
rrrmc_status_t rrrmc_replay(rrrmc_state_t *s, int64_t replica, int sampler, double beta, int64_t iters, int64_t step,
                            const uint8_t *draw_kind, const int64_t *draw_ival, const double *draw_fval, int64_t ndraws,
                            const rrrmc_opts_t *opts, double *Es, int64_t Es_cap, rrrmc_run_info_t *info);
/* Replay of wtmMC (RRRMC.jl:376-430: rand() draws only — N for the initial heap, then one for the moved spin and one per
 * neighbour, WaitingTimes.jl:25-51) and of extremal_opt (RRRMC.jl:468-521: rand() for the rank, rand(1:n) for the class
 * member) from a dumped draw stream; one chain. Es: [Es_cap] energies at the sampling instants. The extremal_opt replay
 * also returns the chain's Emin, itmin and Cmin[nchunks]. */
rrrmc_status_t rrrmc_replay_wtm(rrrmc_state_t *s, int64_t replica, double beta, int64_t samples, double step,
                                const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                                double *Es, int64_t Es_cap, rrrmc_run_info_t *info);
rrrmc_status_t rrrmc_replay_extremal_opt(rrrmc_state_t *s, int64_t replica, const double *ftau, int64_t iters, int64_t step,
                                         const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                                         double *Emin_out, int64_t *itmin_out, uint64_t *Cmin_chunks,
                                         double *Es, int64_t Es_cap, rrrmc_run_info_t *info);

/* Device-resident sweep loop without host round trips (what bench.py times as `value`):
 * runs `nsweeps` checkerboard sweeps starting at sweep counter `sweep0`. thr64: per-class
 * fixed-point acceptance table floor(exp(-β·ΔE_c)·2^64), ΔE_c = allΔE[c], c = 1..nclasses-1. */
rrrmc_status_t rrrmc_checkerboard_sweeps(rrrmc_state_t *s, const uint64_t *thr64, int nthr, int planes_K, int planes_M,
                                         uint64_t seed, uint64_t sweep0, int64_t nsweeps);

/* The same loop with the sparse procedure. tbl: count tables, class 1 (ΔE=4) first: 33 entries for Bin(32,p1)
 * (one count per 32-replica word), then 129 entries per class c=2..D for Bin(128,pc) (one count per 128-replica
 * task); entry k = round(P(count <= k)·2^32) - 1, last entry 2^32-1: more than k lanes pass iff x > tbl[k] for a
 * 32-bit uniform x. rrrmc_checkerboard_sparse_tables builds them from the 64-bit fixed-point probabilities. */
rrrmc_status_t rrrmc_checkerboard_sparse_tables(const uint64_t *thr64, int nthr, uint32_t *tbl, int tbl_len);
rrrmc_status_t rrrmc_checkerboard_sweeps_sparse(rrrmc_state_t *s, const uint32_t *tbl, int tbl_len,
                                                uint64_t seed, uint64_t sweep0, int64_t nsweeps);

/* The same loop with the poisson procedure (ea_poisson.cu). Every lane carries independent Poisson hit processes,
 * level-l hits at rate lam_l - lam_{l+1} with lam_c = -log(1 - p_c); a lane of class c (ΔE = 4c) flips iff it got a hit
 * of level >= c. tbl = TA[64] | TB0[32] | TB[32] | TC[32], entry k = round(P(count <= k)·2^32) - 1 for the per-task
 * (128-lane) counts of level-1, level-2 (rescaled to [0, TC[0]] and plain) and level-3 hits: more than k hits iff
 * x > T[k]. NW = static position words (1, 2, 4 or 6; 4·NW-1 level-1 hits are placed without branching);
 * rrrmc_checkerboard_poisson_nw returns the smallest NW whose overflow probability per task is <= tol (0: none;
 * tol <= 0: the measured per-NW defaults that AUTO uses). */
rrrmc_status_t rrrmc_checkerboard_poisson_tables(const uint64_t *thr64, int nthr, uint32_t *tbl, int tbl_len);
int rrrmc_checkerboard_poisson_nw(const uint32_t *tbl, double tol);
rrrmc_status_t rrrmc_checkerboard_sweeps_poisson(rrrmc_state_t *s, const uint32_t *tbl, int tbl_len, int NW,
                                                 uint64_t seed, uint64_t sweep0, int64_t nsweeps);
/* The same sweeps on a β ladder: tbls[ngroups][160], one table set per group of 128 consecutive replicas (ngroups =
 * replicas / 128), one NW for all (the warmest group's). Same Philox counters and procedure as the one-β entry: a group
 * evolves exactly as it would in a one-β batch with its table. Runs on the multi-sweep brick kernel only (D = 3, L a
 * multiple of 8, replicas a multiple of 1024), else RRRMC_ERR_UNSUPPORTED. This is what rrrmc_standard_mc runs when
 * beta[] differs between groups (parallel tempering on the checkerboard schedule; the reference tempers by running one
 * standardMC per β, RRRMC.jl:81-127). CPU restatement: oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_poisson_ladder. */
rrrmc_status_t rrrmc_checkerboard_sweeps_poisson_ladder(rrrmc_state_t *s, const uint32_t *tbls, int ngroups, int NW,
                                                        uint64_t seed, uint64_t sweep0, int64_t nsweeps);

/* Parallel-tempering exchange for a β ladder laid over the 128-replica groups of a ±J GraphEA batch (tempering.cu): lane l
 * of every group is one ladder, group g its rung at beta_group[g]. One round: energies on the device, then for the pairs
 * (g, g+1) with g ≡ round (mod 2) and every lane l: ΔS = (β_g − β_{g+1})(E_b − E_a), accept iff ΔS <= 0 or u < exp(−ΔS)
 * with u = 53 bits of Philox4x32-10(counter = (round_lo, round_hi, g, l), key = seed); accepted pairs exchange their
 * configurations (β stays constant inside a group, which the ladder sweeps need). Nothing crosses PCIe but beta_group.
 * accepted: NULL, or [ngroups-1] = exchanges accepted per pair since the last call that read them (reading synchronises).
 * The reference has no tempering (one standardMC per β, RRRMC.jl:81-127); this is the north-star's optional swap step.
 * CPU restatement of the decisions: oracle/rrrmc_oracle.c:orc_tempering_decide. */
rrrmc_status_t rrrmc_tempering_exchange(rrrmc_state_t *s, const double *beta_group, int ngroups, uint64_t seed, uint64_t round,
                                        int64_t *accepted);

/* Checkerboard Metropolis for continuous couplings: GraphEANormal (EA.jl:534-680) on the replica batch (ea_normal.cu).
 * Per (site, replica): ΔE = -2·lf with lf accumulated in Float64 in the slot order of energy() (EA.jl:590-603), i.e. the
 * value delta_energy() (EA.jl:665-672) returns on freshly initialised caches, then accept(-βΔE) of RRRMC.jl:39 with a
 * 53-bit uniform from Philox4x32-10(counter = (sweep_hi<<16, site, replica, sweep_lo), key = seed). beta[R] is per
 * replica. Also reached through rrrmc_standard_mc with schedule = RRRMC_SCHED_CHECKERBOARD on a GraphEANormal lattice
 * (even L). CPU restatement: oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_f64. */
rrrmc_status_t rrrmc_checkerboard_sweeps_f64(rrrmc_state_t *s, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps);

/* ---- dense GraphSKNormal path (BASELINE config 4) ----------------------------------------------
 * Local-field initialisation for the whole batch = the energy(X, C) contraction of SK.jl:212-237,
 * lfields[r][i] = 2 σ_ri Σ_j J_ij σ_rj. use_tensor_cores=1: exact fixed-point INT8 digit-plane GEMMs on tcgen05
 * (|error| <= N·2^-(P+1) from quantising J to 40 bits); 0: CUDA cores in the reference's summation order (bit-exact).
 * E_out[R] (may be NULL): energies; device_ms (may be NULL): CUDA-event time of the field kernels. */
rrrmc_status_t rrrmc_sk_fields_init(rrrmc_state_t *s, int use_tensor_cores, double *E_out, float *device_ms);
rrrmc_status_t rrrmc_sk_get_fields(rrrmc_state_t *s, double *lf_out /* [R*N] */);
/* Lock-step Metropolis sweeps (new engine): every replica attempts sites 1..N in order, accept() of RRRMC.jl:39 with
 * ΔE_i = lfields[i] (SK.jl:278-284) and the update_cache! axpy of SK.jl:252-265 on acceptance. U for (sweep t, site i,
 * replica r) is Philox4x32-10(ctr=(i, r, t_lo, t_hi^'SKLS'), key=seed). E_out[R], accepted_out[R] may be NULL;
 * the accepted counters accumulate over calls. */
rrrmc_status_t rrrmc_sk_metropolis_sweeps(rrrmc_state_t *s, const double *beta, uint64_t seed, uint64_t sweep0,
                                          int64_t nsweeps, double *E_out, int64_t *accepted_out);

#ifdef __cplusplus
}
#endif
#endif
