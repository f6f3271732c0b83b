// Sequential samplers on the chain layout: every replica is an independent Markov chain that runs the
// reference's loop with the reference's data structures (local-field cache, ΔE classes as ArraySets,
// Wong-Easton dynamic sampler).  Chains are latency-bound; they are spread one per warp across the SMs while
// R is small so that divergent chains never serialise each other.  For the fully connected families (SK and
// GraphQuant over SK) a chain owns a whole warp: lane 0 runs the sampler, the other lanes wait in a helper loop
// and join it for the O(N) local-field update of an accepted flip.
//
// Reference map: standardMC RRRMC.jl:81-127; rrrMC RRRMC.jl:131-290; bklMC RRRMC.jl:294-359;
// DeltaECache DeltaE.jl:63-295; ArraySet ArraySets.jl:58-85; DeltaECacheCont DeltaE.jl:297-410;
// DynamicSampler DynamicSamplers.jl:84-176; GraphEA cache EA.jl:195-275, 584-663; GraphSKNormal SK.jl:212-284;
// GraphSK SK.jl:62-140; GraphQT QT.jl:68-111; GraphQuant QT.jl:172-199, 270-321.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "chain.cuh"
#include "kernels.cuh"
#include "philox.cuh"
#include <type_traits>

#include "chain_kernel.cuh"

// replay mode: k_chain_run<src_trace> is compiled in chain_trace.cu
void chain_launch_run_trace(const chain_params &P, unsigned grid, cudaStream_t st);

// ------------------------------------------------------------------------------------------------
// energy(X, C) on the chain layout: (re)initialises the local fields (Interface.jl:103), then the per-chain sum
// in site order (sequential, so Float64 energies round exactly like the reference's loop)
// ------------------------------------------------------------------------------------------------
__global__ void k_chain_lfields(chain_params P)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.R * P.N) return;
    const int64_t r = tid / P.N; const int x = (int)(tid % P.N);
    const uint64_t *s = P.chunks + r * P.nchunks;
    if (x < P.M) { P.ml[r * P.M + x] = -1; P.sw[r * P.M + x] = 0; }
    if (P.kind == RRRMC_QT) return;
    if (is_ea(P.kind)) { // EA.jl:201-215 / :591-605
        const int sx = 2 * sget(s, x) - 1;
        if (P.kind == RRRMC_EA_F64 || P.kind == RRRMC_EA_DISCR) {   // DISCR: residual couplings (EA.jl:368-384)
            double lf = 0.0;
            for (int k = 0; k < P.twoD; k++) {
                const int y = P.A[(int64_t)x * P.twoD + k];
                const double sy = (double)(2 * sget(s, y) - 1);
                lf = __dsub_rn(lf, __dmul_rn(__dmul_rn(P.Jd[(int64_t)x * P.twoD + k], (double)sx), sy));
            }
            P.lfd[r * 2 * P.N + x] = 2 * lf; P.lfd[r * 2 * P.N + P.N + x] = 0.0;
        }
        if (P.kind != RRRMC_EA_F64) {
            int lf = 0;
            for (int k = 0; k < P.twoD; k++) {
                const int y = P.A[(int64_t)x * P.twoD + k];
                lf -= (int)P.J8[(int64_t)x * P.twoD + k] * sx * (2 * sget(s, y) - 1);
            }
            P.lfi[r * 2 * P.N + x] = 2 * lf; P.lfi[r * 2 * P.N + P.N + x] = 0;
        }
        return;
    }
    // SK family: slice k, site i of the slice (J symmetric: read column i so that consecutive threads coalesce)
    const int skind = P.kind == RRRMC_QUANT ? P.inner : P.kind;
    if (skind == RRRMC_EMPTY) return;
    const int k = x / P.Nk, i = x % P.Nk;
    const int64_t off = (int64_t)k * P.Nk, cur = r * 2 * P.N + (int64_t)k * 2 * P.Nk;
    const int si = sget(s, (int)(off + i));
    if (skind == RRRMC_EA_F64) { // GraphEANormal slice, EA.jl:591-605
        double lf = 0.0;
        for (int q = 0; q < P.twoD; q++) {
            const int y = P.A[(int64_t)i * P.twoD + q];
            const double sy = (double)(2 * sget(s, (int)(off + y)) - 1);
            lf = __dsub_rn(lf, __dmul_rn(__dmul_rn(P.Jd[(int64_t)i * P.twoD + q], (double)(2 * si - 1)), sy));
        }
        P.lfd[cur + i] = 2 * lf; P.lfd[cur + P.Nk + i] = 0.0;
        return;
    }
    if (skind == RRRMC_SK_F64) { // SK.jl:218-231
        double lf = 0.0;
        for (int j = 0; j < P.Nk; j++) lf = __dadd_rn(lf, __dmul_rn((double)(1 - 2 * (si ^ sget(s, (int)(off + j)))), P.Jd[(int64_t)j * P.Nk + i]));
        P.lfd[cur + i] = 2 * lf; P.lfd[cur + P.Nk + i] = 0.0;
    } else {                     // SK.jl:68-76
        int sc = 0;
        for (int j = 0; j < P.Nk; j++) sc += (int)P.Jb[(int64_t)j * P.Nk + i] ^ sget(s, (int)(off + j));
        const int lf = -(2 * si - 1) * (P.Nk - 1 - 2 * sc);
        P.lfi[cur + i] = 2 * (-lf + 2 * si); P.lfi[cur + P.Nk + i] = 0;
    }
}
// energy of one SK slice from its fresh fields, summed in site order (SK.jl:212-237 / :62-94)
__device__ double sk_slice_energy(const chain_params &P, int skind, int64_t r, int k, const uint64_t *s)
{
    const int64_t cur = r * 2 * P.N + (int64_t)k * 2 * P.Nk;
    if (skind == RRRMC_SK_F64) {
        double n = 0.0;
        for (int i = 0; i < P.Nk; i++) n = __dsub_rn(n, P.lfd[cur + i] / 2);
        return n / 2;
    }
    if (skind == RRRMC_EA_F64) {   // EA.jl:606-611
        double e = 0.0;
        for (int i = 0; i < P.Nk; i++) e = __dadd_rn(e, P.lfd[cur + i] / 2);
        return e / 2;
    }
    if (skind == RRRMC_SK_BIN) {
        long long sums = 0, n;
        for (int i = 0; i < P.Nk; i++) sums += sget(s, (int)((int64_t)k * P.Nk + i));
        n = -2 * sums;
        for (int i = 0; i < P.Nk; i++) { const int si = sget(s, (int)((int64_t)k * P.Nk + i)); n += -(P.lfi[cur + i] / 2 - 2 * si); }
        n /= 2;
        return (double)n / P.sN;
    }
    return 0.0;
}
__device__ long long qt_energy0(const chain_params &P, const uint64_t *s) // QT.jl:68-82
{
    long long n = 0;
    for (int i = 0; i < P.Nk; i++) {
        int sj = sget(s, i + (P.M - 1) * P.Nk);
        for (int k = 0; k < P.M; k++) { const int sk = sget(s, i + k * P.Nk); n -= 1 - 2 * (sk ^ sj); sj = sk; }
    }
    return n;
}
// mode 0: energy(X, C) -> E_out[r];  3: same and start the chain's tracked energy from it;
// 1: Renergies -> E_out[r*M + k];  2: energy0 -> E_out[r]
__global__ void k_chain_energy_sum(chain_params P, double *E_out, int mode)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) return;
    const uint64_t *s = P.chunks + r * P.nchunks;
    if (mode == 1) { for (int k = 0; k < P.M; k++) E_out[r * P.M + k] = sk_slice_energy(P, P.inner, r, k, s); return; }
    if (mode == 2) { E_out[r] = (double)qt_energy0(P, s); return; }
    double E;
    if (P.kind == RRRMC_EA_F64) { // EA.jl:606-611
        double e = 0.0;
        for (int x = 0; x < P.N; x++) e = __dadd_rn(e, P.lfd[r * 2 * P.N + x] / 2);
        E = e / 2;
    } else if (is_ea(P.kind)) {   // EA.jl:216-221
        long long n = 0;
        for (int x = 0; x < P.N; x++) n += P.lfi[r * 2 * P.N + x] / 2;
        E = (double)n / 2.0;
        if (P.kind == RRRMC_EA_DISCR) {   // EA.jl:362-388: E0 + E1, E1 as in GraphEANormal on the residuals
            double e = 0.0;
            for (int x = 0; x < P.N; x++) e = __dadd_rn(e, P.lfd[r * 2 * P.N + x] / 2);
            E = __dadd_rn(E, e / 2);
        }
    } else if (is_sk(P.kind)) E = sk_slice_energy(P, P.kind, r, 0, s);
    else if (P.kind == RRRMC_QT) E = (double)qt_energy0(P, s) * P.fourK / 4; // QT.jl:84
    else {                        // GraphQuant, QT.jl:185-199
        E = (double)qt_energy0(P, s) * (P.fourK_r ? P.fourK_r[r] : P.fourK) / 4;
        for (int k = 0; k < P.M; k++) E = __dadd_rn(E, sk_slice_energy(P, P.inner, r, k, s) / (double)P.M);
    }
    E_out[r] = E;
    if (mode == 3) P.hdr[r].E = E;
}
__global__ void k_chain_hdr_reset(chain_params P, int keep_rng)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) return;
    chain_hdr &h = P.hdr[r];
    h.acc_rate = 0.5; h.z = 0; h.pdE = 0;
    h.it = 0; h.accepted = 0; h.staged_its = 0; h.nextstep = P.step; h.skip = 0;
    if (!keep_rng) h.rng_n = 0;
    h.pending = 0; h.pmove = 0; h.status = 0; h.built = 0; h.trefresh = 0; h.done = 0;
    h.wt_next = P.wt_step;
    h.Emin = 0; h.itmin = 0;
}
// delta_energy straight from the caches: what = 0 delta_energy(X,C,i), 1 residual; all replicas of one site
__global__ void k_chain_delta_site(chain_params P, int site, int what, double *out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) return;
    const gview X = make_view(P, r);
    out[r] = what ? gv_delta_residual(X, site) : gv_delta_energy(X, site);
}
// delta_energy(X, C, i) of every spin of every chain: out[r][i]
__global__ void k_chain_delta_all(chain_params P, double *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t r = blockIdx.y;
    if (x >= P.N) return;
    const gview X = make_view(P, P.chain0 + r);
    out[(P.chain0 + r) * P.N + x] = gv_delta_energy(X, x);
}
__global__ void k_chain_delta_replica(chain_params P, int64_t r, double *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.N) return;
    const gview X = make_view(P, r);
    out[x] = gv_delta_energy(X, x);
}
// overlaps(X) (QT.jl:213-251): out[r*(M/2) + d-1] = normalised Σ over slice pairs at Trotter distance d of Σ_i σσ'
__global__ void k_chain_overlaps(chain_params P, double *out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) return;
    const uint64_t *s = P.chunks + r * P.nchunks;
    const int M = P.M, Nk = P.Nk, H = M / 2;
    for (int d = 0; d < H; d++) out[r * H + d] = 0.0;
    for (int k1 = 0; k1 < M - 1; k1++)
        for (int k2 = k1 + 1; k2 < M; k2++) {
            long long sum = 0;
            for (int i = 0; i < Nk; i++) sum += sget(s, k1 * Nk + i) ^ sget(s, k2 * Nk + i);
            const int dl = k2 - k1 < M + k1 - k2 ? k2 - k1 : M + k1 - k2;
            out[r * H + dl - 1] += (double)(Nk - 2 * sum);
        }
    for (int d = 1; d <= (M - 1) / 2; d++) out[r * H + d - 1] /= (double)((long long)M * Nk);
    if (M % 2 == 0) out[r * H + H - 1] /= (double)((long long)M * Nk) / 2; // multiplicity M/2 at the antipodal distance (QT.jl:240-249)
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void chain_free(rrrmc_state *s)
{
    chain_store *c = s->chain;
    if (!c) return;
    cudaFree(c->lfi); cudaFree(c->lfd); cudaFree(c->ml); cudaFree(c->sw); cudaFree(c->hdr);
    cudaFree(c->av); cudaFree(c->apos); cudaFree(c->cls); cudaFree(c->dEs); cudaFree(c->dv); cudaFree(c->dps);
    cudaFree(c->csj); cudaFree(c->csdE); cudaFree(c->csp);
    cudaFree(c->d_Es); cudaFree(c->d_DE); cudaFree(c->d_beta); cudaFree(c->d_E); cudaFree(c->d_aux);
    cudaFree(c->d_tkind); cudaFree(c->d_tival); cudaFree(c->d_tfval);
    cudaFree(c->ea_lf); cudaFree(c->ea_apos); cudaFree(c->ea_av);
    cudaFree(c->wt_v); cudaFree(c->wt_node); cudaFree(c->wt_pos);
    cudaFree(c->eo_ftau); cudaFree(c->eo_cmin);
    delete c;
    s->chain = nullptr;
}

rrrmc_status_t chain_sync_to_multispin(rrrmc_state *s)
{
    if (s->ms_valid) return RRRMC_OK;
    RR_TRY(launch_upload_transpose(s, 0, s->R));
    s->ms_valid = true;
    return RRRMC_OK;
}
rrrmc_status_t chain_sync_from_multispin(rrrmc_state *s)
{
    if (s->chain_valid) return RRRMC_OK;
    if (!s->d_chunks) RR_CUDA(cudaMalloc(&s->d_chunks, sizeof(uint64_t) * s->R * s->nchunks));
    RR_TRY(launch_download_transpose(s, 0, s->R));
    s->chain_valid = true;
    return RRRMC_OK;
}

static bool graph_f64_fields(const rrrmc_graph *g)
{
    return g->kind == RRRMC_EA_F64 || g->kind == RRRMC_SK_F64 || (g->kind == RRRMC_QUANT && (g->inner == RRRMC_SK_F64 || g->inner == RRRMC_EA_F64));
}
static bool graph_has_fields(const rrrmc_graph *g)
{
    return !(g->kind == RRRMC_QT || (g->kind == RRRMC_QUANT && g->inner == RRRMC_EMPTY));
}

// cache: 0 none, 1 discrete (ΔE classes), 2 continuous (Wong-Easton tree)
static rrrmc_status_t chain_ensure(rrrmc_state *s, int cache)
{
    rrrmc_graph *g = s->g;
    RR_ARG(g->kind == RRRMC_QT || g->kind == RRRMC_QUANT || is_sk_kind(g->kind) || g->twoD <= MAXDEG,
           "2D = %d exceeds the chain kernels' limit %d", g->twoD, MAXDEG);
    if (!s->chain) {
        chain_store *c = new chain_store();
        c->R = s->R; c->N = g->N; c->f64 = graph_f64_fields(g); c->nDE = (int)g->allDE.size();
        s->chain = c;
        const size_t RN = (size_t)s->R * g->N;
        if (graph_has_fields(g)) {
            if (c->f64 || g->kind == RRRMC_EA_DISCR) RR_CUDA(cudaMalloc(&c->lfd, RN * 2 * 8));
            if (!c->f64 || g->kind == RRRMC_EA_DISCR) RR_CUDA(cudaMalloc(&c->lfi, RN * 2 * 4));
        }
        RR_CUDA(cudaMalloc(&c->ml, sizeof(int32_t) * s->R * g->M));
        RR_CUDA(cudaMalloc(&c->sw, (size_t)s->R * g->M));
        RR_CUDA(cudaMalloc(&c->hdr, sizeof(chain_hdr) * s->R));
        RR_CUDA(cudaMemsetAsync(c->hdr, 0, sizeof(chain_hdr) * s->R, g->ctx->stream));
        RR_CUDA(cudaMalloc(&c->d_beta, 8 * s->R));
        RR_CUDA(cudaMalloc(&c->d_E, 8 * s->R));
        if (c->nDE) {
            RR_CUDA(cudaMalloc(&c->d_DE, 8 * c->nDE));
            RR_CUDA(cudaMemcpyAsync(c->d_DE, g->allDE.data(), 8 * c->nDE, cudaMemcpyHostToDevice, g->ctx->stream));
        }
    }
    chain_store *c = s->chain;
    const size_t RN = (size_t)s->R * g->N;
    if (cache == 1 && !c->disc_ready) {
        RR_ARG(c->nDE >= 1 && c->nDE <= MAXL, "|allΔE| = %d exceeds the discrete cache limit %d", c->nDE, MAXL);
        RR_CUDA(cudaMalloc(&c->av, RN * 4 * 2 * c->nDE));
        RR_CUDA(cudaMalloc(&c->apos, RN * 4));
        RR_CUDA(cudaMalloc(&c->cls, RN));
        c->disc_ready = true;
    }
    if (cache == 3 && !c->wtm_ready) {
        RR_CUDA(cudaMalloc(&c->wt_v, RN * 8));
        RR_CUDA(cudaMalloc(&c->wt_node, RN * 4));
        RR_CUDA(cudaMalloc(&c->wt_pos, RN * 4));
        c->wtm_ready = true;
    }
    if (cache == 2 && !c->cont_ready) {
        c->levs = 0; while (((int64_t)1 << c->levs) < g->N) c->levs++;
        c->N2 = (int64_t)1 << c->levs;
        RR_CUDA(cudaMalloc(&c->dEs, RN * 8));
        RR_CUDA(cudaMalloc(&c->dv, (size_t)s->R * (c->N2 + 1) * 8));
        RR_CUDA(cudaMalloc(&c->dps, (size_t)s->R * (c->N2 + 1) * 8));
        RR_CUDA(cudaMalloc(&c->csj, (size_t)s->R * (g->N + 1) * 4));
        RR_CUDA(cudaMalloc(&c->csdE, (size_t)s->R * (g->N + 1) * 8));
        RR_CUDA(cudaMalloc(&c->csp, (size_t)s->R * (g->N + 1) * 8));
        c->cont_ready = true;
    }
    return RRRMC_OK;
}
static rrrmc_status_t chain_aux(rrrmc_state *s, int64_t n, double **out)
{
    chain_store *c = s->chain;
    if (c->aux_len < n) { cudaFree(c->d_aux); c->d_aux = nullptr; RR_CUDA(cudaMalloc(&c->d_aux, 8 * n)); c->aux_len = n; }
    *out = c->d_aux;
    return RRRMC_OK;
}

static void chain_fill_params(rrrmc_state *s, chain_params &P)
{
    rrrmc_graph *g = s->g; chain_store *c = s->chain;
    memset(&P, 0, sizeof P);
    P.kind = g->kind; P.N = (int)g->N; P.twoD = g->twoD; P.nDE = c->nDE; P.levs = c->levs; P.N2 = c->N2;
    P.Nk = (int)g->Nk; P.M = (int)g->M; P.inner = g->inner; P.nz = g->nz_neighbors ? 1 : 0; P.fourK = g->fourK; P.sN = g->sN;
    P.fourK_r = g->kind == RRRMC_QUANT ? s->d_q_fourK : nullptr;
    P.R = s->R; P.nchunks = s->nchunks; P.chain0 = 0;
    P.A = g->d_A; P.J8 = g->d_J8; P.Jd = g->d_Jd; P.Jb = g->d_Jb;
    P.chunks = s->d_chunks;
    P.lfi = c->lfi; P.lfd = c->lfd; P.ml = c->ml; P.sw = c->sw;
    P.hdr = c->hdr; P.av = c->av; P.apos = c->apos; P.cls = c->cls;
    P.dEs = c->dEs; P.dv = c->dv; P.dps = c->dps; P.csj = c->csj; P.csdE = c->csdE; P.csp = c->csp;
    P.DE = c->d_DE; P.beta = c->d_beta;
    P.wt_v = c->wt_v; P.wt_node = c->wt_node; P.wt_pos = c->wt_pos;
    P.step = 1;
    // a warp per chain where an accepted flip costs O(N) (SK and GraphQuant over SK)
    P.coop = (is_sk_kind(g->kind) || (g->kind == RRRMC_QUANT && g->inner != RRRMC_EMPTY && g->inner != RRRMC_EA_F64)) ? 1 : 0;
}

static rrrmc_status_t chain_energy_init(rrrmc_state *s, chain_params &P, bool start_run = false)
{
    rrrmc_ctx *ctx = s->g->ctx;
    k_chain_lfields<<<div_up(P.R * P.N, 256), 256, 0, ctx->stream>>>(P);
    k_chain_energy_sum<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, s->chain->d_E, start_run ? 3 : 0);
    ctx->launches += 2;
    RR_CUDA(cudaGetLastError());
    s->chain_fields_valid = true;
    return RRRMC_OK;
}

rrrmc_status_t chain_energy(rrrmc_state *s, double *E_out)
{
    RR_TRY(chain_ensure(s, 0));
    RR_TRY(chain_sync_from_multispin(s));
    chain_params P; chain_fill_params(s, P);
    RR_TRY(chain_energy_init(s, P));
    RR_CUDA(cudaMemcpyAsync(E_out, s->chain->d_E, 8 * s->R, cudaMemcpyDeviceToHost, s->g->ctx->stream));
    RR_CUDA(cudaStreamSynchronize(s->g->ctx->stream));
    return RRRMC_OK;
}
// delta_energy / residual read the caches, like the reference: (re)build them when the configuration changed
static rrrmc_status_t chain_fields_current(rrrmc_state *s, chain_params &P)
{
    RR_TRY(chain_ensure(s, 0));
    RR_TRY(chain_sync_from_multispin(s));
    chain_fill_params(s, P);
    if (!s->chain_fields_valid) RR_TRY(chain_energy_init(s, P));
    return RRRMC_OK;
}
rrrmc_status_t chain_delta_energy_site(rrrmc_state *s, int64_t site0, int what, double *out)
{
    chain_params P; RR_TRY(chain_fields_current(s, P));
    rrrmc_ctx *ctx = s->g->ctx;
    k_chain_delta_site<<<div_up(P.R, 128), 128, 0, ctx->stream>>>(P, (int)site0, what, s->chain->d_E);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    RR_CUDA(cudaMemcpyAsync(out, s->chain->d_E, 8 * s->R, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    return RRRMC_OK;
}
rrrmc_status_t chain_delta_energy_replica(rrrmc_state *s, int64_t replica, double *out)
{
    chain_params P; RR_TRY(chain_fields_current(s, P));
    rrrmc_ctx *ctx = s->g->ctx;
    double *d_tmp; RR_TRY(chain_aux(s, P.N, &d_tmp));
    k_chain_delta_replica<<<div_up(P.N, 128), 128, 0, ctx->stream>>>(P, replica, d_tmp);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    RR_CUDA(cudaMemcpyAsync(out, d_tmp, 8 * P.N, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    return RRRMC_OK;
}
// GraphQuant observables (QT.jl:113-121, 201-268). what: 0 transverse_mag(β=arg), 1 Qenergy, 2 Renergies, 3 overlaps
rrrmc_status_t chain_quant_observable(rrrmc_state *s, int what, double arg, double *out)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_TRY(chain_ensure(s, 0));
    RR_TRY(chain_sync_from_multispin(s));
    chain_params P; chain_fill_params(s, P);
    const int64_t R = s->R, M = g->M, H = M / 2;
    if (what == 3) {
        double *d; RR_TRY(chain_aux(s, R * std::max<int64_t>(H, 1), &d));
        k_chain_overlaps<<<div_up(R, 64), 64, 0, ctx->stream>>>(P, d);
        ctx->launches++;
        RR_CUDA(cudaGetLastError());
        RR_CUDA(cudaMemcpyAsync(out, d, 8 * R * H, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
        return RRRMC_OK;
    }
    double *d; RR_TRY(chain_aux(s, R * (M + 1), &d));
    std::vector<double> e0(R), re;
    k_chain_energy_sum<<<div_up(R, 64), 64, 0, ctx->stream>>>(P, d, 2);
    ctx->launches++;
    RR_CUDA(cudaMemcpyAsync(e0.data(), d, 8 * R, cudaMemcpyDeviceToHost, ctx->stream));
    if (what != 0) { // slice energies need fresh fields (energy(X1[k], C1[k]) resets them, QT.jl:204-209, :262-265)
        RR_TRY(chain_energy_init(s, P));
        k_chain_energy_sum<<<div_up(R, 64), 64, 0, ctx->stream>>>(P, d + R, 1);
        ctx->launches++;
        re.resize(R * M);
        RR_CUDA(cudaMemcpyAsync(re.data(), d + R, 8 * R * M, cudaMemcpyDeviceToHost, ctx->stream));
    }
    RR_CUDA(cudaGetLastError());
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (what == 2) { memcpy(out, re.data(), 8 * R * M); return RRRMC_OK; }
    for (int64_t r = 0; r < R; r++) {
        const double beta = what == 0 ? arg : (s->q_beta.empty() ? g->beta : s->q_beta[r]);
        const double fourK = s->q_fourK.empty() ? g->fourK : s->q_fourK[r];
        const double p = -e0[r] / (double)g->N, x = beta * fourK / 2;
        const double tm = cosh(x) - p * sinh(x);                 // QT.jl:113-121
        if (what == 0) { out[r] = tm; continue; }
        double E = -g->Gamma * tm;                               // QT.jl:253-268
        for (int64_t k = 0; k < M; k++) E += re[r * M + k] / (double)g->N;
        out[r] = E;
    }
    return RRRMC_OK;
}

template <class SRC>
static rrrmc_status_t chain_drive(rrrmc_state *s, chain_params &P, rrrmc_hook_fn hook, void *user,
                                  double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_ctx *ctx = s->g->ctx; chain_store *c = s->chain;
    const int64_t total_rows = P.iters / P.step;
    const int64_t want_rows = Es ? std::min(Es_cap, total_rows) : 0;
    // device sample buffer: at most 32 MiB per launch
    int64_t rows_per_launch = hook ? 1 : std::max<int64_t>(1, std::min<int64_t>(std::max<int64_t>(total_rows, 1), ((int64_t)32 << 20) / (8 * P.R)));
    if (c->Es_rows < rows_per_launch) {
        cudaFree(c->d_Es);
        RR_CUDA(cudaMalloc(&c->d_Es, 8 * P.R * rows_per_launch));
        c->Es_rows = rows_per_launch;
    }
    P.Es = c->d_Es; P.Es_rows = rows_per_launch; P.quota = rows_per_launch;
    // spread chains: one per warp while they fit on the chip's schedulers
    const int64_t warps = (int64_t)ctx->sm_count * 16;
    P.cpw = P.coop ? 1 : (int)std::min<int64_t>(32, std::max<int64_t>(1, (P.R + warps - 1) / warps));
    const unsigned grid = div_up(P.R, P.cpw);
    std::vector<double> row((size_t)P.R * rows_per_launch);
    std::vector<chain_hdr> hh(P.R);
    std::vector<int64_t> acc(P.R);
    std::vector<double> emin(P.sampler == CHAIN_EO ? P.R : 0);
    const uint64_t l0 = ctx->launches;
    event_pair ev;                     // destroyed on every return path
    RR_CUDA(ev.create());
    cudaEvent_t e0 = ev.e0, e1 = ev.e1;
    RR_CUDA(cudaEventRecord(e0, ctx->stream));
    int64_t nsamples = 0; bool stop = false;
    for (int guard = 0; !stop; guard++) {
        if (P.fast == 2) RR_TRY(chain_warp_launch(s, P));
        else if (P.fast) RR_TRY(chain_ea_launch(s, P));
        else if constexpr (std::is_same<SRC, src_trace>::value) chain_launch_run_trace(P, grid, ctx->stream);
        else k_chain_run<SRC><<<grid, 32, 0, ctx->stream>>>(P);
        ctx->launches++;
        s->ms_valid = false; s->chain_valid = true; // the chains own the configuration (a hook may have re-synced)
        RR_CUDA(cudaGetLastError());
        RR_CUDA(cudaMemcpyAsync(hh.data(), c->hdr + P.chain0, sizeof(chain_hdr) * P.R, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaMemcpyAsync(row.data(), c->d_Es, 8 * P.R * rows_per_launch, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
        bool all_done = true;
        for (int64_t r = 0; r < P.R; r++) {
            if (hh[r].status) { rrrmc_set_error("chain %lld: draw source error %d (trace exhausted/mismatched or sampler precision loss)", (long long)r, hh[r].status); return RRRMC_ERR_STATE; }
            all_done &= hh[r].done != 0;
        }
        // rows emitted by this launch: every chain emits the same number
        int64_t emitted = all_done ? std::min<int64_t>(rows_per_launch, total_rows - nsamples) : rows_per_launch;
        if (emitted < 0) emitted = 0;
        for (int64_t k = 0; k < emitted && !stop; k++) {
            if (nsamples < want_rows) memcpy(Es + nsamples * P.R, row.data() + k * P.R, 8 * P.R);
            nsamples++;
            if (hook) {
                if (P.sampler == CHAIN_EO) {   // hook(it, X, C, E, Emin), RRRMC.jl:499
                    for (int64_t r = 0; r < P.R; r++) emin[r] = hh[r].Emin;
                    if (!reinterpret_cast<rrrmc_eo_hook_fn>(hook)(user, nsamples * P.step, row.data() + k * P.R, emin.data(), P.R)) stop = true;
                    continue;
                }
                for (int64_t r = 0; r < P.R; r++) acc[r] = hh[r].accepted;
                if (!hook(user, nsamples * P.step, row.data() + k * P.R, acc.data(), P.R)) stop = true;
            }
        }
        if (all_done) break;
    }
    RR_CUDA(cudaEventRecord(e1, ctx->stream));
    RR_CUDA(cudaEventSynchronize(e1));
    float ms = 0; RR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (info) {
        int64_t itmax = 0, accs = 0;
        for (int64_t r = 0; r < P.R; r++) { itmax = std::max<int64_t>(itmax, hh[r].it); accs += hh[r].accepted; }
        info->nsamples = std::min(nsamples, want_rows); info->iters_done = itmax;
        info->launches = (int64_t)(ctx->launches - l0); info->device_ms = ms; info->accepted_total = accs;
    }
    return RRRMC_OK;
}

// which cache a sampler builds for this graph: rrrMC on inner_graph(X), bklMC on X (DeltaE.jl:116, RRRMC.jl:171,240,325)
static rrrmc_status_t sampler_cache(const rrrmc_graph *g, int sampler, int *cache)
{
    const bool discr_full = g->kind == RRRMC_EA_PM1 || g->kind == RRRMC_EA_INT || g->kind == RRRMC_QT;
    if (sampler == CHAIN_STANDARD) *cache = 0;
    else if (sampler == CHAIN_RRR) *cache = (discr_full || g->kind == RRRMC_QUANT || g->kind == RRRMC_EA_DISCR) ? 1 : 2;
    else *cache = discr_full ? 1 : 2;
    return RRRMC_OK;
}
static double default_staged_thr(const rrrmc_graph *g)
{
    const bool simple = g->kind == RRRMC_EA_F64 || is_sk_kind(g->kind);
    return simple ? 0.8 : 0.5; // RRRMC.jl:163-165 (SimpleGraph 0.8, else 0.5), :226 (DoubleGraph 0.5)
}

// replay mode: one chain fed a dumped typed draw stream instead of the Philox counter stream
static rrrmc_status_t chain_use_trace(rrrmc_state *s, chain_params &P, const chain_trace_in &tr)
{
    rrrmc_ctx *ctx = s->g->ctx; chain_store *c = s->chain;
    RR_ARG(tr.kind && tr.ival && tr.fval && tr.n >= 0, "replay: NULL draw trace");
    RR_ARG(tr.replica >= 0 && tr.replica < s->R, "replica out of range");
    if (c->tcap < tr.n) {
        cudaFree(c->d_tkind); cudaFree(c->d_tival); cudaFree(c->d_tfval);
        c->tcap = std::max<int64_t>(tr.n, 1);
        RR_CUDA(cudaMalloc(&c->d_tkind, c->tcap)); RR_CUDA(cudaMalloc(&c->d_tival, 8 * c->tcap)); RR_CUDA(cudaMalloc(&c->d_tfval, 8 * c->tcap));
    }
    RR_CUDA(cudaMemcpyAsync(c->d_tkind, tr.kind, tr.n, cudaMemcpyHostToDevice, ctx->stream));
    RR_CUDA(cudaMemcpyAsync(c->d_tival, tr.ival, 8 * tr.n, cudaMemcpyHostToDevice, ctx->stream));
    RR_CUDA(cudaMemcpyAsync(c->d_tfval, tr.fval, 8 * tr.n, cudaMemcpyHostToDevice, ctx->stream));
    P.tkind = c->d_tkind; P.tival = c->d_tival; P.tfval = c->d_tfval; P.tlen = tr.n;
    P.chain0 = tr.replica; P.R = 1;       // run only the requested chain
    return RRRMC_OK;
}

rrrmc_status_t chain_run(rrrmc_state *s, int sampler, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                         rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(beta, "beta is NULL");
    for (int64_t r = 0; r < s->R; r++) RR_ARG(std::isfinite(beta[r]), "β must be finite, given: %g", beta[r]); // RRRMC.jl:159
    int cache; RR_TRY(sampler_cache(g, sampler, &cache));
    RR_TRY(chain_ensure(s, cache));
    RR_TRY(chain_sync_from_multispin(s));
    chain_store *c = s->chain;
    chain_params P; chain_fill_params(s, P);
    P.sampler = sampler; P.iters = iters; P.step = step; P.seed = seed;
    P.staged_thr = std::isnan(o->staged_thr) ? default_staged_thr(g) : o->staged_thr;
    P.staged_thr_fact = o->staged_thr_fact;
    RR_CUDA(cudaMemcpyAsync(c->d_beta, beta, 8 * s->R, cudaMemcpyHostToDevice, ctx->stream));
    k_chain_hdr_reset<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, seed == 0);
    ctx->launches++;
    RR_TRY(chain_energy_init(s, P, true));
    s->ms_valid = false; // chains now own the configuration
    sk_dense_invalidate(s);
    if (o->site_pick == RRRMC_PICK_RANK) {
        if (!chain_warp_eligible(s, sampler)) {
            rrrmc_set_error("site_pick = RANK (the warp-cooperative kernel) takes rrrMC / bklMC on ±J graphs — GraphEA lattices with L >= 3, "
                            "GraphRRG with 2..6 distinct neighbours and no zero coupling — whose chain state fits shared memory (N <= ~110000)");
            return RRRMC_ERR_UNSUPPORTED;
        }
        P.fast = 2; P.jcode = g->d_jcode; P.latL = g->L;
        s->chain_fields_valid = false;
    } else {
        RR_ARG(o->site_pick == RRRMC_PICK_REFERENCE, "unknown site_pick %d", o->site_pick);
        if (chain_ea_eligible(s, sampler)) { RR_TRY(chain_ea_prepare(s, P)); s->chain_fields_valid = false; }
    }
    RR_TRY(chain_drive<src_philox>(s, P, hook, user, Es, Es_cap, info));
    return RRRMC_OK;
}

rrrmc_status_t chain_run_wtm(rrrmc_state *s, const double *beta, int64_t samples, double step, uint64_t seed,
                             rrrmc_hook_fn hook, void *user, double *Es, int64_t Es_cap, rrrmc_run_info_t *info,
                             const chain_trace_in *tr)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(beta, "beta is NULL");
    for (int64_t r = 0; r < s->R; r++) RR_ARG(std::isfinite(beta[r]), "β must be finite, given: %g", beta[r]);
    RR_ARG(samples >= 0, "samples must be >= 0, given %lld", (long long)samples);
    RR_ARG(std::isfinite(step) && step > 0, "step must be a positive global-time interval, given %g", step);
    RR_TRY(chain_ensure(s, 3));
    RR_TRY(chain_sync_from_multispin(s));
    chain_store *c = s->chain;
    chain_params P; chain_fill_params(s, P);
    P.sampler = CHAIN_WTM; P.iters = samples; P.step = 1; P.seed = seed;   // one row of Es per sample
    P.wt_step = step / (double)g->N;                                        // step /= N, RRRMC.jl:394
    P.wt_tmax = P.wt_step * (double)samples;                                // tmax = step * samples, :397
    RR_CUDA(cudaMemcpyAsync(c->d_beta, beta, 8 * s->R, cudaMemcpyHostToDevice, ctx->stream));
    k_chain_hdr_reset<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, seed == 0);
    ctx->launches++;
    RR_TRY(chain_energy_init(s, P, true));
    s->ms_valid = false;
    sk_dense_invalidate(s);
    if (samples == 0) { if (info) { info->nsamples = 0; info->iters_done = 0; info->launches = 0; info->device_ms = 0; info->accepted_total = 0; } return RRRMC_OK; }
    if (tr) { RR_TRY(chain_use_trace(s, P, *tr)); RR_TRY(chain_drive<src_trace>(s, P, nullptr, nullptr, Es, Es_cap, info)); return RRRMC_OK; }
    RR_TRY(chain_drive<src_philox>(s, P, hook, user, Es, Es_cap, info));
    return RRRMC_OK;
}

rrrmc_status_t chain_run_eo(rrrmc_state *s, const double *ftau, int64_t ftau_stride, int64_t iters, int64_t step, uint64_t seed,
                            rrrmc_eo_hook_fn hook, void *user, double *Emin_out, int64_t *itmin_out, uint64_t *Cmin_chunks,
                            double *Es, int64_t Es_cap, rrrmc_run_info_t *info, const chain_trace_in *tr)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    const bool discr_full = g->kind == RRRMC_EA_PM1 || g->kind == RRRMC_EA_INT || g->kind == RRRMC_QT;
    const bool simple_f64 = g->kind == RRRMC_EA_F64 || g->kind == RRRMC_SK_F64;   // EOCacheCont (DeltaE.jl:555-635)
    if (!discr_full && !simple_f64) {
        rrrmc_set_error("extremal_opt: DiscrGraph models (GraphEA / GraphRRG with integer levels, GraphQT) and the Float64 SimpleGraphs "
                        "(GraphEANormal, GraphRRGNormal, GraphSKNormal) are on this path");
        return RRRMC_ERR_UNSUPPORTED;
    }
    RR_ARG(ftau, "ftau is NULL");
    RR_ARG(ftau_stride == 0 || ftau_stride >= g->N, "ftau_stride must be 0 (one table for all chains) or >= N, given %lld", (long long)ftau_stride);
    const int64_t ntab = ftau_stride ? s->R : 1, tstride = ftau_stride ? ftau_stride : 0;
    for (int64_t q = 0; q < ntab; q++) {
        const double *f = ftau + q * tstride;
        RR_ARG(std::isfinite(f[g->N - 1]) && f[0] > 0, "ftau must hold the positive cumulative sums of j^-tau");
        for (int64_t j = 1; j < g->N; j++) RR_ARG(f[j] >= f[j - 1], "ftau must be non-decreasing (cumsum of j^-tau)");
    }
    RR_TRY(chain_ensure(s, discr_full ? 1 : 2));
    RR_TRY(chain_sync_from_multispin(s));
    chain_store *c = s->chain;
    if (c->eo_ftau_len < ntab * g->N) {
        cudaFree(c->eo_ftau); c->eo_ftau = nullptr;
        RR_CUDA(cudaMalloc(&c->eo_ftau, 8 * ntab * g->N)); c->eo_ftau_len = ntab * g->N;
    }
    if (!c->eo_cmin) RR_CUDA(cudaMalloc(&c->eo_cmin, 8 * s->R * s->nchunks));
    if (tstride == 0 || tstride == g->N) RR_CUDA(cudaMemcpyAsync(c->eo_ftau, ftau, 8 * ntab * g->N, cudaMemcpyHostToDevice, ctx->stream));
    else RR_CUDA(cudaMemcpy2DAsync(c->eo_ftau, 8 * g->N, ftau, 8 * tstride, 8 * g->N, ntab, cudaMemcpyHostToDevice, ctx->stream));
    chain_params P; chain_fill_params(s, P);
    P.sampler = CHAIN_EO; P.iters = iters; P.step = step; P.seed = seed;
    P.eo_ftau = c->eo_ftau; P.eo_stride = ftau_stride ? g->N : 0; P.eo_cmin = c->eo_cmin;
    k_chain_hdr_reset<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, seed == 0);
    ctx->launches++;
    RR_TRY(chain_energy_init(s, P, true));
    s->ms_valid = false;
    sk_dense_invalidate(s);
    if (!discr_full) {
        // EOCacheCont ctor (DeltaE.jl:560-568): ΔEs = [delta_energy(X, C, i) for i = 1:N] on the device, rank = sortperm(ΔEs)
        // on the host (once per run; the kernel keeps the order with an insertion pass per move)
        const int64_t R = s->R, N = g->N;
        k_chain_delta_all<<<dim3(div_up(N, 128), (unsigned)R), 128, 0, ctx->stream>>>(P, c->dEs);
        ctx->launches++;
        RR_CUDA(cudaGetLastError());
        std::vector<double> hd((size_t)R * N);
        RR_CUDA(cudaMemcpyAsync(hd.data(), c->dEs, 8 * (size_t)R * N, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
        std::vector<int32_t> hr((size_t)R * (N + 1), 0);
        for (int64_t r = 0; r < R; r++) {
            int32_t *rk = hr.data() + r * (N + 1);
            const double *v = hd.data() + r * N;
            for (int64_t i = 0; i < N; i++) rk[i] = (int32_t)i;
            std::stable_sort(rk, rk + N, [v](int32_t a, int32_t b) { return v[a] < v[b]; });
        }
        RR_CUDA(cudaMemcpyAsync(c->csj, hr.data(), 4 * (size_t)R * (N + 1), cudaMemcpyHostToDevice, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (tr) { RR_TRY(chain_use_trace(s, P, *tr)); RR_TRY(chain_drive<src_trace>(s, P, nullptr, nullptr, Es, Es_cap, info)); }
    else RR_TRY(chain_drive<src_philox>(s, P, reinterpret_cast<rrrmc_hook_fn>(hook), user, Es, Es_cap, info));
    std::vector<chain_hdr> hh(s->R);
    RR_CUDA(cudaMemcpyAsync(hh.data(), c->hdr, sizeof(chain_hdr) * s->R, cudaMemcpyDeviceToHost, ctx->stream));
    if (Cmin_chunks) RR_CUDA(cudaMemcpyAsync(Cmin_chunks, c->eo_cmin, 8 * s->R * s->nchunks, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t r = 0; r < s->R; r++) {
        if (Emin_out) Emin_out[r] = hh[r].Emin;
        if (itmin_out) itmin_out[r] = hh[r].itmin;
    }
    return RRRMC_OK;
}

rrrmc_status_t chain_replay(rrrmc_state *s, int64_t replica, int sampler, double beta, int64_t iters, int64_t step,
                            const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                            const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(std::isfinite(beta), "β must be finite, given: %g", beta);
    int cache; RR_TRY(sampler_cache(g, sampler, &cache));
    RR_TRY(chain_ensure(s, cache));
    RR_TRY(chain_sync_from_multispin(s));
    chain_store *c = s->chain;
    chain_params P; chain_fill_params(s, P);
    P.sampler = sampler; P.iters = iters; P.step = step; P.seed = 0;
    P.staged_thr = std::isnan(o->staged_thr) ? default_staged_thr(g) : o->staged_thr;
    P.staged_thr_fact = o->staged_thr_fact;
    std::vector<double> b(s->R, beta);
    RR_CUDA(cudaMemcpyAsync(c->d_beta, b.data(), 8 * s->R, cudaMemcpyHostToDevice, ctx->stream));
    k_chain_hdr_reset<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, 0);
    ctx->launches++;
    RR_TRY(chain_energy_init(s, P, true));
    s->ms_valid = false;
    sk_dense_invalidate(s);
    const chain_trace_in tr{ replica, kind, ival, fval, ndraws };
    RR_TRY(chain_use_trace(s, P, tr));
    P.beta = c->d_beta; // all equal
    RR_TRY(chain_drive<src_trace>(s, P, nullptr, nullptr, Es, Es_cap, info));
    return RRRMC_OK;
}
