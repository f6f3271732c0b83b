// C-ABI entry points (include/rrrmc_b200.h): handle management, host-side graph logic, drivers.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include "common.cuh"
#include "kernels.cuh"
#include "cb_params.cuh"
#include "chain.cuh"
#include "ea_tma.cuh"
#include "ea_normal.cuh"

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void rrrmc_set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
extern "C" const char *rrrmc_last_error(void) { return g_err; }
extern "C" const char *rrrmc_version(void) { return "rrrmc_b200 0.1 (sm_100a)"; }

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" rrrmc_status_t rrrmc_ctx_create(int device, void *cuda_stream, rrrmc_ctx_t **out)
{
    RR_ARG(out != nullptr, "rrrmc_ctx_create: out is NULL");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        rrrmc_set_error("rrrmc_ctx_create: no CUDA device (%s); this engine has no CPU fallback", cudaGetErrorString(e));
        return RRRMC_ERR_CUDA;
    }
    RR_ARG(device >= 0 && device < ndev, "rrrmc_ctx_create: device %d out of range (0..%d)", device, ndev - 1);
    RR_CUDA(cudaSetDevice(device));
    rrrmc_ctx *c = new rrrmc_ctx();
    c->device = device;
    if (cuda_stream) c->stream = (cudaStream_t)cuda_stream;
    else { RR_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)); c->own_stream = true; }
    RR_CUDA(cudaEventCreate(&c->ev0));
    RR_CUDA(cudaEventCreate(&c->ev1));
    RR_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    *out = c;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_ctx_destroy(rrrmc_ctx_t *c)
{
    if (!c) return RRRMC_OK;
    cudaSetDevice(c->device);
    if (c->flush_buf) cudaFree(c->flush_buf);
    if (c->d_cbp_bucket) cudaFree(c->d_cbp_bucket);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->stream);
    delete c;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_ctx_sync(rrrmc_ctx_t *c)
{
    RR_ARG(c, "ctx is NULL");
    RR_CUDA(cudaStreamSynchronize(c->stream));
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_ctx_timer_start(rrrmc_ctx_t *c)
{
    RR_ARG(c, "ctx is NULL");
    RR_CUDA(cudaEventRecord(c->ev0, c->stream));
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_ctx_timer_stop(rrrmc_ctx_t *c, float *ms)
{
    RR_ARG(c && ms, "ctx/ms is NULL");
    RR_CUDA(cudaEventRecord(c->ev1, c->stream));
    RR_CUDA(cudaEventSynchronize(c->ev1));
    RR_CUDA(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_ctx_launch_count(rrrmc_ctx_t *c, uint64_t *count)
{
    RR_ARG(c && count, "ctx/count is NULL");
    *count = c->launches;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_ctx_flush_l2(rrrmc_ctx_t *c)
{
    RR_ARG(c, "ctx is NULL");
    RR_CUDA(cudaSetDevice(c->device));
    if (!c->flush_buf) {
        c->flush_bytes = (size_t)256 << 20; // 2x the 126 MB L2
        RR_CUDA(cudaMalloc(&c->flush_buf, c->flush_bytes));
    }
    return launch_flush(c);
}

// ------------------------------------------------------------------------------------------------
// EA lattice host logic
// ------------------------------------------------------------------------------------------------
namespace {
struct nb_slot { int64_t site; int code; }; // code = 2d (own forward bond) or 2d+1 (bond owned by the i-e_d neighbour)

// neighbours of 0-based site i in the reference's slot order: ascending index (EA.jl:40); for L=2 the
// two bonds to the same neighbour are ordered "lower site's forward bond first" (gen_J fill order, EA.jl:50-59)
void lattice_slots(int L, int D, int64_t i, nb_slot *out)
{
    int64_t stride = 1, rem = i;
    for (int d = 0; d < D; d++) {
        const int64_t c = rem % L; rem /= L;
        const int64_t up = i + ((c + 1 == L ? 0 : c + 1) - c) * stride;
        const int64_t dn = i + ((c == 0 ? L - 1 : c - 1) - c) * stride;
        out[2 * d] = { up, 2 * d };
        out[2 * d + 1] = { dn, 2 * d + 1 };
        stride *= L;
    }
    std::stable_sort(out, out + 2 * D, [i](const nb_slot &a, const nb_slot &b) {
        if (a.site != b.site) return a.site < b.site;
        const bool a_first = (a.code & 1) == (i < a.site ? 0 : 1); // i<nbr: own forward first; else the other's
        const bool b_first = (b.code & 1) == (i < b.site ? 0 : 1);
        return a_first && !b_first;
    });
}
int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }
} // namespace

extern "C" rrrmc_status_t rrrmc_gen_ea_adjacency(int L, int D, int64_t *A_out)
{
    RR_ARG(L >= 2, "L must be >= 2, given: %d", L);   // EA.jl:25
    RR_ARG(D >= 1 && D <= 8, "D must be in 1..8, given: %d", D);
    RR_ARG(A_out, "A_out is NULL");
    const int64_t N = ipow(L, D);
    nb_slot sl[16];
    for (int64_t i = 0; i < N; i++) {
        lattice_slots(L, D, i, sl);
        for (int k = 0; k < 2 * D; k++) A_out[i * 2 * D + k] = sl[k].site + 1;
    }
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_graph_ea_create(rrrmc_ctx_t *ctx, int L, int D, int kind,
                                                const int64_t *A, const void *J, rrrmc_graph_t **out)
{
    RR_ARG(ctx && A && J && out, "rrrmc_graph_ea_create: NULL argument");
    RR_ARG(L >= 2, "L must be >= 2, given: %d", L);
    RR_ARG(D >= 1 && D <= 4, "D must be in 1..4, given: %d", D);
    RR_ARG(kind == RRRMC_EA_PM1 || kind == RRRMC_EA_INT || kind == RRRMC_EA_F64, "unknown coupling kind %d", kind);
    const int twoD = 2 * D;
    const int64_t N = ipow(L, D);
    RR_ARG(N >= 2 && N < ((int64_t)1 << 31), "N = L^D = %lld out of range", (long long)N);
    rrrmc_graph *g = new rrrmc_graph();
    g->ctx = ctx; g->kind = kind; g->L = L; g->D = D; g->twoD = twoD; g->N = N;
    g->Nk = N; g->M = 1; g->max_deg = twoD;
    g->bipartite = (L % 2 == 0);
    g->A0.resize(N * twoD);
    std::vector<int8_t> code(N * twoD);
    nb_slot sl[16];
    for (int64_t i = 0; i < N; i++) {
        lattice_slots(L, D, i, sl);
        for (int k = 0; k < twoD; k++) {
            if (A[i * twoD + k] != sl[k].site + 1) {
                rrrmc_set_error("invalid A, does not look like an EA graph: A[%lld][%d] = %lld, expected %lld",
                                (long long)i + 1, k + 1, (long long)A[i * twoD + k], (long long)sl[k].site + 1);
                delete g; return RRRMC_ERR_ARG;
            }
            g->A0[i * twoD + k] = (int32_t)sl[k].site;
            code[i * twoD + k] = (int8_t)sl[k].code;
        }
    }
    // unique neighbours (EA.jl:158, :548)
    g->uA0.resize(N * twoD); g->nuA.resize(N);
    for (int64_t i = 0; i < N; i++) {
        int n = 0;
        for (int k = 0; k < twoD; k++) {
            const int32_t y = g->A0[i * twoD + k];
            if (n == 0 || g->uA0[i * twoD + n - 1] != y) g->uA0[i * twoD + n++] = y;
        }
        g->nuA[i] = n;
    }
    // couplings, slot-aligned with A; symmetry check through the (site, bond) codes
    if (kind == RRRMC_EA_F64) g->Jd.assign((const double *)J, (const double *)J + N * twoD);
    else g->Ji.assign((const int64_t *)J, (const int64_t *)J + N * twoD);
    auto Jat = [&](int64_t idx) { return kind == RRRMC_EA_F64 ? g->Jd[idx] : (double)g->Ji[idx]; };
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < twoD; k++) {
            const int c = code[i * twoD + k];
            if (c & 1) continue;                       // own forward bond (i -> up); find it at the other end
            const int64_t up = g->A0[i * twoD + k];
            int l = -1;
            for (int m = 0; m < twoD; m++) if (g->A0[up * twoD + m] == i && code[up * twoD + m] == (c | 1)) l = m;
            if (l < 0 || Jat(i * twoD + k) != Jat(up * twoD + l)) {
                rrrmc_set_error("J is not symmetric at bond (%lld,%lld)", (long long)i + 1, (long long)up + 1);
                delete g; return RRRMC_ERR_ARG;
            }
        }
    std::set<int64_t> levels;
    if (kind != RRRMC_EA_F64) {
        for (int64_t v : g->Ji) {
            if (kind == RRRMC_EA_PM1 && v != 1 && v != -1) {
                rrrmc_set_error("the given J is incompatible with levels (-1, 1): found %lld", (long long)v);
                delete g; return RRRMC_ERR_ARG;
            }
            if (v < -127 || v > 127) { rrrmc_set_error("integer couplings must fit int8, found %lld", (long long)v); delete g; return RRRMC_ERR_ARG; }
            levels.insert(v);
        }
        // allΔE (EA.jl:293-309): all sums of 2D signed levels
        std::set<int64_t> es = { 0 };
        if (kind == RRRMC_EA_PM1) levels = { -1, 1 };
        for (int n = 0; n < twoD; n++) {
            std::set<int64_t> nw;
            for (int64_t e : es) for (int64_t l : levels) { nw.insert(e + l); nw.insert(e - l); }
            es.swap(nw);
        }
        std::set<int64_t> de;
        for (int64_t e : es) de.insert(2 * (e < 0 ? -e : e));
        for (int64_t d : de) g->allDE.push_back((double)d);
        if (g->allDE.size() > 64) { rrrmc_set_error("too many ΔE classes (%zu > 64)", g->allDE.size()); delete g; return RRRMC_ERR_UNSUPPORTED; }
    }
    // device copies
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_CUDA(cudaMalloc(&g->d_A, sizeof(int32_t) * N * twoD));
    RR_CUDA(cudaMemcpy(g->d_A, g->A0.data(), sizeof(int32_t) * N * twoD, cudaMemcpyHostToDevice));
    if (kind == RRRMC_EA_F64) {
        RR_CUDA(cudaMalloc(&g->d_Jd, sizeof(double) * N * twoD));
        RR_CUDA(cudaMemcpy(g->d_Jd, g->Jd.data(), sizeof(double) * N * twoD, cudaMemcpyHostToDevice));
    } else {
        std::vector<int8_t> j8(N * twoD);
        for (int64_t k = 0; k < N * twoD; k++) j8[k] = (int8_t)g->Ji[k];
        RR_CUDA(cudaMalloc(&g->d_J8, N * twoD));
        RR_CUDA(cudaMemcpy(g->d_J8, j8.data(), N * twoD, cudaMemcpyHostToDevice));
    }
    if (kind == RRRMC_EA_PM1 && D <= 3) {
        std::vector<uint8_t> jc(N, 0);
        for (int64_t i = 0; i < N; i++)
            for (int k = 0; k < twoD; k++)
                if (g->Ji[i * twoD + k] < 0) jc[i] |= (uint8_t)(1u << code[i * twoD + k]);
        RR_CUDA(cudaMalloc(&g->d_jcode, N));
        RR_CUDA(cudaMemcpy(g->d_jcode, jc.data(), N, cudaMemcpyHostToDevice));
        // the same signs expanded to whole-word masks [N][8] (two unused): the sparse checkerboard kernel is bound
        // by its ALU pipe, so it loads the six masks (32 B per site, shared by all replicas) instead of expanding bits
        std::vector<uint32_t> jm(N * 8, 0u);
        for (int64_t i = 0; i < N; i++)
            for (int k = 0; k < 6; k++) jm[i * 8 + k] = ((jc[i] >> k) & 1u) ? 0xffffffffu : 0u;
        RR_CUDA(cudaMalloc(&g->d_jmask, sizeof(uint32_t) * N * 8));
        RR_CUDA(cudaMemcpy(g->d_jmask, jm.data(), sizeof(uint32_t) * N * 8, cudaMemcpyHostToDevice));
    }
    RR_CUDA(cudaDeviceSynchronize()); // default-stream uploads are complete before any non-blocking stream uses them
    *out = g;
    return RRRMC_OK;
}

// GraphRRG{Int,LEV,K}(A, J) (RRG.jl:112-137) / GraphRRGNormal (RRG.jl, continuous couplings): a K-regular graph with an
// explicit adjacency (the reference draws it with the Bollobás pairing model, gen_RRG RRG.jl:27-68). energy,
// update_cache!, delta_energy and neighbors (RRG.jl:165-250) are those of GraphEA on a general adjacency when all
// neighbours of a site are distinct, so the graph runs on the chain engine's EA kinds; the one difference is that
// neighbors() lists only the entries with a non-zero coupling (uA, RRG.jl:133), which the engine applies as a filter.
extern "C" rrrmc_status_t rrrmc_graph_rrg_create(rrrmc_ctx_t *ctx, int64_t N, int K, int kind,
                                                 const int64_t *A, const void *J, rrrmc_graph_t **out)
{
    RR_ARG(ctx && A && J && out, "rrrmc_graph_rrg_create: NULL argument");
    RR_ARG(kind == RRRMC_EA_PM1 || kind == RRRMC_EA_INT || kind == RRRMC_EA_F64, "unknown coupling kind %d", kind);
    RR_ARG(K >= 1 && K <= 8, "K must be in 1..8 on this engine, given: %d", K);
    RR_ARG(N >= 2 && N < ((int64_t)1 << 31) && (N * K) % 2 == 0, "N * K must be even and N in range, given N=%lld, K=%d", (long long)N, K); // RRG.jl:29
    rrrmc_graph *g = new rrrmc_graph();
    g->ctx = ctx; g->kind = kind; g->L = 0; g->D = 0; g->twoD = K; g->N = N; g->Nk = N; g->M = 1; g->max_deg = K;
    g->bipartite = false;
    g->A0.resize(N * K); g->uA0.resize(N * K); g->nuA.assign(N, K);
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < K; k++) {
            const int64_t y = A[i * K + k];
            if (y < 1 || y > N || y == i + 1 || (k > 0 && y <= A[i * K + k - 1])) {
                rrrmc_set_error("invalid A: row %lld must hold K distinct neighbours in ascending order, none equal to the site", (long long)i + 1);
                delete g; return RRRMC_ERR_ARG;
            }
            g->A0[i * K + k] = (int32_t)(y - 1); g->uA0[i * K + k] = (int32_t)(y - 1);
        }
    if (kind == RRRMC_EA_F64) g->Jd.assign((const double *)J, (const double *)J + N * K);
    else g->Ji.assign((const int64_t *)J, (const int64_t *)J + N * K);
    auto Jat = [&](int64_t idx) { return kind == RRRMC_EA_F64 ? g->Jd[idx] : (double)g->Ji[idx]; };
    for (int64_t i = 0; i < N; i++)
        for (int k = 0; k < K; k++) {
            const int64_t y = g->A0[i * K + k];
            int l = -1;
            for (int m = 0; m < K; m++) if (g->A0[y * K + m] == i) l = m;
            if (l < 0 || Jat(i * K + k) != Jat(y * K + l)) {
                rrrmc_set_error("A / J are not symmetric at bond (%lld,%lld)", (long long)i + 1, (long long)y + 1);
                delete g; return RRRMC_ERR_ARG;
            }
        }
    if (kind != RRRMC_EA_F64) {
        std::set<int64_t> levels;
        for (int64_t v : g->Ji) {
            if ((kind == RRRMC_EA_PM1 && v != 1 && v != -1) || v < -127 || v > 127) {
                rrrmc_set_error("the given J is incompatible with the levels of this kind (int8; ±1 for PM1): found %lld", (long long)v);
                delete g; return RRRMC_ERR_ARG;
            }
            levels.insert(v);
        }
        if (kind == RRRMC_EA_PM1) levels = { -1, 1 };
        if (levels.count(0)) {                             // neighbors() = the entries with a non-zero coupling (RRG.jl:133)
            g->nz_neighbors = true;
            for (int64_t i = 0; i < N; i++) {
                int n = 0;
                for (int k = 0; k < K; k++) if (g->Ji[i * K + k] != 0) g->uA0[i * K + n++] = g->A0[i * K + k];
                g->nuA[i] = n;
            }
        }
        std::set<int64_t> es = { 0 };                      // allΔE, RRG.jl:252-270: sums of K signed levels
        for (int n = 0; n < K; n++) {
            std::set<int64_t> nw;
            for (int64_t e : es) for (int64_t l : levels) { nw.insert(e + l); nw.insert(e - l); }
            es.swap(nw);
        }
        std::set<int64_t> de;
        for (int64_t e : es) de.insert(2 * (e < 0 ? -e : e));
        for (int64_t d : de) g->allDE.push_back((double)d);
        if (g->allDE.size() > 64) { rrrmc_set_error("too many ΔE classes (%zu > 64)", g->allDE.size()); delete g; return RRRMC_ERR_UNSUPPORTED; }
    }
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_CUDA(cudaMalloc(&g->d_A, sizeof(int32_t) * N * K));
    RR_CUDA(cudaMemcpy(g->d_A, g->A0.data(), sizeof(int32_t) * N * K, cudaMemcpyHostToDevice));
    if (kind == RRRMC_EA_F64) {
        RR_CUDA(cudaMalloc(&g->d_Jd, sizeof(double) * N * K));
        RR_CUDA(cudaMemcpy(g->d_Jd, g->Jd.data(), sizeof(double) * N * K, cudaMemcpyHostToDevice));
    } else {
        std::vector<int8_t> j8(N * K);
        for (int64_t k = 0; k < N * K; k++) j8[k] = (int8_t)g->Ji[k];
        RR_CUDA(cudaMalloc(&g->d_J8, N * K));
        RR_CUDA(cudaMemcpy(g->d_J8, j8.data(), N * K, cudaMemcpyHostToDevice));
    }
    RR_CUDA(cudaDeviceSynchronize());
    *out = g;
    return RRRMC_OK;
}

// ------------------------------------------------------------------------------------------------
// SK / QT / GraphQuant host logic
// ------------------------------------------------------------------------------------------------
static rrrmc_status_t upload_sk_couplings(rrrmc_graph *g, int64_t n, int kind, const void *J)
{
    // validation as in the constructors (SK.jl:32-46, :185-196): square, symmetric, zero diagonal
    if (kind == RRRMC_SK_F64) {
        const double *Jd = (const double *)J;
        for (int64_t i = 0; i < n; i++) {
            RR_ARG(Jd[i * n + i] == 0, "invalid J: diagonal entry J[%lld][%lld] = %g, expected 0", (long long)i + 1, (long long)i + 1, Jd[i * n + i]);
            for (int64_t j = i + 1; j < n; j++)
                RR_ARG(Jd[i * n + j] == Jd[j * n + i], "invalid J: not symmetric at (%lld,%lld)", (long long)i + 1, (long long)j + 1);
        }
        RR_CUDA(cudaMalloc(&g->d_Jd, sizeof(double) * n * n));
        RR_CUDA(cudaMemcpy(g->d_Jd, Jd, sizeof(double) * n * n, cudaMemcpyHostToDevice));
    } else {
        const uint8_t *Jb = (const uint8_t *)J;
        for (int64_t i = 0; i < n; i++) {
            RR_ARG(Jb[i * n + i] == 0, "invalid J: diagonal bit J[%lld][%lld] set, expected 0", (long long)i + 1, (long long)i + 1);
            for (int64_t j = 0; j < n; j++) {
                RR_ARG(Jb[i * n + j] <= 1, "invalid J: entries must be 0/1 bits");
                RR_ARG(Jb[i * n + j] == Jb[j * n + i], "invalid J: not symmetric at (%lld,%lld)", (long long)i + 1, (long long)j + 1);
            }
        }
        RR_CUDA(cudaMalloc(&g->d_Jb, n * n));
        RR_CUDA(cudaMemcpy(g->d_Jb, Jb, n * n, cudaMemcpyHostToDevice));
    }
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_graph_sk_create(rrrmc_ctx_t *ctx, int64_t N, int kind, const void *J, rrrmc_graph_t **out)
{
    RR_ARG(ctx && J && out, "rrrmc_graph_sk_create: NULL argument");
    RR_ARG(kind == RRRMC_SK_F64 || kind == RRRMC_SK_BIN, "coupling kind must be RRRMC_SK_F64 or RRRMC_SK_BIN, given %d", kind);
    RR_ARG(N >= 1 && N <= 46340, "N = %lld out of range 1..46340", (long long)N);
    rrrmc_graph *g = new rrrmc_graph();
    g->ctx = ctx; g->kind = kind; g->N = N; g->Nk = N; g->M = 1; g->max_deg = (int)N - 1;
    g->sN = sqrt((double)N); // SK.jl:47
    if (kind == RRRMC_SK_F64) g->Jd.assign((const double *)J, (const double *)J + N * N);
    RR_CUDA(cudaSetDevice(ctx->device));
    rrrmc_status_t st = upload_sk_couplings(g, N, kind, J);
    if (st != RRRMC_OK) { delete g; return st; }
    RR_CUDA(cudaDeviceSynchronize()); // default-stream uploads are complete before any non-blocking stream uses them
    *out = g;
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_graph_qt_create(rrrmc_ctx_t *ctx, int64_t N, int64_t M, double fourK, rrrmc_graph_t **out)
{
    RR_ARG(ctx && out, "rrrmc_graph_qt_create: NULL argument");
    RR_ARG(M > 2, "M must be greater than 2, given: %lld", (long long)M);                     // QT.jl:47
    RR_ARG(N >= M && N % M == 0, "N must be divisible by M, given: N=%lld M=%lld", (long long)N, (long long)M); // QT.jl:48
    RR_ARG(N < ((int64_t)1 << 31), "N = %lld out of range", (long long)N);
    rrrmc_graph *g = new rrrmc_graph();
    g->ctx = ctx; g->kind = RRRMC_QT; g->N = N; g->M = M; g->Nk = N / M; g->fourK = fourK; g->max_deg = 2;
    g->allDE = { 0.0, fourK };                                                                 // QT.jl:111
    RR_CUDA(cudaDeviceSynchronize()); // default-stream uploads are complete before any non-blocking stream uses them
    *out = g;
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_graph_quant_create(rrrmc_ctx_t *ctx, int64_t Nk, int64_t M, double Gamma, double beta,
                                                   int inner, const void *J_inner, rrrmc_graph_t **out)
{
    RR_ARG(ctx && out, "rrrmc_graph_quant_create: NULL argument");
    RR_ARG(Gamma >= 0, "Γ must be non-negative, given: %g", Gamma);                           // QT.jl:164
    RR_ARG(M > 2, "M must be greater than 2, given: %lld", (long long)M);
    RR_ARG(inner == RRRMC_SK_F64 || inner == RRRMC_SK_BIN || inner == RRRMC_EMPTY, "unsupported inner graph kind %d", inner);
    RR_ARG(inner == RRRMC_EMPTY || J_inner, "J_inner is NULL");
    RR_ARG(Nk >= 1 && Nk <= 46340 && Nk * M < ((int64_t)1 << 31), "Nk = %lld, M = %lld out of range", (long long)Nk, (long long)M);
    RR_ARG(std::isfinite(beta) && beta > 0, "β must be finite and positive, given: %g", beta);
    rrrmc_graph *g = new rrrmc_graph();
    g->ctx = ctx; g->kind = RRRMC_QUANT; g->N = Nk * M; g->Nk = Nk; g->M = M; g->inner = inner;
    g->Gamma = Gamma; g->beta = beta; g->sN = sqrt((double)Nk);
    g->fourK = nearbyint(2.0 / beta * log(1.0 / tanh(beta * Gamma / (double)M)) * 1e8) / 1e8; // round(·, digits=8), QT.jl:165
    g->allDE = { 0.0, g->fourK };                                                              // allΔE of inner_graph(X), QT.jl:111
    g->max_deg = 2 + (inner == RRRMC_EMPTY ? 0 : (int)Nk - 1);
    RR_CUDA(cudaSetDevice(ctx->device));
    if (inner != RRRMC_EMPTY) {
        rrrmc_status_t st = upload_sk_couplings(g, Nk, inner, J_inner);
        if (st != RRRMC_OK) { delete g; return st; }
    }
    RR_CUDA(cudaDeviceSynchronize()); // default-stream uploads are complete before any non-blocking stream uses them
    *out = g;
    return RRRMC_OK;
}
// GraphQEAT (QAliases.jl:51-81): GraphQuant(N, M, Γ, β, GraphEANormal{2D}, L, A, J) — M Trotter slices of one
// GraphEANormal instance (QT.jl:139-147: the slices share A and J and keep separate local-field caches).
extern "C" rrrmc_status_t rrrmc_graph_quant_ea_create(rrrmc_ctx_t *ctx, int L, int D, int64_t M, double Gamma, double beta,
                                                      const int64_t *A, const double *J, rrrmc_graph_t **out)
{
    RR_ARG(ctx && A && J && out, "rrrmc_graph_quant_ea_create: NULL argument");
    RR_ARG(Gamma >= 0, "Γ must be non-negative, given: %g", Gamma);                           // QT.jl:164
    RR_ARG(M > 2, "M must be greater than 2, given: %lld", (long long)M);
    RR_ARG(std::isfinite(beta) && beta > 0, "β must be finite and positive, given: %g", beta);
    rrrmc_graph *e = nullptr;
    RR_TRY(rrrmc_graph_ea_create(ctx, L, D, RRRMC_EA_F64, A, J, &e));   // validates A and J, uploads d_A and d_Jd
    const int64_t Nk = e->N;
    if (!(Nk * M < ((int64_t)1 << 31))) { rrrmc_graph_destroy(e); rrrmc_set_error("Nk = %lld, M = %lld out of range", (long long)Nk, (long long)M); return RRRMC_ERR_ARG; }
    e->kind = RRRMC_QUANT; e->inner = RRRMC_EA_F64; e->Nk = Nk; e->M = M; e->N = Nk * M;
    e->Gamma = Gamma; e->beta = beta; e->sN = sqrt((double)Nk); e->bipartite = false;
    e->fourK = nearbyint(2.0 / beta * log(1.0 / tanh(beta * Gamma / (double)M)) * 1e8) / 1e8;  // QT.jl:165
    e->allDE = { 0.0, e->fourK };                                                              // QT.jl:111
    e->max_deg = 2 + e->twoD;
    *out = e;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_state_set_quant_betas(rrrmc_state_t *s, const double *beta, double *fourK_out)
{
    RR_ARG(s, "state is NULL");
    rrrmc_graph *g = s->g;
    RR_ARG(g->kind == RRRMC_QUANT, "a per-replica β ladder applies to GraphQuant only (fourK is a function of β, QT.jl:165)");
    RR_CUDA(cudaSetDevice(g->ctx->device));
    s->energy_valid = false;
    if (!beta) { s->q_beta.clear(); s->q_fourK.clear(); cudaFree(s->d_q_fourK); s->d_q_fourK = nullptr; return RRRMC_OK; }
    std::vector<double> b(beta, beta + s->R), fk(s->R);
    for (int64_t r = 0; r < s->R; r++) {
        RR_ARG(std::isfinite(b[r]) && b[r] > 0, "β must be finite and positive, given: %g (replica %lld)", b[r], (long long)r);
        fk[r] = nearbyint(2.0 / b[r] * log(1.0 / tanh(b[r] * g->Gamma / (double)g->M)) * 1e8) / 1e8;   // QT.jl:165
    }
    if (!s->d_q_fourK) RR_CUDA(cudaMalloc(&s->d_q_fourK, 8 * s->R));
    RR_CUDA(cudaMemcpyAsync(s->d_q_fourK, fk.data(), 8 * s->R, cudaMemcpyHostToDevice, g->ctx->stream));
    RR_CUDA(cudaStreamSynchronize(g->ctx->stream));
    if (fourK_out) memcpy(fourK_out, fk.data(), 8 * s->R);
    s->q_beta.swap(b); s->q_fourK.swap(fk);
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_graph_fourK(const rrrmc_graph_t *g, double *fourK)
{
    RR_ARG(g && fourK, "NULL argument");
    RR_ARG(g->kind == RRRMC_QT || g->kind == RRRMC_QUANT, "fourK is a parameter of GraphQT / GraphQuant only");
    *fourK = g->fourK;
    return RRRMC_OK;
}

// GraphEANormalDiscretized{Int,LEV,2D} (EA.jl:311-344) from explicit continuous couplings cJ: discretize (Common.jl:38-49:
// nearest level, the first wins ties) splits every coupling into a level dJ — the inner GraphEA{Int,LEV} — and a residual
// rJ = cJ - dJ (Float64). allΔE is that of the inner graph for the FULL level tuple (EA.jl:295-309), used or not.
extern "C" rrrmc_status_t rrrmc_graph_ea_discretized_create(rrrmc_ctx_t *ctx, int L, int D, const int64_t *A, const double *cJ,
                                                            const int64_t *lev, int nlev, rrrmc_graph_t **out)
{
    RR_ARG(ctx && A && cJ && lev && out, "rrrmc_graph_ea_discretized_create: NULL argument");
    RR_ARG(nlev >= 1 && nlev <= 16, "LEV must hold 1..16 levels, given %d", nlev);
    RR_ARG(L >= 2, "L must be >= 2, given: %d", L);
    RR_ARG(D >= 1 && D <= 4, "D must be in 1..4, given: %d", D);
    for (int l = 0; l < nlev; l++) RR_ARG(lev[l] >= -127 && lev[l] <= 127, "levels must fit int8, given %lld", (long long)lev[l]);
    const int twoD = 2 * D;
    const int64_t N = ipow(L, D);
    std::vector<int64_t> dJ((size_t)N * twoD);
    std::vector<double> rJ((size_t)N * twoD);
    for (int64_t a = 0; a < N * twoD; a++) {
        const double x = cJ[a];
        RR_ARG(std::isfinite(x), "cJ[%lld] is not finite", (long long)a);
        int64_t d = lev[0]; double r = x - (double)d;
        for (int l = 1; l < nlev; l++) {
            const double r1 = x - (double)lev[l];
            if (fabs(r1) < fabs(r)) { d = lev[l]; r = r1; }
        }
        dJ[a] = d; rJ[a] = r;
    }
    rrrmc_graph *g = nullptr;
    RR_TRY(rrrmc_graph_ea_create(ctx, L, D, RRRMC_EA_INT, A, dJ.data(), &g));   // validates A and the symmetry of dJ
    g->kind = RRRMC_EA_DISCR; g->M = 2;   // two local-field caches per chain: inner (integer) and residual (Float64)
    g->Jd = rJ;
    std::set<int64_t> es = { 0 };
    for (int n = 0; n < twoD; n++) {
        std::set<int64_t> nw;
        for (int64_t e : es) for (int l = 0; l < nlev; l++) { nw.insert(e + lev[l]); nw.insert(e - lev[l]); }
        es.swap(nw);
    }
    std::set<int64_t> de;
    for (int64_t e : es) de.insert(2 * (e < 0 ? -e : e));
    g->allDE.clear();
    for (int64_t d : de) g->allDE.push_back((double)d);
    if (g->allDE.size() > 64) { rrrmc_set_error("too many ΔE classes (%zu > 64)", g->allDE.size()); rrrmc_graph_destroy(g); return RRRMC_ERR_UNSUPPORTED; }
    RR_CUDA(cudaMalloc(&g->d_Jd, sizeof(double) * N * twoD));
    RR_CUDA(cudaMemcpy(g->d_Jd, g->Jd.data(), sizeof(double) * N * twoD, cudaMemcpyHostToDevice));
    RR_CUDA(cudaDeviceSynchronize());
    *out = g;
    return RRRMC_OK;
}

// GraphRRGNormalDiscretized{Int,LEV,K} (RRG.jl:274-310): GraphEANormalDiscretized's construction over a K-regular adjacency.
// neighbors(X) is the whole row (RRG.jl:499); neighbors(inner_graph(X)) skips the couplings discretised to zero (:133).
extern "C" rrrmc_status_t rrrmc_graph_rrg_discretized_create(rrrmc_ctx_t *ctx, int64_t N, int K, const int64_t *A, const double *cJ,
                                                             const int64_t *lev, int nlev, rrrmc_graph_t **out)
{
    RR_ARG(ctx && A && cJ && lev && out, "rrrmc_graph_rrg_discretized_create: NULL argument");
    RR_ARG(nlev >= 1 && nlev <= 16, "LEV must hold 1..16 levels, given %d", nlev);
    RR_ARG(K >= 1 && K <= 8 && N >= 2, "K must be in 1..8 and N >= 2");
    for (int l = 0; l < nlev; l++) RR_ARG(lev[l] >= -127 && lev[l] <= 127, "levels must fit int8, given %lld", (long long)lev[l]);
    std::vector<int64_t> dJ((size_t)N * K);
    std::vector<double> rJ((size_t)N * K);
    for (int64_t a = 0; a < N * K; a++) {
        const double x = cJ[a];
        RR_ARG(std::isfinite(x), "cJ[%lld] is not finite", (long long)a);
        int64_t d = lev[0]; double r = x - (double)d;
        for (int l = 1; l < nlev; l++) {
            const double r1 = x - (double)lev[l];
            if (fabs(r1) < fabs(r)) { d = lev[l]; r = r1; }
        }
        dJ[a] = d; rJ[a] = r;
    }
    rrrmc_graph *g = nullptr;
    RR_TRY(rrrmc_graph_rrg_create(ctx, N, K, RRRMC_EA_INT, A, dJ.data(), &g));
    g->kind = RRRMC_EA_DISCR; g->M = 2;
    g->nz_neighbors = true;                                // harmless when no coupling was discretised to zero
    for (int64_t i = 0; i < N; i++) {                      // neighbors(X, i) = A[i] (RRG.jl:499)
        for (int k = 0; k < K; k++) g->uA0[i * K + k] = g->A0[i * K + k];
        g->nuA[i] = K;
    }
    g->Jd = rJ;
    std::set<int64_t> es = { 0 };
    for (int n = 0; n < K; n++) {
        std::set<int64_t> nw;
        for (int64_t e : es) for (int l = 0; l < nlev; l++) { nw.insert(e + lev[l]); nw.insert(e - lev[l]); }
        es.swap(nw);
    }
    std::set<int64_t> de;
    for (int64_t e : es) de.insert(2 * (e < 0 ? -e : e));
    g->allDE.clear();
    for (int64_t d : de) g->allDE.push_back((double)d);
    if (g->allDE.size() > 64) { rrrmc_set_error("too many ΔE classes (%zu > 64)", g->allDE.size()); rrrmc_graph_destroy(g); return RRRMC_ERR_UNSUPPORTED; }
    RR_CUDA(cudaMalloc(&g->d_Jd, sizeof(double) * N * K));
    RR_CUDA(cudaMemcpy(g->d_Jd, g->Jd.data(), sizeof(double) * N * K, cudaMemcpyHostToDevice));
    RR_CUDA(cudaDeviceSynchronize());
    *out = g;
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_graph_destroy(rrrmc_graph_t *g)
{
    if (!g) return RRRMC_OK;
    cudaSetDevice(g->ctx->device);
    cudaFree(g->d_jcode); cudaFree(g->d_jmask); cudaFree(g->d_A); cudaFree(g->d_J8); cudaFree(g->d_Jd); cudaFree(g->d_Jb);
    delete g;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_getN(const rrrmc_graph_t *g, int64_t *N)
{
    RR_ARG(g && N, "NULL argument");
    *N = g->N;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_neighbors(const rrrmc_graph_t *g, int64_t site, int64_t *out, int *n)
{
    RR_ARG(g && out && n, "NULL argument");
    RR_ARG(site >= 1 && site <= g->N, "site %lld out of range 1..%lld", (long long)site, (long long)g->N);
    int m = 0;
    if (g->kind == RRRMC_SK_F64 || g->kind == RRRMC_SK_BIN) {        // AllButOne, Common.jl:78-92
        for (int64_t j = 1; j <= g->N; j++) if (j != site) out[m++] = j;
    } else if (g->kind == RRRMC_QT || g->kind == RRRMC_QUANT) {      // QT.jl:105-108; QNeighbIter QT.jl:288-321
        out[m++] = site - g->Nk + (site <= g->Nk ? g->N : 0);
        out[m++] = site + g->Nk - (site + g->Nk > g->N ? g->N : 0);
        if (g->kind == RRRMC_QUANT && g->inner != RRRMC_EMPTY) {
            const int64_t k = (site - 1) / g->Nk, i = (site - 1) % g->Nk + 1;
            if (g->inner == RRRMC_EA_F64)            // GraphQEAT: the slice graph's uA row, shifted to the slice
                for (int q = 0; q < g->nuA[i - 1]; q++) out[m++] = k * g->Nk + g->uA0[(i - 1) * g->twoD + q] + 1;
            else
                for (int64_t j = 1; j <= g->Nk; j++) if (j != i) out[m++] = k * g->Nk + j;
        }
    } else {
        m = g->nuA[site - 1];
        for (int k = 0; k < m; k++) out[k] = g->uA0[(site - 1) * g->twoD + k] + 1;
    }
    *n = m;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_max_neighbors(const rrrmc_graph_t *g, int64_t *n)
{
    RR_ARG(g && n, "NULL argument");
    *n = g->max_deg;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_allDE(const rrrmc_graph_t *g, double *out, int *n)
{
    RR_ARG(g && out && n, "NULL argument");
    if (g->allDE.empty()) { rrrmc_set_error("allΔE is only defined for DiscrGraph types"); return RRRMC_ERR_UNSUPPORTED; }
    *n = (int)g->allDE.size();
    for (int k = 0; k < *n; k++) out[k] = g->allDE[k];
    return RRRMC_OK;
}

// ------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------
extern "C" rrrmc_status_t rrrmc_state_create(rrrmc_graph_t *g, int64_t R, rrrmc_state_t **out)
{
    RR_ARG(g && out, "NULL argument");
    RR_ARG(R >= 1 && R <= ((int64_t)1 << 24), "n_replicas %lld out of range", (long long)R);
    rrrmc_state *s = new rrrmc_state();
    s->g = g; s->R = R; s->W = (R + 31) / 32; s->nchunks = (g->N + 63) / 64;
    RR_CUDA(cudaSetDevice(g->ctx->device));
    RR_CUDA(cudaMalloc(&s->d_spins, sizeof(uint32_t) * g->N * s->W));
    RR_CUDA(cudaMemsetAsync(s->d_spins, 0, sizeof(uint32_t) * g->N * s->W, g->ctx->stream));
    s->ibuf_len = std::max<int64_t>(g->N, s->W * 32);
    RR_CUDA(cudaMalloc(&s->d_ibuf, sizeof(int32_t) * s->ibuf_len));
    RR_CUDA(cudaMalloc(&s->d_acc, sizeof(long long) * s->W * 32));
    *out = s;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_state_destroy(rrrmc_state_t *s)
{
    if (!s) return RRRMC_OK;
    cudaSetDevice(s->g->ctx->device);
    cudaStreamSynchronize(s->g->ctx->stream);
    cudaFree(s->d_spins); cudaFree(s->d_chunks); cudaFree(s->d_ibuf); cudaFree(s->d_acc);
    cudaFree(s->d_flips); cudaFree(s->d_mask); cudaFree(s->d_q_fourK); cudaFree(s->d_beta);
    cudaFree(s->d_pt_beta); cudaFree(s->d_pt_masks); cudaFree(s->d_pt_acc);
    chain_free(s);
    sk_dense_free(s);
    checkerboard_tma_free(s);
    delete s;
    return RRRMC_OK;
}
static rrrmc_status_t ensure_chunks(rrrmc_state *s)
{
    if (!s->d_chunks) RR_CUDA(cudaMalloc(&s->d_chunks, sizeof(uint64_t) * s->R * s->nchunks));
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_state_randomize(rrrmc_state_t *s, uint64_t seed)
{
    RR_ARG(s, "state is NULL");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    // the fresh configuration is written in the multispin layout: that copy is current, the chain copy is stale
    s->energy_valid = false; s->ms_valid = true; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return launch_randomize(s, seed);
}
extern "C" rrrmc_status_t rrrmc_state_upload(rrrmc_state_t *s, int64_t first, int64_t count, const uint64_t *chunks)
{
    RR_ARG(s && chunks, "NULL argument");
    RR_ARG(first >= 0 && count >= 1 && first + count <= s->R, "replica range [%lld,%lld) outside 0..%lld",
           (long long)first, (long long)(first + count), (long long)s->R);
    rrrmc_ctx *ctx = s->g->ctx;
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_TRY(ensure_chunks(s));
    RR_TRY(chain_sync_to_multispin(s));
    RR_CUDA(cudaMemcpyAsync(s->d_chunks, chunks, sizeof(uint64_t) * count * s->nchunks, cudaMemcpyHostToDevice, ctx->stream));
    RR_TRY(launch_upload_transpose(s, first, count));
    RR_CUDA(cudaStreamSynchronize(ctx->stream)); // the caller may free `chunks` on return
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_state_download(rrrmc_state_t *s, int64_t first, int64_t count, uint64_t *chunks)
{
    RR_ARG(s && chunks, "NULL argument");
    RR_ARG(first >= 0 && count >= 1 && first + count <= s->R, "replica range [%lld,%lld) outside 0..%lld",
           (long long)first, (long long)(first + count), (long long)s->R);
    rrrmc_ctx *ctx = s->g->ctx;
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_TRY(ensure_chunks(s));
    RR_TRY(chain_sync_to_multispin(s));
    RR_TRY(launch_download_transpose(s, first, count));
    RR_CUDA(cudaMemcpyAsync(chunks, s->d_chunks, sizeof(uint64_t) * count * s->nchunks, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    s->chain_valid = (first == 0 && count == s->R); // d_chunks doubles as the chain layout
    return RRRMC_OK;
}

// ------------------------------------------------------------------------------------------------
// Interface queries
// ------------------------------------------------------------------------------------------------
static rrrmc_status_t energy_to_host(rrrmc_state *s, double *E_out)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    if (g->kind == RRRMC_EA_PM1 && g->d_jcode) {
        RR_TRY(launch_energy_pm1(s, s->d_ibuf));
        std::vector<int> h(s->W * 32);
        RR_CUDA(cudaMemcpyAsync(h.data(), s->d_ibuf, sizeof(int) * s->W * 32, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
        const double base = -(double)g->D * (double)g->N;
        for (int64_t r = 0; r < s->R; r++) E_out[r] = base + 2.0 * (double)h[r];
        return RRRMC_OK;
    }
    return chain_energy(s, E_out);
}
extern "C" rrrmc_status_t rrrmc_energy(rrrmc_state_t *s, double *E_out)
{
    RR_ARG(s && E_out, "NULL argument");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    RR_TRY(energy_to_host(s, E_out));
    s->energy_valid = true;
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_delta_energy(rrrmc_state_t *s, int64_t site, double *dE_out)
{
    RR_ARG(s && dE_out, "NULL argument");
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(site >= 1 && site <= g->N, "site %lld out of range 1..%lld", (long long)site, (long long)g->N);
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    if (g->kind == RRRMC_EA_PM1 && g->d_jcode) {
        RR_TRY(launch_delta_energy_site(s, site - 1, s->d_ibuf));
        std::vector<int> h(s->W * 32);
        RR_CUDA(cudaMemcpyAsync(h.data(), s->d_ibuf, sizeof(int) * s->W * 32, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int64_t r = 0; r < s->R; r++) dE_out[r] = (double)h[r];
        return RRRMC_OK;
    }
    return chain_delta_energy_site(s, site - 1, 0, dE_out);
}
extern "C" rrrmc_status_t rrrmc_all_delta_energy(rrrmc_state_t *s, int64_t replica, double *dE_out)
{
    RR_ARG(s && dE_out, "NULL argument");
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(replica >= 0 && replica < s->R, "replica %lld out of range 0..%lld", (long long)replica, (long long)s->R - 1);
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    if (g->kind == RRRMC_EA_PM1 && g->d_jcode) {
        RR_TRY(launch_delta_energy_replica(s, replica, s->d_ibuf));
        std::vector<int> h(g->N);
        RR_CUDA(cudaMemcpyAsync(h.data(), s->d_ibuf, sizeof(int) * g->N, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int64_t i = 0; i < g->N; i++) dE_out[i] = (double)h[i];
        return RRRMC_OK;
    }
    return chain_delta_energy_replica(s, replica, dE_out);
}
extern "C" rrrmc_status_t rrrmc_spinflip(rrrmc_state_t *s, int64_t site, const uint32_t *replica_mask)
{
    RR_ARG(s, "state is NULL");
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(site >= 1 && site <= g->N, "site %lld out of range 1..%lld", (long long)site, (long long)g->N);
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    const uint32_t *d_mask = nullptr;
    if (replica_mask) {
        if (!s->d_mask) RR_CUDA(cudaMalloc(&s->d_mask, sizeof(uint32_t) * s->W));
        RR_CUDA(cudaMemcpyAsync(s->d_mask, replica_mask, sizeof(uint32_t) * s->W, cudaMemcpyHostToDevice, ctx->stream));
        d_mask = s->d_mask;
    }
    RR_TRY(launch_flip_site(s, site - 1, d_mask));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_magnetization(rrrmc_state_t *s, double *m_out)
{
    RR_ARG(s && m_out, "NULL argument");
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    RR_CUDA(cudaMemsetAsync(s->d_acc, 0, sizeof(long long) * s->W * 32, ctx->stream));
    RR_TRY(launch_count_lanes(ctx, s->d_spins, g->N, (int)s->W, s->d_acc));
    std::vector<long long> h(s->W * 32);
    RR_CUDA(cudaMemcpyAsync(h.data(), s->d_acc, sizeof(long long) * s->W * 32, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t r = 0; r < s->R; r++) m_out[r] = 2.0 * (double)h[r] - (double)g->N;
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_delta_energy_residual(rrrmc_state_t *s, int64_t site, double *dE_out)
{
    RR_ARG(s && dE_out, "NULL argument");
    rrrmc_graph *g = s->g;
    RR_ARG(site >= 1 && site <= g->N, "site %lld out of range 1..%lld", (long long)site, (long long)g->N);
    RR_CUDA(cudaSetDevice(g->ctx->device));
    if (g->kind != RRRMC_QUANT && g->kind != RRRMC_EA_DISCR) { for (int64_t r = 0; r < s->R; r++) dE_out[r] = 0.0; return RRRMC_OK; } // Interface.jl:261
    RR_TRY(chain_sync_to_multispin(s));
    return chain_delta_energy_site(s, site - 1, 1, dE_out);
}
static rrrmc_status_t quant_observable(rrrmc_state_t *s, int what, double arg, double *out)
{
    RR_ARG(s && out, "NULL argument");
    rrrmc_graph *g = s->g;
    RR_ARG(g->kind == RRRMC_QUANT || (what == 0 && g->kind == RRRMC_QT), "this observable is defined for GraphQuant only");
    RR_CUDA(cudaSetDevice(g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    return chain_quant_observable(s, what, arg, out);
}
extern "C" rrrmc_status_t rrrmc_transverse_mag(rrrmc_state_t *s, double beta, double *out) { return quant_observable(s, 0, beta, out); }
extern "C" rrrmc_status_t rrrmc_Qenergy(rrrmc_state_t *s, double *out) { return quant_observable(s, 1, 0, out); }
extern "C" rrrmc_status_t rrrmc_Renergies(rrrmc_state_t *s, double *out) { return quant_observable(s, 2, 0, out); }
extern "C" rrrmc_status_t rrrmc_overlaps(rrrmc_state_t *s, double *out) { return quant_observable(s, 3, 0, out); }

extern "C" rrrmc_status_t rrrmc_sk_fields_init(rrrmc_state_t *s, int use_tensor_cores, double *E_out, float *device_ms)
{
    RR_ARG(s, "state is NULL");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    return sk_dense_fields_init(s, use_tensor_cores, E_out, device_ms);
}
extern "C" rrrmc_status_t rrrmc_sk_get_fields(rrrmc_state_t *s, double *lf_out)
{
    RR_ARG(s && lf_out, "NULL argument");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    return sk_dense_get_fields(s, lf_out);
}
extern "C" rrrmc_status_t rrrmc_sk_metropolis_sweeps(rrrmc_state_t *s, const double *beta, uint64_t seed, uint64_t sweep0,
                                                     int64_t nsweeps, double *E_out, int64_t *accepted_out)
{
    RR_ARG(s && beta, "NULL argument");
    RR_ARG(nsweeps >= 0 && nsweeps < ((int64_t)1 << 31), "nsweeps out of range");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    return sk_dense_sweeps(s, beta, seed, sweep0, nsweeps, E_out, accepted_out);
}

// ------------------------------------------------------------------------------------------------
// samplers
// ------------------------------------------------------------------------------------------------
extern "C" rrrmc_status_t rrrmc_opts_default(rrrmc_opts_t *o)
{
    RR_ARG(o, "opts is NULL");
    memset(o, 0, sizeof *o);
    o->schedule = RRRMC_SCHED_RANDOM_SITE;   // the reference order and sampling contract; lattice sweeps are opt-in
    o->planes_K = 5;
    o->planes_M = 4;
    o->cb_method = RRRMC_CB_AUTO;
    o->count_accepted = 1;
    o->staged_thr = NAN;
    o->staged_thr_fact = 5.0;
    return RRRMC_OK;
}

static uint64_t fixed64(double p) // floor(p * 2^64) clamped to 2^64-1, p in [0,1]
{
    if (!(p > 0)) return 0;
    const double v = ldexp(p, 64);
    if (v >= 18446744073709551616.0) return ~0ull;
    return (uint64_t)v;
}

static rrrmc_status_t fill_cb_params(rrrmc_state *s, const uint64_t *thr64, int nthr, int K, int M, uint64_t seed, cb_params &p)
{
    rrrmc_graph *g = s->g;
    if (!(g->kind == RRRMC_EA_PM1 && g->d_jcode)) {
        rrrmc_set_error("checkerboard sweeps need a ±J GraphEA with D<=3");
        return RRRMC_ERR_UNSUPPORTED;
    }
    if (!g->bipartite) {
        rrrmc_set_error("checkerboard sweeps need even L (a two-colourable lattice), given L=%d; use schedule=RANDOM_SITE", g->L);
        return RRRMC_ERR_UNSUPPORTED;
    }
    RR_ARG(nthr == g->D, "expected %d acceptance thresholds (ΔE=4..%d), given %d", g->D, 4 * g->D, nthr);
    RR_ARG(K >= 0 && M >= 0 && K + M <= CB_MAXK, "planes_K and planes_M must be >= 0 with planes_K + planes_M <= %d, given %d + %d", CB_MAXK, K, M);
    RR_ARG(M % 4 == 0, "planes_M must be a multiple of 4 (one Philox call serves four merged planes), given %d", M);
    memset(&p, 0, sizeof p);
    p.spins = s->d_spins; p.flips = nullptr; p.jcode = g->d_jcode;
    p.L = g->L; p.Lh = g->L / 2; p.W = (int)s->W; p.G = (int)((s->W + 3) / 4);
    p.k0 = (uint32_t)seed; p.k1 = (uint32_t)(seed >> 32);
    p.K = K; p.M = M;
    RR_ARG((int64_t)g->N * s->W < ((int64_t)1 << 31), "N*W = %lld words exceeds the kernel's 32-bit indexing", (long long)(g->N * s->W));
    p.invG = 1.0f / (float)p.G;
    { const char *v = getenv("RRRMC_CB_VARIANT"); p.variant = v ? atoi(v) : 0; }
    auto after = [](uint64_t t, int used) { return (uint32_t)((used ? (t << used) : t) >> 32); }; // next 32 bits
    for (int c = 0; c < nthr; c++) {
        for (int q = 0; q < K + M; q++) p.plane[q][c] = ((thr64[c] >> (63 - q)) & 1ull) ? 0xffffffffu : 0u;
        p.rem[c] = after(thr64[c], K);
        p.remM[c] = after(thr64[c], K + M);
    }
    for (int q = 0; q < K + M; q++) {
        int ones = 0;
        for (int c = 0; c < nthr; c++) ones += p.plane[q][c] != 0;
        p.planeop[q] = ones == 0 ? 0 : (ones == nthr ? 1 : 2);
    }
    p.Ku = 0;
    while (p.Ku < K && p.planeop[p.Ku] != 2) p.Ku++;
    p.Kz = 0;
    while (p.Kz < p.Ku && p.planeop[p.Kz] == 0) p.Kz++;
    for (int r = 0; r < 10; r++) { p.rk[r][0] = p.k0 + (uint32_t)r * 0x9E3779B9u; p.rk[r][1] = p.k1 + (uint32_t)r * 0xBB67AE85u; }
    return RRRMC_OK;
}

static rrrmc_status_t run_sweep(rrrmc_state *s, cb_params &p, uint64_t t)
{
    rrrmc_graph *g = s->g;
    p.t_lo = (uint32_t)t; p.t_hi16 = (uint32_t)(t >> 32) << 16;
    RR_TRY(launch_checkerboard(g->ctx, p, g->D, 0));
    RR_TRY(launch_checkerboard(g->ctx, p, g->D, 1));
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_checkerboard_sweeps(rrrmc_state_t *s, const uint64_t *thr64, int nthr, int K, int M,
                                                    uint64_t seed, uint64_t sweep0, int64_t nsweeps)
{
    RR_ARG(s && thr64, "NULL argument");
    RR_ARG(nsweeps >= 0, "nsweeps must be >= 0");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    cb_params p;
    RR_TRY(fill_cb_params(s, thr64, nthr, K, M, seed, p));
    for (int64_t k = 0; k < nsweeps; k++) RR_TRY(run_sweep(s, p, sweep0 + (uint64_t)k));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return RRRMC_OK;
}

// ---- "sparse" acceptance procedure: binomial-count tables and sweeps (ea_multispin.cu:k_checkerboard_sparse)
extern "C" rrrmc_status_t rrrmc_checkerboard_sparse_tables(const uint64_t *thr64, int nthr, uint32_t *tbl, int tbl_len)
{
    RR_ARG(thr64 && tbl, "NULL argument");
    RR_ARG(nthr >= 1 && nthr <= 3, "expected 1..3 acceptance thresholds, given %d", nthr);
    RR_ARG(tbl_len >= CBS_T1 + (nthr - 1) * CBS_TC, "table buffer too small: %d < %d", tbl_len, CBS_T1 + (nthr - 1) * CBS_TC);
    for (int c = 1; c <= nthr; c++) {
        const int n = c == 1 ? 32 : 128;
        uint32_t *T = c == 1 ? tbl : tbl + CBS_T1 + (c - 2) * CBS_TC;
        const long double p = (long double)thr64[c - 1] / 18446744073709551616.0L, q = 1.0L - p;
        long double pk = powl(q, (long double)n), cdf = 0.0L;   // P(Bin(n,p) = k), running CDF
        for (int k = 0; k <= n; k++) {
            cdf += pk;
            const long double v = rintl(cdf * 4294967296.0L);    // more than k lanes pass iff x > T[k]
            T[k] = (k == n || v >= 4294967296.0L) ? 0xffffffffu : (v < 1.0L ? 0u : (uint32_t)(v - 1.0L));
            pk = q > 0.0L ? pk * (long double)(n - k) / (long double)(k + 1) * (p / q) : (k + 1 == n ? 1.0L : 0.0L);
        }
    }
    return RRRMC_OK;
}

static rrrmc_status_t fill_cbs_params(rrrmc_state *s, const uint32_t *tbl, int tbl_len, uint64_t seed, cbs_params &p)
{
    rrrmc_graph *g = s->g;
    if (!(g->kind == RRRMC_EA_PM1 && g->d_jcode)) {
        rrrmc_set_error("checkerboard sweeps need a ±J GraphEA with D<=3");
        return RRRMC_ERR_UNSUPPORTED;
    }
    if (!g->bipartite) {
        rrrmc_set_error("checkerboard sweeps need even L (a two-colourable lattice), given L=%d; use schedule=RANDOM_SITE", g->L);
        return RRRMC_ERR_UNSUPPORTED;
    }
    const int need = CBS_T1 + (g->D - 1) * CBS_TC;
    RR_ARG(tbl_len == need, "expected a %d-entry count table for D=%d, given %d", need, g->D, tbl_len);
    RR_ARG((int64_t)g->N * s->W < ((int64_t)1 << 31), "N*W = %lld words exceeds the kernel's 32-bit indexing", (long long)(g->N * s->W));
    RR_ARG(tbl[CBS_T1 - 1] == 0xffffffffu, "count table of class 1 must end with 2^32-1");
    for (int c = 2; c <= g->D; c++) RR_ARG(tbl[CBS_T1 + (c - 1) * CBS_TC - 1] == 0xffffffffu, "count table of class %d must end with 2^32-1", c);
    memset(&p, 0, sizeof p);
    p.spins = s->d_spins; p.flips = nullptr; p.jcode = g->d_jcode; p.jmask = reinterpret_cast<const uint4 *>(g->d_jmask);
    p.L = g->L; p.Lh = g->L / 2; p.W = (int)s->W; p.G = (int)((s->W + 3) / 4);
    p.invG = 1.0f / (float)p.G;
    p.Gshift = -1;
    for (int b = 0; b < 30; b++) if (p.G == (1 << b)) p.Gshift = b;
    { const char *v = getenv("RRRMC_CB_VARIANT"); p.variant = v ? atoi(v) : 0; }
    memcpy(p.tbl, tbl, sizeof(uint32_t) * need);
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { p.rk[r][0] = k0 + (uint32_t)r * 0x9E3779B9u; p.rk[r][1] = k1 + (uint32_t)r * 0xBB67AE85u; }
    return RRRMC_OK;
}
static rrrmc_status_t run_sweep_sparse(rrrmc_state *s, cbs_params &p, uint64_t t)
{
    rrrmc_graph *g = s->g;
    p.t_lo = (uint32_t)t; p.t_hi16 = (uint32_t)(t >> 32) << 16;
    RR_TRY(launch_checkerboard_sparse(g->ctx, p, g->D, 0));
    RR_TRY(launch_checkerboard_sparse(g->ctx, p, g->D, 1));
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_checkerboard_sweeps_sparse(rrrmc_state_t *s, const uint32_t *tbl, int tbl_len,
                                                           uint64_t seed, uint64_t sweep0, int64_t nsweeps)
{
    RR_ARG(s && tbl, "NULL argument");
    RR_ARG(nsweeps >= 0, "nsweeps must be >= 0");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    cbs_params p;
    RR_TRY(fill_cbs_params(s, tbl, tbl_len, seed, p));
    for (int64_t k = 0; k < nsweeps; k++) RR_TRY(run_sweep_sparse(s, p, sweep0 + (uint64_t)k));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return RRRMC_OK;
}
// ---- "poisson" acceptance procedure: Poisson-count tables and sweeps (ea_poisson.cu:k_checkerboard_poisson)
static void poisson_table(long double mu, long double scale, uint32_t last, uint32_t *T, int n)
{
    long double pk = expl(-mu), cdf = 0.0L;
    for (int k = 0; k < n; k++) {
        cdf += pk;
        const long double v = rintl(cdf * scale);            // more than k hits iff x > T[k]
        T[k] = (k == n - 1 || v > (long double)last) ? last : (v < 1.0L ? 0u : (uint32_t)(v - 1.0L));
        pk = pk * mu / (long double)(k + 1);
    }
}
extern "C" rrrmc_status_t rrrmc_checkerboard_poisson_tables(const uint64_t *thr64, int nthr, uint32_t *tbl, int tbl_len)
{
    RR_ARG(thr64 && tbl, "NULL argument");
    RR_ARG(nthr >= 1 && nthr <= 3, "expected 1..3 acceptance thresholds, given %d", nthr);
    RR_ARG(tbl_len >= CBP_LEN, "table buffer too small: %d < %d", tbl_len, CBP_LEN);
    long double lam[5] = { 0, 0, 0, 0, 0 };
    for (int c = 1; c <= nthr; c++) lam[c] = -log1pl(-(long double)thr64[c - 1] / 18446744073709551616.0L);
    uint32_t *TA = tbl, *TB0 = TA + CBP_KA, *TB = TB0 + CBP_KR, *TC = TB + CBP_KR;
    poisson_table(128.0L * (lam[1] - lam[2]), 4294967296.0L, 0xffffffffu, TA, CBP_KA);
    poisson_table(128.0L * (lam[2] - lam[3]), 4294967296.0L, 0xffffffffu, TB, CBP_KR);
    poisson_table(128.0L * lam[3], 4294967296.0L, 0xffffffffu, TC, CBP_KR);
    poisson_table(128.0L * (lam[2] - lam[3]), (long double)TC[0] + 1.0L, TC[0], TB0, CBP_KR);
    return RRRMC_OK;
}
// Number of static position words: the smallest NW in (1, 2, 4, 6) whose overflow probability per task (level-1 count
// above 4·NW-1) is <= tol; 0 if none. tol <= 0 selects the measured defaults: the second tier is cheap, so a word of
// static slots (four to eight one-hot masks for every task) only pays off when it is needed often.
extern "C" int rrrmc_checkerboard_poisson_nw(const uint32_t *tbl, double tol)
{
    if (!tbl) return 0;
    if (tbl[CBP_KA - 2] != 0xffffffffu) return 0;   // the 64-entry table does not cover the count distribution
    const int nws[4] = { 1, 2, 4, 6 };
    // (NW = 6: with 1.2 % of the tasks overflowing, a fifth of the warps take the second tier; beyond that — β < 0.54 in 3D —
    // the bit-plane kernel is faster: measured 2.5 vs 3.0·10¹² at β = 0.5, 3.1 vs 2.9 at 0.55, profiles/r2x_beta_sweep.jsonl)
    const double dflt[4] = { 0.03, 0.2, 0.06, 0.012 };
    for (int k = 0; k < 4; k++)
        if (1.0 - ((double)tbl[4 * nws[k] - 1] + 1.0) / 4294967296.0 <= (tol > 0 ? tol : dflt[k])) return nws[k];
    return 0;
}

// level-1 count lookup on the top 10 bits of the uniform: inside bucket e the count is a0 + (x > T); a bucket that holds
// two or more table entries stores {2^32-1, 64 + a0} (count >= a0: second tier)
static void cbp_build_bucket(const uint32_t *TA, uint2 *bk)
{
    for (uint32_t e = 0; e < (uint32_t)CBP_BUCKETS; e++) {
        const uint32_t lo = e << 22, hi = lo + ((1u << 22) - 1u);
        uint32_t below = 0, inside = 0, T = 0xffffffffu;
        for (int k = 0; k < CBP_KA; k++) {
            if (TA[k] < lo) below++;
            else if (TA[k] < hi) { inside++; T = TA[k]; }
        }
        bk[e] = inside <= 1 ? make_uint2(T, below) : make_uint2(0xffffffffu, 64u + below);   // ambiguous: base count only
    }
}
static rrrmc_status_t validate_cbp_tables(const rrrmc_state *s, const uint32_t *tbl, int tbl_len, int NW)
{
    const rrrmc_graph *g = s->g;
    if (!(g->kind == RRRMC_EA_PM1 && g->d_jmask)) {
        rrrmc_set_error("checkerboard sweeps need a ±J GraphEA with D<=3");
        return RRRMC_ERR_UNSUPPORTED;
    }
    if (!g->bipartite) {
        rrrmc_set_error("checkerboard sweeps need even L (a two-colourable lattice), given L=%d; use schedule=RANDOM_SITE", g->L);
        return RRRMC_ERR_UNSUPPORTED;
    }
    RR_ARG(tbl_len == CBP_LEN, "expected a %d-entry count table, given %d", CBP_LEN, tbl_len);
    RR_ARG(NW == 1 || NW == 2 || NW == 4 || NW == 6, "static position words NW must be 1, 2, 4 or 6, given %d", NW);
    RR_ARG((int64_t)g->N * s->W < ((int64_t)1 << 31), "N*W = %lld words exceeds the kernel's 32-bit indexing", (long long)(g->N * s->W));
    const uint32_t *TA = tbl, *TB0 = TA + CBP_KA, *TB = TB0 + CBP_KR, *TC = TB + CBP_KR;
    RR_ARG(TA[CBP_KA - 1] == 0xffffffffu && TB[CBP_KR - 1] == 0xffffffffu && TC[CBP_KR - 1] == 0xffffffffu,
           "count tables TA, TB, TC must end with 2^32-1");
    RR_ARG(TB0[CBP_KR - 1] == TC[0], "count table TB0 must end with TC[0]");
    for (int k = 1; k < CBP_KA; k++) RR_ARG(TA[k] >= TA[k - 1], "count table TA must be non-decreasing");
    if (g->D < 3) RR_ARG(TC[0] == 0xffffffffu, "D=%d has no level-3 hits: TC[0] must be 2^32-1", g->D);
    if (g->D < 2) RR_ARG(TB0[0] == 0xffffffffu && TB[0] == 0xffffffffu, "D=1 has no level-2 hits: TB0[0], TB[0] must be 2^32-1");
    return RRRMC_OK;
}
static rrrmc_status_t fill_cbp_params(rrrmc_state *s, const uint32_t *tbl, int tbl_len, int NW, uint64_t seed, cbp_params &p)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_TRY(validate_cbp_tables(s, tbl, tbl_len, NW));
    const uint32_t *TA = tbl, *TB0 = TA + CBP_KA, *TC = TB0 + 2 * CBP_KR;
    memset(&p, 0, sizeof p);
    p.spins = s->d_spins; p.flips = nullptr; p.jmask = reinterpret_cast<const uint4 *>(g->d_jmask);
    p.L = g->L; p.Lh = g->L / 2; p.W = (int)s->W; p.G = (int)((s->W + 3) / 4); p.NW = NW;
    p.invG = 1.0f / (float)p.G;
    p.Gshift = -1;
    for (int b = 0; b < 30; b++) if (p.G == (1 << b)) p.Gshift = b;
    { const char *v = getenv("RRRMC_CB_VARIANT"); p.variant = v ? atoi(v) : 0; }
    memcpy(p.tbl, tbl, sizeof(uint32_t) * CBP_LEN);
    p.tb0_0 = TB0[0]; p.tb0_1 = TB0[1]; p.tc0 = TC[0]; p.one = 1u;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { p.rk[r][0] = k0 + (uint32_t)r * 0x9E3779B9u; p.rk[r][1] = k1 + (uint32_t)r * 0xBB67AE85u; }
    // level-1 count lookup on the top 10 bits of the uniform: count = a0 + (x > T) inside a bucket
    if (!ctx->d_cbp_bucket) RR_CUDA(cudaMalloc(&ctx->d_cbp_bucket, sizeof(uint2) * CBP_BUCKETS));
    if (ctx->cbp_bucket_key.size() != (size_t)CBP_KA || memcmp(ctx->cbp_bucket_key.data(), TA, sizeof(uint32_t) * CBP_KA) != 0) {
        std::vector<uint2> bk(CBP_BUCKETS);
        cbp_build_bucket(TA, bk.data());
        RR_CUDA(cudaStreamSynchronize(ctx->stream));   // no launch may still be reading the previous lookup
        RR_CUDA(cudaMemcpy(ctx->d_cbp_bucket, bk.data(), sizeof(uint2) * CBP_BUCKETS, cudaMemcpyHostToDevice));
        ctx->cbp_bucket_key.assign(TA, TA + CBP_KA);
    }
    p.bucket = ctx->d_cbp_bucket;
    return RRRMC_OK;
}
// The poisson procedure has two kernels with identical results: the TMA-staged brick kernel (ea_tma.cu: 3D, whole
// 1024-replica slabs, L a multiple of 8) and the cp.async kernels of ea_poisson.cu (everything else).
// RRRMC_CB_VARIANT bit 11 (2048) forbids the TMA kernel (A/B timing, tests of the other path).
struct cbp_run {
    cbp_params p;
    cbt_params *tma = nullptr;      // non-null: launch the TMA kernel
    std::vector<cbp_group> groups;  // β ladder: one table set per 128-replica group (multi-sweep kernel only)
    std::vector<uint2> gbucket;
    ~cbp_run() { delete tma; }
};
static rrrmc_status_t prepare_poisson_run(rrrmc_state *s, const uint32_t *tbl, int tbl_len, int NW, uint64_t seed, cbp_run &run)
{
    RR_TRY(fill_cbp_params(s, tbl, tbl_len, NW, seed, run.p));
    if (checkerboard_tma_eligible(s) && !(run.p.variant & 2048)) {
        run.tma = new cbt_params();
        RR_TRY(checkerboard_tma_prepare(s, run.p, *run.tma));
    }
    return RRRMC_OK;
}
static rrrmc_status_t run_sweep_poisson(rrrmc_state *s, cbp_run &run, uint64_t t)
{
    rrrmc_graph *g = s->g;
    cbp_params &p = run.tma ? run.tma->p : run.p;
    p.t_lo = (uint32_t)t; p.t_hi16 = (uint32_t)(t >> 32) << 16;
    if (run.tma) {
        RR_TRY(launch_checkerboard_tma(g->ctx, *run.tma, 0));
        RR_TRY(launch_checkerboard_tma(g->ctx, *run.tma, 1));
        return RRRMC_OK;
    }
    RR_TRY(launch_checkerboard_poisson(g->ctx, p, g->D, 0));
    RR_TRY(launch_checkerboard_poisson(g->ctx, p, g->D, 1));
    return RRRMC_OK;
}
// n whole sweeps: one launch of the multi-sweep kernel when the TMA path applies (RRRMC_CB_VARIANT bit 12 (4096) keeps
// the per-colour launches for A/B timing and tests), else two launches per sweep.
static rrrmc_status_t run_sweeps_poisson(rrrmc_state *s, cbp_run &run, uint64_t t0, int64_t n)
{
    if (!run.groups.empty())
        return launch_checkerboard_flow(s, *run.tma, t0, n, run.groups.data(), run.gbucket.empty() ? nullptr : run.gbucket.data(), (int)run.groups.size());
    if (run.tma && !(run.p.variant & 4096) && !s->tma->flow_unavailable) {
        const rrrmc_status_t st = launch_checkerboard_flow(s, *run.tma, t0, n, nullptr, nullptr, 0);
        if (!(st == RRRMC_ERR_UNSUPPORTED && s->tma->flow_unavailable)) return st;
    }
    for (int64_t k = 0; k < n; k++) RR_TRY(run_sweep_poisson(s, run, t0 + (uint64_t)k));
    return RRRMC_OK;
}
// β ladder: tbls[ngroups][CBP_LEN], one validated table set per 128-replica group. Only the multi-sweep TMA kernel
// reads per-group tables (3D, L a multiple of 8, whole 1024-replica slabs).
static rrrmc_status_t prepare_poisson_ladder(rrrmc_state *s, const uint32_t *tbls, int ngroups, int NW, uint64_t seed, cbp_run &run)
{
    RR_ARG(ngroups == (int)((s->W + 3) / 4), "expected one table set per 128-replica group (%d), given %d", (int)((s->W + 3) / 4), ngroups);
    if (!checkerboard_tma_eligible(s)) {
        rrrmc_set_error("a β ladder on the checkerboard schedule needs the brick kernel: D=3, L a multiple of 8, replicas a multiple of 1024 "
                        "(given L=%d, D=%d, R=%lld)", s->g->L, s->g->D, (long long)s->R);
        return RRRMC_ERR_UNSUPPORTED;
    }
    for (int gI = 1; gI < ngroups; gI++) RR_TRY(validate_cbp_tables(s, tbls + (size_t)gI * CBP_LEN, CBP_LEN, NW));
    RR_TRY(fill_cbp_params(s, tbls, CBP_LEN, NW, seed, run.p));
    run.tma = new cbt_params();
    RR_TRY(checkerboard_tma_prepare(s, run.p, *run.tma));
    run.groups.resize(ngroups);
    // the device copy of the ladder's tables survives between calls (a tempering loop sends the same ladder every round)
    cb_tma_store *c = s->tma;
    if (c->ladder_key.size() == (size_t)ngroups * CBP_LEN && memcmp(c->ladder_key.data(), tbls, sizeof(uint32_t) * CBP_LEN * ngroups) == 0) return RRRMC_OK;
    run.gbucket.resize((size_t)ngroups * CBP_BUCKETS);
    for (int gI = 0; gI < ngroups; gI++) {
        const uint32_t *T = tbls + (size_t)gI * CBP_LEN;
        cbp_group &G = run.groups[gI];
        memcpy(G.tbl, T, sizeof(uint32_t) * CBP_LEN);
        G.tb0_0 = T[CBP_KA]; G.tb0_1 = T[CBP_KA + 1]; G.tc0 = T[CBP_KA + 2 * CBP_KR]; G.pad = 0;
        cbp_build_bucket(T, run.gbucket.data() + (size_t)gI * CBP_BUCKETS);
    }
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_checkerboard_sweeps_poisson_ladder(rrrmc_state_t *s, const uint32_t *tbls, int ngroups, int NW,
                                                                   uint64_t seed, uint64_t sweep0, int64_t nsweeps)
{
    RR_ARG(s && tbls, "NULL argument");
    RR_ARG(nsweeps >= 0, "nsweeps must be >= 0");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    cbp_run run;
    RR_TRY(prepare_poisson_ladder(s, tbls, ngroups, NW, seed, run));
    RR_TRY(run_sweeps_poisson(s, run, sweep0, nsweeps));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_checkerboard_sweeps_poisson(rrrmc_state_t *s, const uint32_t *tbl, int tbl_len, int NW,
                                                            uint64_t seed, uint64_t sweep0, int64_t nsweeps)
{
    RR_ARG(s && tbl, "NULL argument");
    RR_ARG(nsweeps >= 0, "nsweeps must be >= 0");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    cbp_run run;
    RR_TRY(prepare_poisson_run(s, tbl, tbl_len, NW, seed, run));
    RR_TRY(run_sweeps_poisson(s, run, sweep0, nsweeps));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return RRRMC_OK;
}
// ---- parallel-tempering exchange on the device (tempering.cu)
extern "C" rrrmc_status_t rrrmc_tempering_exchange(rrrmc_state_t *s, const double *beta_group, int ngroups, uint64_t seed,
                                                   uint64_t round, int64_t *accepted)
{
    RR_ARG(s && beta_group, "NULL argument");
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    if (!(g->kind == RRRMC_EA_PM1 && g->d_jcode)) {
        rrrmc_set_error("rrrmc_tempering_exchange: implemented for ±J GraphEA lattices in the multispin layout");
        return RRRMC_ERR_UNSUPPORTED;
    }
    RR_ARG(s->R % 128 == 0, "the ladder lies over whole 128-replica groups: replicas must be a multiple of 128, given %lld", (long long)s->R);
    const int G = (int)(s->W / 4);
    RR_ARG(ngroups == G, "expected one β per 128-replica group (%d), given %d", G, ngroups);
    for (int k = 0; k < G; k++) RR_ARG(std::isfinite(beta_group[k]) && beta_group[k] >= 0, "β must be finite and >= 0, given: %g (group %d)", beta_group[k], k);
    RR_CUDA(cudaSetDevice(ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    if (!s->d_pt_beta) {
        RR_CUDA(cudaMalloc(&s->d_pt_beta, sizeof(double) * G));
        RR_CUDA(cudaMalloc(&s->d_pt_masks, sizeof(uint32_t) * 4 * G));
        RR_CUDA(cudaMalloc(&s->d_pt_acc, sizeof(long long) * G));
        RR_CUDA(cudaMemsetAsync(s->d_pt_acc, 0, sizeof(long long) * G, ctx->stream));
    }
    RR_CUDA(cudaMemsetAsync(s->d_pt_masks, 0, sizeof(uint32_t) * 4 * G, ctx->stream));
    // (a pageable source of a few bytes is staged before cudaMemcpyAsync returns: no synchronisation needed)
    RR_CUDA(cudaMemcpyAsync(s->d_pt_beta, beta_group, sizeof(double) * G, cudaMemcpyHostToDevice, ctx->stream));
    RR_TRY(launch_tempering_exchange(s, s->d_pt_beta, s->d_pt_masks, s->d_pt_acc, seed, round));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    if (accepted) {      // exchanges accepted per pair (g, g+1) since the last read; reading synchronises
        std::vector<long long> h(G);
        RR_CUDA(cudaMemcpyAsync(h.data(), s->d_pt_acc, sizeof(long long) * G, cudaMemcpyDeviceToHost, ctx->stream));
        RR_CUDA(cudaMemsetAsync(s->d_pt_acc, 0, sizeof(long long) * G, ctx->stream));
        RR_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int k = 0; k + 1 < G; k++) accepted[k] = (int64_t)h[k];
    }
    return RRRMC_OK;
}
// ---- continuous couplings (GraphEANormal): per-lane Float64 checkerboard sweeps (ea_normal.cu)
static rrrmc_status_t fill_cbn_params(rrrmc_state *s, const double *beta, uint64_t seed, cbn_params &p)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    if (!(g->kind == RRRMC_EA_F64 && g->L > 0 && g->d_Jd && g->d_A)) {
        rrrmc_set_error("continuous-coupling checkerboard sweeps need a GraphEANormal lattice");
        return RRRMC_ERR_UNSUPPORTED;
    }
    if (!g->bipartite) {
        rrrmc_set_error("checkerboard sweeps need even L (a two-colourable lattice), given L=%d; use schedule=RANDOM_SITE", g->L);
        return RRRMC_ERR_UNSUPPORTED;
    }
    RR_ARG(beta, "beta is NULL");
    for (int64_t r = 0; r < s->R; r++)
        RR_ARG(std::isfinite(beta[r]) && beta[r] >= 0, "β must be finite and >= 0, given: %g (replica %lld)", beta[r], (long long)r);
    RR_ARG(g->N * s->W < ((int64_t)1 << 40), "N*W too large");
    if (!s->d_beta) RR_CUDA(cudaMalloc(&s->d_beta, sizeof(double) * s->W * 32));
    RR_CUDA(cudaMemcpyAsync(s->d_beta, beta, sizeof(double) * s->R, cudaMemcpyHostToDevice, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));      // `beta` is a caller buffer
    memset(&p, 0, sizeof p);
    p.spins = s->d_spins; p.flips = nullptr; p.A = g->d_A; p.J = g->d_Jd; p.beta = s->d_beta;
    p.L = g->L; p.D = g->D; p.twoD = g->twoD; p.W = (int)s->W; p.R = s->R;
    p.nwg = (int)((s->W + 3) / 4); p.ntasks = (g->N / 2) * p.nwg;
    p.k0 = (uint32_t)seed; p.k1 = (uint32_t)(seed >> 32);
    return RRRMC_OK;
}
static rrrmc_status_t run_sweep_f64(rrrmc_state *s, cbn_params &p, uint64_t t)
{
    p.t_lo = (uint32_t)t; p.t_hi16 = (uint32_t)(t >> 32) << 16;
    RR_TRY(launch_checkerboard_f64(s->g->ctx, p, 0));
    RR_TRY(launch_checkerboard_f64(s->g->ctx, p, 1));
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_checkerboard_sweeps_f64(rrrmc_state_t *s, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps)
{
    RR_ARG(s, "state is NULL");
    RR_ARG(nsweeps >= 0, "nsweeps must be >= 0");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    RR_TRY(chain_sync_to_multispin(s));
    cbn_params p;
    RR_TRY(fill_cbn_params(s, beta, seed, p));
    for (int64_t k = 0; k < nsweeps; k++) RR_TRY(run_sweep_f64(s, p, sweep0 + (uint64_t)k));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    return RRRMC_OK;
}
// standardMC with schedule = CHECKERBOARD on a GraphEANormal lattice: whole sweeps, per-replica β, samples every
// ceil(step/N) sweeps (energies through the chain layout's sequential sum, which rounds like the reference's loop).
static rrrmc_status_t standard_mc_checkerboard_f64(rrrmc_state *s, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                                                   rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *o,
                                                   double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_TRY(chain_sync_to_multispin(s));
    cbn_params p;
    RR_TRY(fill_cbn_params(s, beta, seed, p));
    const int64_t N = g->N;
    const int64_t nsweeps = (iters + N - 1) / N, step_sw = std::max<int64_t>(1, (step + N - 1) / N);
    const bool count = o->count_accepted != 0;
    if (count) {
        if (!s->d_flips) RR_CUDA(cudaMalloc(&s->d_flips, sizeof(uint32_t) * N * s->W));
        p.flips = s->d_flips;
        RR_CUDA(cudaMemsetAsync(s->d_acc, 0, sizeof(long long) * s->W * 32, ctx->stream));
    }
    const uint64_t l0 = ctx->launches;
    std::vector<double> E(s->R);
    std::vector<long long> acc_h(s->W * 32);
    std::vector<int64_t> acc(s->R, -1);
    int64_t nsamples = 0, done = 0;
    event_pair ev;
    RR_CUDA(ev.create());
    RR_CUDA(cudaEventRecord(ev.e0, ctx->stream));
    for (int64_t sw = 1; sw <= nsweeps; sw++) {
        p.spins = s->d_spins;
        RR_TRY(run_sweep_f64(s, p, (uint64_t)(sw - 1)));
        s->ms_valid = true; s->chain_valid = false; s->chain_fields_valid = false; s->energy_valid = false;
        if (count) {
            // both colours wrote their accept masks into disjoint halves of d_flips during this sweep
            RR_TRY(launch_count_lanes(ctx, s->d_flips, N, (int)s->W, s->d_acc));
        }
        done = sw;
        if (sw % step_sw == 0 && (hook || (Es && nsamples < Es_cap))) {
            RR_TRY(chain_energy(s, E.data()));
            if (count) {
                RR_CUDA(cudaMemcpyAsync(acc_h.data(), s->d_acc, sizeof(long long) * s->W * 32, cudaMemcpyDeviceToHost, ctx->stream));
                RR_CUDA(cudaStreamSynchronize(ctx->stream));
                for (int64_t r = 0; r < s->R; r++) acc[r] = acc_h[r];
            }
            if (Es && nsamples < Es_cap) memcpy(Es + nsamples * s->R, E.data(), sizeof(double) * s->R);
            nsamples++;
            if (hook && !hook(user, sw * N, E.data(), acc.data(), s->R)) break;
            RR_TRY(chain_sync_to_multispin(s));      // a hook may have run a chain-layout query
        }
    }
    RR_CUDA(cudaEventRecord(ev.e1, ctx->stream));
    RR_CUDA(cudaEventSynchronize(ev.e1));
    float ms = 0; RR_CUDA(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    if (info) { info->nsamples = std::min(nsamples, Es ? Es_cap : nsamples); info->iters_done = done * N; info->launches = (int64_t)(ctx->launches - l0); info->device_ms = ms; info->accepted_total = -1; }
    return RRRMC_OK;
}

// AUTO: the sparse procedure wins while few lanes pass (expected passing lanes per 32-lane word <= 1.5)
static bool cb_use_sparse(const rrrmc_opts_t *o, double p1)
{
    if (o->cb_method == RRRMC_CB_SPARSE) return true;
    if (o->cb_method == RRRMC_CB_PLANES) return false;
    return 32.0 * p1 <= 1.5;
}

// β of every 128-replica group (the ±J acceptance procedures draw one hit count per 128-lane task, so β must be
// constant inside a group); ladder = the groups differ
static rrrmc_status_t group_betas(const rrrmc_state *s, const double *beta, std::vector<double> &gb, bool &ladder)
{
    RR_ARG(beta, "beta is NULL");
    gb.assign((size_t)((s->R + 127) / 128), 0.0);
    ladder = false;
    for (int64_t r = 0; r < s->R; r++) {
        RR_ARG(std::isfinite(beta[r]) && beta[r] >= 0, "β must be finite and >= 0, given: %g (replica %lld)", beta[r], (long long)r);
        if (r % 128 == 0) gb[(size_t)(r / 128)] = beta[r];
        else if (beta[r] != gb[(size_t)(r / 128)]) {
            rrrmc_set_error("checkerboard schedule on a ±J lattice: β must be constant inside each group of 128 consecutive replicas "
                            "(replica %lld has %g, its group %g)", (long long)r, beta[r], gb[(size_t)(r / 128)]);
            return RRRMC_ERR_UNSUPPORTED;
        }
        if (beta[r] != beta[0]) ladder = true;
    }
    return RRRMC_OK;
}

static rrrmc_status_t standard_mc_checkerboard(rrrmc_state *s, const std::vector<double> &gb, bool ladder, int64_t iters, int64_t step, uint64_t seed,
                                               rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *o,
                                               double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    if (!(g->kind == RRRMC_EA_PM1 && g->d_jcode)) {
        rrrmc_set_error("checkerboard sweeps need a ±J GraphEA lattice with D<=3; use schedule=RANDOM_SITE");
        return RRRMC_ERR_UNSUPPORTED;
    }
    if (!g->bipartite) {
        rrrmc_set_error("checkerboard sweeps need even L (a two-colourable lattice), given L=%d; use schedule=RANDOM_SITE", g->L);
        return RRRMC_ERR_UNSUPPORTED;
    }
    RR_TRY(chain_sync_to_multispin(s));
    const double beta = gb[0];
    uint64_t thr[3];
    for (int c = 1; c <= g->D; c++) thr[c - 1] = fixed64(exp(-beta * 4.0 * c));
    RR_ARG(o->cb_method >= RRRMC_CB_AUTO && o->cb_method <= RRRMC_CB_POISSON, "unknown cb_method %d", o->cb_method);
    cb_params p; cbs_params ps; cbp_run pp;
    // AUTO: poisson while its static position slots cover the level-1 hit count (β >~ 0.5), else sparse / planes
    uint32_t ptbl[CBP_LEN];
    RR_TRY(rrrmc_checkerboard_poisson_tables(thr, g->D, ptbl, CBP_LEN));
    int NW = rrrmc_checkerboard_poisson_nw(ptbl, 0.0);
    std::vector<uint32_t> ltbl;                         // β ladder: one table set per group, one NW (the warmest group's)
    if (ladder) {
        if (o->cb_method != RRRMC_CB_AUTO && o->cb_method != RRRMC_CB_POISSON) {
            rrrmc_set_error("a β ladder on the checkerboard schedule runs the poisson procedure only (cb_method AUTO or POISSON)");
            return RRRMC_ERR_UNSUPPORTED;
        }
        ltbl.resize(gb.size() * CBP_LEN);
        for (size_t gI = 0; gI < gb.size(); gI++) {
            uint64_t th[3];
            for (int c = 1; c <= g->D; c++) th[c - 1] = fixed64(exp(-gb[gI] * 4.0 * c));
            RR_TRY(rrrmc_checkerboard_poisson_tables(th, g->D, ltbl.data() + gI * CBP_LEN, CBP_LEN));
            const int nw = rrrmc_checkerboard_poisson_nw(ltbl.data() + gI * CBP_LEN, 0.0);
            if (nw == 0) {
                rrrmc_set_error("β ladder: β=%g (group %d) is too warm for the poisson procedure's static position slots", gb[gI], (int)gI);
                return RRRMC_ERR_UNSUPPORTED;
            }
            if (gI == 0 || nw > NW) NW = nw;
        }
    }
    if (o->cb_method == RRRMC_CB_POISSON && NW == 0) {
        rrrmc_set_error("cb_method POISSON: β=%g is too warm for the procedure's static position slots (use AUTO, SPARSE or PLANES)", beta);
        return RRRMC_ERR_UNSUPPORTED;
    }
    const bool poisson = ladder || o->cb_method == RRRMC_CB_POISSON || (o->cb_method == RRRMC_CB_AUTO && NW > 0);
    const bool sparse = !poisson && cb_use_sparse(o, exp(-beta * 4.0));
    if (ladder) RR_TRY(prepare_poisson_ladder(s, ltbl.data(), (int)gb.size(), NW, seed, pp));
    else if (poisson) RR_TRY(prepare_poisson_run(s, ptbl, CBP_LEN, NW, seed, pp));
    else if (sparse) {
        uint32_t tbl[CBS_T1 + 2 * CBS_TC];
        RR_TRY(rrrmc_checkerboard_sparse_tables(thr, g->D, tbl, CBS_T1 + 2 * CBS_TC));
        RR_TRY(fill_cbs_params(s, tbl, CBS_T1 + (g->D - 1) * CBS_TC, seed, ps));
    } else RR_TRY(fill_cb_params(s, thr, g->D, o->planes_K, o->planes_M, seed, p));
    const int64_t N = g->N;
    const int64_t nsweeps = (iters + N - 1) / N, step_sw = std::max<int64_t>(1, (step + N - 1) / N);
    const bool count = o->count_accepted != 0;
    if (count) {
        if (!s->d_flips) RR_CUDA(cudaMalloc(&s->d_flips, sizeof(uint32_t) * N * s->W));
        p.flips = ps.flips = pp.p.flips = s->d_flips;
        if (pp.tma) pp.tma->p.flips = s->d_flips;
        RR_CUDA(cudaMemsetAsync(s->d_acc, 0, sizeof(long long) * s->W * 32, ctx->stream));
    }
    const uint64_t l0 = ctx->launches;
    std::vector<double> E(s->R);
    std::vector<long long> acc_h(s->W * 32);
    std::vector<int64_t> acc(s->R, -1);
    int64_t nsamples = 0, done = 0;
    event_pair ev;                     // destroyed on every return path (a hook that stops the run, an error)
    RR_CUDA(ev.create());
    cudaEvent_t e0 = ev.e0, e1 = ev.e1;
    RR_CUDA(cudaEventRecord(e0, ctx->stream));
    // sweeps between two samples go out as one batch (one launch of the multi-sweep kernel) unless accepted moves
    // are counted, which reads the flip masks after every sweep
    const bool sample = hook || Es;
    for (int64_t sw = 1; sw <= nsweeps; sw++) {
        if (poisson && !count) {
            int64_t last = sample ? std::min(nsweeps, ((sw - 1) / step_sw + 1) * step_sw) : nsweeps;
            RR_TRY(run_sweeps_poisson(s, pp, (uint64_t)(sw - 1), last - sw + 1));
            sw = last;
        }
        else if (poisson) RR_TRY(run_sweeps_poisson(s, pp, (uint64_t)(sw - 1), 1));
        else if (sparse) RR_TRY(run_sweep_sparse(s, ps, (uint64_t)(sw - 1)));
        else RR_TRY(run_sweep(s, p, (uint64_t)(sw - 1)));
        if (count) RR_TRY(launch_count_lanes(ctx, s->d_flips, N, (int)s->W, s->d_acc));
        done = sw;
        if (sw % step_sw == 0 && (hook || (Es && nsamples < Es_cap))) {
            RR_TRY(energy_to_host(s, E.data()));
            if (count) {
                RR_CUDA(cudaMemcpyAsync(acc_h.data(), s->d_acc, sizeof(long long) * s->W * 32, cudaMemcpyDeviceToHost, ctx->stream));
                RR_CUDA(cudaStreamSynchronize(ctx->stream));
                for (int64_t r = 0; r < s->R; r++) acc[r] = acc_h[r];
            }
            if (Es && nsamples < Es_cap) memcpy(Es + nsamples * s->R, E.data(), sizeof(double) * s->R);
            nsamples++;
            if (hook && !hook(user, sw * N, E.data(), acc.data(), s->R)) break;
        }
    }
    RR_CUDA(cudaEventRecord(e1, ctx->stream));
    RR_CUDA(cudaEventSynchronize(e1));
    float ms = 0; RR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    s->energy_valid = false; s->chain_valid = false; s->chain_fields_valid = false; sk_dense_invalidate(s);
    if (info) { info->nsamples = std::min(nsamples, Es ? Es_cap : nsamples); info->iters_done = done * N; info->launches = (int64_t)(ctx->launches - l0); info->device_ms = ms; info->accepted_total = -1; }
    return RRRMC_OK;
}

extern "C" rrrmc_status_t rrrmc_standard_mc(rrrmc_state_t *s, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                                            rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *opts,
                                            double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s, "state is NULL");
    RR_ARG(iters >= 0, "iters must be >= 0, given %lld", (long long)iters);
    RR_ARG(step >= 1, "step must be >= 1, given %lld", (long long)step);
    rrrmc_opts_t o;
    if (opts) o = *opts; else rrrmc_opts_default(&o);
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    if (o.schedule == RRRMC_SCHED_CHECKERBOARD && s->g->kind == RRRMC_EA_F64 && s->g->L > 0)
        return standard_mc_checkerboard_f64(s, beta, iters, step, seed, hook, user, &o, Es, Es_cap, info);
    if (o.schedule == RRRMC_SCHED_CHECKERBOARD) {
        std::vector<double> gb; bool ladder = false;
        RR_TRY(group_betas(s, beta, gb, ladder));
        return standard_mc_checkerboard(s, gb, ladder, iters, step, seed, hook, user, &o, Es, Es_cap, info);
    }
    if (o.schedule == RRRMC_SCHED_RANDOM_SITE)
        return chain_run(s, CHAIN_STANDARD, beta, iters, step, seed, hook, user, &o, Es, Es_cap, info);
    rrrmc_set_error("unknown schedule %d", o.schedule);
    return RRRMC_ERR_ARG;
}

extern "C" rrrmc_status_t rrrmc_rrr_mc(rrrmc_state_t *s, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                                       rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *opts,
                                       double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s, "state is NULL");
    RR_ARG(iters >= 0 && step >= 1, "iters must be >= 0 and step >= 1");
    rrrmc_opts_t o;
    if (opts) o = *opts; else rrrmc_opts_default(&o);
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    return chain_run(s, CHAIN_RRR, beta, iters, step, seed, hook, user, &o, Es, Es_cap, info);
}
extern "C" rrrmc_status_t rrrmc_bkl_mc(rrrmc_state_t *s, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                                       rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *opts,
                                       double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s, "state is NULL");
    RR_ARG(iters >= 0 && step >= 1, "iters must be >= 0 and step >= 1");
    rrrmc_opts_t o;
    if (opts) o = *opts; else rrrmc_opts_default(&o);
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    return chain_run(s, CHAIN_BKL, beta, iters, step, seed, hook, user, &o, Es, Es_cap, info);
}
extern "C" rrrmc_status_t rrrmc_wtm_mc(rrrmc_state_t *s, const double *beta, int64_t samples, double step, uint64_t seed,
                                       rrrmc_hook_fn hook, void *user, double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s, "state is NULL");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    return chain_run_wtm(s, beta, samples, step, seed, hook, user, Es, Es_cap, info);
}
extern "C" rrrmc_status_t rrrmc_extremal_opt(rrrmc_state_t *s, const double *ftau, int64_t ftau_stride, int64_t iters, int64_t step,
                                             uint64_t seed, rrrmc_eo_hook_fn hook, void *user,
                                             double *Emin_out, int64_t *itmin_out, uint64_t *Cmin_chunks,
                                             double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s, "state is NULL");
    RR_ARG(iters >= 0 && step >= 1, "iters must be >= 0 and step >= 1");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    return chain_run_eo(s, ftau, ftau_stride, iters, step, seed, hook, user, Emin_out, itmin_out, Cmin_chunks, Es, Es_cap, info);
}
extern "C" rrrmc_status_t rrrmc_replay_wtm(rrrmc_state_t *s, int64_t replica, double beta, int64_t samples, double step,
                                           const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                                           double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s && kind && ival && fval, "NULL argument");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    std::vector<double> b(s->R, beta);
    const chain_trace_in tr{ replica, kind, ival, fval, ndraws };
    return chain_run_wtm(s, b.data(), samples, step, 1, nullptr, nullptr, Es, Es_cap, info, &tr);
}
extern "C" rrrmc_status_t rrrmc_replay_extremal_opt(rrrmc_state_t *s, int64_t replica, const double *ftau, int64_t iters, int64_t step,
                                                    const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                                                    double *Emin_out, int64_t *itmin_out, uint64_t *Cmin_chunks,
                                                    double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s && kind && ival && fval, "NULL argument");
    RR_ARG(iters >= 0 && step >= 1, "iters must be >= 0 and step >= 1");
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    const chain_trace_in tr{ replica, kind, ival, fval, ndraws };
    std::vector<double> emin(s->R); std::vector<int64_t> itmin(s->R); std::vector<uint64_t> cmin((size_t)s->R * s->nchunks);
    RR_TRY(chain_run_eo(s, ftau, 0, iters, step, 1, nullptr, nullptr, emin.data(), itmin.data(), cmin.data(), Es, Es_cap, info, &tr));
    if (Emin_out) *Emin_out = emin[replica];
    if (itmin_out) *itmin_out = itmin[replica];
    if (Cmin_chunks) memcpy(Cmin_chunks, cmin.data() + (size_t)replica * s->nchunks, 8 * s->nchunks);
    return RRRMC_OK;
}
extern "C" rrrmc_status_t rrrmc_replay(rrrmc_state_t *s, int64_t replica, int sampler, double beta, int64_t iters, int64_t step,
                                       const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                                       const rrrmc_opts_t *opts, double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    RR_ARG(s && kind && ival && fval, "NULL argument");
    RR_ARG(replica >= 0 && replica < s->R, "replica out of range");
    RR_ARG(sampler >= 0 && sampler <= 2, "sampler must be 0 (standardMC), 1 (rrrMC) or 2 (bklMC)");
    rrrmc_opts_t o;
    if (opts) o = *opts; else rrrmc_opts_default(&o);
    RR_CUDA(cudaSetDevice(s->g->ctx->device));
    if (info) memset(info, 0, sizeof *info);
    return chain_replay(s, replica, sampler, beta, iters, step, kind, ival, fval, ndraws, &o, Es, Es_cap, info);
}
