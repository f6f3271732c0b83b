// Sequential samplers on the chain layout: every replica is an independent Markov chain that runs the
// reference's loop with the reference's data structures (local-field cache, ΔE classes as ArraySets,
// Wong-Easton dynamic sampler), one chain per active lane.  Chains are latency-bound; they are spread one
// per warp across the SMs while R is small so that divergent chains never serialise each other.
//
// Reference map: standardMC RRRMC.jl:81-127; rrrMC RRRMC.jl:131-219; bklMC RRRMC.jl:294-359;
// DeltaECache DeltaE.jl:63-295; ArraySet ArraySets.jl:58-85; DeltaECacheCont DeltaE.jl:297-410;
// DynamicSampler DynamicSamplers.jl:84-176; GraphEA cache EA.jl:195-275, 584-663.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "chain.cuh"
#include "kernels.cuh"
#include "philox.cuh"

constexpr int MAXL = 16;  // |allΔE| supported by the discrete cache
constexpr int MAXDEG = 8; // 2D <= 8

struct chain_hdr {
    double E, acc_rate, z, pdE;
    double T[2 * MAXL + 1];
    long long it, accepted, staged_its, nextstep, skip, rng_n;
    int t[2 * MAXL + 1];
    int move_last, pending, pmove, status, built, trefresh, done, pad;
};

struct chain_store {
    int64_t R = 0, N = 0, N2 = 0;
    int levs = 0, nDE = 0;
    bool f64 = false;
    int32_t *lfi = nullptr, *lfi_last = nullptr;
    double *lfd = nullptr, *lfd_last = nullptr;
    chain_hdr *hdr = nullptr;
    int32_t *av = nullptr, *apos = nullptr;
    uint8_t *cls = nullptr;
    double *dEs = nullptr, *dv = nullptr, *dps = nullptr;
    double *d_Es = nullptr; int64_t Es_rows = 0;
    double *d_DE = nullptr, *d_beta = nullptr, *d_E = nullptr;
    uint8_t *d_tkind = nullptr; int64_t *d_tival = nullptr; double *d_tfval = nullptr; int64_t tcap = 0;
};

struct chain_params {
    int kind, N, twoD, sampler, nDE, levs, cpw;
    int64_t R, N2, nchunks, chain0;
    const int32_t *A; const int8_t *J8; const double *Jd;
    uint64_t *chunks;
    int32_t *lfi, *lfi_last; double *lfd, *lfd_last;
    chain_hdr *hdr;
    int32_t *av, *apos; uint8_t *cls;
    double *dEs, *dv, *dps;
    const double *DE, *beta;
    double *Es; int64_t Es_rows, quota;
    long long iters, step;
    uint64_t seed;
    double staged_thr, staged_thr_fact;
    const uint8_t *tkind; const int64_t *tival; const double *tfval; int64_t tlen;
};

// ------------------------------------------------------------------------------------------------
// draw sources
// ------------------------------------------------------------------------------------------------
struct src_philox {
    chain_rng r; int err;
    __device__ src_philox(const chain_params &P, int64_t chain, long long n) { r.seed = P.seed; r.chain = (uint64_t)chain; r.n = (uint64_t)n; r.tag = 0; err = 0; }
    __device__ double f64() { return r.f64(); }
    __device__ long long range(long long n) { return r.range(n); }
    __device__ long long pos() const { return (long long)r.n; }
};
struct src_trace { // typed draw stream dumped from the reference
    const uint8_t *kind; const int64_t *iv; const double *fv; long long p, len; int err;
    __device__ src_trace(const chain_params &P, int64_t, long long n) : kind(P.tkind), iv(P.tival), fv(P.tfval), p(n), len(P.tlen), err(0) {}
    __device__ double f64() { if (p >= len || kind[p] != 1) { err = 1; return 0.5; } return fv[p++]; }
    __device__ long long range(long long n) { if (p >= len || kind[p] != 0 || iv[p] < 1 || iv[p] > n) { err = 1; return 1; } return iv[p++]; }
    __device__ long long pos() const { return p; }
};

// ------------------------------------------------------------------------------------------------
// chain view of a GraphEA (int or Float64 couplings)
// ------------------------------------------------------------------------------------------------
struct cview {
    int f64, N, twoD;
    const int32_t *A; const int8_t *J8; const double *Jd;
    uint64_t *s;
    int32_t *lfi, *lfi_last; double *lfd, *lfd_last;
    int move_last; // 0-based site, -1 = none
};
__device__ __forceinline__ int sget(const uint64_t *s, int i) { return (int)((s[i >> 6] >> (i & 63)) & 1ull); }

__device__ __forceinline__ double cv_delta_energy(const cview &c, int i) // EA.jl:266-275 / :655-663
{
    return c.f64 ? -c.lfd[i] : -(double)c.lfi[i];
}
__device__ int cv_neighbors(const cview &c, int i, int *out) // uA[i], EA.jl:292
{
    int n = 0;
    for (int k = 0; k < c.twoD; k++) {
        const int y = c.A[(int64_t)i * c.twoD + k];
        if (n == 0 || out[n - 1] != y) out[n++] = y;
    }
    return n;
}
__device__ void cv_spinflip(cview &c, int i) // Interface.jl:89-92 + update_cache! EA.jl:224-264 / :613-653
{
    c.s[i >> 6] ^= 1ull << (i & 63);
    int U[MAXDEG]; const int nU = cv_neighbors(c, i, U);
    if (c.f64) {
        if (c.move_last == i) {
            for (int k = 0; k < nU; k++) { const double t = c.lfd[U[k]]; c.lfd[U[k]] = c.lfd_last[U[k]]; c.lfd_last[U[k]] = t; }
            c.lfd[i] = -c.lfd[i]; c.lfd_last[i] = -c.lfd_last[i];
            return;
        }
        for (int k = 0; k < nU; k++) c.lfd_last[U[k]] = c.lfd[U[k]];
        const int sx = sget(c.s, i);
        for (int k = 0; k < c.twoD; k++) {
            const int y = c.A[(int64_t)i * c.twoD + k];
            const double f = (double)(4 * (1 - 2 * (sx ^ sget(c.s, y))));
            c.lfd[y] = __dsub_rn(c.lfd[y], __dmul_rn(f, c.Jd[(int64_t)i * c.twoD + k]));
        }
        const double lfm = c.lfd[i];
        c.lfd_last[i] = lfm; c.lfd[i] = -lfm;
    } else {
        if (c.move_last == i) {
            for (int k = 0; k < nU; k++) { const int t = c.lfi[U[k]]; c.lfi[U[k]] = c.lfi_last[U[k]]; c.lfi_last[U[k]] = t; }
            c.lfi[i] = -c.lfi[i]; c.lfi_last[i] = -c.lfi_last[i];
            return;
        }
        for (int k = 0; k < nU; k++) c.lfi_last[U[k]] = c.lfi[U[k]];
        const int sx = sget(c.s, i);
        for (int k = 0; k < c.twoD; k++) {
            const int y = c.A[(int64_t)i * c.twoD + k];
            c.lfi[y] -= 4 * (1 - 2 * (sx ^ sget(c.s, y))) * (int)c.J8[(int64_t)i * c.twoD + k];
        }
        const int lfm = c.lfi[i];
        c.lfi_last[i] = lfm; c.lfi[i] = -lfm;
    }
    c.move_last = i;
}

// ------------------------------------------------------------------------------------------------
// discrete ΔE-class cache (DeltaE.jl:63-295) with ArraySets (ArraySets.jl:58-85)
// ------------------------------------------------------------------------------------------------
struct dcache {
    int N, L;
    const double *DE;
    double ft[MAXL];
    double Ta[2 * MAXL + 1], Tb[2 * MAXL + 1];
    double *T, *Tp;
    double z, zp;
    int *t;            // class sizes (hdr)
    int32_t *av, *apos; uint8_t *cls;
    int st[MAXDEG + 1][3], nst;
};
__device__ __forceinline__ int dc_findk(const dcache &c, double dE) // DeltaE.jl:28-60
{
    dE = fabs(dE);
    for (int k = 1; k <= c.L; k++) if (c.DE[k - 1] == dE) return k;
    return 0;
}
__device__ __forceinline__ double dc_f(const dcache &c, int k) { return k > c.L ? c.ft[k - c.L - 1] : 1.0; }
__device__ __forceinline__ void as_push(dcache &c, int k, int i) { c.av[(int64_t)(k - 1) * c.N + c.t[k]] = i; c.t[k]++; c.apos[i] = c.t[k]; }
__device__ __forceinline__ void as_delete(dcache &c, int k, int i)
{
    const int p = c.apos[i];
    const int last = c.av[(int64_t)(k - 1) * c.N + c.t[k] - 1];
    c.av[(int64_t)(k - 1) * c.N + p - 1] = last;
    c.apos[last] = p;
    c.apos[i] = 0;
    c.t[k]--;
}
__device__ __forceinline__ int dc_class_of(const dcache &c, const cview &X, int j)
{
    const double dE = cv_delta_energy(X, j);
    const int up = dE > 0 || (dE == 0 && sget(X.s, j) == 1);
    return dc_findk(c, dE) + c.L * up;
}
__device__ void dc_build(dcache &c, const cview &X, double beta) // DeltaE.jl:74-104
{
    for (int k = 0; k <= 2 * c.L; k++) c.t[k] = 0;
    for (int i = 0; i < c.N; i++) {
        const int ki = dc_class_of(c, X, i);
        c.cls[i] = (uint8_t)ki;
        as_push(c, ki, i);
    }
    c.z = 0.0;
    for (int k = 1; k <= 2 * c.L; k++) { const double x = (double)c.t[k] * dc_f(c, k); c.z += x; c.T[k] = x; }
    c.zp = c.z;
}
template <class SRC> __device__ long long dc_rand_skip(const dcache &c, SRC &d) // DeltaE.jl:141-144
{
    return (long long)floor(log1p(-d.f64()) / log1p(-c.z / (double)c.N));
}
template <class SRC> __device__ int dc_rand_move(const dcache &c, SRC &d, double &dE) // DeltaE.jl:146-167
{
    const int L = c.L;
    const double r = d.f64() * c.z;
    double cT = 0.0;
    int k = 1; bool broke = false;
    for (; k <= 2 * L; k++) { cT += c.T[k]; if (r < cT) { broke = true; break; } }
    if (!broke) k = 2 * L;
    if (!(r < cT)) while (c.T[k] == 0) k--;
    dE = k <= L ? -c.DE[k - 1] : c.DE[k - L - 1];
    const long long p = d.range(c.t[k]);
    return c.av[(int64_t)(k - 1) * c.N + p - 1];
}
__device__ void dc_compute_staged(dcache &c, cview &X, int i) // DeltaE.jl:202-230
{
    cv_spinflip(X, i);
    c.nst = 0;
    int nb[MAXDEG]; const int n = cv_neighbors(X, i, nb);
    for (int a = 0; a < n; a++) {
        const int j = nb[a], k0 = c.cls[j], k1 = dc_class_of(c, X, j);
        if (k0 == k1) continue;
        c.st[c.nst][0] = j; c.st[c.nst][1] = k0; c.st[c.nst][2] = k1; c.nst++;
    }
    const int k0 = c.cls[i], k1 = k0 - c.L * (2 * (k0 > c.L) - 1);
    c.st[c.nst][0] = i; c.st[c.nst][1] = k0; c.st[c.nst][2] = k1; c.nst++;
    cv_spinflip(X, i);
}
__device__ double dc_reverse(dcache &c) // DeltaE.jl:184-200
{
    double zp = c.z;
    for (int k = 0; k <= 2 * c.L; k++) c.Tp[k] = c.T[k];
    for (int a = 0; a < c.nst; a++) {
        const int k0 = c.st[a][1], k1 = c.st[a][2];
        const double f0 = dc_f(c, k0), f1 = dc_f(c, k1);
        c.Tp[k0] -= f0; c.Tp[k1] += f1;
        zp += f1 - f0;
    }
    c.zp = zp;
    return zp;
}
__device__ void dc_apply_staged(dcache &c) // DeltaE.jl:169-182
{
    for (int a = 0; a < c.nst; a++) {
        const int j = c.st[a][0], k0 = c.st[a][1], k1 = c.st[a][2];
        as_delete(c, k0, j); as_push(c, k1, j); c.cls[j] = (uint8_t)k1;
    }
    double *tmp = c.T; c.T = c.Tp; c.Tp = tmp; c.z = c.zp;
}
__device__ double dc_apply_move(dcache &c, cview &X, int move) // DeltaE.jl:232-295
{
    cv_spinflip(X, move);
    double zp = c.z;
    int nb[MAXDEG]; const int n = cv_neighbors(X, move, nb);
    for (int a = 0; a <= n; a++) {
        int j, k0, k1;
        if (a < n) { j = nb[a]; k0 = c.cls[j]; k1 = dc_class_of(c, X, j); if (k0 == k1) continue; }
        else { j = move; k0 = c.cls[move]; k1 = k0 - c.L * (2 * (k0 > c.L) - 1); }
        const double f0 = dc_f(c, k0), f1 = dc_f(c, k1);
        c.T[k0] -= f0; c.T[k1] += f1;
        zp += f1 - f0;
        as_delete(c, k0, j); as_push(c, k1, j); c.cls[j] = (uint8_t)k1;
    }
    const double cc = c.z / zp;
    c.z = zp;
    return cc;
}

// ------------------------------------------------------------------------------------------------
// continuous cache (DeltaE.jl:297-410) on the Wong-Easton sampler (DynamicSamplers.jl:84-176)
// ------------------------------------------------------------------------------------------------
struct ccache {
    int N, levs; long long N2;
    double *v, *ps, *dEs;  // v,ps 1-based
    double z, beta;
    int trefresh;
    int sj[MAXDEG + 1]; double sdE[MAXDEG + 1], sp[MAXDEG + 1]; int nst;
};
__device__ __forceinline__ double prior(double x) { return x > 0 ? exp(-x) : 1.0; } // DeltaE.jl:297
__device__ void ds_add_path(ccache &c, int i1, double x)
{
    long long k = 0, off = 1, u = c.levs > 0 ? 1ll << (c.levs - 1) : 0; const long long i0 = i1 - 1;
    for (int lev = 1; lev <= c.levs; lev++) {
        if ((i0 & u) == 0) { c.ps[off + k] += x; k *= 2; } else k = 2 * k + 1;
        u >>= 1; off *= 2;
    }
}
__device__ void ds_refresh(ccache &c) // DynamicSamplers.jl:84-98
{
    double z = 0.0;
    for (long long i = 1; i <= c.N2; i++) z += c.v[i];
    c.z = z;
    for (long long i = 0; i <= c.N2; i++) c.ps[i] = 0.0;
    for (int i = 1; i <= c.N; i++) ds_add_path(c, i, c.v[i]);
    c.trefresh = 0;
}
__device__ int ds_getel(ccache &c, double x, int &err) // DynamicSamplers.jl:130-152
{
    for (int guard = 0; guard < 3; guard++) {
        x *= c.z;
        long long k = 0, off = 1;
        for (int lev = 1; lev <= c.levs; lev++) {
            const double p = c.ps[off + k];
            k *= 2;
            if (x > p) { x -= p; k += 1; }
            off *= 2;
        }
        if (k >= c.N || c.v[k + 1] == 0) {
            if (!(c.trefresh > 0)) { err = 2; return 1; }
            ds_refresh(c);
            continue; // sic: the reference re-enters with the scaled residual x
        }
        return (int)k + 1;
    }
    err = 2;
    return 1;
}
__device__ void ds_set(ccache &c, int i1, double x) // DynamicSamplers.jl:159-176
{
    if (c.trefresh >= (c.N > 100 ? c.N : 100)) ds_refresh(c);
    c.trefresh++;
    const double d = x - c.v[i1];
    c.v[i1] = x;
    c.z += d;
    ds_add_path(c, i1, d);
}
__device__ void cc_build(ccache &c, const cview &X) // DeltaE.jl:304-311 + DynamicSamplers.jl:35-51
{
    for (long long i = 0; i <= c.N2; i++) c.v[i] = 0.0;
    for (int i = 0; i < c.N; i++) { c.dEs[i] = cv_delta_energy(X, i); c.v[i + 1] = prior(c.beta * c.dEs[i]); }
    ds_refresh(c);
}
template <class SRC> __device__ long long cc_rand_skip(const ccache &c, SRC &d) // DeltaE.jl:319-324
{
    double b = c.z / (double)c.N;
    b = fmin(fmax(b, 2.2250738585072014e-308), 1.0);
    return (long long)floor(log1p(-d.f64()) / log1p(-b));
}
__device__ void cc_compute_staged(ccache &c, cview &X, int i) // DeltaE.jl:356-373
{
    cv_spinflip(X, i);
    double dE = cv_delta_energy(X, i);
    c.sj[0] = i; c.sdE[0] = dE; c.sp[0] = prior(c.beta * dE); c.nst = 1;
    int nb[MAXDEG]; const int n = cv_neighbors(X, i, nb);
    for (int a = 0; a < n; a++) {
        dE = cv_delta_energy(X, nb[a]);
        c.sj[c.nst] = nb[a]; c.sdE[c.nst] = dE; c.sp[c.nst] = prior(c.beta * dE); c.nst++;
    }
    cv_spinflip(X, i);
}
__device__ double cc_reverse(const ccache &c) // DeltaE.jl:344-354
{
    double z = c.z;
    for (int a = 0; a < c.nst; a++) z += c.sp[a] - c.v[c.sj[a] + 1];
    return fmin(fmax(z, 2.2250738585072014e-308), (double)c.N);
}
__device__ void cc_apply_staged(ccache &c) // DeltaE.jl:334-342
{
    for (int a = 0; a < c.nst; a++) { c.dEs[c.sj[a]] = c.sdE[a]; ds_set(c, c.sj[a] + 1, c.sp[a]); }
}
__device__ double cc_apply_move(ccache &c, cview &X, int move) // DeltaE.jl:378-410
{
    cv_spinflip(X, move);
    const double z = c.z;
    double dE = cv_delta_energy(X, move);
    c.dEs[move] = dE; ds_set(c, move + 1, prior(c.beta * dE));
    int nb[MAXDEG]; const int n = cv_neighbors(X, move, nb);
    for (int a = 0; a < n; a++) {
        dE = cv_delta_energy(X, nb[a]);
        c.dEs[nb[a]] = dE; ds_set(c, nb[a] + 1, prior(c.beta * dE));
    }
    return z / c.z;
}

// ------------------------------------------------------------------------------------------------
// the sampler kernel (resumable: pauses after `quota` samples so that the host can run the hook)
// ------------------------------------------------------------------------------------------------
template <class SRC>
__global__ void __launch_bounds__(32) k_chain_run(chain_params P)
{
    const int lane = threadIdx.x;
    if (lane >= P.cpw) return;
    const int64_t r = P.chain0 + (int64_t)blockIdx.x * P.cpw + lane;
    if (r >= P.chain0 + P.R) return;
    chain_hdr h = P.hdr[r];
    if (h.done) return;
    const int N = P.N;
    cview X;
    X.f64 = P.kind == RRRMC_EA_F64; X.N = N; X.twoD = P.twoD; X.A = P.A; X.J8 = P.J8; X.Jd = P.Jd;
    X.s = P.chunks + r * P.nchunks;
    X.lfi = P.lfi ? P.lfi + r * N : nullptr; X.lfi_last = P.lfi_last ? P.lfi_last + r * N : nullptr;
    X.lfd = P.lfd ? P.lfd + r * N : nullptr; X.lfd_last = P.lfd_last ? P.lfd_last + r * N : nullptr;
    X.move_last = h.move_last;
    SRC src(P, r, h.rng_n);
    const double beta = P.beta[r];
    const bool discr = !X.f64;
    long long emitted = 0;
    const long long iters = P.iters, step = P.step;
    double *Es = P.Es;

    dcache dc; ccache cc;
    if (P.sampler != CHAIN_STANDARD) {
        if (discr) {
            dc.N = N; dc.L = P.nDE; dc.DE = P.DE; dc.t = h.t; dc.T = dc.Ta; dc.Tp = dc.Tb;
            dc.av = P.av + r * (int64_t)(2 * P.nDE) * N; dc.apos = P.apos + r * N; dc.cls = P.cls + r * N;
            for (int k = 0; k < dc.L; k++) dc.ft[k] = exp(-beta * dc.DE[k]);
            for (int k = 0; k <= 2 * dc.L; k++) dc.T[k] = h.T[k];
            dc.z = h.z; dc.zp = h.z; dc.nst = 0;
            if (!h.built) { dc_build(dc, X, beta); h.built = 1; }
        } else {
            cc.N = N; cc.levs = P.levs; cc.N2 = P.N2; cc.beta = beta;
            cc.v = P.dv + r * (P.N2 + 1); cc.ps = P.dps + r * (P.N2 + 1); cc.dEs = P.dEs + r * N;
            cc.z = h.z; cc.trefresh = h.trefresh; cc.nst = 0;
            if (!h.built) { cc_build(cc, X); h.built = 1; }
        }
    }
#define EMIT_SAMPLE()                                                         \
    do {                                                                      \
        if (Es && emitted < P.Es_rows) Es[emitted * P.R + (r - P.chain0)] = h.E; \
        emitted++;                                                            \
    } while (0)

    if (P.sampler == CHAIN_STANDARD) { // RRRMC.jl:100-119
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { h.done = 1; break; }
                h.it++;
                if (h.it % step == 0) { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 1; break; } }
            }
            h.pending = 0;
            const int i = (int)src.range(N) - 1;
            const double dE = cv_delta_energy(X, i);
            const double x = -beta * dE;
            if (!(x >= 0 || src.f64() < exp(x))) continue; // accept(), RRRMC.jl:39
            cv_spinflip(X, i);
            h.E += dE;
            h.accepted++;
        }
    } else if (P.sampler == CHAIN_RRR) { // RRRMC.jl:180-211
        const double lambda = P.staged_thr_fact / (double)N;
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { h.done = 1; break; }
                h.it++;
                if (h.it % step == 0) { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 1; break; } }
            }
            h.pending = 0;
            int acc = 0;
            if (h.acc_rate < P.staged_thr) {
                h.staged_its++;
                double z, zp, dE; int move;
                if (discr) { z = dc.z; move = dc_rand_move(dc, src, dE); dc_compute_staged(dc, X, move); zp = dc_reverse(dc); }
                else { z = cc.z; move = ds_getel(cc, src.f64(), src.err) - 1; dE = cc.dEs[move]; cc_compute_staged(cc, X, move); zp = cc_reverse(cc); }
                const double c = z / zp;
                if (src.f64() < c) {
                    cv_spinflip(X, move);
                    if (discr) dc_apply_staged(dc); else cc_apply_staged(cc);
                    h.E += dE; h.accepted++; acc = 1;
                }
            } else {
                double dE; int move;
                if (discr) move = dc_rand_move(dc, src, dE); else { move = ds_getel(cc, src.f64(), src.err) - 1; dE = cc.dEs[move]; }
                const double c = discr ? dc_apply_move(dc, X, move) : cc_apply_move(cc, X, move);
                if (src.f64() < c) { h.E += dE; h.accepted++; acc = 1; }
                else { if (discr) dc_apply_move(dc, X, move); else cc_apply_move(cc, X, move); }
            }
            h.acc_rate = h.acc_rate * (1 - lambda) + acc * lambda;
            if (src.err) { h.done = 1; break; }
        }
    } else { // bklMC, RRRMC.jl:332-350
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { h.done = 1; break; }
                h.skip = discr ? dc_rand_skip(dc, src) : cc_rand_skip(cc, src);
                if (discr) h.pmove = dc_rand_move(dc, src, h.pdE); else { h.pmove = ds_getel(cc, src.f64(), src.err) - 1; h.pdE = cc.dEs[h.pmove]; }
                h.pending = 1;
            }
            bool out = false, paused = false;
            while (h.it + h.skip + 1 >= h.nextstep) {
                if (h.pending == 2) h.pending = 1; // resuming right after the hook of this sample
                else { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 2; paused = true; break; } }
                h.nextstep += step;
                if (h.nextstep > iters) { out = true; break; }
            }
            if (paused) break;
            if (out || src.err) { h.done = 1; break; }
            if (discr) dc_apply_move(dc, X, h.pmove); else cc_apply_move(cc, X, h.pmove);
            h.it += h.skip + 1;
            h.E += h.pdE;
            h.accepted++;
            h.pending = 0;
        }
    }
#undef EMIT_SAMPLE
    if (P.sampler != CHAIN_STANDARD) {
        if (discr) { for (int k = 0; k <= 2 * dc.L; k++) h.T[k] = dc.T[k]; h.z = dc.z; }
        else { h.z = cc.z; h.trefresh = cc.trefresh; }
    }
    h.move_last = X.move_last;
    h.rng_n = src.pos();
    if (src.err) h.status = src.err;
    P.hdr[r] = h;
}

// ------------------------------------------------------------------------------------------------
// energy(X, C) on the chain layout: local fields (EA.jl:201-215 / :591-605), then the per-chain sum in
// site order (sequential, so Float64 energies round exactly like the reference's loop)
// ------------------------------------------------------------------------------------------------
__global__ void k_chain_lfields(chain_params P)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.R * P.N) return;
    const int64_t r = tid / P.N; const int x = (int)(tid % P.N);
    const uint64_t *s = P.chunks + r * P.nchunks;
    const int sx = 2 * sget(s, x) - 1;
    if (P.kind == RRRMC_EA_F64) {
        double lf = 0.0;
        for (int k = 0; k < P.twoD; k++) {
            const int y = P.A[(int64_t)x * P.twoD + k];
            const double sy = (double)(2 * sget(s, y) - 1);
            lf = __dsub_rn(lf, __dmul_rn(__dmul_rn(P.Jd[(int64_t)x * P.twoD + k], (double)sx), sy));
        }
        P.lfd[tid] = 2 * lf; P.lfd_last[tid] = 0.0;
    } else {
        int lf = 0;
        for (int k = 0; k < P.twoD; k++) {
            const int y = P.A[(int64_t)x * P.twoD + k];
            lf -= (int)P.J8[(int64_t)x * P.twoD + k] * sx * (2 * sget(s, y) - 1);
        }
        P.lfi[tid] = 2 * lf; P.lfi_last[tid] = 0;
    }
}
__global__ void k_chain_energy_sum(chain_params P, double *E_out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) return;
    double E;
    if (P.kind == RRRMC_EA_F64) {
        double e = 0.0;
        for (int x = 0; x < P.N; x++) e = __dadd_rn(e, P.lfd[r * P.N + x] / 2);
        E = e / 2;
    } else {
        long long n = 0;
        for (int x = 0; x < P.N; x++) n += P.lfi[r * P.N + x] / 2;
        E = (double)n / 2.0;
    }
    E_out[r] = E;
    chain_hdr &h = P.hdr[r];
    h.E = E; h.move_last = -1;
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
}
// naive ΔE straight from the spins (EA.jl:277-289 commented form == -lfields of a fresh cache)
__global__ void k_chain_delta_site(chain_params P, int site, double *out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.R) return;
    const uint64_t *s = P.chunks + r * P.nchunks;
    const int sx = 2 * sget(s, site) - 1;
    double lf = 0.0;
    for (int k = 0; k < P.twoD; k++) {
        const int y = P.A[(int64_t)site * P.twoD + k];
        const double J = P.kind == RRRMC_EA_F64 ? P.Jd[(int64_t)site * P.twoD + k] : (double)P.J8[(int64_t)site * P.twoD + k];
        lf = __dsub_rn(lf, __dmul_rn(__dmul_rn(J, (double)sx), (double)(2 * sget(s, y) - 1)));
    }
    out[r] = -(2 * lf);
}
__global__ void k_chain_delta_replica(chain_params P, int64_t r, double *out)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= P.N) return;
    const uint64_t *s = P.chunks + r * P.nchunks;
    const int sx = 2 * sget(s, x) - 1;
    double lf = 0.0;
    for (int k = 0; k < P.twoD; k++) {
        const int y = P.A[(int64_t)x * P.twoD + k];
        const double J = P.kind == RRRMC_EA_F64 ? P.Jd[(int64_t)x * P.twoD + k] : (double)P.J8[(int64_t)x * P.twoD + k];
        lf = __dsub_rn(lf, __dmul_rn(__dmul_rn(J, (double)sx), (double)(2 * sget(s, y) - 1)));
    }
    out[x] = -(2 * lf);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void chain_free(rrrmc_state *s)
{
    chain_store *c = s->chain;
    if (!c) return;
    cudaFree(c->lfi); cudaFree(c->lfi_last); cudaFree(c->lfd); cudaFree(c->lfd_last); cudaFree(c->hdr);
    cudaFree(c->av); cudaFree(c->apos); cudaFree(c->cls); cudaFree(c->dEs); cudaFree(c->dv); cudaFree(c->dps);
    cudaFree(c->d_Es); cudaFree(c->d_DE); cudaFree(c->d_beta); cudaFree(c->d_E);
    cudaFree(c->d_tkind); cudaFree(c->d_tival); cudaFree(c->d_tfval);
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

static rrrmc_status_t chain_ensure(rrrmc_state *s, bool need_cache)
{
    rrrmc_graph *g = s->g;
    RR_ARG(g->twoD <= MAXDEG, "2D = %d exceeds the chain kernels' limit %d", g->twoD, MAXDEG);
    if (!s->chain) {
        chain_store *c = new chain_store();
        c->R = s->R; c->N = g->N; c->f64 = g->kind == RRRMC_EA_F64; c->nDE = (int)g->allDE.size();
        s->chain = c;
        const size_t RN = (size_t)s->R * g->N;
        if (c->f64) { RR_CUDA(cudaMalloc(&c->lfd, RN * 8)); RR_CUDA(cudaMalloc(&c->lfd_last, RN * 8)); }
        else { RR_CUDA(cudaMalloc(&c->lfi, RN * 4)); RR_CUDA(cudaMalloc(&c->lfi_last, RN * 4)); }
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
    if (need_cache) {
        const size_t RN = (size_t)s->R * g->N;
        if (!c->f64 && !c->av) {
            RR_ARG(c->nDE >= 1 && c->nDE <= MAXL, "|allΔE| = %d exceeds the discrete cache limit %d", c->nDE, MAXL);
            RR_CUDA(cudaMalloc(&c->av, RN * 4 * 2 * c->nDE));
            RR_CUDA(cudaMalloc(&c->apos, RN * 4));
            RR_CUDA(cudaMalloc(&c->cls, RN));
        }
        if (c->f64 && !c->dv) {
            c->levs = 0; while (((int64_t)1 << c->levs) < g->N) c->levs++;
            c->N2 = (int64_t)1 << c->levs;
            RR_CUDA(cudaMalloc(&c->dEs, RN * 8));
            RR_CUDA(cudaMalloc(&c->dv, (size_t)s->R * (c->N2 + 1) * 8));
            RR_CUDA(cudaMalloc(&c->dps, (size_t)s->R * (c->N2 + 1) * 8));
        }
    }
    return RRRMC_OK;
}

static void chain_fill_params(rrrmc_state *s, chain_params &P)
{
    rrrmc_graph *g = s->g; chain_store *c = s->chain;
    memset(&P, 0, sizeof P);
    P.kind = g->kind; P.N = (int)g->N; P.twoD = g->twoD; P.nDE = c->nDE; P.levs = c->levs; P.N2 = c->N2;
    P.R = s->R; P.nchunks = s->nchunks; P.chain0 = 0;
    P.A = g->d_A; P.J8 = g->d_J8; P.Jd = g->d_Jd;
    P.chunks = s->d_chunks;
    P.lfi = c->lfi; P.lfi_last = c->lfi_last; P.lfd = c->lfd; P.lfd_last = c->lfd_last;
    P.hdr = c->hdr; P.av = c->av; P.apos = c->apos; P.cls = c->cls;
    P.dEs = c->dEs; P.dv = c->dv; P.dps = c->dps; P.DE = c->d_DE; P.beta = c->d_beta;
    P.step = 1;
}

static rrrmc_status_t chain_energy_init(rrrmc_state *s, chain_params &P)
{
    rrrmc_ctx *ctx = s->g->ctx;
    k_chain_lfields<<<div_up(P.R * P.N, 256), 256, 0, ctx->stream>>>(P);
    k_chain_energy_sum<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, s->chain->d_E);
    ctx->launches += 2;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

rrrmc_status_t chain_energy(rrrmc_state *s, double *E_out)
{
    RR_TRY(chain_ensure(s, false));
    RR_TRY(chain_sync_from_multispin(s));
    chain_params P; chain_fill_params(s, P);
    RR_TRY(chain_energy_init(s, P));
    RR_CUDA(cudaMemcpyAsync(E_out, s->chain->d_E, 8 * s->R, cudaMemcpyDeviceToHost, s->g->ctx->stream));
    RR_CUDA(cudaStreamSynchronize(s->g->ctx->stream));
    return RRRMC_OK;
}
rrrmc_status_t chain_delta_energy_site(rrrmc_state *s, int64_t site0, double *out)
{
    RR_TRY(chain_ensure(s, false));
    RR_TRY(chain_sync_from_multispin(s));
    chain_params P; chain_fill_params(s, P);
    rrrmc_ctx *ctx = s->g->ctx;
    k_chain_delta_site<<<div_up(P.R, 128), 128, 0, ctx->stream>>>(P, (int)site0, s->chain->d_E);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    RR_CUDA(cudaMemcpyAsync(out, s->chain->d_E, 8 * s->R, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    return RRRMC_OK;
}
rrrmc_status_t chain_delta_energy_replica(rrrmc_state *s, int64_t replica, double *out)
{
    RR_TRY(chain_ensure(s, false));
    RR_TRY(chain_sync_from_multispin(s));
    chain_params P; chain_fill_params(s, P);
    rrrmc_ctx *ctx = s->g->ctx;
    double *d_tmp = nullptr;
    RR_CUDA(cudaMalloc(&d_tmp, 8 * P.N));
    k_chain_delta_replica<<<div_up(P.N, 128), 128, 0, ctx->stream>>>(P, replica, d_tmp);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    RR_CUDA(cudaMemcpyAsync(out, d_tmp, 8 * P.N, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_tmp);
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
    P.cpw = (int)std::min<int64_t>(32, std::max<int64_t>(1, (P.R + warps - 1) / warps));
    const unsigned grid = div_up(P.R, P.cpw);
    std::vector<double> row((size_t)P.R * rows_per_launch);
    std::vector<chain_hdr> hh(P.R);
    std::vector<int64_t> acc(P.R);
    const uint64_t l0 = ctx->launches;
    cudaEvent_t e0, e1;
    RR_CUDA(cudaEventCreate(&e0)); RR_CUDA(cudaEventCreate(&e1));
    RR_CUDA(cudaEventRecord(e0, ctx->stream));
    int64_t nsamples = 0; bool stop = false;
    for (int guard = 0; !stop; guard++) {
        k_chain_run<SRC><<<grid, 32, 0, ctx->stream>>>(P);
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
                for (int64_t r = 0; r < P.R; r++) acc[r] = hh[r].accepted;
                if (!hook(user, nsamples * P.step, row.data() + k * P.R, acc.data(), P.R)) stop = true;
            }
        }
        if (all_done) break;
    }
    RR_CUDA(cudaEventRecord(e1, ctx->stream));
    RR_CUDA(cudaEventSynchronize(e1));
    float ms = 0; RR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (info) {
        int64_t itmax = 0; for (int64_t r = 0; r < P.R; r++) itmax = std::max<int64_t>(itmax, hh[r].it);
        info->nsamples = std::min(nsamples, want_rows); info->iters_done = itmax;
        info->launches = (int64_t)(ctx->launches - l0); info->device_ms = ms;
    }
    return RRRMC_OK;
}

rrrmc_status_t chain_run(rrrmc_state *s, int sampler, const double *beta, int64_t iters, int64_t step, uint64_t seed,
                         rrrmc_hook_fn hook, void *user, const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(beta, "beta is NULL");
    for (int64_t r = 0; r < s->R; r++) RR_ARG(std::isfinite(beta[r]), "β must be finite, given: %g", beta[r]); // RRRMC.jl:159
    RR_TRY(chain_ensure(s, sampler != CHAIN_STANDARD));
    RR_TRY(chain_sync_from_multispin(s));
    chain_store *c = s->chain;
    chain_params P; chain_fill_params(s, P);
    P.sampler = sampler; P.iters = iters; P.step = step; P.seed = seed;
    const bool discr = g->kind != RRRMC_EA_F64;
    P.staged_thr = std::isnan(o->staged_thr) ? (discr ? 0.5 : 0.8) : o->staged_thr; // RRRMC.jl:163-165
    P.staged_thr_fact = o->staged_thr_fact;
    RR_CUDA(cudaMemcpyAsync(c->d_beta, beta, 8 * s->R, cudaMemcpyHostToDevice, ctx->stream));
    k_chain_hdr_reset<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, seed == 0);
    ctx->launches++;
    RR_TRY(chain_energy_init(s, P));
    s->ms_valid = false; // chains now own the configuration
    RR_TRY(chain_drive<src_philox>(s, P, hook, user, Es, Es_cap, info));
    return RRRMC_OK;
}

rrrmc_status_t chain_replay(rrrmc_state *s, int64_t replica, int sampler, double beta, int64_t iters, int64_t step,
                            const uint8_t *kind, const int64_t *ival, const double *fval, int64_t ndraws,
                            const rrrmc_opts_t *o, double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_ARG(std::isfinite(beta), "β must be finite, given: %g", beta);
    RR_TRY(chain_ensure(s, sampler != CHAIN_STANDARD));
    RR_TRY(chain_sync_from_multispin(s));
    chain_store *c = s->chain;
    if (c->tcap < ndraws) {
        cudaFree(c->d_tkind); cudaFree(c->d_tival); cudaFree(c->d_tfval);
        c->tcap = std::max<int64_t>(ndraws, 1);
        RR_CUDA(cudaMalloc(&c->d_tkind, c->tcap)); RR_CUDA(cudaMalloc(&c->d_tival, 8 * c->tcap)); RR_CUDA(cudaMalloc(&c->d_tfval, 8 * c->tcap));
    }
    RR_CUDA(cudaMemcpyAsync(c->d_tkind, kind, ndraws, cudaMemcpyHostToDevice, ctx->stream));
    RR_CUDA(cudaMemcpyAsync(c->d_tival, ival, 8 * ndraws, cudaMemcpyHostToDevice, ctx->stream));
    RR_CUDA(cudaMemcpyAsync(c->d_tfval, fval, 8 * ndraws, cudaMemcpyHostToDevice, ctx->stream));
    chain_params P; chain_fill_params(s, P);
    P.sampler = sampler; P.iters = iters; P.step = step; P.seed = 0;
    const bool discr = g->kind != RRRMC_EA_F64;
    P.staged_thr = std::isnan(o->staged_thr) ? (discr ? 0.5 : 0.8) : o->staged_thr;
    P.staged_thr_fact = o->staged_thr_fact;
    P.tkind = c->d_tkind; P.tival = c->d_tival; P.tfval = c->d_tfval; P.tlen = ndraws;
    std::vector<double> b(s->R, beta);
    RR_CUDA(cudaMemcpyAsync(c->d_beta, b.data(), 8 * s->R, cudaMemcpyHostToDevice, ctx->stream));
    k_chain_hdr_reset<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, 0);
    ctx->launches++;
    RR_TRY(chain_energy_init(s, P));
    s->ms_valid = false;
    // run only the requested chain
    P.chain0 = replica; P.R = 1;
    P.beta = c->d_beta; // all equal
    RR_TRY(chain_drive<src_trace>(s, P, nullptr, nullptr, Es, Es_cap, info));
    return RRRMC_OK;
}
