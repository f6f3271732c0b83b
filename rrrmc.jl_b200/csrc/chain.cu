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
// chain view of a graph: the Interface of src/Interface.jl:87-270 on one replica
// ------------------------------------------------------------------------------------------------
struct gview {
    int kind, N, twoD;
    const int32_t *A; const int8_t *J8; const double *Jd; const uint8_t *Jb;
    uint64_t *s;
    int32_t *lfi; double *lfd;   // EA: [2][N] = (lfields, lfields_last). SK family: [M][2][Nk], halves swapped by sw[k]
    int32_t *ml; uint8_t *sw;
    int Nk, M, inner, nz;
    double fourK, sN;
    int coop;                    // 1: a warp serves this chain (lane 0 leads)
};
__device__ __forceinline__ int sget(const uint64_t *s, int i) { return (int)((s[i >> 6] >> (i & 63)) & 1ull); }
__device__ __forceinline__ bool is_ea(int kind) { return kind == RRRMC_EA_PM1 || kind == RRRMC_EA_INT || kind == RRRMC_EA_F64 || kind == RRRMC_EA_DISCR; }
__device__ __forceinline__ bool is_sk(int kind) { return kind == RRRMC_SK_F64 || kind == RRRMC_SK_BIN; }

// -- SK slice k of the view (SK proper: k = 0): current / last halves of the field pair
__device__ __forceinline__ int64_t sk_cur_off(const gview &c, int k) { return ((int64_t)k * 2 + c.sw[k]) * c.Nk; }
__device__ __forceinline__ int64_t sk_last_off(const gview &c, int k) { return ((int64_t)k * 2 + (c.sw[k] ^ 1)) * c.Nk; }
__device__ __forceinline__ double sk_delta(const gview &c, int skind, int k, int i) // SK.jl:278-284 / :135-140
{
    if (skind == RRRMC_SK_F64) return c.lfd[sk_cur_off(c, k) + i];
    if (skind == RRRMC_SK_BIN) return (double)c.lfi[sk_cur_off(c, k) + i] / c.sN;
    if (skind == RRRMC_EA_F64) return -c.lfd[(int64_t)k * 2 * c.Nk + i]; // GraphEANormal slice (GraphQEAT), EA.jl:655-663
    return 0.0; // GraphEmpty (Empty.jl:28-31)
}
// part of update_cache! that every lane of the serving warp runs: sites j = lane, lane+nl, ... of slice k
// (SK.jl:252-265 / :109-122). `si` is the new spin of site i; spins of the slice sit at bit offset k*Nk.
__device__ __forceinline__ void sk_update_part(const gview &c, int skind, int k, int i, int si, int lane, int nl)
{
    const int64_t cur = sk_cur_off(c, k), last = sk_last_off(c, k);
    const int64_t off = (int64_t)k * c.Nk;
    constexpr int UB = 8; // loads of UB sites are issued together: the loop is bound by memory latency, not arithmetic
    if (skind == RRRMC_SK_F64) {
        const double *Ji = c.Jd + (int64_t)i * c.Nk;
        double *lf = c.lfd + cur, *lfl = c.lfd + last;
        for (int j0 = lane; j0 < c.Nk; j0 += nl * UB) {
            double Jv[UB], lv[UB]; int sv[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int j = j0 + u * nl;
                if (j < c.Nk) { Jv[u] = Ji[j]; lv[u] = lf[j]; sv[u] = sget(c.s, (int)(off + j)); }
            }
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int j = j0 + u * nl;
                if (j < c.Nk) {
                    const double Js = __dmul_rn((double)(1 - 2 * (si ^ sv[u])), Jv[u]);
                    lfl[j] = lv[u];
                    lf[j] = __dadd_rn(lv[u], 4 * Js);
                }
            }
        }
    } else {
        const uint8_t *Ji = c.Jb + (int64_t)i * c.Nk;
        int32_t *lf = c.lfi + cur, *lfl = c.lfi + last;
        for (int j0 = lane; j0 < c.Nk; j0 += nl * UB) {
            int Jv[UB], lv[UB], sv[UB];
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int j = j0 + u * nl;
                if (j < c.Nk) { Jv[u] = (int)Ji[j]; lv[u] = lf[j]; sv[u] = sget(c.s, (int)(off + j)); }
            }
#pragma unroll
            for (int u = 0; u < UB; u++) {
                const int j = j0 + u * nl;
                if (j < c.Nk) {
                    const int Js = si ^ sv[u] ^ Jv[u];
                    lfl[j] = lv[u];
                    lf[j] = lv[u] + 8 * Js - 4;
                }
            }
        }
    }
}
enum { COOP_EXIT = 0, COOP_SK_UPDATE = 1 };
// update_cache! of one SK slice after s_i flipped (SK.jl:239-276 / :96-133); called by the chain's leader
__device__ void sk_update_cache(gview &c, int skind, int k, int i)
{
    if (skind == RRRMC_EMPTY) return;
    if (skind == RRRMC_EA_F64) {   // GraphEANormal slice of a GraphQEAT (QAliases.jl:51): update_cache! EA.jl:613-653
        double *lf = c.lfd + (int64_t)k * 2 * c.Nk, *lfl = lf + c.Nk;
        const int64_t off = (int64_t)k * c.Nk;
        int U[MAXDEG], nU = 0;
        for (int q = 0; q < c.twoD; q++) {
            const int y = c.A[(int64_t)i * c.twoD + q];
            if (nU == 0 || U[nU - 1] != y) U[nU++] = y;
        }
        if (c.ml[k] == i) {
            for (int q = 0; q < nU; q++) { const double t = lf[U[q]]; lf[U[q]] = lfl[U[q]]; lfl[U[q]] = t; }
            lf[i] = -lf[i]; lfl[i] = -lfl[i];
            return;
        }
        for (int q = 0; q < nU; q++) lfl[U[q]] = lf[U[q]];
        const int sx = sget(c.s, (int)(off + i));
        for (int q = 0; q < c.twoD; q++) {
            const int y = c.A[(int64_t)i * c.twoD + q];
            const double f = (double)(4 * (1 - 2 * (sx ^ sget(c.s, (int)(off + y)))));
            lf[y] = __dsub_rn(lf[y], __dmul_rn(f, c.Jd[(int64_t)i * c.twoD + q]));
        }
        const double lfm = lf[i];
        lfl[i] = lfm; lf[i] = -lfm;
        c.ml[k] = i;
        return;
    }
    if (c.ml[k] == i) { c.sw[k] ^= 1; return; } // swap lfields <-> lfields_last (SK.jl:247-250)
    const int si = sget(c.s, (int)((int64_t)k * c.Nk + i));
    double lfm_d = 0; int lfm_i = 0;
    if (skind == RRRMC_SK_F64) lfm_d = c.lfd[sk_cur_off(c, k) + i]; else lfm_i = c.lfi[sk_cur_off(c, k) + i];
    if (c.coop) {
        __threadfence_block();
        __shfl_sync(FULLMASK, (int)COOP_SK_UPDATE, 0); __shfl_sync(FULLMASK, k, 0); __shfl_sync(FULLMASK, i, 0); __shfl_sync(FULLMASK, si, 0);
        sk_update_part(c, skind, k, i, si, 0, 32);
        __syncwarp(FULLMASK);
    } else sk_update_part(c, skind, k, i, si, 0, 1);
    if (skind == RRRMC_SK_F64) { c.lfd[sk_last_off(c, k) + i] = lfm_d; c.lfd[sk_cur_off(c, k) + i] = -lfm_d; }
    else { c.lfi[sk_last_off(c, k) + i] = lfm_i; c.lfi[sk_cur_off(c, k) + i] = -lfm_i; }
    c.ml[k] = i;
}
// helper lanes of a cooperative chain: serve the leader until it says exit
__device__ void coop_helper_loop(const gview &c, int lane)
{
    const int skind = c.kind == RRRMC_QUANT ? c.inner : c.kind;
    for (;;) {
        const int cmd = __shfl_sync(FULLMASK, 0, 0);
        if (cmd == COOP_EXIT) return;
        const int k = __shfl_sync(FULLMASK, 0, 0), i = __shfl_sync(FULLMASK, 0, 0), si = __shfl_sync(FULLMASK, 0, 0);
        sk_update_part(c, skind, k, i, si, lane, 32);
        __threadfence_block();
        __syncwarp(FULLMASK);
    }
}

// -- GraphQT (QT.jl:86-108)
__device__ __forceinline__ void qt_neighbors(const gview &c, int i, int &k1, int &k2)
{
    k1 = i - c.Nk + (i < c.Nk ? c.N : 0);
    k2 = i + c.Nk - (i + c.Nk >= c.N ? c.N : 0);
}
__device__ __forceinline__ double qt_delta(const gview &c, int i)
{
    int k1, k2; qt_neighbors(c, i, k1, k2);
    const int sk = sget(c.s, i), s1 = sget(c.s, k1), s2 = sget(c.s, k2);
    return (double)((sk == s1) - (sk != s2)) * c.fourK;
}

// delta_energy(X, C, i): `inner` selects inner_graph(X) for a DoubleGraph (Interface.jl:239-240)
__device__ __forceinline__ double gv_delta_energy(const gview &c, int i, bool inner = false)
{
    switch (c.kind) {
    case RRRMC_EA_F64: return -c.lfd[i];                       // EA.jl:655-663
    case RRRMC_EA_PM1: case RRRMC_EA_INT: return -(double)c.lfi[i]; // EA.jl:266-275
    case RRRMC_EA_DISCR:                                       // EA.jl:519-523: convert(Float64, ΔE0 + ΔE1)
        return inner ? -(double)c.lfi[i] : __dadd_rn(-(double)c.lfi[i], -c.lfd[i]);
    case RRRMC_SK_F64: case RRRMC_SK_BIN: return sk_delta(c, c.kind, 0, i);
    case RRRMC_QT: return qt_delta(c, i);
    case RRRMC_QUANT: {                                        // QT.jl:283-286, residual :270-281
        const double d0 = qt_delta(c, i);
        if (inner) return d0;
        return d0 + sk_delta(c, c.inner, i / c.Nk, i % c.Nk) / (double)c.M;
    }
    }
    return 0.0;
}
__device__ __forceinline__ double gv_delta_residual(const gview &c, int i) // Interface.jl:254-261; QT.jl:270-281
{
    if (c.kind == RRRMC_EA_DISCR) return -c.lfd[i];            // EA.jl:489-497
    if (c.kind != RRRMC_QUANT) return 0.0;
    return sk_delta(c, c.inner, i / c.Nk, i % c.Nk) / (double)c.M;
}
// neighbors(X, i) in the reference's iteration order (EA.jl:292 uA; Common.jl:78-92 AllButOne; QT.jl:105-108, :288-321)
template <class F> __device__ __forceinline__ void gv_for_neighbors(const gview &c, int i, bool inner, F f)
{
    if (is_ea(c.kind)) {
        // GraphRRG (RRG.jl:133, :261): the integer graph's neighbours are the entries with a non-zero coupling; the
        // DoubleGraph over it (GraphRRGNormalDiscretized, RRG.jl:499) lists the whole row
        const bool skip0 = c.nz && (c.kind == RRRMC_EA_INT || (c.kind == RRRMC_EA_DISCR && inner));
        int prev = -1;
        for (int k = 0; k < c.twoD; k++) {
            const int y = c.A[(int64_t)i * c.twoD + k];
            if (y != prev && !(skip0 && c.J8[(int64_t)i * c.twoD + k] == 0)) f(y);
            prev = y;
        }
    } else if (is_sk(c.kind)) {
        for (int j = 0; j < c.N; j++) if (j != i) f(j);
    } else {
        int k1, k2; qt_neighbors(c, i, k1, k2);
        f(k1); f(k2);
        if (c.kind == RRRMC_QUANT && !inner && c.inner != RRRMC_EMPTY) {
            const int base = (i / c.Nk) * c.Nk, ii = i - base;
            if (c.inner == RRRMC_EA_F64) {              // neighbors(X1[k], j) = uA[j], shifted to the slice (QT.jl:288-321)
                int prev = -1;
                for (int q = 0; q < c.twoD; q++) {
                    const int y = c.A[(int64_t)ii * c.twoD + q];
                    if (y != prev) f(base + y);
                    prev = y;
                }
            } else
                for (int j = 0; j < c.Nk; j++) if (j != ii) f(base + j);
        }
    }
}
// spinflip!(X, C, i) = flip + update_cache! (Interface.jl:89-92)
__device__ void gv_spinflip(gview &c, int i, bool inner = false)
{
    c.s[i >> 6] ^= 1ull << (i & 63);
    if (c.kind == RRRMC_QT || (c.kind == RRRMC_QUANT && inner)) return;   // Interface.jl:87: no cache
    if (c.kind == RRRMC_QUANT) { sk_update_cache(c, c.inner, i / c.Nk, i % c.Nk); return; } // QT.jl:172-183
    if (is_sk(c.kind)) { sk_update_cache(c, c.kind, 0, i); return; }
    // GraphEA update_cache! EA.jl:224-264 / :613-653. GraphEANormalDiscretized (EA.jl:390-450) = the integer update of
    // its inner GraphEA, then (unless only inner_graph(X) is being flipped) update_cache_residual! (:452-487), which is
    // the GraphEANormal update on the residual couplings; each cache keeps its own move_last (ml[0], ml[1]).
    int U[MAXDEG], nU = 0;
    for (int k = 0; k < c.twoD; k++) {
        const int y = c.A[(int64_t)i * c.twoD + k];
        if (nU == 0 || U[nU - 1] != y) U[nU++] = y;
    }
    const int N = c.N;
    if (c.kind != RRRMC_EA_F64) {
        int32_t *lf = c.lfi, *lfl = c.lfi + N;
        if (c.ml[0] == i) {
            for (int k = 0; k < nU; k++) { const int t = lf[U[k]]; lf[U[k]] = lfl[U[k]]; lfl[U[k]] = t; }
            lf[i] = -lf[i]; lfl[i] = -lfl[i];
        } else {
            for (int k = 0; k < nU; k++) lfl[U[k]] = lf[U[k]];
            const int sx = sget(c.s, i);
            for (int k = 0; k < c.twoD; k++) {
                const int y = c.A[(int64_t)i * c.twoD + k];
                lf[y] -= 4 * (1 - 2 * (sx ^ sget(c.s, y))) * (int)c.J8[(int64_t)i * c.twoD + k];
            }
            const int lfm = lf[i];
            lfl[i] = lfm; lf[i] = -lfm;
            c.ml[0] = i;
        }
    }
    if (c.kind == RRRMC_EA_F64 || (c.kind == RRRMC_EA_DISCR && !inner)) {
        int32_t &ml = c.ml[c.kind == RRRMC_EA_DISCR ? 1 : 0];
        double *lf = c.lfd, *lfl = c.lfd + N;
        if (ml == i) {
            for (int k = 0; k < nU; k++) { const double t = lf[U[k]]; lf[U[k]] = lfl[U[k]]; lfl[U[k]] = t; }
            lf[i] = -lf[i]; lfl[i] = -lfl[i];
            return;
        }
        for (int k = 0; k < nU; k++) lfl[U[k]] = lf[U[k]];
        const int sx = sget(c.s, i);
        for (int k = 0; k < c.twoD; k++) {
            const int y = c.A[(int64_t)i * c.twoD + k];
            const double f = (double)(4 * (1 - 2 * (sx ^ sget(c.s, y))));
            lf[y] = __dsub_rn(lf[y], __dmul_rn(f, c.Jd[(int64_t)i * c.twoD + k]));
        }
        const double lfm = lf[i];
        lfl[i] = lfm; lf[i] = -lfm;
        ml = i;
    }
}

// ------------------------------------------------------------------------------------------------
// discrete ΔE-class cache (DeltaE.jl:63-295) with ArraySets (ArraySets.jl:58-85); built on inner_graph(X)
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
__device__ __forceinline__ int dc_class_of(const dcache &c, const gview &X, int j)
{
    const double dE = gv_delta_energy(X, j, true);
    const int up = dE > 0 || (dE == 0 && sget(X.s, j) == 1);
    return dc_findk(c, dE) + c.L * up;
}
__device__ void dc_build(dcache &c, const gview &X, double beta) // DeltaE.jl:74-104
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
__device__ void dc_compute_staged(dcache &c, gview &X, int i) // DeltaE.jl:202-230 (on the inner graph)
{
    gv_spinflip(X, i, true);
    c.nst = 0;
    gv_for_neighbors(X, i, true, [&](int j) {
        const int k0 = c.cls[j], k1 = dc_class_of(c, X, j);
        if (k0 == k1) return;
        c.st[c.nst][0] = j; c.st[c.nst][1] = k0; c.st[c.nst][2] = k1; c.nst++;
    });
    const int k0 = c.cls[i], k1 = k0 - c.L * (2 * (k0 > c.L) - 1);
    c.st[c.nst][0] = i; c.st[c.nst][1] = k0; c.st[c.nst][2] = k1; c.nst++;
    gv_spinflip(X, i, true);
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
__device__ double dc_apply_move(dcache &c, gview &X, int move) // DeltaE.jl:232-295 (flip on X, classes on inner_graph(X))
{
    gv_spinflip(X, move, false);
    double zp = c.z;
    auto reclass = [&](int j, int k0, int k1) {
        const double f0 = dc_f(c, k0), f1 = dc_f(c, k1);
        c.T[k0] -= f0; c.T[k1] += f1;
        zp += f1 - f0;
        as_delete(c, k0, j); as_push(c, k1, j); c.cls[j] = (uint8_t)k1;
    };
    gv_for_neighbors(X, move, true, [&](int j) {
        const int k0 = c.cls[j], k1 = dc_class_of(c, X, j);
        if (k0 != k1) reclass(j, k0, k1);
    });
    { const int k0 = c.cls[move]; reclass(move, k0, k0 - c.L * (2 * (k0 > c.L) - 1)); }
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
    int32_t *sj; double *sdE, *sp; int nst; // staged list, up to N entries
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
__device__ void cc_build(ccache &c, const gview &X, bool inner) // DeltaE.jl:304-311 + DynamicSamplers.jl:35-51
{
    for (long long i = 0; i <= c.N2; i++) c.v[i] = 0.0;
    for (int i = 0; i < c.N; i++) { c.dEs[i] = gv_delta_energy(X, i, inner); c.v[i + 1] = prior(c.beta * c.dEs[i]); }
    ds_refresh(c);
}
template <class SRC> __device__ long long cc_rand_skip(const ccache &c, SRC &d) // DeltaE.jl:319-324
{
    double b = c.z / (double)c.N;
    b = fmin(fmax(b, 2.2250738585072014e-308), 1.0);
    return (long long)floor(log1p(-d.f64()) / log1p(-b));
}
__device__ void cc_compute_staged(ccache &c, gview &X, int i) // DeltaE.jl:356-373 (SingleGraph only)
{
    gv_spinflip(X, i);
    double dE = gv_delta_energy(X, i);
    c.sj[0] = i; c.sdE[0] = dE; c.sp[0] = prior(c.beta * dE); c.nst = 1;
    gv_for_neighbors(X, i, false, [&](int j) {
        const double d = gv_delta_energy(X, j);
        c.sj[c.nst] = j; c.sdE[c.nst] = d; c.sp[c.nst] = prior(c.beta * d); c.nst++;
    });
    gv_spinflip(X, i);
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
__device__ double cc_apply_move(ccache &c, gview &X, int move, bool inner) // DeltaE.jl:378-410
{
    gv_spinflip(X, move);
    const double z = c.z;
    double dE = gv_delta_energy(X, move, inner);
    c.dEs[move] = dE; ds_set(c, move + 1, prior(c.beta * dE));
    gv_for_neighbors(X, move, inner, [&](int j) {
        const double d = gv_delta_energy(X, j, inner);
        c.dEs[j] = d; ds_set(c, j + 1, prior(c.beta * d));
    });
    return z / c.z;
}

// ------------------------------------------------------------------------------------------------
// wtmMC (RRRMC.jl:376-430, WaitingTimes.jl): mutable binary min-heap of the spins' next flip times. The reference
// uses DataStructures.MutableBinaryMinHeap; only its minimum and update! are observable, so any min-heap reproduces
// the trajectory (flip times are continuous: ties have probability zero).
// ------------------------------------------------------------------------------------------------
struct wheap { int N; double *v; int32_t *node, *pos; };
__device__ __forceinline__ void wh_swap(wheap &H, int a, int b)
{
    const double tv = H.v[a]; H.v[a] = H.v[b]; H.v[b] = tv;
    const int tn = H.node[a]; H.node[a] = H.node[b]; H.node[b] = tn;
    H.pos[H.node[a]] = a; H.pos[H.node[b]] = b;
}
__device__ void wh_up(wheap &H, int h) { while (h > 0) { const int q = (h - 1) / 2; if (!(H.v[h] < H.v[q])) break; wh_swap(H, h, q); h = q; } }
__device__ void wh_down(wheap &H, int h)
{
    for (;;) {
        const int l = 2 * h + 1, r = l + 1; int m = h;
        if (l < H.N && H.v[l] < H.v[m]) m = l;
        if (r < H.N && H.v[r] < H.v[m]) m = r;
        if (m == h) break;
        wh_swap(H, h, m); h = m;
    }
}
__device__ void wh_update(wheap &H, int site, double val)
{
    const int h = H.pos[site];
    const double old = H.v[h];
    H.v[h] = val;
    if (val < old) wh_up(H, h); else wh_down(H, h);
}
__device__ __forceinline__ double wt_tau(double beta, double dE) { const double e = exp(__dmul_rn(beta, dE)); return e > 1.0 ? e : 1.0; } // WaitingTimes.jl:15

__device__ __forceinline__ gview make_view(const chain_params &P, int64_t r)
{
    gview X;
    X.kind = P.kind; X.N = P.N; X.twoD = P.twoD; X.A = P.A; X.J8 = P.J8; X.Jd = P.Jd; X.Jb = P.Jb;
    X.s = P.chunks + r * P.nchunks;
    X.lfi = P.lfi ? P.lfi + r * 2 * (int64_t)P.N : nullptr;
    X.lfd = P.lfd ? P.lfd + r * 2 * (int64_t)P.N : nullptr;
    X.ml = P.ml + r * P.M; X.sw = P.sw + r * P.M;
    X.Nk = P.Nk; X.M = P.M; X.inner = P.inner; X.nz = P.nz; X.fourK = P.fourK_r ? P.fourK_r[r] : P.fourK; X.sN = P.sN;
    X.coop = P.coop;
    return X;
}

// ------------------------------------------------------------------------------------------------
// the sampler kernel (resumable: pauses after `quota` samples so that the host can run the hook)
// ------------------------------------------------------------------------------------------------
template <class SRC>
__global__ void __launch_bounds__(32) k_chain_run(chain_params P)
{
    const int lane = threadIdx.x;
    int64_t r;
    if (P.coop) r = P.chain0 + blockIdx.x;
    else {
        if (lane >= P.cpw) return;
        r = P.chain0 + (int64_t)blockIdx.x * P.cpw + lane;
        if (r >= P.chain0 + P.R) return;
    }
    gview X = make_view(P, r);
    if (P.coop && lane != 0) { coop_helper_loop(X, lane); return; }
    chain_hdr h = P.hdr[r];
    if (h.done) { if (P.coop) __shfl_sync(FULLMASK, (int)COOP_EXIT, 0); return; }
    const int N = P.N;
    SRC src(P, r, h.rng_n);
    const double beta = P.beta[r];
    const bool dbl = P.kind == RRRMC_QUANT || P.kind == RRRMC_EA_DISCR;            // DoubleGraph
    const bool discr_full = P.kind == RRRMC_EA_PM1 || P.kind == RRRMC_EA_INT || P.kind == RRRMC_QT; // X <: DiscrGraph
    // rrrMC builds its cache on inner_graph(X) (RRRMC.jl:170-171, :239-240); bklMC on X itself (:325)
    const bool discr = P.sampler == CHAIN_RRR ? (discr_full || dbl) : discr_full;
    const bool cc_inner = P.sampler == CHAIN_RRR;
    long long emitted = 0;
    const long long iters = P.iters, step = P.step;
    double *Es = P.Es;

    dcache dc; ccache cc; double de_q[2];
    if (P.sampler != CHAIN_STANDARD && P.sampler != CHAIN_WTM && P.sampler != CHAIN_EO) {
        if (discr) {
            dc.N = N; dc.L = P.nDE; dc.DE = P.DE; dc.t = h.t; dc.T = dc.Ta; dc.Tp = dc.Tb;
            if (P.fourK_r) { de_q[0] = 0.0; de_q[1] = X.fourK; dc.DE = de_q; }   // allΔE of this replica's GraphQT, QT.jl:111
            dc.av = P.av + r * (int64_t)(2 * P.nDE) * N; dc.apos = P.apos + r * N; dc.cls = P.cls + r * N;
            for (int k = 0; k < dc.L; k++) dc.ft[k] = exp(-beta * dc.DE[k]);
            for (int k = 0; k <= 2 * dc.L; k++) dc.T[k] = h.T[k];
            dc.z = h.z; dc.zp = h.z; dc.nst = 0;
            if (!h.built) { dc_build(dc, X, beta); h.built = 1; }
        } else {
            cc.N = N; cc.levs = P.levs; cc.N2 = P.N2; cc.beta = beta;
            cc.v = P.dv + r * (P.N2 + 1); cc.ps = P.dps + r * (P.N2 + 1); cc.dEs = P.dEs + r * N;
            cc.sj = P.csj + r * ((int64_t)N + 1); cc.sdE = P.csdE + r * ((int64_t)N + 1); cc.sp = P.csp + r * ((int64_t)N + 1);
            cc.z = h.z; cc.trefresh = h.trefresh; cc.nst = 0;
            if (!h.built) { cc_build(cc, X, cc_inner); h.built = 1; }
        }
    }
#define EMIT_SAMPLE()                                                         \
    do {                                                                      \
        if (Es && emitted < P.Es_rows) Es[emitted * P.R + (r - P.chain0)] = h.E; \
        emitted++;                                                            \
    } while (0)
    // accept(c, x) of RRRMC.jl:40-44 (DoubleGraph residual filter)
    auto accept2 = [&](double c, double x) -> bool {
        if (c >= 1 && x >= 0) return true;
        const double a = c * exp(x);
        return a >= 1 || src.f64() < a;
    };

    if (P.sampler == CHAIN_STANDARD) { // RRRMC.jl:100-119
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { h.done = 1; break; }
                h.it++;
                if (h.it % step == 0) { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 1; break; } }
            }
            h.pending = 0;
            const int i = (int)src.range(N) - 1;
            const double dE = gv_delta_energy(X, i);
            const double x = -beta * dE;
            if (src.err) { h.done = 1; break; }
            if (!(x >= 0 || src.f64() < exp(x))) continue; // accept(), RRRMC.jl:39
            gv_spinflip(X, i);
            h.E += dE;
            h.accepted++;
        }
    } else if (P.sampler == CHAIN_RRR) { // RRRMC.jl:180-211 (SingleGraph), :249-282 (DoubleGraph)
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
                double z, zp, dE0, dE1 = 0.0; int move;
                if (discr) { z = dc.z; move = dc_rand_move(dc, src, dE0); dc_compute_staged(dc, X, move); zp = dc_reverse(dc); }
                else { z = cc.z; move = ds_getel(cc, src.f64(), src.err) - 1; dE0 = cc.dEs[move]; cc_compute_staged(cc, X, move); zp = cc_reverse(cc); }
                const double c = z / zp;
                bool ok;
                if (dbl) { dE1 = gv_delta_residual(X, move); ok = accept2(c, -beta * dE1); }
                else ok = src.f64() < c;
                if (ok) {
                    gv_spinflip(X, move);
                    if (discr) dc_apply_staged(dc); else cc_apply_staged(cc);
                    h.E += dE0 + dE1; h.accepted++; acc = 1;
                }
            } else {
                double dE0, dE1 = 0.0; int move;
                if (discr) move = dc_rand_move(dc, src, dE0); else { move = ds_getel(cc, src.f64(), src.err) - 1; dE0 = cc.dEs[move]; }
                if (dbl) dE1 = gv_delta_residual(X, move);
                const double c = discr ? dc_apply_move(dc, X, move) : cc_apply_move(cc, X, move, true);
                const bool ok = dbl ? accept2(c, -beta * dE1) : (src.f64() < c);
                if (ok) { h.E += dE0 + dE1; h.accepted++; acc = 1; }
                else { if (discr) dc_apply_move(dc, X, move); else cc_apply_move(cc, X, move, true); }
            }
            h.acc_rate = h.acc_rate * (1 - lambda) + acc * lambda;
            if (src.err) { h.done = 1; break; }
        }
    } else if (P.sampler == CHAIN_WTM) { // RRRMC.jl:389-422
        wheap H; H.N = N; H.v = P.wt_v + r * (int64_t)N; H.node = P.wt_node + r * (int64_t)N; H.pos = P.wt_pos + r * (int64_t)N;
        auto gen_wt = [&](double tau) -> double { return __dmul_rn(-tau, log1p(-src.f64())); };   // WaitingTimes.jl:17-21
        if (!h.built) {                  // THeap(X, C, β): all τ first, then N draws in site order (WaitingTimes.jl:25-35)
            for (int i = 0; i < N; i++) H.v[i] = wt_tau(beta, gv_delta_energy(X, i));
            for (int i = 0; i < N; i++) { H.v[i] = gen_wt(H.v[i]); H.node[i] = i; H.pos[i] = i; wh_up(H, i); }
            h.built = 1;
        }
        for (;;) {
            const double tp = H.v[0]; const int move = H.node[0];   // top_with_handle
            bool out = false, paused = false;
            while (tp >= h.wt_next) {
                if (h.pending == 2) h.pending = 1; // resuming right after the hook of this sample
                else { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 2; paused = true; break; } }
                h.wt_next = __dadd_rn(h.wt_next, P.wt_step);
                if (h.wt_next > P.wt_tmax + 1e-10) { out = true; break; }
            }
            if (paused) break;
            if (out || src.err) { h.done = 1; break; }
            h.pending = 0;
            const double dE = gv_delta_energy(X, move);               // update_heap!, WaitingTimes.jl:39-51
            gv_spinflip(X, move);
            wh_update(H, move, __dadd_rn(tp, gen_wt(wt_tau(beta, -dE))));
            gv_for_neighbors(X, move, false, [&](int j) {
                wh_update(H, j, __dadd_rn(tp, gen_wt(wt_tau(beta, gv_delta_energy(X, j)))));
            });
            h.E += dE;
            h.accepted++; h.it++;
        }
    } else if (P.sampler == CHAIN_EO) { // extremal_opt, RRRMC.jl:494-513 on EOCache (DeltaE.jl:413-543)
        // classes in ascending ΔE (findks, DeltaE.jl:413-422): K = 2L - has_zero, ΔE = 0 is one class
        dc.N = N; dc.L = P.nDE; dc.DE = P.DE; dc.t = h.t;
        dc.av = P.av + r * (int64_t)(2 * P.nDE) * N; dc.apos = P.apos + r * N; dc.cls = P.cls + r * N;
        const int L = dc.L, hz = dc.DE[0] == 0.0 ? 1 : 0, K = 2 * L - hz;
        const double *ft = P.eo_ftau + r * P.eo_stride;
        const double z = ft[N - 1];
        uint64_t *cmin = P.eo_cmin + r * P.nchunks;
        auto findks = [&](int j) -> int {
            const double dE = gv_delta_energy(X, j);
            const int ak = dc_findk(dc, dE);
            return dE >= 0 ? ak + L - hz : L + 1 - ak;
        };
        if (!h.built) {                  // EOCache ctor, DeltaE.jl:433-441; Emin = E, Cmin = copy(C), RRRMC.jl:480-482
            for (int k = 0; k <= 2 * L; k++) dc.t[k] = 0;
            for (int i = 0; i < N; i++) { const int ki = findks(i); dc.cls[i] = (uint8_t)ki; as_push(dc, ki, i); }
            h.Emin = h.E; h.itmin = 0;
            for (int64_t w = 0; w < P.nchunks; w++) cmin[w] = X.s[w];
            h.built = 1;
        }
        // copy!(Cmin, C) (RRRMC.jl:508-512) is deferred while the chain keeps improving: during a descent every move
        // lowers Emin and only the last configuration of the streak survives, so Cmin is written when the chain is
        // about to leave its minimum (or the kernel returns), not at every improvement
        bool at_min = false;
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { h.done = 1; break; }
                h.it++;
                if (h.it % step == 0) { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 1; break; } }
            }
            h.pending = 0;
            const double rr = (1 - src.f64()) * z;                      // rand_move, DeltaE.jl:480-517
            int lo = 0, hi = N;                                         // searchsortedfirst(fτ, r)
            while (lo < hi) { const int m = (lo + hi) >> 1; if (ft[m] < rr) lo = m + 1; else hi = m; }
            const int i = lo + 1;
            if (i > N || src.err) { if (!src.err) h.status = 3; h.done = 1; break; }
            int k = 0, t = 0;
            while (i > t && k < K) { k++; t += dc.t[k]; }
            const double dE = k <= L ? -dc.DE[L - k] : dc.DE[k - L + hz - 1];
            const int move = dc.av[(int64_t)(k - 1) * N + src.range(dc.t[k]) - 1];
            if (src.err) { h.done = 1; break; }
            if (at_min && !(h.E + dE < h.Emin)) { for (int64_t w = 0; w < P.nchunks; w++) cmin[w] = X.s[w]; at_min = false; }
            gv_spinflip(X, move);                                       // apply_move!, DeltaE.jl:519-543
            auto reclass = [&](int j) {
                const int k0 = dc.cls[j], k1 = findks(j);
                if (k0 == k1) return;
                as_delete(dc, k0, j); as_push(dc, k1, j); dc.cls[j] = (uint8_t)k1;
            };
            gv_for_neighbors(X, move, false, reclass);
            reclass(move);
            h.E += dE;
            h.accepted++;
            if (h.E < h.Emin) { h.Emin = h.E; h.itmin = h.it; at_min = true; }
        }
        if (at_min) for (int64_t w = 0; w < P.nchunks; w++) cmin[w] = X.s[w];
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
            if (discr) dc_apply_move(dc, X, h.pmove); else cc_apply_move(cc, X, h.pmove, false);
            h.it += h.skip + 1;
            h.E += h.pdE;
            h.accepted++;
            h.pending = 0;
        }
    }
#undef EMIT_SAMPLE
    if (P.coop) __shfl_sync(FULLMASK, (int)COOP_EXIT, 0);
    if (P.sampler != CHAIN_STANDARD && P.sampler != CHAIN_WTM && P.sampler != CHAIN_EO) {
        if (discr) { for (int k = 0; k <= 2 * dc.L; k++) h.T[k] = dc.T[k]; h.z = dc.z; }
        else { h.z = cc.z; h.trefresh = cc.trefresh; }
    }
    h.rng_n = src.pos();
    if (src.err) h.status = src.err;
    P.hdr[r] = h;
}

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
    cudaEvent_t e0, e1;
    RR_CUDA(cudaEventCreate(&e0)); RR_CUDA(cudaEventCreate(&e1));
    RR_CUDA(cudaEventRecord(e0, ctx->stream));
    int64_t nsamples = 0; bool stop = false;
    for (int guard = 0; !stop; guard++) {
        if (P.fast) RR_TRY(chain_ea_launch(s, P)); else k_chain_run<SRC><<<grid, 32, 0, ctx->stream>>>(P);
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
    cudaEventDestroy(e0); cudaEventDestroy(e1);
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
    if (chain_ea_eligible(s, sampler)) { RR_TRY(chain_ea_prepare(s, P)); s->chain_fields_valid = false; }
    RR_TRY(chain_drive<src_philox>(s, P, hook, user, Es, Es_cap, info));
    return RRRMC_OK;
}

rrrmc_status_t chain_run_wtm(rrrmc_state *s, const double *beta, int64_t samples, double step, uint64_t seed,
                             rrrmc_hook_fn hook, void *user, double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
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
    RR_TRY(chain_drive<src_philox>(s, P, hook, user, Es, Es_cap, info));
    return RRRMC_OK;
}

rrrmc_status_t chain_run_eo(rrrmc_state *s, const double *ftau, int64_t ftau_stride, int64_t iters, int64_t step, uint64_t seed,
                            rrrmc_eo_hook_fn hook, void *user, double *Emin_out, int64_t *itmin_out, uint64_t *Cmin_chunks,
                            double *Es, int64_t Es_cap, rrrmc_run_info_t *info)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    const bool discr_full = g->kind == RRRMC_EA_PM1 || g->kind == RRRMC_EA_INT || g->kind == RRRMC_QT;
    if (!discr_full) {   // gen_EOcache(X::AbstractGraph) = EOCacheCont re-sorts all N spins per move (DeltaE.jl:545-635)
        rrrmc_set_error("extremal_opt: only DiscrGraph models (GraphEA / GraphRRG with integer levels, GraphQT) are on this path");
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
    RR_TRY(chain_ensure(s, 1));
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
    RR_TRY(chain_drive<src_philox>(s, P, reinterpret_cast<rrrmc_hook_fn>(hook), user, Es, Es_cap, info));
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
    P.staged_thr = std::isnan(o->staged_thr) ? default_staged_thr(g) : o->staged_thr;
    P.staged_thr_fact = o->staged_thr_fact;
    P.tkind = c->d_tkind; P.tival = c->d_tival; P.tfval = c->d_tfval; P.tlen = ndraws;
    std::vector<double> b(s->R, beta);
    RR_CUDA(cudaMemcpyAsync(c->d_beta, b.data(), 8 * s->R, cudaMemcpyHostToDevice, ctx->stream));
    k_chain_hdr_reset<<<div_up(P.R, 64), 64, 0, ctx->stream>>>(P, 0);
    ctx->launches++;
    RR_TRY(chain_energy_init(s, P, true));
    s->ms_valid = false;
    sk_dense_invalidate(s);
    // run only the requested chain
    P.chain0 = replica; P.R = 1;
    P.beta = c->d_beta; // all equal
    RR_TRY(chain_drive<src_trace>(s, P, nullptr, nullptr, Es, Es_cap, info));
    return RRRMC_OK;
}
