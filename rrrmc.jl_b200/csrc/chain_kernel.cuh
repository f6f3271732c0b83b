// Device code of the sequential-sampler kernel k_chain_run<SRC> (see chain.cu for the description and the reference
// map). It lives in a header because the kernel is instantiated in two translation units that compile in parallel:
// chain.cu (SRC = src_philox, the samplers) and chain_trace.cu (SRC = src_trace, replay mode) — one unit took ~7 min.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include "chain.cuh"
#include "kernels.cuh"
#include "philox.cuh"

namespace {   // internal linkage: the header is compiled into two translation units

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
        if ((c.Nk & 1) == 0) {
            // two sites per lane and step: 16-byte coupling and field loads, two spin bits from one chunk (off + j is even:
            // the pair never straddles a chunk), 16-byte stores. Per element the same two roundings as the scalar loop.
            constexpr int UV = 4;
            for (int j0 = 2 * lane; j0 < c.Nk; j0 += 2 * nl * UV) {
                double2 Jv[UV], lv[UV]; uint32_t sb[UV];
#pragma unroll
                for (int u = 0; u < UV; u++) {
                    const int j = j0 + u * 2 * nl;
                    if (j < c.Nk) {
                        Jv[u] = *reinterpret_cast<const double2 *>(Ji + j);
                        lv[u] = *reinterpret_cast<const double2 *>(lf + j);
                        const int64_t b = off + j;
                        sb[u] = (uint32_t)(c.s[b >> 6] >> (b & 63));
                    }
                }
#pragma unroll
                for (int u = 0; u < UV; u++) {
                    const int j = j0 + u * 2 * nl;
                    if (j < c.Nk) {
                        *reinterpret_cast<double2 *>(lfl + j) = lv[u];
                        double2 nv;
                        nv.x = __dadd_rn(lv[u].x, 4 * __dmul_rn((double)(1 - 2 * (si ^ (int)(sb[u] & 1u))), Jv[u].x));
                        nv.y = __dadd_rn(lv[u].y, 4 * __dmul_rn((double)(1 - 2 * (si ^ (int)((sb[u] >> 1) & 1u))), Jv[u].y));
                        *reinterpret_cast<double2 *>(lf + j) = nv;
                    }
                }
            }
            return;
        }
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
        if ((c.Nk & 3) == 0) {
            // four sites per lane and step: one 4-byte coupling load, one 16-byte field load, four spin bits from one word
            // (off + j is a multiple of 4: the bits never straddle a chunk), two 16-byte stores. Integer arithmetic: the
            // same values as the scalar loop below in any order.
            constexpr int UV = 2;
            for (int j0 = 4 * lane; j0 < c.Nk; j0 += 4 * nl * UV) {
                uchar4 Jv[UV]; int4 lv[UV]; uint32_t sb[UV];
#pragma unroll
                for (int u = 0; u < UV; u++) {
                    const int j = j0 + u * 4 * nl;
                    if (j < c.Nk) {
                        Jv[u] = *reinterpret_cast<const uchar4 *>(Ji + j);
                        lv[u] = *reinterpret_cast<const int4 *>(lf + j);
                        const int64_t b = off + j;
                        sb[u] = (uint32_t)(c.s[b >> 6] >> (b & 63));
                    }
                }
#pragma unroll
                for (int u = 0; u < UV; u++) {
                    const int j = j0 + u * 4 * nl;
                    if (j < c.Nk) {
                        *reinterpret_cast<int4 *>(lfl + j) = lv[u];
                        int4 nv;
                        nv.x = lv[u].x + 8 * (si ^ (int)(sb[u] & 1u) ^ (int)Jv[u].x) - 4;
                        nv.y = lv[u].y + 8 * (si ^ (int)((sb[u] >> 1) & 1u) ^ (int)Jv[u].y) - 4;
                        nv.z = lv[u].z + 8 * (si ^ (int)((sb[u] >> 2) & 1u) ^ (int)Jv[u].z) - 4;
                        nv.w = lv[u].w + 8 * (si ^ (int)((sb[u] >> 3) & 1u) ^ (int)Jv[u].w) - 4;
                        *reinterpret_cast<int4 *>(lf + j) = nv;
                    }
                }
            }
            return;
        }
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
    } else if (P.sampler == CHAIN_EO && !discr_full) {
        // extremal_opt on EOCacheCont (DeltaE.jl:555-635): graphs that are not DiscrGraph. ΔEs of every spin and their
        // sorted order `rank` (set up by the host: ΔEs from k_chain_delta_all, the initial sortperm on the host);
        // rand_move is ONE draw — rank[searchsortedfirst(fτ, (1 - rand())·z)] (:575-587); apply_move! refreshes ΔEs of
        // the move and its neighbours and re-sorts (sortperm!(rank, ΔEs, initialized=true), :589-606): an adaptive
        // insertion pass over the nearly sorted permutation. With continuous couplings no two spins share a ΔE, so the
        // order is unique and rankshuffle! (:608-633) does nothing; equal values keep their order here.
        double *dEs = P.dEs + r * N;
        int32_t *rank = P.csj + r * (int64_t)(N + 1);
        const double *ft = P.eo_ftau + r * P.eo_stride;
        const double z = ft[N - 1];
        uint64_t *cmin = P.eo_cmin + r * P.nchunks;
        if (!h.built) { h.Emin = h.E; h.itmin = 0; for (int64_t w = 0; w < P.nchunks; w++) cmin[w] = X.s[w]; h.built = 1; }
        bool at_min = false;
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { h.done = 1; break; }
                h.it++;
                if (h.it % step == 0) { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 1; break; } }
            }
            h.pending = 0;
            const double rr = (1 - src.f64()) * z;
            int lo = 0, hi = N;
            while (lo < hi) { const int m = (lo + hi) >> 1; if (ft[m] < rr) lo = m + 1; else hi = m; }
            const int i = lo + 1;
            if (i > N || src.err) { if (!src.err) h.status = 3; h.done = 1; break; }
            const int move = rank[i - 1];
            const double dE = dEs[move];
            if (at_min && !(h.E + dE < h.Emin)) { for (int64_t w = 0; w < P.nchunks; w++) cmin[w] = X.s[w]; at_min = false; }
            gv_spinflip(X, move);
            dEs[move] = gv_delta_energy(X, move);
            gv_for_neighbors(X, move, false, [&](int j) { dEs[j] = gv_delta_energy(X, j); });
            for (int p = 1; p < N; p++) {
                const int key = rank[p]; const double kv = dEs[key];
                int q = p - 1;
                while (q >= 0 && dEs[rank[q]] > kv) { rank[q + 1] = rank[q]; q--; }
                rank[q + 1] = key;
            }
            h.E += dE;
            h.accepted++;
            if (h.E < h.Emin) { h.Emin = h.E; h.itmin = h.it; at_min = true; }
        }
        if (at_min) for (int64_t w = 0; w < P.nchunks; w++) cmin[w] = X.s[w];
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

} // namespace
