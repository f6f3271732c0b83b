// Device code shared by the "poisson" checkerboard kernels (ea_poisson.cu: cp.async staging; ea_tma.cu: TMA staging):
// the Philox call of a task, the hit masks of a task (three tiers), the flip mask of a 32-lane word.
// The procedure itself is described at the top of ea_poisson.cu and restated on the CPU in
// oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_poisson.
#pragma once
#include "common.cuh"
#include "philox.cuh"
#include "cb_params.cuh"

template <int LUT> __device__ __forceinline__ uint32_t lop3p(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
constexpr int P_XOR3 = 0x96, P_MAJ = 0xE8, P_OR3 = 0xFE;
// 1 << amt with amounts above 31 (including "negative" ones) giving 0
__device__ __forceinline__ uint32_t shl_clamp(uint32_t amt)
{
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(1u), "r"(amt));
    return r;
}
// one-hot of position amt in 128 lanes (four words), 0 for amt >= 128: two 64-bit shifts (shl.b64 clamps amounts above
// 63; SASS: SHF.L.U32 + SHF.L.U64.HI each), so a hit costs four ALU-pipe instructions and ONE subtraction, which is
// written amt·one - 64 to stay an IMAD (fma pipe) — the four 32-bit shifts it replaces needed three
__device__ __forceinline__ void onehot128(uint32_t amt, uint32_t one, uint32_t (&o)[4])
{
    uint64_t lo, hi;
    asm("shl.b64 %0, %1, %2;" : "=l"(lo) : "l"(1ull), "r"(amt));
    asm("shl.b64 %0, %1, %2;" : "=l"(hi) : "l"(1ull), "r"(amt * one - 64u));
    o[0] = (uint32_t)lo; o[1] = (uint32_t)(lo >> 32); o[2] = (uint32_t)hi; o[3] = (uint32_t)(hi >> 32);
}

// What a task's hit generation reads besides the Philox round keys: the sweep counter and the count tables. The
// one-half-sweep kernels fill it from the launch parameters (cbp_env_of); the multi-sweep kernel (ea_flow.cu) advances
// the counter itself and, for a β ladder, points every 128-replica group at its own tables.
struct cbp_env {
    uint32_t t_lo, t_hi16;              // sweep counter: low 32 bits, (high bits) << 16
    uint32_t tb0_0, tb0_1, tc0;         // TB0[0], TB0[1], TC[0]
    const uint32_t *tbl;                // TA | TB0 | TB | TC (second and third tier only)
    const uint2 *bucket;                // level-1 count lookup (shared or global memory)
};
__device__ __forceinline__ cbp_env cbp_env_of(const cbp_params &p, const uint2 *bucket)
{
    cbp_env e;
    e.t_lo = p.t_lo; e.t_hi16 = p.t_hi16; e.tb0_0 = p.tb0_0; e.tb0_1 = p.tb0_1; e.tc0 = p.tc0; e.tbl = p.tbl; e.bucket = bucket;
    return e;
}

__device__ __forceinline__ philox_out cbp_philox(const cbp_params &p, const cbp_env &e, uint32_t ctr0, uint32_t c1, uint32_t c2)
{
    uint32_t c0 = ctr0 | e.t_hi16, c3 = e.t_lo;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ p.rk[r][0];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ p.rk[r][1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    philox_out o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// hit masks of a task: m = lanes with a hit of level >= 1, g = level >= 2, h = level 3
struct cbp_hits { uint32_t m[4], g[4], h[4]; };

// The complete procedure, as the oracle states it (rare path: the fast path's masks are discarded).
template <int NW>
__device__ __noinline__ cbp_hits cbp_slow(const cbp_params &p, const uint32_t *tbl, uint32_t t_lo, uint32_t t_hi16,
                                          uint32_t c1, uint32_t c2, uint32_t X0, uint32_t X1,
                                          uint32_t P0, uint32_t P1, uint32_t P2, uint32_t P3, uint32_t P4, uint32_t P5)
{
    constexpr int NS = 4 * NW - 1;
    const uint32_t *TA = tbl, *TB0 = TA + CBP_KA, *TB = TB0 + CBP_KR, *TC = TB + CBP_KR;
    cbp_hits r;
#pragma unroll
    for (int w = 0; w < 4; w++) r.m[w] = r.g[w] = r.h[w] = 0u;
    uint32_t sw1 = 0, sw2 = 0, sw3 = 0, Y = 0, call = NW > 2 ? 2u : 1u;
    int used = 0; bool loaded = false;
    auto fetch = [&]() {
        const philox_out o = philox4x32_10(call | t_hi16, c1, c2, t_lo, p.rk[0][0], p.rk[0][1]);
        if (!loaded) Y = o.x;
        sw1 = o.y; sw2 = o.z; sw3 = o.w; loaded = true; used = 0; call++;
    };
    auto slot = [&]() -> uint32_t {
        if (!loaded || used == 12) fetch();
        const uint32_t w = used < 4 ? sw1 : (used < 8 ? sw2 : sw3);
        const uint32_t v = (w >> (8 * (used & 3))) & 127u;
        used++;
        return v;
    };
    auto stat = [&](int j) -> uint32_t {
        const int q = j >> 2;
        const uint32_t w = q == 0 ? P0 : (q == 1 ? P1 : (q == 2 ? P2 : (q == 3 ? P3 : (q == 4 ? P4 : P5))));
        return (w >> (8 * (j & 3))) & 127u;
    };
    auto mark = [&](uint32_t pos, int level) {
        const uint32_t bit = 1u << (pos & 31u);
        const int ww = (int)(pos >> 5);
#pragma unroll
        for (int w = 0; w < 4; w++)
            if (w == ww) { r.m[w] |= bit; if (level >= 2) r.g[w] |= bit; if (level >= 3) r.h[w] |= bit; }
    };
    int a = 0, b = 0, c = 0;
    while (X0 > TA[a]) a++;
    for (int j = 0; j < a; j++) mark(j < NS ? stat(j) : slot(), 1);
    if (X1 <= TC[0]) { while (X1 > TB0[b]) b++; }
    else {
        while (X1 > TC[c]) c++;
        if (!loaded) fetch();
        while (Y > TB[b]) b++;
    }
    for (int j = 0; j < b; j++) mark(j == 0 ? stat(NS) : slot(), 2);
    for (int j = 0; j < c; j++) mark(slot(), 3);
    return r;
}

// flip mask of one 32-lane word from the bond planes b_k = s_i ^ s_k ^ neg_k (1 = unsatisfied) and the hit masks:
// flip iff u + [m] + [g] >= D with u = Σ b_k (a hit of level l counts l times; for D = 3 the caller makes one bond of
// a lane with a level-3 hit unsatisfied, which lifts u + 2 to D)
template <int D>
__device__ __forceinline__ uint32_t cbp_flip_planes(const uint32_t (&b)[2 * D], uint32_t m, uint32_t g)
{
    if (D == 1) return lop3p<P_OR3>(b[0], b[1], m);                          // u + m >= 1
    if (D == 2) {                                                             // u + m + g >= 2
        const uint32_t s1 = lop3p<P_XOR3>(b[0], b[1], b[2]), k1 = lop3p<P_MAJ>(b[0], b[1], b[2]);
        const uint32_t s3 = lop3p<P_XOR3>(s1, b[3], m), k3 = lop3p<P_MAJ>(s1, b[3], m);
        return lop3p<P_OR3>(k1, k3, s3 & g);
    }
    // D == 3: u + m + g >= 3 with u = s1 + s2 + 2(k1 + k2)
    const uint32_t s1 = lop3p<P_XOR3>(b[0], b[1], b[2]), k1 = lop3p<P_MAJ>(b[0], b[1], b[2]);
    const uint32_t s2 = lop3p<P_XOR3>(b[3], b[4], b[5]), k2 = lop3p<P_MAJ>(b[3], b[4], b[5]);
    const uint32_t s3 = lop3p<P_XOR3>(s1, s2, m), k3 = lop3p<P_MAJ>(s1, s2, m);
    const uint32_t ks = lop3p<P_XOR3>(k1, k2, k3), kc = lop3p<P_MAJ>(k1, k2, k3);
    return lop3p<0xF8>(kc, ks, s3 | g);                                       // kc | (ks & (s3 | g))
}

// The same decision in two halves for D = 3: flip = kc | tt. A caller that does not need the flip mask forms the new
// spin word with one more LOP3, sc ^ (kc | tt), instead of two (flip, then xor).
__device__ __forceinline__ void cbp_flip_parts(const uint32_t (&b)[6], uint32_t m, uint32_t g, uint32_t &kc, uint32_t &tt)
{
    const uint32_t s1 = lop3p<P_XOR3>(b[0], b[1], b[2]), k1 = lop3p<P_MAJ>(b[0], b[1], b[2]);
    const uint32_t s2 = lop3p<P_XOR3>(b[3], b[4], b[5]), k2 = lop3p<P_MAJ>(b[3], b[4], b[5]);
    const uint32_t s3 = lop3p<P_XOR3>(s1, s2, m), k3 = lop3p<P_MAJ>(s1, s2, m);
    const uint32_t ks = lop3p<P_XOR3>(k1, k2, k3);
    kc = lop3p<P_MAJ>(k1, k2, k3);
    tt = lop3p<0xE0>(ks, s3, g);                                              // ks & (s3 | g)
}

// the rare part of a task's flip: lanes with a level-3 hit flip unconditionally. Out of line on purpose (see the caller);
// everything by value, so that the call moves registers and not a stack frame. Returns kc | tt | h.
static __device__ __noinline__ uint4 cbp_merge_level3(uint4 kc, uint4 tt, uint4 h)
{
    return make_uint4(kc.x | tt.x | h.x, kc.y | tt.y | h.y, kc.z | tt.z | h.z, kc.w | tt.w | h.w);
}

// Spin-independent half of a task: the hit masks of (site c1, group c2). Returns true when the task left the fast
// path (only then can h, the level-3 hits, be non-zero).
// Three tiers. (1) The fast path above: branch free. (2) A lane with more hits than static slots would stall its whole
// warp, so the second tier is entered by the WHOLE warp (one uniform branch when any lane needs it, ~15 % of the warps
// at β = 1): every lane computes the first overflow call, and the extra hits — up to twelve per lane, level-1 hits
// first, then level 2, then level 3, the oracle's order — are placed by a loop whose trip count is the warp's maximum
// (one or two). (3) What is left (a count past the twelve overflow slots, an ambiguous lookup bucket whose base count
// is below the static slots; probability < 1e-6) runs the complete scalar procedure cbp_slow().
// the random words of a task's fast path: call 0 = (X0, X1, P[0], P[1]), call 1 = P[2..5] when NW > 2
template <int NW> struct cbp_words { philox_out A; uint32_t P[NW > 2 ? 6 : 2]; };
template <int NW>
__device__ __forceinline__ cbp_words<NW> cbp_draw(const cbp_params &p, const cbp_env &e, uint32_t c1, uint32_t c2)
{
    cbp_words<NW> r;
    r.A = cbp_philox(p, e, 0u, c1, c2);
    r.P[0] = r.A.z; r.P[1] = r.A.w;
    if (NW > 2) { const philox_out B = cbp_philox(p, e, 1u, c1, c2); r.P[2] = B.x; r.P[3] = B.y; r.P[4] = B.z; r.P[5] = B.w; }
    return r;
}

template <int D, int NW>
__device__ __forceinline__ bool cbp_task_hits(const cbp_params &p, const cbp_env &env, uint32_t c1, uint32_t c2,
                                              const cbp_words<NW> &rw, uint32_t (&m)[4], uint32_t (&g)[4], uint32_t (&h)[4],
                                              bool *warp_slow = nullptr)
{
    constexpr int NS = 4 * NW - 1;
    const philox_out A = rw.A;
    uint32_t P[6] = { 0u, 0u, 0u, 0u, 0u, 0u };
#pragma unroll
    for (int q = 0; q < (NW > 2 ? 6 : 2); q++) P[q] = rw.P[q];
    const uint2 e = env.bucket[A.x >> 22];
    const uint32_t a = e.y + (A.x > e.x ? 1u : 0u);            // level-1 count (>= 64: ambiguous bucket, slow path)
    bool slow = a > (uint32_t)NS;
    if (D >= 2) slow = slow || A.y > env.tb0_1;
    const uint32_t one = p.one;                                  // 1, opaque to ptxas: keeps amt·1 - 32w an IMAD (fma pipe)
    const uint32_t kv = (128u - a) * 0x01010101u;
    uint32_t f[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) {
        const uint32_t X = kv + (0x03020100u + (uint32_t)q * 0x04040404u);   // byte j: bit 7 iff slot 4q+j >= a
        f[q] = lop3p<0xD8>(P[q], X, q == NW - 1 ? 0x00808080u : 0x80808080u); // (P & ~mask) | (X & mask)
    }
    // last static slot: first level-2 hit, valid iff b >= 1
    if (D >= 2) { if (!(A.y > env.tb0_0)) f[NW - 1] |= 0x80000000u; else f[NW - 1] &= 0x7fffffffu; }
    else f[NW - 1] |= 0x80000000u;
    uint32_t acc[4][2] = { { 0u, 0u }, { 0u, 0u }, { 0u, 0u }, { 0u, 0u } };
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const uint32_t amt = (j & 3) == 3 ? f[j >> 2] >> 24 : __byte_perm(f[j >> 2], 0u, 0x4440u + (j & 3));
        uint32_t o[4];
        onehot128(amt, one, o);
        // OR tree three inputs at a time: a pending one-hot waits in acc[w][1]
#pragma unroll
        for (int w = 0; w < 4; w++) {
            if (j == 1) { acc[w][0] = acc[w][1] | o[w]; acc[w][1] = 0u; }
            else if (j & 1) { acc[w][0] = lop3p<P_OR3>(acc[w][0], acc[w][1], o[w]); acc[w][1] = 0u; }
            else acc[w][1] = o[w];
        }
    }
#pragma unroll
    for (int w = 0; w < 4; w++) { g[w] = 0u; h[w] = 0u; }
    if (D >= 2) {
        const uint32_t amt = f[NW - 1] >> 24;
        onehot128(amt, one, g);
    }
#pragma unroll
    for (int w = 0; w < 4; w++) m[w] = NS == 1 ? (acc[w][1] | g[w]) : lop3p<P_OR3>(acc[w][0], acc[w][1], g[w]);
    const bool any_slow = __any_sync(__activemask(), slow);   // warp-uniform: some lane left the fast path
    if (warp_slow) *warp_slow = any_slow;
    if (any_slow) {
        const uint32_t *TA = env.tbl, *TB0 = TA + CBP_KA, *TB = TB0 + CBP_KR, *TC = TB + CBP_KR;
        const philox_out S = cbp_philox(p, env, NW > 2 ? 2u : 1u, c1, c2);   // first overflow call: Y, then twelve byte slots
        uint32_t na = 0, nb = 0, nc = 0;
        bool full = false;
        if (slow) {
            uint32_t at = a;
            if (e.y >= 64u) {                     // ambiguous bucket: count on from the bucket's base
                at = e.y - 64u;
                if (at < (uint32_t)NS) full = true;
                else while (A.x > TA[at]) at++;
            }
            na = at > (uint32_t)NS ? at - (uint32_t)NS : 0u;
            if (D >= 2) {
                uint32_t b = A.y > env.tb0_0 ? 1u : 0u;
                if (D == 3 && A.y > env.tc0) {      // level-3 hits: their count from X1, the level-2 count from Y
                    nc = 1u; while (A.y > TC[nc]) nc++;
                    b = 0u; while (S.x > TB[b]) b++;
                    if (b == 0u) { g[0] = g[1] = g[2] = g[3] = 0u; }   // the static level-2 slot was not a hit after all
                } else if (A.y > env.tb0_1) { b = 2u; while (A.y > TB0[b]) b++; }
                nb = b > 1u ? b - 1u : 0u;
            }
            if (na + nb + nc > 12u) full = true;
            if (full) na = nb = nc = 0u;
        }
        const uint32_t n1 = na, n2 = na + nb, n3 = na + nb + nc;
        const uint32_t nmax = __reduce_max_sync(__activemask(), n3);
        uint32_t xm[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
        for (int sl = 0; sl < 12; sl++) {
            if ((uint32_t)sl >= nmax) break;
            const uint32_t wsl = sl < 4 ? S.y : (sl < 8 ? S.z : S.w);
            const uint32_t pos = (wsl >> (8 * (sl & 3))) & 127u;
            const uint32_t amt = (uint32_t)sl < n3 ? pos : 255u;
            uint32_t o[4];
            onehot128(amt, 1u, o);
#pragma unroll
            for (int w = 0; w < 4; w++) {
                xm[w] |= o[w];
                if ((uint32_t)sl >= n1) g[w] |= o[w];
                if ((uint32_t)sl >= n2) h[w] |= o[w];
            }
        }
#pragma unroll
        for (int w = 0; w < 4; w++) m[w] = acc[w][0] | acc[w][1] | g[w] | xm[w];
        if (full) {
            const cbp_hits r = cbp_slow<NW>(p, env.tbl, env.t_lo, env.t_hi16, c1, c2, A.x, A.y, P[0], P[1], P[2], P[3], P[4], P[5]);
#pragma unroll
            for (int w = 0; w < 4; w++) { m[w] = r.m[w]; g[w] = r.g[w]; h[w] = r.h[w]; }
        }
    }
    return slow;
}
