// Checkerboard Metropolis half-sweep for GraphEA ±J, "poisson" acceptance procedure.
//
// Same multispin layout and task decomposition as ea_multispin.cu: one task = (site of the active colour, group of
// 128 replicas = four 32-lane words). ΔE (EA.jl:266-275) is evaluated bit-sliced; lanes with ΔE <= 0 always flip and a
// lane of class c (ΔE = 4c > 0) flips with probability p_c = exp(-4βc) (accept(), RRRMC.jl:39).
//
// The filter is sampled per TASK instead of per lane. Every lane carries D independent Poisson hit processes: level-l
// hits arrive with rate lam_l - lam_{l+1}, lam_c = -log(1 - p_c) (lam_{D+1} = 0), and a lane of class c flips iff it
// received a hit of level >= c: probability 1 - exp(-lam_c) = p_c, independently across lanes. Per task the number of
// hits of a level is Poisson (inverse CDF of one 32-bit uniform against a host-built table) and each hit lands on a
// uniform 7-bit lane position WITH replacement — two hits on one lane are harmless, so nothing is ever redrawn and the
// common case has no data-dependent control flow at all:
//   call 0 = (X0, X1, P[0], P[1]), call 1 = (P[2..5]) when NW > 2          (Philox4x32-10, counter (call, site, group, sweep))
//   X0 -> a = number of level-1 hits; static slots j < NS = 4·NW-1: byte j&3 of P[j>>2] (low 7 bits)
//   X1 -> (b, c) = numbers of level-2 / level-3 hits; the first level-2 hit sits in the last static slot
//   a > NS, b > 1 or c > 0 (probability ~1e-3 per task): overflow stream, cbp_slow() below.
// Restated on the CPU in oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_poisson; the two agree bit for bit.
//
// The kernel is bound by the ALU pipe (one 32-lane integer instruction per 2 cycles per SM sub-partition), so the
// fast path is written instruction by instruction:
//   * the count a comes from a 1024-bucket lookup on the top 10 bits of X0 (one L1-resident 8-byte load + one compare);
//   * slot validity is byte arithmetic: byte j of (128 - a)·0x01010101 + 0x03020100 has bit 7 set iff j >= a, and one
//     LOP3 merges it over the 7 random position bits, so an invalid slot is simply a position >= 128;
//   * a position becomes a one-hot 128-bit mask by four clamped shifts `1 << (pos - 32w)` (PTX shl clamps amounts
//     above 31 to "all bits out", so words that do not own the position, and invalid slots, get 0);
//   * the flip word is the threshold function [u + m + g >= D] of the six bond planes and the hit masks (16 LOP3).
#include "common.cuh"
#include "philox.cuh"
#include "kernels.cuh"
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

__device__ __forceinline__ philox_out cbp_philox(const cbp_params &p, uint32_t ctr0, uint32_t c1, uint32_t c2)
{
    uint32_t c0 = ctr0 | p.t_hi16, c3 = p.t_lo;
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
__device__ __noinline__ cbp_hits cbp_slow(const cbp_params &p, uint32_t c1, uint32_t c2, uint32_t X0, uint32_t X1,
                                          uint32_t P0, uint32_t P1, uint32_t P2, uint32_t P3, uint32_t P4, uint32_t P5)
{
    constexpr int NS = 4 * NW - 1;
    const uint32_t *TA = p.tbl, *TB0 = TA + CBP_KA, *TB = TB0 + CBP_KR, *TC = TB + CBP_KR;
    cbp_hits r;
#pragma unroll
    for (int w = 0; w < 4; w++) r.m[w] = r.g[w] = r.h[w] = 0u;
    uint32_t sw1 = 0, sw2 = 0, sw3 = 0, Y = 0, call = NW > 2 ? 2u : 1u;
    int used = 0; bool loaded = false;
    auto fetch = [&]() {
        const philox_out o = philox4x32_10(call | p.t_hi16, c1, c2, p.t_lo, p.rk[0][0], p.rk[0][1]);
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
__device__ __forceinline__ cbp_words<NW> cbp_draw(const cbp_params &p, uint32_t c1, uint32_t c2)
{
    cbp_words<NW> r;
    r.A = cbp_philox(p, 0u, c1, c2);
    r.P[0] = r.A.z; r.P[1] = r.A.w;
    if (NW > 2) { const philox_out B = cbp_philox(p, 1u, c1, c2); r.P[2] = B.x; r.P[3] = B.y; r.P[4] = B.z; r.P[5] = B.w; }
    return r;
}

template <int D, int NW>
__device__ __forceinline__ bool cbp_task_hits(const cbp_params &p, const uint2 *__restrict__ bucket, uint32_t c1, uint32_t c2,
                                              const cbp_words<NW> &rw, uint32_t (&m)[4], uint32_t (&g)[4], uint32_t (&h)[4])
{
    constexpr int NS = 4 * NW - 1;
    const philox_out A = rw.A;
    uint32_t P[6] = { 0u, 0u, 0u, 0u, 0u, 0u };
#pragma unroll
    for (int q = 0; q < (NW > 2 ? 6 : 2); q++) P[q] = rw.P[q];
    const uint2 e = bucket[A.x >> 22];
    const uint32_t a = e.y + (A.x > e.x ? 1u : 0u);            // level-1 count (>= 64: ambiguous bucket, slow path)
    bool slow = a > (uint32_t)NS;
    if (D >= 2) slow = slow || A.y > p.tb0_1;
    const uint32_t one = p.one;                                  // 1, opaque to ptxas: keeps amt·1 - 32w an IMAD (fma pipe)
    const uint32_t kv = (128u - a) * 0x01010101u;
    uint32_t f[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) {
        const uint32_t X = kv + (0x03020100u + (uint32_t)q * 0x04040404u);   // byte j: bit 7 iff slot 4q+j >= a
        f[q] = lop3p<0xD8>(P[q], X, q == NW - 1 ? 0x00808080u : 0x80808080u); // (P & ~mask) | (X & mask)
    }
    // last static slot: first level-2 hit, valid iff b >= 1
    if (D >= 2) { if (!(A.y > p.tb0_0)) f[NW - 1] |= 0x80000000u; else f[NW - 1] &= 0x7fffffffu; }
    else f[NW - 1] |= 0x80000000u;
    uint32_t acc[4][2] = { { 0u, 0u }, { 0u, 0u }, { 0u, 0u }, { 0u, 0u } };
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const uint32_t amt = (j & 3) == 3 ? f[j >> 2] >> 24 : __byte_perm(f[j >> 2], 0u, 0x4440u + (j & 3));
        const uint32_t o[4] = { shl_clamp(amt), shl_clamp(amt * one - 32u), shl_clamp(amt * one - 64u), shl_clamp(amt * one - 96u) };
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
        g[0] = shl_clamp(amt); g[1] = shl_clamp(amt * one - 32u); g[2] = shl_clamp(amt * one - 64u); g[3] = shl_clamp(amt * one - 96u);
    }
#pragma unroll
    for (int w = 0; w < 4; w++) m[w] = NS == 1 ? (acc[w][1] | g[w]) : lop3p<P_OR3>(acc[w][0], acc[w][1], g[w]);
    if (__any_sync(__activemask(), slow)) {
        const uint32_t *TA = p.tbl, *TB0 = TA + CBP_KA, *TB = TB0 + CBP_KR, *TC = TB + CBP_KR;
        const philox_out S = cbp_philox(p, NW > 2 ? 2u : 1u, c1, c2);   // first overflow call: Y, then twelve byte slots
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
                uint32_t b = A.y > p.tb0_0 ? 1u : 0u;
                if (D == 3 && A.y > p.tc0) {      // level-3 hits: their count from X1, the level-2 count from Y
                    nc = 1u; while (A.y > TC[nc]) nc++;
                    b = 0u; while (S.x > TB[b]) b++;
                    if (b == 0u) { g[0] = g[1] = g[2] = g[3] = 0u; }   // the static level-2 slot was not a hit after all
                } else if (A.y > p.tb0_1) { b = 2u; while (A.y > TB0[b]) b++; }
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
            const uint32_t o[4] = { shl_clamp(amt), shl_clamp(amt - 32u), shl_clamp(amt - 64u), shl_clamp(amt - 96u) };
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
            const cbp_hits r = cbp_slow<NW>(p, c1, c2, A.x, A.y, P[0], P[1], P[2], P[3], P[4], P[5]);
#pragma unroll
            for (int w = 0; w < 4; w++) { m[w] = r.m[w]; g[w] = r.g[w]; h[w] = r.h[w]; }
        }
    }
    return slow;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// One task per thread, any D <= 3, any R (multiple of 32). With whole 128-replica groups (FULL) the 2D+1 spin
// words of the task are staged through shared memory with cp.async, issued before the hit masks are generated: the
// L2 latency runs under the spin-independent half of the task without holding 4·(2D+1) registers, and each thread
// reads back only its own slots (no block barrier).
template <int D, bool FULL, int NW, int MINB>
__global__ void __launch_bounds__(256, MINB) k_checkerboard_poisson(const __grid_constant__ cbp_params p, int colour)
{
    __shared__ uint4 stage[FULL ? 2 * D + 1 : 1][256];
    const int L = p.L;
    int x, y, z, g;
    if (D == 3 && p.brick) {
        // brick mapping: the block owns the active sites of a bx·by·bz brick for all G groups, so that the neighbour
        // words shared by its tasks are fetched from L2 once and served from L1 afterwards
        const int t0 = threadIdx.x, a = t0 >> p.Gshift;
        g = t0 & (p.G - 1);
        const int ax = a & ((1 << p.sh_hbx) - 1), ay = (a >> p.sh_hbx) & ((1 << p.sh_by) - 1), az = a >> (p.sh_hbx + p.sh_by);
        y = (blockIdx.y << p.sh_by) + ay; z = blockIdx.z * p.bz + az;
        x = (blockIdx.x << (p.sh_hbx + 1)) + 2 * ax + ((y + z + colour) & 1);
    } else {
        const int row_tid = blockIdx.x * blockDim.x + threadIdx.x;
        if (row_tid >= p.Lh * p.G) return;
        int xh;
        if (p.Gshift >= 0) { xh = row_tid >> p.Gshift; g = row_tid & (p.G - 1); }
        else { xh = __float2int_rz(((float)row_tid + 0.5f) * p.invG); g = row_tid - xh * p.G; } // exact for row_tid < 2^22
        y = (D >= 2) ? blockIdx.y : 0; z = (D >= 3) ? blockIdx.z : 0;
        x = 2 * xh + ((y + z + colour) & 1);
    }
    const uint32_t row = (uint32_t)L * (uint32_t)(y + L * z);
    const uint32_t i = row + x;
    uint32_t nb[2 * D];
    nb[0] = row + (x + 1 == L ? 0 : x + 1);
    nb[1] = row + (x == 0 ? L - 1 : x - 1);
    if (D >= 2) {
        nb[2] = y + 1 == L ? i - (uint32_t)(L - 1) * L : i + L;
        nb[3] = y == 0 ? i + (uint32_t)(L - 1) * L : i - L;
    }
    if (D >= 3) {
        const uint32_t LL = (uint32_t)L * L;
        nb[4] = z + 1 == L ? i - (uint32_t)(L - 1) * LL : i + LL;
        nb[5] = z == 0 ? i + (uint32_t)(L - 1) * LL : i - LL;
    }
    const uint32_t W = p.W, W4 = W >> 2;
    const int t = threadIdx.x;
    if (FULL) {
        const uint4 *sp4 = reinterpret_cast<const uint4 *>(p.spins);
        cp_async16(&stage[0][t], sp4 + (i * W4 + g));
#pragma unroll
        for (int k = 0; k < 2 * D; k++) cp_async16(&stage[1 + k][t], sp4 + (nb[k] * W4 + g));
    }
    uint32_t neg[2 * D];
    {
        const uint4 ja = __ldg(p.jmask + 2 * (size_t)i);
        neg[0] = ja.x; neg[1] = ja.y;
        if (D >= 2) { neg[2] = ja.z; neg[3] = ja.w; }
        if (D >= 3) { const uint2 jb = __ldg(reinterpret_cast<const uint2 *>(p.jmask + 2 * (size_t)i + 1)); neg[4] = jb.x; neg[5] = jb.y; }
    }
    uint32_t m[4], gg[4], h[4];
    const bool slow = cbp_task_hits<D, NW>(p, p.bucket, i, (uint32_t)g, cbp_draw<NW>(p, i, (uint32_t)g), m, gg, h);

    uint32_t sc[4], b[4][2 * D];
    if (FULL) {
        cp_async_wait_all();
        const uint4 c = stage[0][t];
        sc[0] = c.x; sc[1] = c.y; sc[2] = c.z; sc[3] = c.w;
#pragma unroll
        for (int k = 0; k < 2 * D; k++) {
            const uint4 v = stage[1 + k][t];
            b[0][k] = lop3p<P_XOR3>(sc[0], v.x, neg[k]); b[1][k] = lop3p<P_XOR3>(sc[1], v.y, neg[k]);
            b[2][k] = lop3p<P_XOR3>(sc[2], v.z, neg[k]); b[3][k] = lop3p<P_XOR3>(sc[3], v.w, neg[k]);
        }
    } else {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const bool ok = 4 * g + w < W;
            sc[w] = ok ? p.spins[i * W + 4 * g + w] : 0u;
#pragma unroll
            for (int k = 0; k < 2 * D; k++) b[w][k] = lop3p<P_XOR3>(sc[w], ok ? p.spins[nb[k] * W + 4 * g + w] : 0u, neg[k]);
        }
    }
    uint32_t fl[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        if (D == 3 && slow) b[w][0] |= h[w];     // a level-3 hit flips every lane (m = g = 1 there, so u >= 1 suffices)
        fl[w] = cbp_flip_planes<D>(b[w], m[w], gg[w]);
        if (!FULL && !(4 * g + w < W)) fl[w] = 0;
        sc[w] ^= fl[w];
    }
    if (FULL) {
        reinterpret_cast<uint4 *>(p.spins)[i * W4 + g] = make_uint4(sc[0], sc[1], sc[2], sc[3]);
        if (p.flips) reinterpret_cast<uint4 *>(p.flips)[i * W4 + g] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
    } else {
#pragma unroll
        for (int w = 0; w < 4; w++)
            if (4 * g + w < W) {
                p.spins[i * W + 4 * g + w] = sc[w];
                if (p.flips) p.flips[i * W + 4 * g + w] = fl[w];
            }
    }
}

// Persistent kernel (3D, whole groups, brick mapping): the grid is one wave of blocks and every block walks over
// bricks b = blockIdx.x, blockIdx.x + gridDim.x, ... One iteration of a thread is one task, software-pipelined over a
// single shared-memory stage: the spin words of task k+1 are requested (cp.async) as soon as those of task k have
// been read into registers, so they travel under the flip logic of task k and the whole hit generation of task k+1.
// The kernel is launched with programmatic stream serialization: its blocks may become resident while the previous
// half-sweep drains, run the spin-independent hit generation of their first task, and only then wait for the
// previous grid (griddepcontrol.wait) before touching the spins.
template <int NW, int MINB>
__global__ void __launch_bounds__(256, MINB) k_checkerboard_poisson_persist(const __grid_constant__ cbp_params p, int colour)
{
    constexpr int D = 3;
    __shared__ uint4 stage[2 * D + 1][256];
    __shared__ uint2 sbucket[CBP_BUCKETS];   // level-1 count lookup: 8 KB, read once per block
    const int t = threadIdx.x, L = p.L;
#pragma unroll
    for (int k = 0; k < CBP_BUCKETS / 256; k++) sbucket[t + 256 * k] = __ldg(p.bucket + t + 256 * k);
    __syncthreads();
    const int g = t & (p.G - 1), a = t >> p.Gshift;
    const int ax = a & ((1 << p.sh_hbx) - 1), ay = (a >> p.sh_hbx) & ((1 << p.sh_by) - 1), az = a >> (p.sh_hbx + p.sh_by);
    const uint32_t W4 = (uint32_t)p.W >> 2, LL = (uint32_t)L * L;
    const uint4 *sp4 = reinterpret_cast<const uint4 *>(p.spins);
    uint32_t i, nb[2 * D], neg[2 * D];
    auto locate = [&](int b) {          // site and neighbours of this thread's task in brick b
        const int q = __float2int_rz(((float)b + 0.5f) * p.inv_nbx);          // b / nbx, exact for b < 2^22
        const int X = b - q * p.nbx;
        const int Z = __float2int_rz(((float)q + 0.5f) * p.inv_nby), Y = q - Z * p.nby;
        const int y = (Y << p.sh_by) + ay, z = Z * p.bz + az;
        const int x = (X << (p.sh_hbx + 1)) + 2 * ax + ((y + z + colour) & 1);
        const uint32_t row = (uint32_t)L * (uint32_t)(y + L * z);
        i = row + x;
        nb[0] = row + (x + 1 == L ? 0 : x + 1);
        nb[1] = row + (x == 0 ? L - 1 : x - 1);
        nb[2] = y + 1 == L ? i - (uint32_t)(L - 1) * L : i + L;
        nb[3] = y == 0 ? i + (uint32_t)(L - 1) * L : i - L;
        nb[4] = z + 1 == L ? i - (uint32_t)(L - 1) * LL : i + LL;
        nb[5] = z == 0 ? i + (uint32_t)(L - 1) * LL : i - LL;
    };
    auto request = [&]() {              // spins of the located task -> shared-memory stage; bond signs -> registers
        cp_async16(&stage[0][t], sp4 + (i * W4 + g));
#pragma unroll
        for (int k = 0; k < 2 * D; k++) cp_async16(&stage[1 + k][t], sp4 + (nb[k] * W4 + g));
    };
    auto signs = [&]() {
        const uint4 ja = __ldg(p.jmask + 2 * (size_t)i);
        const uint2 jb = __ldg(reinterpret_cast<const uint2 *>(p.jmask + 2 * (size_t)i + 1));
        neg[0] = ja.x; neg[1] = ja.y; neg[2] = ja.z; neg[3] = ja.w; neg[4] = jb.x; neg[5] = jb.y;
    };
    int b = blockIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (b >= p.nbricks) return;
    locate(b);
    signs();
    bool first = true;
    for (;;) {
        uint32_t m[4], gg[4], h[4];
        const bool slow = cbp_task_hits<D, NW>(p, sbucket, i, (uint32_t)g, cbp_draw<NW>(p, i, (uint32_t)g), m, gg, h);
        if (first) {
            asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous half-sweep is complete and visible
            request();
            first = false;
        }
        cp_async_wait_all();
        uint32_t sc[4], bp[4][2 * D];
        {
            const uint4 c = stage[0][t];
            sc[0] = c.x; sc[1] = c.y; sc[2] = c.z; sc[3] = c.w;
#pragma unroll
            for (int k = 0; k < 2 * D; k++) {
                const uint4 v = stage[1 + k][t];
                bp[0][k] = lop3p<P_XOR3>(sc[0], v.x, neg[k]); bp[1][k] = lop3p<P_XOR3>(sc[1], v.y, neg[k]);
                bp[2][k] = lop3p<P_XOR3>(sc[2], v.z, neg[k]); bp[3][k] = lop3p<P_XOR3>(sc[3], v.w, neg[k]);
            }
        }
        const uint32_t icur = i;
        b += gridDim.x;
        const bool more = b < p.nbricks;
        if (more) { locate(b); request(); signs(); }
        uint32_t fl[4];
#pragma unroll
        for (int w = 0; w < 4; w++) {
            if (slow) bp[w][0] |= h[w];          // a level-3 hit flips every lane (m = g = 1 there, so u >= 1 suffices)
            fl[w] = cbp_flip_planes<D>(bp[w], m[w], gg[w]);
            sc[w] ^= fl[w];
        }
        reinterpret_cast<uint4 *>(p.spins)[icur * W4 + g] = make_uint4(sc[0], sc[1], sc[2], sc[3]);
        if (p.flips) reinterpret_cast<uint4 *>(p.flips)[icur * W4 + g] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
        if (!more) break;
    }
}

// Two tasks per thread: the same site for two replica groups (g and g + G/2). A third of a task's instructions is
// not Monte Carlo at all — brick decoding, neighbour indices, 64-bit address formation, the loop — and all of it is a
// function of the site only, so the pair shares it (the second task's addresses are the first's plus a constant).
// 128 threads per block cover the same 256-task brick as k_checkerboard_poisson_persist; same results bit for bit.
template <int NW, int MINB>
__global__ void __launch_bounds__(128, MINB) k_checkerboard_poisson_persist2(const __grid_constant__ cbp_params p, int colour)
{
    constexpr int D = 3, NT = 128;
    __shared__ uint4 stage[2][2 * D + 1][NT];
    __shared__ uint2 sbucket[CBP_BUCKETS];
    const int t = threadIdx.x, L = p.L;
#pragma unroll
    for (int k = 0; k < CBP_BUCKETS / NT; k++) sbucket[t + NT * k] = __ldg(p.bucket + t + NT * k);
    __syncthreads();
    const int Gh = p.G >> 1;                                   // groups per half: the thread owns g0 and g0 + Gh
    const int g0 = t & (Gh - 1), a = t >> (p.Gshift - 1);
    const int ax = a & ((1 << p.sh_hbx) - 1), ay = (a >> p.sh_hbx) & ((1 << p.sh_by) - 1), az = a >> (p.sh_hbx + p.sh_by);
    const uint32_t W4 = (uint32_t)p.W >> 2, LL = (uint32_t)L * L;
    const uint4 *sp4 = reinterpret_cast<const uint4 *>(p.spins);
    uint32_t i, nb[2 * D], neg[2 * D];
    auto locate = [&](int b) {
        const int q = __float2int_rz(((float)b + 0.5f) * p.inv_nbx);
        const int X = b - q * p.nbx;
        const int Z = __float2int_rz(((float)q + 0.5f) * p.inv_nby), Y = q - Z * p.nby;
        const int y = (Y << p.sh_by) + ay, z = Z * p.bz + az;
        const int x = (X << (p.sh_hbx + 1)) + 2 * ax + ((y + z + colour) & 1);
        const uint32_t row = (uint32_t)L * (uint32_t)(y + L * z);
        i = row + x;
        nb[0] = row + (x + 1 == L ? 0 : x + 1);
        nb[1] = row + (x == 0 ? L - 1 : x - 1);
        nb[2] = y + 1 == L ? i - (uint32_t)(L - 1) * L : i + L;
        nb[3] = y == 0 ? i + (uint32_t)(L - 1) * L : i - L;
        nb[4] = z + 1 == L ? i - (uint32_t)(L - 1) * LL : i + LL;
        nb[5] = z == 0 ? i + (uint32_t)(L - 1) * LL : i - LL;
    };
    auto request = [&]() {
        const uint4 *c = sp4 + (i * W4 + g0);
        cp_async16(&stage[0][0][t], c); cp_async16(&stage[1][0][t], c + Gh);
#pragma unroll
        for (int k = 0; k < 2 * D; k++) {
            const uint4 *n = sp4 + (nb[k] * W4 + g0);
            cp_async16(&stage[0][1 + k][t], n); cp_async16(&stage[1][1 + k][t], n + Gh);
        }
    };
    auto signs = [&]() {
        const uint4 ja = __ldg(p.jmask + 2 * (size_t)i);
        const uint2 jb = __ldg(reinterpret_cast<const uint2 *>(p.jmask + 2 * (size_t)i + 1));
        neg[0] = ja.x; neg[1] = ja.y; neg[2] = ja.z; neg[3] = ja.w; neg[4] = jb.x; neg[5] = jb.y;
    };
    int b = blockIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (b >= p.nbricks) return;
    locate(b);
    signs();
    bool first = true;
    for (;;) {
        uint32_t m[2][4], gg[2][4], h[2][4];
        bool slow[2];
        // both Philox chains in one basic block: their rounds interleave (each chain alone waits on its own latency)
        const cbp_words<NW> rw0 = cbp_draw<NW>(p, i, (uint32_t)g0), rw1 = cbp_draw<NW>(p, i, (uint32_t)(g0 + Gh));
        slow[0] = cbp_task_hits<D, NW>(p, sbucket, i, (uint32_t)g0, rw0, m[0], gg[0], h[0]);
        slow[1] = cbp_task_hits<D, NW>(p, sbucket, i, (uint32_t)(g0 + Gh), rw1, m[1], gg[1], h[1]);
        if (first) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            request();
            first = false;
        }
        cp_async_wait_all();
        uint4 *out = reinterpret_cast<uint4 *>(p.spins) + (i * W4 + g0);
        uint4 *outf = p.flips ? reinterpret_cast<uint4 *>(p.flips) + (i * W4 + g0) : nullptr;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint4 c = stage[j][0][t];
            uint32_t sc[4] = { c.x, c.y, c.z, c.w }, bp[4][2 * D], fl[4];
#pragma unroll
            for (int k = 0; k < 2 * D; k++) {
                const uint4 v = stage[j][1 + k][t];
                bp[0][k] = lop3p<P_XOR3>(sc[0], v.x, neg[k]); bp[1][k] = lop3p<P_XOR3>(sc[1], v.y, neg[k]);
                bp[2][k] = lop3p<P_XOR3>(sc[2], v.z, neg[k]); bp[3][k] = lop3p<P_XOR3>(sc[3], v.w, neg[k]);
            }
#pragma unroll
            for (int w = 0; w < 4; w++) {
                if (slow[j]) bp[w][0] |= h[j][w];
                fl[w] = cbp_flip_planes<D>(bp[w], m[j][w], gg[j][w]);
                sc[w] ^= fl[w];
            }
            out[j * Gh] = make_uint4(sc[0], sc[1], sc[2], sc[3]);
            if (outf) outf[j * Gh] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
        }
        b += gridDim.x;
        if (b >= p.nbricks) break;
        locate(b); request(); signs();
    }
}

template <int NW, int MINB>
static cudaError_t launch_persist2_one(const cbp_params &p, int colour, int sm_count, cudaStream_t st)
{
    static int occ = 0;
    if (!occ) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_checkerboard_poisson_persist2<NW, MINB>, 128, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    int grid = sm_count * occ;
    if (p.variant & 128) grid = 2;
    if (grid > p.nbricks) grid = p.nbricks;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (p.variant & 32) ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, k_checkerboard_poisson_persist2<NW, MINB>, p, colour);
}
template <int MINB>
static cudaError_t launch_persist2(const cbp_params &p, int colour, int sm_count, cudaStream_t st)
{
    switch (p.NW) {
    case 1: return launch_persist2_one<1, MINB>(p, colour, sm_count, st);
    case 2: return launch_persist2_one<2, MINB>(p, colour, sm_count, st);
    case 4: return launch_persist2_one<4, MINB>(p, colour, sm_count, st);
    default: return launch_persist2_one<6, MINB>(p, colour, sm_count, st);
    }
}

template <int NW, int MINB>
static cudaError_t launch_persist_one(const cbp_params &p, int colour, int sm_count, cudaStream_t st)
{
    static int occ = 0;   // resident blocks per SM of this instantiation
    if (!occ) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_checkerboard_poisson_persist<NW, MINB>, 256, 0);
        if (e != cudaSuccess) return e;
        if (occ < 1) occ = 1;
    }
    int grid = sm_count * occ;
    if (p.variant & 128) grid = 2;   // tests: many iterations per block
    if (grid > p.nbricks) grid = p.nbricks;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (p.variant & 32) ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, k_checkerboard_poisson_persist<NW, MINB>, p, colour);
}
template <int MINB>
static cudaError_t launch_persist(const cbp_params &p, int colour, int sm_count, cudaStream_t st)
{
    switch (p.NW) {
    case 1: return launch_persist_one<1, MINB>(p, colour, sm_count, st);
    case 2: return launch_persist_one<2, MINB>(p, colour, sm_count, st);
    case 4: return launch_persist_one<4, MINB>(p, colour, sm_count, st);
    default: return launch_persist_one<6, MINB>(p, colour, sm_count, st);
    }
}

template <int D, bool FULL, int MINB>
static void launch_nw(const cbp_params &p, int colour, dim3 grid, dim3 block, cudaStream_t st)
{
    switch (p.NW) {
    case 1: k_checkerboard_poisson<D, FULL, 1, MINB><<<grid, block, 0, st>>>(p, colour); break;
    case 2: k_checkerboard_poisson<D, FULL, 2, MINB><<<grid, block, 0, st>>>(p, colour); break;
    case 4: k_checkerboard_poisson<D, FULL, 4, MINB><<<grid, block, 0, st>>>(p, colour); break;
    default: k_checkerboard_poisson<D, FULL, 6, MINB><<<grid, block, 0, st>>>(p, colour); break;
    }
}

rrrmc_status_t launch_checkerboard_poisson(rrrmc_ctx *ctx, cbp_params &p, int D, int colour)
{
    const bool full = (p.W % 4) == 0;
    if (!(p.NW == 1 || p.NW == 2 || p.NW == 4 || p.NW == 6)) { rrrmc_set_error("checkerboard poisson: NW=%d unsupported (1, 2, 4, 6)", p.NW); return RRRMC_ERR_ARG; }
    dim3 block(256), grid(div_up((int64_t)p.Lh * p.G, 256), D >= 2 ? p.L : 1, D >= 3 ? p.L : 1);
    // brick mapping (3D, whole groups, G a power of two <= 256): 256/G active sites = a brick of 512/G sites, as cubic
    // as the lattice side allows; every side is a power of two that divides L
    p.brick = 0;
    if (D == 3 && full && p.Gshift >= 0 && p.G <= 128 && !(p.variant & 16)) {
        int sh[3] = { 1, 0, 0 }, left = 9 - p.Gshift - 1;   // log2 sides; bx >= 2
        auto fits = [&](int s) { return (p.L % (1 << s)) == 0; };
        bool ok = fits(1);
        while (ok && left > 0) {
            int best = -1;
            for (int d = 2; d >= 0; d--) if (fits(sh[d] + 1) && (best < 0 || sh[d] < sh[best])) best = d;
            if (best < 0) { ok = false; break; }
            sh[best]++; left--;
        }
        if (ok) {
            p.brick = 1; p.sh_hbx = sh[0] - 1; p.sh_by = sh[1]; p.bz = 1 << sh[2];
            grid = dim3(p.L >> sh[0], p.L >> sh[1], p.L >> sh[2]);
            p.nbx = (int)grid.x; p.nby = (int)grid.y; p.nbricks = (int)(grid.x * grid.y * grid.z);
            p.inv_nbx = 1.0f / (float)p.nbx; p.inv_nby = 1.0f / (float)p.nby;
        }
    }
    if (p.brick && p.nbricks < (1 << 22) && !(p.variant & 8)) {   // persistent, software-pipelined kernel
        const int mb = p.variant & 3;   // RRRMC_CB_VARIANT (tuning): resident blocks per SM; 3 (80 registers) measured best
        // two tasks (groups g, g + G/2) per thread: measured 3-4 % faster where the hit generation is long (NW >= 2),
        // slower where it is short. RRRMC_CB_VARIANT bit 9 forces it, bit 10 forbids it (tuning, tests).
        if (p.G >= 2 && !(p.variant & 1024) && (p.NW >= 2 || (p.variant & 512))) {
            cudaError_t e2 = mb == 1 ? launch_persist2<6>(p, colour, ctx->sm_count, ctx->stream)
                           : mb == 2 ? launch_persist2<5>(p, colour, ctx->sm_count, ctx->stream)
                           : mb == 3 ? launch_persist2<3>(p, colour, ctx->sm_count, ctx->stream)
                                     : launch_persist2<4>(p, colour, ctx->sm_count, ctx->stream);
            ctx->launches++;
            RR_CUDA(e2);
            RR_CUDA(cudaGetLastError());
            return RRRMC_OK;
        }
        cudaError_t e = mb == 1 ? launch_persist<5>(p, colour, ctx->sm_count, ctx->stream)
                      : mb == 2 ? launch_persist<2>(p, colour, ctx->sm_count, ctx->stream)
                      : mb == 3 ? launch_persist<4>(p, colour, ctx->sm_count, ctx->stream)
                                : launch_persist<3>(p, colour, ctx->sm_count, ctx->stream);
        ctx->launches++;
        RR_CUDA(e);
        RR_CUDA(cudaGetLastError());
        return RRRMC_OK;
    }
    if (D == 1) { if (full) launch_nw<1, true, 1>(p, colour, grid, block, ctx->stream); else launch_nw<1, false, 1>(p, colour, grid, block, ctx->stream); }
    else if (D == 2) { if (full) launch_nw<2, true, 1>(p, colour, grid, block, ctx->stream); else launch_nw<2, false, 1>(p, colour, grid, block, ctx->stream); }
    else if (D == 3) {
        if (!full) launch_nw<3, false, 1>(p, colour, grid, block, ctx->stream);
        else if ((p.variant & 3) == 1) launch_nw<3, true, 5>(p, colour, grid, block, ctx->stream);   // RRRMC_CB_VARIANT: tuning
        else if ((p.variant & 3) == 2) launch_nw<3, true, 6>(p, colour, grid, block, ctx->stream);
        else if ((p.variant & 3) == 3) launch_nw<3, true, 3>(p, colour, grid, block, ctx->stream);
        else launch_nw<3, true, 4>(p, colour, grid, block, ctx->stream);
    }
    else { rrrmc_set_error("checkerboard: D=%d unsupported (1..3)", D); return RRRMC_ERR_UNSUPPORTED; }
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}
