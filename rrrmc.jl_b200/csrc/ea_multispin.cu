// Multispin (bit-packed over replicas) kernels for GraphEA ±J on the periodic hyper-cubic lattice.
//
// Layout in HBM: spins[N][W] uint32, site-major / replica-minor; bit b of word w of site i is the
// spin s∈{0,1} of replica 32w+b (σ=2s-1, Interface.jl:34-37); site i = x + L*y + L*L*z is the
// reference's 1-based index minus one (EA.jl:31-35). One 128-bit load serves 128 replicas.
// Couplings are shared by all replicas: jcode[i] holds the sign bits of the 2D bonds of site i.
//
// ΔE (EA.jl:266-275, naive form :277-289) is evaluated bit-sliced: per bond k the plane
// b_k = s_i ^ s_k ^ neg_k is 1 where the bond is unsatisfied; with u = Σ_k b_k (carry-save adders),
// ΔE = 4(D-u) ∈ {-4D..4D}, i.e. exactly -lfields[i] of the reference cache.
#include "common.cuh"
#include "philox.cuh"
#include "kernels.cuh"
#include "cb_params.cuh"

// ------------------------------------------------------------------------------------------------
// bit-sliced helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void full_add(uint32_t a, uint32_t b, uint32_t c, uint32_t &s, uint32_t &cy)
{
    s = a ^ b ^ c;
    cy = (a & b) | (c & (a ^ b));
}

template <int D> struct site_geom {
    int64_t i;            // site
    int64_t nb[2 * D];    // neighbour sites in (d+, d-) order
};

template <int D>
__device__ __forceinline__ site_geom<D> make_geom(int L, int x, int y, int z)
{
    site_geom<D> s;
    const int64_t row = (int64_t)L * (y + (int64_t)L * z);
    s.i = row + x;
    s.nb[0] = row + (x + 1 == L ? 0 : x + 1);
    s.nb[1] = row + (x == 0 ? L - 1 : x - 1);
    if (D >= 2) {
        s.nb[2] = s.i + (y + 1 == L ? -(int64_t)(L - 1) * L : (int64_t)L);
        s.nb[3] = s.i + (y == 0 ? (int64_t)(L - 1) * L : -(int64_t)L);
    }
    if (D >= 3) {
        const int64_t LL = (int64_t)L * L;
        s.nb[4] = s.i + (z + 1 == L ? -(int64_t)(L - 1) * LL : LL);
        s.nb[5] = s.i + (z == 0 ? (int64_t)(L - 1) * LL : -LL);
    }
    return s;
}

// number of unsatisfied bonds among the 2D bonds of a site, as bit planes (u2,u1,u0)
template <int D>
__device__ __forceinline__ void unsat_planes(uint32_t sc, const uint32_t (&sn)[2 * D], uint32_t jc,
                                             uint32_t &u0, uint32_t &u1, uint32_t &u2)
{
    uint32_t b[2 * D];
#pragma unroll
    for (int k = 0; k < 2 * D; k++) {
        const uint32_t neg = 0u - ((jc >> k) & 1u);
        b[k] = sc ^ sn[k] ^ neg;
    }
    if (D == 1) { u0 = b[0] ^ b[1]; u1 = b[0] & b[1]; u2 = 0; }
    if (D == 2) {
        uint32_t s1, c1; full_add(b[0], b[1], b[2], s1, c1);
        u0 = s1 ^ b[3];
        const uint32_t c2 = s1 & b[3];
        u1 = c1 ^ c2; u2 = c1 & c2;
    }
    if (D == 3) {
        uint32_t s1, c1, s2, c2; full_add(b[0], b[1], b[2], s1, c1); full_add(b[3], b[4], b[5], s2, c2);
        u0 = s1 ^ s2;
        const uint32_t c3 = s1 & s2;
        full_add(c1, c2, c3, u1, u2);
    }
}

// class masks: mc[c-1] = lanes with ΔE = 4c > 0, i.e. u = D-c
template <int D>
__device__ __forceinline__ void class_masks(uint32_t u0, uint32_t u1, uint32_t u2, uint32_t (&mc)[3])
{
    if (D == 1) { mc[0] = ~u1 & ~u0; mc[1] = 0; mc[2] = 0; }                        // u=0
    if (D == 2) { mc[0] = ~u2 & ~u1 & u0; mc[1] = ~u2 & ~u1 & ~u0; mc[2] = 0; }     // u=1, u=0
    if (D == 3) { mc[0] = ~u2 & u1 & ~u0; mc[1] = ~u2 & ~u1 & u0; mc[2] = ~(u2 | u1 | u0); } // u=2,1,0
}

// unsat_planes with the bond sign masks already expanded
template <int D>
__device__ __forceinline__ void unsat_planes_neg(uint32_t sc, const uint32_t (&sn)[2 * D], const uint32_t (&neg)[2 * D],
                                                 uint32_t &u0, uint32_t &u1, uint32_t &u2)
{
    uint32_t b[2 * D];
#pragma unroll
    for (int k = 0; k < 2 * D; k++) b[k] = sc ^ sn[k] ^ neg[k];
    if (D == 1) { u0 = b[0] ^ b[1]; u1 = b[0] & b[1]; u2 = 0; }
    if (D == 2) {
        uint32_t s1, c1; full_add(b[0], b[1], b[2], s1, c1);
        u0 = s1 ^ b[3];
        const uint32_t c2 = s1 & b[3];
        u1 = c1 ^ c2; u2 = c1 & c2;
    }
    if (D == 3) {
        uint32_t s1, c1, s2, c2; full_add(b[0], b[1], b[2], s1, c1); full_add(b[3], b[4], b[5], s2, c2);
        u0 = s1 ^ s2;
        const uint32_t c3 = s1 & s2;
        full_add(c1, c2, c3, u1, u2);
    }
}

// ------------------------------------------------------------------------------------------------
// Checkerboard Metropolis half-sweep. One thread = one task (site of the active colour, group of
// 128 replicas). Acceptance is Metropolis (RRRMC.jl:39: ΔE<=0 always, else U<exp(-βΔE)) with U built
// from Philox bits — procedure documented in DESIGN.md §5 and restated on the CPU in
// oracle/rrrmc_oracle.c:orc_checkerboard_sweeps (the two must agree bit for bit).
//
// Cost model (ncu, profiles/): the kernel is integer-issue bound (LOP3 on the ALU pipe for the bit-sliced
// logic + Philox xors, IMAD.WIDE on the FMA pipe for the Philox multiplies), not HBM bound. Hence:
//  * the leading planes whose threshold bit is the same for every class do not depend on the spins: they
//    run first, while the seven 128-bit loads of the task are still in flight;
//  * once undecided lanes are sparse the four words of the task are overlaid on one ("merged planes"), so
//    one Philox call serves four planes instead of one;
//  * the per-lane 32-bit tail is a rare slow path;
//  * index arithmetic is 32-bit with a float-reciprocal split of (site, group).
// ------------------------------------------------------------------------------------------------
// Philox4x32-10 with the ten round keys precomputed on the host (they are kernel-uniform): the xors read them
// straight from the constant bank instead of re-deriving them on the uniform datapath in every call.
template <class PARAMS>
__device__ __forceinline__ philox_out philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PARAMS &p)
{
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
#define CB_PHILOX(ctr0) philox4x32_10_rk((uint32_t)(ctr0) | p.t_hi16, c1, c2, p.t_lo, p)

template <int D, bool FULL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_checkerboard(const __grid_constant__ cb_params p, int colour)
{
    const int row_tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (row_tid >= p.Lh * p.G) return;
    const int xh = __float2int_rz(((float)row_tid + 0.5f) * p.invG);   // row_tid / G, exact for row_tid < 2^22
    const int g = row_tid - xh * p.G;
    const int y = (D >= 2) ? blockIdx.y : 0, z = (D >= 3) ? blockIdx.z : 0;
    const int L = p.L;
    const int x = 2 * xh + ((y + z + colour) & 1);
    // site index and neighbour sites (32-bit; N*W < 2^31 words is checked on the host)
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
    const uint32_t W = p.W;

    uint32_t sc[4], sn[4][2 * D];
    if (FULL) {
        const uint4 *sp = reinterpret_cast<const uint4 *>(p.spins);
        const uint32_t W4 = W >> 2;
        const uint4 c = sp[i * W4 + g];
        sc[0] = c.x; sc[1] = c.y; sc[2] = c.z; sc[3] = c.w;
#pragma unroll
        for (int k = 0; k < 2 * D; k++) {
            const uint4 v = sp[nb[k] * W4 + g];
            sn[0][k] = v.x; sn[1][k] = v.y; sn[2][k] = v.z; sn[3][k] = v.w;
        }
    } else {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const bool ok = 4 * g + w < W;
            sc[w] = ok ? p.spins[i * W + 4 * g + w] : 0u;
#pragma unroll
            for (int k = 0; k < 2 * D; k++) sn[w][k] = ok ? p.spins[nb[k] * W + 4 * g + w] : 0u;
        }
    }
    const uint32_t jc = p.jcode[i];
    const uint32_t c1 = i, c2 = (uint32_t)g;

    // ΔE classes of the four words (spin dependent)
    uint32_t mc[4][3], up[4], eq[4], lt[4];
    // acc = "U < thr decided", und = "U == thr so far", before the spins are known
    uint32_t acc[4] = { 0u, 0u, 0u, 0u }, und[4] = { ~0u, ~0u, ~0u, ~0u };
    auto classes = [&]() {
        // keep the spin-dependent work behind the spin-independent planes: the loads issued above stay in flight
        // while the first Philox calls run (ptxas would otherwise consume them first to free registers).
        // p.zero is always 0; the data dependency on `und` is what pins the order.
#pragma unroll
        for (int w = 0; w < 4; w++) sc[w] ^= und[w] & p.zero;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t u0, u1, u2;
            unsat_planes<D>(sc[w], sn[w], jc, u0, u1, u2);
            class_masks<D>(u0, u1, u2, mc[w]);
            up[w] = mc[w][0] | mc[w][1] | mc[w][2];
            if (!FULL && !(4 * g + w < W)) { up[w] = 0; mc[w][0] = mc[w][1] = mc[w][2] = 0; }
            eq[w] = up[w] & und[w]; lt[w] = up[w] & acc[w];
        }
    };
    // merged planes (see below): mg = lanes of each word that own their bit position, em/lm = eq/lt of the overlay
    uint32_t mg[4] = { 0u, 0u, 0u, 0u }, mm[3] = { 0u, 0u, 0u }, em = 0u, lm = 0u;

    // plane phase, part 1: bit q (from the MSB) of U for every lane of the task comes from Philox call q. While the
    // threshold bit is class-independent (q < Ku) the comparison needs no spins.
    for (int q = 0; q < p.Kz; q++) {   // leading zeros of every threshold: lanes with U bit 1 are rejected
        const philox_out r = CB_PHILOX(q);
        und[0] &= ~r.x; und[1] &= ~r.y; und[2] &= ~r.z; und[3] &= ~r.w;
    }
    for (int q = p.Kz; q < p.Ku; q++) {
        const philox_out r = CB_PHILOX(q);
        const uint32_t rr[4] = { r.x, r.y, r.z, r.w };
        if (p.planeop[q] == 0) {
#pragma unroll
            for (int w = 0; w < 4; w++) und[w] &= ~rr[w];
        } else {                       // threshold bit 1: lanes with U bit 0 are accepted
#pragma unroll
            for (int w = 0; w < 4; w++) { acc[w] |= und[w] & ~rr[w]; und[w] &= rr[w]; }
        }
    }
    classes();
    // plane phase, part 2: the remaining full planes
    for (int q = p.Ku; q < p.K; q++) {
        const philox_out r = CB_PHILOX(q);
        const uint32_t rr[4] = { r.x, r.y, r.z, r.w };
        const int op = p.planeop[q];
        if (op == 0) {
#pragma unroll
            for (int w = 0; w < 4; w++) eq[w] &= ~rr[w];
        } else if (op == 1) {
#pragma unroll
            for (int w = 0; w < 4; w++) { lt[w] |= eq[w] & ~rr[w]; eq[w] &= rr[w]; }
        } else {
            const uint32_t B0 = p.plane[q][0], B1 = p.plane[q][1], B2 = p.plane[q][2];
#pragma unroll
            for (int w = 0; w < 4; w++) {
                const uint32_t thr = (mc[w][0] & B0) | (mc[w][1] & B1) | (mc[w][2] & B2);
                lt[w] |= eq[w] & ~rr[w] & thr;
                eq[w] &= ~(rr[w] ^ thr);
            }
        }
    }
    // merged planes: undecided lanes are sparse now, so overlay the four words on one. At every bit position the
    // lowest word with an undecided lane owns the merged lane; bit K+j of its U is that bit of word j%4 of call
    // K+j/4. Lanes shadowed at their position stay undecided with K bits consumed (slow path).
    const int ncalls = p.K + (p.M >> 2);
    if (p.M > 0) {
        mg[0] = eq[0]; mg[1] = eq[1] & ~eq[0]; mg[2] = eq[2] & ~(eq[0] | eq[1]); mg[3] = eq[3] & ~(eq[0] | eq[1] | eq[2]);
        em = eq[0] | eq[1] | eq[2] | eq[3];
#pragma unroll
        for (int c = 0; c < 3; c++) mm[c] = (mc[0][c] & mg[0]) | (mc[1][c] & mg[1]) | (mc[2][c] & mg[2]) | (mc[3][c] & mg[3]);
        for (int call = p.K; call < ncalls; call++) {
            const philox_out r = CB_PHILOX(call);
            const uint32_t rr[4] = { r.x, r.y, r.z, r.w };
            const int q0 = p.K + ((call - p.K) << 2);
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const uint32_t thr = (mm[0] & p.plane[q0 + jj][0]) | (mm[1] & p.plane[q0 + jj][1]) | (mm[2] & p.plane[q0 + jj][2]);
                lm |= em & ~rr[jj] & thr;
                em &= ~(rr[jj] ^ thr);
            }
        }
#pragma unroll
        for (int w = 0; w < 4; w++) { lt[w] |= lm & mg[w]; eq[w] &= em | ~mg[w]; }
    }

    // slow path (rare): the n-th lane still undecided, in ascending (word, bit) order, takes word n%4 of call
    // K+M/4+n/4 against the next 32 bits of its threshold.
    if (eq[0] | eq[1] | eq[2] | eq[3]) {
        int n = 0;
        philox_out r = CB_PHILOX(ncalls);
#pragma unroll
        for (int w = 0; w < 4; w++) {
            uint32_t e = eq[w];
            while (e) {
                const uint32_t bit = e & (0u - e);
                e ^= bit;
                const int c = (mc[w][0] & bit) ? 0 : ((mc[w][1] & bit) ? 1 : 2);
                const uint32_t rem = (mg[w] & bit) ? p.remM[c] : p.rem[c];
                if (n >= 4 && (n & 3) == 0) r = CB_PHILOX(ncalls + (n >> 2));
                const int m = n & 3;
                const uint32_t V = m == 0 ? r.x : (m == 1 ? r.y : (m == 2 ? r.z : r.w));
                if (V < rem) lt[w] |= bit;
                n++;
            }
        }
    }
    uint32_t fl[4];
#pragma unroll
    for (int w = 0; w < 4; w++) {
        fl[w] = ~up[w] | lt[w];
        if (!FULL && !(4 * g + w < W)) fl[w] = 0;
        sc[w] ^= fl[w];
    }
    if (FULL) {
        const uint32_t W4 = W >> 2;
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
#undef CB_PHILOX

rrrmc_status_t launch_checkerboard(rrrmc_ctx *ctx, const cb_params &p, int D, int colour)
{
    const bool full = (p.W % 4) == 0;
    const int bs = (p.variant & 8) ? 128 : 256;
    dim3 block(bs), grid(div_up((int64_t)p.Lh * p.G, bs), D >= 2 ? p.L : 1, D >= 3 ? p.L : 1);
#define LAUNCH(DD, FF, MB) k_checkerboard<DD, FF, MB><<<grid, block, 0, ctx->stream>>>(p, colour)
    if (D == 1) { if (full) LAUNCH(1, true, 1); else LAUNCH(1, false, 1); }
    else if (D == 2) { if (full) LAUNCH(2, true, 1); else LAUNCH(2, false, 1); }
    else if (D == 3) {
        if (!full) LAUNCH(3, false, 1);
        else if ((p.variant & 7) == 1) LAUNCH(3, true, 5);
        else if ((p.variant & 7) == 2) LAUNCH(3, true, 6);
        else if ((p.variant & 7) == 3) LAUNCH(3, true, 3);
        else LAUNCH(3, true, 4);
    }
    else { rrrmc_set_error("checkerboard: D=%d unsupported (1..3)", D); return RRRMC_ERR_UNSUPPORTED; }
#undef LAUNCH
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

// ------------------------------------------------------------------------------------------------
// Checkerboard Metropolis half-sweep, "sparse" acceptance procedure. Same task decomposition as k_checkerboard.
// At low temperature almost every lane with ΔE>0 is rejected, so instead of comparing one uniform per lane the
// task samples the SET of passing lanes of each ΔE class directly: the number of passing lanes is binomial
// (inverse CDF on one 32-bit uniform against a host-built table), their positions are uniform and distinct
// (duplicates redrawn). Lane l of class c flips iff l is in the set of class c, i.e. with probability p_c,
// independently across lanes, exactly as accept() of RRRMC.jl:39 prescribes. Random words of a task:
//   call 0: words 0..3 = count uniforms of class 1 (ΔE=4) for the four 32-lane words of the task
//   call 1: S = word0 | word1<<32: bits [0,40) eight 5-bit static slots (word w draws 2w, 2w+1), bits [40,61) the
//           first three 7-bit slots of the overflow stream; word 2 / 3 = count uniform of class 2 / 3 (128 lanes)
//   call 2+k: words 0,1 = nine more 7-bit overflow slots (rare)
// The overflow stream serves the 3rd.. draws and redraws of class-1 words 0..3, then class 2, then class 3.
// Restated on the CPU in oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_sparse (bit-for-bit).
// The common case (<= 2 passing lanes per word, no duplicate, no class-2/3 lane) is branch free.
// ------------------------------------------------------------------------------------------------
struct cbs_stream { uint32_t y0, y1, call; int left; };
// refill of the overflow window: rare, kept out of line so that the hot path stays small
__device__ __noinline__ uint2 cbs_refill(uint32_t ctr0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
    const philox_out r = philox4x32_10(ctr0, c1, c2, c3, k0, k1);
    return make_uint2(r.x, r.y);
}

// one LOP3 with an explicit truth table (inputs a=0xF0, b=0xCC, c=0xAA): the bit-sliced ΔE logic is the ALU-pipe
// bottleneck of this kernel, so its operation count is pinned by hand instead of left to the optimiser
template <int LUT> __device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
    return r;
}
constexpr int LUT_XOR3 = 0x96, LUT_MAJ = 0xE8;

// flip mask of one 32-lane word: lanes with ΔE<=0 plus the class-1 lanes (ΔE=4) that are in the pass set m.
// 3D: with (s1,k1), (s2,k2) the full-adder outputs of bonds 0-2 and 3-5, u = s1+s2+2(k1+k2) and
//     [u>=3] | ([u==2] & m) = MAJ(k1, k2, s1|s2|m) | (s1&s2&m)  — 14 LOP3 per word including the six bond planes.
template <int D>
__device__ __forceinline__ uint32_t cbs_flip_word(uint32_t sc, const uint32_t (&sn)[2 * D], const uint32_t (&neg)[2 * D], uint32_t m)
{
    uint32_t b[2 * D];
#pragma unroll
    for (int k = 0; k < 2 * D; k++) b[k] = lop3<LUT_XOR3>(sc, sn[k], neg[k]);
    if (D == 1) return lop3<0xFE>(b[0], b[1], m);                       // u>=1 free; u==0 needs m
    if (D == 2) {                                                        // u>=2 free; u==1 needs m
        const uint32_t s1 = lop3<LUT_XOR3>(b[0], b[1], b[2]), k1 = lop3<LUT_MAJ>(b[0], b[1], b[2]);
        return k1 | lop3<LUT_MAJ>(s1, b[3], m);
    }
    const uint32_t s1 = lop3<LUT_XOR3>(b[0], b[1], b[2]), k1 = lop3<LUT_MAJ>(b[0], b[1], b[2]);
    const uint32_t s2 = lop3<LUT_XOR3>(b[3], b[4], b[5]), k2 = lop3<LUT_MAJ>(b[3], b[4], b[5]);
    const uint32_t f1 = lop3<0xFE>(s1, s2, m), f2 = lop3<0x80>(s1, s2, m);
    return lop3<LUT_MAJ>(k1, k2, f1) | f2;
}

#define CB_PHILOX(ctr0) philox4x32_10_rk((uint32_t)(ctr0) | p.t_hi16, c1, c2, p.t_lo, p)
// One task of the sparse procedure: the flip masks fl[4] of (site c1, group c2) from the centre words sc, the
// neighbour words sn and the bond sign masks neg. Shared by the generic and the row-chunk kernels.
template <int D>
__device__ __forceinline__ void cbs_task(const cbs_params &p, uint32_t c1, uint32_t c2, const uint32_t (&sc)[4],
                                         const uint32_t (&sn)[4][2 * D], const uint32_t (&neg)[2 * D], uint32_t (&fl)[4])
{
    // ---- spin-independent part: the sets of passing lanes of class 1
    const philox_out A = CB_PHILOX(0), B = CB_PHILOX(1);
    const uint32_t xa[4] = { A.x, A.y, A.z, A.w };
    const uint32_t T0 = p.tbl[0], T1 = p.tbl[1], T2 = p.tbl[2];
    uint32_t m[4], need = 0;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        // static slots: bit offsets 10w and 10w+5 of S = B.x | B.y << 32
        const uint32_t q0 = 10 * w < 32 ? __funnelshift_r(B.x, B.y, 10 * w) : B.y >> (10 * w - 32);
        const uint32_t q1 = 10 * w + 5 < 32 ? __funnelshift_r(B.x, B.y, 10 * w + 5) : B.y >> (10 * w + 5 - 32);
        const uint32_t b0 = 1u << (q0 & 31u), b1 = 1u << (q1 & 31u);
        const bool ge1 = xa[w] > T0, ge2 = xa[w] > T1, ge3 = xa[w] > T2;
        m[w] = (ge1 ? b0 : 0u) | (ge2 ? b1 : 0u);
        need |= (ge3 || (ge2 && b0 == b1)) ? (1u << w) : 0u;   // a third draw is due: overflow stream
    }
    cbs_stream st; st.y0 = B.y >> 8; st.y1 = 0u; st.left = 3; st.call = 2;   // S >> 40
    auto slot = [&]() -> uint32_t {     // next 7-bit slot of the task's overflow stream
        if (st.left == 0) {
            const uint2 r = cbs_refill(st.call | p.t_hi16, c1, c2, p.t_lo, p.rk[0][0], p.rk[0][1]);
            st.call++; st.y0 = r.x; st.y1 = r.y; st.left = 9;
        }
        const uint32_t v = st.y0 & 127u;
        st.y0 = __funnelshift_r(st.y0, st.y1, 7); st.y1 >>= 7; st.left--;
        return v;
    };
    if (need) {
#pragma unroll
        for (int w = 0; w < 4; w++)
            if (need & (1u << w)) {
                uint32_t mm = m[w]; int s = __popc(mm);
                while (xa[w] > p.tbl[s]) {  // tbl[32] = 2^32-1 ends the scan
                    const uint32_t nm = mm | (1u << (slot() & 31u));
                    s += nm != mm;          // a duplicate position is redrawn
                    mm = nm;
                }
                m[w] = mm;
            }
    }
    // ---- spin-dependent part
#pragma unroll
    for (int w = 0; w < 4; w++) fl[w] = cbs_flip_word<D>(sc[w], sn[w], neg, m[w]);
    if (D >= 2) {   // classes 2..D: rare at the temperatures where this procedure is selected
        const uint32_t xc[2] = { B.z, B.w };
        bool rare = xc[0] > p.tbl[CBS_T1];
        if (D == 3) rare = rare || xc[1] > p.tbl[CBS_T1 + CBS_TC];
        if (rare) {
            uint32_t pm[2][4] = { { 0u, 0u, 0u, 0u }, { 0u, 0u, 0u, 0u } };
#pragma unroll
            for (int c = 2; c <= D; c++) {
                const uint32_t *T = p.tbl + CBS_T1 + (c - 2) * CBS_TC;
                int s = 0;
                while (xc[c - 2] > T[s]) {  // T[128] = 2^32-1
                    const uint32_t pos = slot();
                    const uint32_t bit = 1u << (pos & 31u);
                    const int ww = (int)(pos >> 5);
                    bool dup = false;
#pragma unroll
                    for (int w = 0; w < 4; w++) if (w == ww) { dup = (pm[c - 2][w] & bit) != 0; pm[c - 2][w] |= bit; }
                    s += !dup;
                }
            }
#pragma unroll
            for (int w = 0; w < 4; w++)
                if (pm[0][w] | pm[1][w]) {   // the ΔE planes are rebuilt only for the words that drew a position
                    uint32_t u0, u1, u2, mc[3];
                    unsat_planes_neg<D>(sc[w], sn[w], neg, u0, u1, u2);
                    class_masks<D>(u0, u1, u2, mc);
                    fl[w] |= (mc[1] & pm[0][w]) | (mc[2] & pm[1][w]);
                }
        }
    }
}
#undef CB_PHILOX

// generic kernel: one task per thread, any D <= 3, any R (multiple of 32)
template <int D, bool FULL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_checkerboard_sparse(const __grid_constant__ cbs_params p, int colour)
{
    const int row_tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (row_tid >= p.Lh * p.G) return;
    int xh, g;
    if (p.Gshift >= 0) { xh = row_tid >> p.Gshift; g = row_tid & (p.G - 1); }
    else { xh = __float2int_rz(((float)row_tid + 0.5f) * p.invG); g = row_tid - xh * p.G; } // exact for row_tid < 2^22
    const int y = (D >= 2) ? blockIdx.y : 0, z = (D >= 3) ? blockIdx.z : 0;
    const int L = p.L;
    const int x = 2 * xh + ((y + z + colour) & 1);
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

    uint32_t sc[4], sn[4][2 * D];
    if (FULL) {
        const uint4 *sp4 = reinterpret_cast<const uint4 *>(p.spins);
        const uint4 c = sp4[i * W4 + g];
        sc[0] = c.x; sc[1] = c.y; sc[2] = c.z; sc[3] = c.w;
#pragma unroll
        for (int k = 0; k < 2 * D; k++) {
            const uint4 v = sp4[nb[k] * W4 + g];
            sn[0][k] = v.x; sn[1][k] = v.y; sn[2][k] = v.z; sn[3][k] = v.w;
        }
    } else {
#pragma unroll
        for (int w = 0; w < 4; w++) {
            const bool ok = 4 * g + w < W;
            sc[w] = ok ? p.spins[i * W + 4 * g + w] : 0u;
#pragma unroll
            for (int k = 0; k < 2 * D; k++) sn[w][k] = ok ? p.spins[nb[k] * W + 4 * g + w] : 0u;
        }
    }
    uint32_t neg[2 * D];
    const uint32_t jc = p.jcode[i];
#pragma unroll
    for (int k = 0; k < 2 * D; k++) neg[k] = 0u - ((jc >> k) & 1u);

    uint32_t fl[4];
    cbs_task<D>(p, i, (uint32_t)g, sc, sn, neg, fl);
#pragma unroll
    for (int w = 0; w < 4; w++) {
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

// Row-chunk kernel (3D, R a multiple of 128): the kernel is bound by integer instruction issue, and a third of a
// one-task thread is index arithmetic, address formation and parameter loads. Here a thread owns T consecutive
// sites of the active colour along x for one replica group: the row geometry is computed once, task j's addresses
// are the chunk base plus compile-time offsets (W4C = words-per-site/4 known at compile time), and the x+1
// neighbour of task j is kept in registers as the x-1 neighbour of task j+1 (six loads per task instead of seven).
// Tasks are the same (site, group) tasks as in the generic kernel: results are identical bit for bit.
template <int T, int W4C, int MINB>
__global__ void __launch_bounds__(64, MINB) k_checkerboard_sparse_row(const __grid_constant__ cbs_params p, int colour)
{
    constexpr int D = 3;
    const int rt = blockIdx.x * blockDim.x + threadIdx.x;      // (chunk, group) within the row
    if (rt >= p.tpr) return;
    const int y = blockIdx.y * blockDim.y + threadIdx.y, z = blockIdx.z;
    int c, g;
    if (p.Gshift >= 0) { c = rt >> p.Gshift; g = rt & (p.G - 1); }
    else { c = rt / p.G; g = rt - c * p.G; }
    const int L = p.L;
    const int W4 = W4C ? W4C : (p.W >> 2);
    const int x0 = 2 * c * T + ((y + z + colour) & 1);
    const uint32_t i0 = (uint32_t)L * (uint32_t)(y + L * z) + x0;
    // neighbour displacements in sites
    const int dyp = y + 1 == L ? -(L - 1) * L : L, dym = y == 0 ? (L - 1) * L : -L;
    const int dzp = z + 1 == L ? -(L - 1) * L * L : L * L, dzm = z == 0 ? (L - 1) * L * L : -L * L;
    const int dxm0 = x0 == 0 ? L - 1 : -1;                       // x-1 of the first task
    const int dxpl = x0 + 2 * T - 1 == L ? -(L - 1) : 1;         // x+1 of the last task, relative to that task's site
    uint4 *P = reinterpret_cast<uint4 *>(p.spins) + ((size_t)i0 * W4 + g);
    const uint4 *Pyp = P + (ptrdiff_t)dyp * W4, *Pym = P + (ptrdiff_t)dym * W4;
    const uint4 *Pzp = P + (ptrdiff_t)dzp * W4, *Pzm = P + (ptrdiff_t)dzm * W4;
    const uint4 *JM = p.jmask + 2 * (size_t)i0;
    uint4 xm = P[(ptrdiff_t)dxm0 * W4];
#pragma unroll
    for (int j = 0; j < T; j++) {
        const int o = 2 * j * W4;
        const uint4 ce = P[o];
        const uint4 xp = j == T - 1 ? P[o + (ptrdiff_t)dxpl * W4] : P[o + W4];
        const uint4 yp = Pyp[o], ym = Pym[o], zp = Pzp[o], zm = Pzm[o];
        const uint4 ja = JM[4 * j];
        const uint2 jb = reinterpret_cast<const uint2 *>(JM)[8 * j + 2];
        const uint32_t neg[6] = { ja.x, ja.y, ja.z, ja.w, jb.x, jb.y };
        uint32_t sc[4] = { ce.x, ce.y, ce.z, ce.w };
        const uint32_t sn[4][6] = { { xp.x, xm.x, yp.x, ym.x, zp.x, zm.x }, { xp.y, xm.y, yp.y, ym.y, zp.y, zm.y },
                                    { xp.z, xm.z, yp.z, ym.z, zp.z, zm.z }, { xp.w, xm.w, yp.w, ym.w, zp.w, zm.w } };
        uint32_t fl[4];
        cbs_task<D>(p, i0 + 2 * j, (uint32_t)g, sc, sn, neg, fl);
        P[o] = make_uint4(sc[0] ^ fl[0], sc[1] ^ fl[1], sc[2] ^ fl[2], sc[3] ^ fl[3]);
        if (p.flips) (reinterpret_cast<uint4 *>(p.flips) + ((size_t)i0 * W4 + g))[o] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
        xm = xp;
        asm volatile("" ::: "memory");   // keep the tasks of a thread in order (registers, not latency, are the constraint)
    }
}

template <int T, int W4C>
static void launch_row(const cbs_params &p, int colour, cudaStream_t stream)
{
    // 64-thread blocks: bx threads along the row (chunks x groups), by rows
    int bx = p.tpr >= 64 ? 64 : p.tpr, by = 1;
    if (p.tpr < 64) { by = 64 / p.tpr; while (by > 1 && p.L % by) by--; }
    dim3 block(bx, by), grid(div_up(p.tpr, bx), p.L / by, p.L);
    if (p.variant & 4) k_checkerboard_sparse_row<T, W4C, 16><<<grid, block, 0, stream>>>(p, colour);
    else k_checkerboard_sparse_row<T, W4C, 12><<<grid, block, 0, stream>>>(p, colour);
}

rrrmc_status_t launch_checkerboard_sparse(rrrmc_ctx *ctx, cbs_params &p, int D, int colour)
{
    const bool full = (p.W % 4) == 0;
    // row-chunk kernel: 3D, whole 128-replica groups, T | L/2. It wins when the rare paths are really rare (its
    // unrolled body is four tasks long: at warmer temperatures the divergent regions start missing the instruction
    // cache), so by default it is used when the class-1 count is zero for > 80 % of the words (32·p1 < 0.22).
    // RRRMC_CB_VARIANT (tuning/tests): bits 0-1 force T = 2, 8, 4; bit 2: 16 blocks/SM; bit 3: force the generic kernel.
    int T = 0;
    if (D == 3 && full && !(p.variant & 8)) {
        const int sel = p.variant & 3;
        const int want = sel == 1 ? 2 : (sel == 2 ? 8 : 4);
        if (sel != 0 || p.tbl[0] > 0xCCCCCCCCu)
            for (int t = want; t >= 2; t >>= 1) if (p.Lh % t == 0) { T = t; break; }
    }
    if (T) {
        p.tpr = (p.Lh / T) * p.G;
        const bool w8 = p.W == 32;
        if (T == 8) { if (w8) launch_row<8, 8>(p, colour, ctx->stream); else launch_row<8, 0>(p, colour, ctx->stream); }
        else if (T == 4) { if (w8) launch_row<4, 8>(p, colour, ctx->stream); else launch_row<4, 0>(p, colour, ctx->stream); }
        else { if (w8) launch_row<2, 8>(p, colour, ctx->stream); else launch_row<2, 0>(p, colour, ctx->stream); }
    } else {
        dim3 block(256), grid(div_up((int64_t)p.Lh * p.G, 256), D >= 2 ? p.L : 1, D >= 3 ? p.L : 1);
#define LAUNCH(DD, FF, MB) k_checkerboard_sparse<DD, FF, MB><<<grid, block, 0, ctx->stream>>>(p, colour)
        if (D == 1) { if (full) LAUNCH(1, true, 1); else LAUNCH(1, false, 1); }
        else if (D == 2) { if (full) LAUNCH(2, true, 1); else LAUNCH(2, false, 1); }
        else if (D == 3) { if (full) LAUNCH(3, true, 4); else LAUNCH(3, false, 1); }
        else { rrrmc_set_error("checkerboard: D=%d unsupported (1..3)", D); return RRRMC_ERR_UNSUPPORTED; }
#undef LAUNCH
    }
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

// ------------------------------------------------------------------------------------------------
// vertical (bit-sliced) counters: cnt[b] holds bit b of a per-lane counter
// ------------------------------------------------------------------------------------------------
template <int B>
__device__ __forceinline__ void vc_add(uint32_t (&cnt)[B], uint32_t plane, int from)
{
#pragma unroll
    for (int b = 0; b < B; b++) {
        if (b < from) continue;
        const uint32_t t = cnt[b] & plane;
        cnt[b] ^= plane;
        plane = t;
    }
}
template <int B>
__device__ __forceinline__ int vc_lane(const uint32_t (&cnt)[B], int lane)
{
    int v = 0;
#pragma unroll
    for (int b = 0; b < B; b++) v |= (int)((cnt[b] >> lane) & 1u) << b;
    return v;
}

// energy(X, C) (EA.jl:195-222) for every replica: E = -Σ_<xy> J σσ = 2·#unsat − D·N (±J).
// Thread = (chunk of ENERGY_S sites, word); counts unsatisfied *forward* bonds per lane. The per-replica counters are few
// (32·W) and every thread adds to 32 of them, so the adds go to a shared-memory copy first (a block walks over many chunks)
// and each block flushes its copy once: with global atomics alone the kernel was bound by same-address atomic throughput
// (4.2 M atomics on 1024 counters at L = 64, R = 1024: 0.40 ms).
constexpr int ENERGY_S = 64;
constexpr int ENERGY_SMEM_W = 64;           // words (32 replicas each) whose counters fit the shared copy
template <int D>
__global__ void __launch_bounds__(128) k_energy_pm1(const uint32_t *__restrict__ spins, const uint8_t *__restrict__ jcode,
                                                    int L, int64_t N, int W, int64_t nthreads, int *__restrict__ unsat_out)
{
    __shared__ int sacc[ENERGY_SMEM_W * 32];
    const bool use_s = W <= ENERGY_SMEM_W;
    if (use_s) {
        for (int k = threadIdx.x; k < W * 32; k += blockDim.x) sacc[k] = 0;
        __syncthreads();
    }
    for (int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; tid < nthreads; tid += (int64_t)gridDim.x * blockDim.x) {
        const int w = (int)(tid % W);
        const int64_t chunk = tid / W;
        if (chunk * ENERGY_S >= N) continue;
        uint32_t cnt[9];
#pragma unroll
        for (int b = 0; b < 9; b++) cnt[b] = 0;
        const int64_t i1 = min(N, (chunk + 1) * ENERGY_S);
        for (int64_t i = chunk * ENERGY_S; i < i1; i++) {
            const int x = (int)(i % L), y = D >= 2 ? (int)((i / L) % L) : 0, z = D >= 3 ? (int)(i / ((int64_t)L * L)) : 0;
            const site_geom<D> sg = make_geom<D>(L, x, y, z);
            const uint32_t sc = spins[i * W + w], jc = jcode[i];
            uint32_t b[D];
#pragma unroll
            for (int d = 0; d < D; d++) b[d] = sc ^ spins[sg.nb[2 * d] * W + w] ^ (0u - ((jc >> (2 * d)) & 1u));
            if (D == 1) vc_add<9>(cnt, b[0], 0);
            if (D == 2) { vc_add<9>(cnt, b[0] ^ b[1], 0); vc_add<9>(cnt, b[0] & b[1], 1); }
            if (D == 3) { uint32_t s, c; full_add(b[0], b[1], b[2], s, c); vc_add<9>(cnt, s, 0); vc_add<9>(cnt, c, 1); }
        }
#pragma unroll 1
        for (int lane = 0; lane < 32; lane++) {
            const int v = vc_lane<9>(cnt, lane);
            if (v) atomicAdd(use_s ? &sacc[32 * w + lane] : &unsat_out[32 * w + lane], v);
        }
    }
    if (use_s) {
        __syncthreads();
        for (int k = threadIdx.x; k < W * 32; k += blockDim.x) { const int v = sacc[k]; if (v) atomicAdd(&unsat_out[k], v); }
    }
}

rrrmc_status_t launch_energy_pm1(rrrmc_state *s, int *d_unsat)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_CUDA(cudaMemsetAsync(d_unsat, 0, sizeof(int) * s->W * 32, ctx->stream));
    const int64_t nthreads = (int64_t)div_up(g->N, ENERGY_S) * s->W;
    const int64_t cap = (int64_t)(ctx->sm_count > 0 ? ctx->sm_count : 148) * 8;    // a block walks over many chunks
    dim3 block(128), grid((unsigned)std::min<int64_t>(div_up(nthreads, 128), cap));
    if (g->D == 1) k_energy_pm1<1><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, g->N, (int)s->W, nthreads, d_unsat);
    else if (g->D == 2) k_energy_pm1<2><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, g->N, (int)s->W, nthreads, d_unsat);
    else k_energy_pm1<3><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, g->N, (int)s->W, nthreads, d_unsat);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

// per-lane popcount over sites of a mask array [N][W] (accepted-move counters, magnetisation)
constexpr int COUNT_S = 128;
__global__ void __launch_bounds__(128) k_count_lanes(const uint32_t *__restrict__ masks, int64_t N, int W,
                                                     long long *__restrict__ out)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int w = (int)(tid % W);
    const int64_t chunk = tid / W;
    if (chunk * COUNT_S >= N) return;
    uint32_t cnt[8];
#pragma unroll
    for (int b = 0; b < 8; b++) cnt[b] = 0;
    const int64_t i1 = min(N, (chunk + 1) * COUNT_S);
    for (int64_t i = chunk * COUNT_S; i < i1; i++) vc_add<8>(cnt, masks[i * W + w], 0);
#pragma unroll 1
    for (int lane = 0; lane < 32; lane++) {
        const int v = vc_lane<8>(cnt, lane);
        if (v) atomicAdd((unsigned long long *)&out[32 * w + lane], (unsigned long long)v);
    }
}

rrrmc_status_t launch_count_lanes(rrrmc_ctx *ctx, const uint32_t *masks, int64_t N, int W, long long *d_out)
{
    const int64_t nthreads = (int64_t)div_up(N, COUNT_S) * W;
    k_count_lanes<<<div_up(nthreads, 128), 128, 0, ctx->stream>>>(masks, N, W, d_out);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

// delta_energy(X, C, i) for all replicas of one site (EA.jl:266-275): out[32w+b] = 4(D-u)
template <int D>
__global__ void k_delta_energy_site(const uint32_t *__restrict__ spins, const uint8_t *__restrict__ jcode,
                                    int L, int W, int64_t site, int *__restrict__ out)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const int x = (int)(site % L), y = D >= 2 ? (int)((site / L) % L) : 0, z = D >= 3 ? (int)(site / ((int64_t)L * L)) : 0;
    const site_geom<D> sg = make_geom<D>(L, x, y, z);
    uint32_t sn[2 * D];
#pragma unroll
    for (int k = 0; k < 2 * D; k++) sn[k] = spins[sg.nb[k] * W + w];
    uint32_t u0, u1, u2;
    unsat_planes<D>(spins[site * W + w], sn, jcode[site], u0, u1, u2);
    for (int b = 0; b < 32; b++) {
        const int u = (int)((u0 >> b) & 1) | (int)((u1 >> b) & 1) << 1 | (int)((u2 >> b) & 1) << 2;
        out[32 * w + b] = 4 * (D - u);
    }
}

// delta_energy(X, C, i), i=1..N, for one replica
template <int D>
__global__ void k_delta_energy_replica(const uint32_t *__restrict__ spins, const uint8_t *__restrict__ jcode,
                                       int L, int64_t N, int W, int64_t replica, int *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int w = (int)(replica >> 5), b = (int)(replica & 31);
    const int x = (int)(i % L), y = D >= 2 ? (int)((i / L) % L) : 0, z = D >= 3 ? (int)(i / ((int64_t)L * L)) : 0;
    const site_geom<D> sg = make_geom<D>(L, x, y, z);
    const uint32_t jc = jcode[i];
    const int sc = (spins[i * W + w] >> b) & 1;
    int u = 0;
#pragma unroll
    for (int k = 0; k < 2 * D; k++) u += sc ^ (int)((spins[sg.nb[k] * W + w] >> b) & 1) ^ (int)((jc >> k) & 1);
    out[i] = 4 * (D - u);
}

rrrmc_status_t launch_delta_energy_site(rrrmc_state *s, int64_t site0, int *d_out)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    dim3 block(128), grid(div_up(s->W, 128));
    if (g->D == 1) k_delta_energy_site<1><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, (int)s->W, site0, d_out);
    else if (g->D == 2) k_delta_energy_site<2><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, (int)s->W, site0, d_out);
    else k_delta_energy_site<3><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, (int)s->W, site0, d_out);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}
rrrmc_status_t launch_delta_energy_replica(rrrmc_state *s, int64_t replica, int *d_out)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    dim3 block(256), grid(div_up(g->N, 256));
    if (g->D == 1) k_delta_energy_replica<1><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, g->N, (int)s->W, replica, d_out);
    else if (g->D == 2) k_delta_energy_replica<2><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, g->N, (int)s->W, replica, d_out);
    else k_delta_energy_replica<3><<<grid, block, 0, ctx->stream>>>(s->d_spins, g->d_jcode, g->L, g->N, (int)s->W, replica, d_out);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

// spinflip!(C, i) on selected replicas (multispin layout has no local-field cache to update)
__global__ void k_flip_site(uint32_t *spins, int W, int64_t site, const uint32_t *mask)
{
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < W) spins[site * W + w] ^= mask ? mask[w] : 0xffffffffu;
}
rrrmc_status_t launch_flip_site(rrrmc_state *s, int64_t site0, const uint32_t *d_mask)
{
    rrrmc_ctx *ctx = s->g->ctx;
    k_flip_site<<<div_up(s->W, 128), 128, 0, ctx->stream>>>(s->d_spins, (int)s->W, site0, d_mask);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

// Config(N) random init (Interface.jl:24-28): word (i,w) = Philox(ctr=(i_lo,i_hi,w,'CNFG'), key=seed).x
__global__ void k_randomize(uint32_t *spins, int64_t N, int W, uint32_t k0, uint32_t k1)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= N * W) return;
    const int64_t i = tid / W; const int w = (int)(tid % W);
    spins[tid] = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)w, 0x434e4647u, k0, k1).x;
}
rrrmc_status_t launch_randomize(rrrmc_state *s, uint64_t seed)
{
    rrrmc_ctx *ctx = s->g->ctx;
    const int64_t n = s->g->N * s->W;
    k_randomize<<<div_up(n, 256), 256, 0, ctx->stream>>>(s->d_spins, s->g->N, (int)s->W, (uint32_t)seed, (uint32_t)(seed >> 32));
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

// ------------------------------------------------------------------------------------------------
// Layout transposes: reference BitVector chunks [count][nchunks] (site-major bits per chain,
// Interface.jl:21-29) <-> multispin words. Thread = (chunk c, word w): a 32x64 bit-matrix transpose.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_chunks_to_multispin(const uint64_t *__restrict__ chunks, int64_t nchunks, int64_t first,
                                                             int64_t count, uint32_t *__restrict__ spins, int64_t N, int W, int w0, int nw)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nchunks * nw) return;
    const int w = w0 + (int)(tid % nw);
    const int64_t c = tid / nw;
    uint64_t ch[32];
    uint32_t lanes = 0;
#pragma unroll
    for (int b = 0; b < 32; b++) {
        const int64_t r = 32 * (int64_t)w + b - first;
        const bool ok = r >= 0 && r < count;
        ch[b] = ok ? chunks[r * nchunks + c] : 0ull;
        lanes |= ok ? (1u << b) : 0u;
    }
    for (int k = 0; k < 64; k++) {
        const int64_t i = 64 * c + k;
        if (i >= N) break;
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 32; b++) word |= (uint32_t)((ch[b] >> k) & 1ull) << b;
        uint32_t *dst = &spins[i * W + w];
        *dst = lanes == 0xffffffffu ? word : ((*dst & ~lanes) | word);
    }
}
__global__ void __launch_bounds__(128) k_multispin_to_chunks(const uint32_t *__restrict__ spins, int64_t N, int W, int64_t first,
                                                             int64_t count, uint64_t *__restrict__ chunks, int64_t nchunks, int w0, int nw)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= nchunks * nw) return;
    const int w = w0 + (int)(tid % nw);
    const int64_t c = tid / nw;
    uint64_t ch[32];
#pragma unroll
    for (int b = 0; b < 32; b++) ch[b] = 0;
    for (int k = 0; k < 64; k++) {
        const int64_t i = 64 * c + k;
        if (i >= N) break;
        const uint32_t word = spins[i * W + w];
#pragma unroll
        for (int b = 0; b < 32; b++) ch[b] |= (uint64_t)((word >> b) & 1u) << k;
    }
#pragma unroll
    for (int b = 0; b < 32; b++) {
        const int64_t r = 32 * (int64_t)w + b - first;
        if (r >= 0 && r < count) chunks[r * nchunks + c] = ch[b];
    }
}

rrrmc_status_t launch_upload_transpose(rrrmc_state *s, int64_t first, int64_t count)
{
    rrrmc_ctx *ctx = s->g->ctx;
    const int w0 = (int)(first / 32), w1 = (int)((first + count + 31) / 32), nw = w1 - w0;
    const int64_t n = s->nchunks * nw;
    k_chunks_to_multispin<<<div_up(n, 128), 128, 0, ctx->stream>>>(s->d_chunks, s->nchunks, first, count, s->d_spins, s->g->N, (int)s->W, w0, nw);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}
rrrmc_status_t launch_download_transpose(rrrmc_state *s, int64_t first, int64_t count)
{
    rrrmc_ctx *ctx = s->g->ctx;
    const int w0 = (int)(first / 32), w1 = (int)((first + count + 31) / 32), nw = w1 - w0;
    const int64_t n = s->nchunks * nw;
    k_multispin_to_chunks<<<div_up(n, 128), 128, 0, ctx->stream>>>(s->d_spins, s->g->N, (int)s->W, first, count, s->d_chunks, s->nchunks, w0, nw);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

__global__ void k_flush(uint32_t *buf, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = (uint32_t)i;
}
rrrmc_status_t launch_flush(rrrmc_ctx *ctx)
{
    k_flush<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>((uint32_t *)ctx->flush_buf, ctx->flush_bytes / 4);
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}
