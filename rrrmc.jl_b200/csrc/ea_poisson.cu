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

#include "ea_poisson_core.cuh"

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
    const bool slow = cbp_task_hits<D, NW>(p, cbp_env_of(p, p.bucket), i, (uint32_t)g, cbp_draw<NW>(p, cbp_env_of(p, p.bucket), i, (uint32_t)g), m, gg, h);

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
        const bool slow = cbp_task_hits<D, NW>(p, cbp_env_of(p, sbucket), i, (uint32_t)g, cbp_draw<NW>(p, cbp_env_of(p, sbucket), i, (uint32_t)g), m, gg, h);
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
        const cbp_env env = cbp_env_of(p, sbucket);
        const cbp_words<NW> rw0 = cbp_draw<NW>(p, env, i, (uint32_t)g0), rw1 = cbp_draw<NW>(p, env, i, (uint32_t)(g0 + Gh));
        slow[0] = cbp_task_hits<D, NW>(p, env, i, (uint32_t)g0, rw0, m[0], gg[0], h[0]);
        slow[1] = cbp_task_hits<D, NW>(p, env, i, (uint32_t)(g0 + Gh), rw1, m[1], gg[1], h[1]);
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
