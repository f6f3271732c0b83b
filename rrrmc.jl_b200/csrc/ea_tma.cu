// Checkerboard Metropolis half-sweep for GraphEA ±J, "poisson" acceptance procedure, TMA-staged bricks.
//
// Same procedure, same Philox counters and therefore the same trajectories, bit for bit, as ea_poisson.cu (the CPU
// restatement is oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_poisson); what changes is how a task gets its seven
// spin words. ea_poisson.cu has every thread form seven 64-bit addresses and issue seven cp.async per task — a third
// of its instructions, on the ALU pipe that bounds the kernel. Here the lattice is a rank-5 tensor
// [z][y][x][slab][32 words] (a slab = 1024 replicas = one 128-byte row per site) and a dedicated producer warp
// brings a whole 8x4x4 brick plus its halo into shared memory with a handful of cp.async.bulk.tensor operations per
// brick; 256 consumer threads then read their words with LDS.128 at addresses that are compile-time offsets from
// one per-thread constant (y and z neighbours) or two per-thread constants (x neighbours), and write the updated
// centre word straight to global memory. No per-task address arithmetic is left beside one IMAD.WIDE for the store.
//
// Shared-memory stage (43 008 B, two stages per block, two blocks per SM):
//   planes   6 planes (z' = z+1 = 0..5) x 48 rows (x = 0..7, y' = y+1 = 0..5) x 128 B. Planes 1..4 are the brick with its
//            y halo: one box (8 x, 6 y) per plane, or two/three boxes when the halo wraps around the lattice (the
//            split is along the slowest box dimension, so the pieces land where the single box would have).
//            Planes 0 and 5 are the z halo: one box (8 x, 4 y) each, at rows y' = 1..4.
//   x faces  2 x 16 rows (y + 4 z): the sites at x0-1 and x0+8 (periodic), one box (1 x, 4 y, 4 z) each
//   bonds    64 x 32 B: the six whole-word sign masks of the brick's 64 active sites, brick-ordered copy of jmask
//            (one 1-D bulk copy)
// A quarter warp (8 lanes) reads the 8 groups of ONE site = one 128-byte row, so every LDS.128 is conflict free
// without swizzling. The producer/consumer handshake is the usual full/empty mbarrier pair per stage; consumers
// generate the (spin-independent) hit masks of a brick's two tasks BEFORE waiting for its data.
#include <cuda.h>
#include <mutex>
#include <vector>
#include "kernels.cuh"
#include "ea_poisson_core.cuh"
#include "ea_tma.cuh"

namespace {

constexpr int BX = CBT_BX, BY = CBT_BY, BZ = CBT_BZ;
constexpr int ROWB = 128;                              // bytes of one site's slab row
constexpr int PLANE_ROWS = BX * (BY + 2);              // 48
constexpr int PLANE_BYTES = PLANE_ROWS * ROWB;         // 6144
constexpr int XF_OFF = PLANE_BYTES * (BZ + 2);         // 36864
constexpr int XF_BYTES = BY * BZ * ROWB;               // 2048 per face
constexpr int JM_OFF = XF_OFF + 2 * XF_BYTES;          // 40960
constexpr int NACT = CBT_NACT;                         // 64 active sites per brick
constexpr int STAGE_BYTES = JM_OFF + NACT * 32;        // 43008
constexpr int TX_BYTES = BZ * PLANE_BYTES + 2 * BX * BY * ROWB + 2 * XF_BYTES + NACT * 32; // bytes landing per brick
constexpr int NSTAGE = 2;
constexpr int NCONS = 256;                             // consumer threads (8 warps); warp 8 is the producer
constexpr int SM_BUCKET = NSTAGE * STAGE_BYTES;
constexpr int SM_BARS = SM_BUCKET + CBP_BUCKETS * 8;
constexpr int SMEM_BYTES = SM_BARS + 64;
static_assert(NACT == 2 * (NCONS / 8), "two tasks per consumer thread and brick");
static_assert(STAGE_BYTES % 128 == 0 && XF_OFF % 128 == 0 && JM_OFF % 128 == 0, "TMA destinations are 128-byte aligned");

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// The consumers' "stage is free" arrive. It must not be issued before the thread's ld.shared of the stage have actually
// RETURNED: ptxas places SYNCS.ARRIVE right behind the (still outstanding) LDS, and once the last thread has arrived the
// issuer refills the stage through the async proxy — measured on B200 (scripts/stress_flow.py): one warp of a block's first
// brick now and then computed its second site from the NEXT brick's bytes. So the barrier address is made to depend on a
// value computed from every word the thread loaded (dep·1 − dep = 0, `one` being a launch parameter opaque to ptxas): the
// arrive waits on the loads' scoreboards like any consumer of their data.
__device__ __forceinline__ void mbar_arrive_after(uint32_t bar, uint32_t dep, uint32_t one)
{
    const uint32_t z = dep * one - dep;
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar + z) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n"
                 "W_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra W_%=;\n\t}" :: "r"(bar), "r"(parity) : "memory");
}
// the helper lanes' wait: backs off between polls, so that a lane that waits for most of a brick's time does not
// compete with the consumer warps of its scheduler for issue slots (measured: 58 polls per brick without it)
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity)
{
    // try_wait with a suspend-time hint: the lane is parked by the hardware until the phase completes (or the hint
    // expires) instead of polling (a __nanosleep(128) loop still issued an instruction every ~10 cycles)
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(1000000u) : "memory");
        if (ok) break;
    }
}
__device__ __forceinline__ void tma_load5(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1, int c2, int c3, int c4)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 :: "r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template <int OFF> __device__ __forceinline__ uint4 lds128(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr), "n"(OFF));
    return v;
}
template <int OFF> __device__ __forceinline__ uint2 lds64(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(OFF));
    return v;
}

// brick b' of this launch -> slab and lattice origin (exact float reciprocals for b' < 2^22, as in ea_poisson.cu)
struct brick_pos { int slab, x0, y0, z0, b; };
__device__ __forceinline__ brick_pos locate_brick(const cbt_params &P, int bb)
{
    brick_pos r;
    r.slab = __float2int_rz(((float)bb + 0.5f) * P.inv_nbricks);
    r.b = bb - r.slab * P.nbricks;
    const int q = __float2int_rz(((float)r.b + 0.5f) * P.inv_nbx);
    const int X = r.b - q * P.nbx;
    const int Z = __float2int_rz(((float)q + 0.5f) * P.inv_nby), Y = q - Z * P.nby;
    r.x0 = X * BX; r.y0 = Y * BY; r.z0 = Z * BZ;
    return r;
}

template <int NW, int MINB>
__global__ void __launch_bounds__(NCONS + 32, MINB) k_checkerboard_tma(const __grid_constant__ cbt_params P, int colour)
{
    constexpr int D = 3;
    extern __shared__ __align__(1024) uint8_t smem[];
    const cbp_params &p = P.p;
    const int t = threadIdx.x, L = p.L;
    const uint32_t sm0 = smem_u32(smem);
    const uint32_t bars = sm0 + SM_BARS;                 // full[s] at bars + 8 s, empty[s] at bars + 16 + 8 s
    uint2 *sbucket = reinterpret_cast<uint2 *>(smem + SM_BUCKET);
    for (int k = t; k < CBP_BUCKETS; k += NCONS + 32) sbucket[k] = __ldg(p.bucket + k);
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 16 + 8 * s, NCONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int total = P.nbricks * P.nslab;

    if (t >= NCONS) {
        // ---------------- producer: one lane issues the bulk copies of every brick this block owns ----------------
        if (t != NCONS) return;
        asm volatile("griddepcontrol.wait;" ::: "memory");   // the previous half-sweep is complete and visible
        int k = 0;
        for (int bb = blockIdx.x; bb < total; bb += gridDim.x, k++) {
            const int s = k & 1;
            const uint32_t sb = sm0 + s * STAGE_BYTES, full = bars + 8 * s, empty = bars + 16 + 8 * s;
            if (k >= NSTAGE) mbar_wait(empty, ((k >> 1) - 1) & 1);
            const brick_pos r = locate_brick(P, bb);
            mbar_expect_tx(full, TX_BYTES);
            const bool wrap_lo = r.y0 == 0, wrap_hi = r.y0 + BY == L;
#pragma unroll
            for (int zz = 0; zz < BZ; zz++) {
                const uint32_t dst = sb + (zz + 1) * PLANE_BYTES;
                const int z = r.z0 + zz;
                if (!wrap_lo && !wrap_hi) tma_load5(dst, &P.m_y6, full, 0, r.slab, r.x0, r.y0 - 1, z);
                else {
                    // rows y' = 0 | 1..4 | 5; a wrapped halo row is its own box, the rest stays one box
                    if (wrap_lo) tma_load5(dst, &P.m_y1, full, 0, r.slab, r.x0, L - 1, z);
                    if (wrap_hi) tma_load5(dst + 5 * BX * ROWB, &P.m_y1, full, 0, r.slab, r.x0, 0, z);
                    if (wrap_lo && wrap_hi) tma_load5(dst + BX * ROWB, &P.m_y4, full, 0, r.slab, r.x0, r.y0, z);
                    else if (wrap_lo) tma_load5(dst + BX * ROWB, &P.m_y5, full, 0, r.slab, r.x0, r.y0, z);
                    else tma_load5(dst, &P.m_y5, full, 0, r.slab, r.x0, r.y0 - 1, z);
                }
            }
            tma_load5(sb + BX * ROWB, &P.m_y4, full, 0, r.slab, r.x0, r.y0, r.z0 == 0 ? L - 1 : r.z0 - 1);
            tma_load5(sb + (BZ + 1) * PLANE_BYTES + BX * ROWB, &P.m_y4, full, 0, r.slab, r.x0, r.y0, r.z0 + BZ == L ? 0 : r.z0 + BZ);
            tma_load5(sb + XF_OFF, &P.m_xf, full, 0, r.slab, r.x0 == 0 ? L - 1 : r.x0 - 1, r.y0, r.z0);
            tma_load5(sb + XF_OFF + XF_BYTES, &P.m_xf, full, 0, r.slab, r.x0 + BX == L ? 0 : r.x0 + BX, r.y0, r.z0);
            bulk_load(sb + JM_OFF, P.jbrick + ((size_t)colour * P.nbricks + r.b) * (NACT * 2), NACT * 32, full);
        }
        return;
    }

    // ---------------- consumers: thread = (active-site slot s and s + 32, group g8) ----------------
    const int g8 = t & 7, sl = t >> 3;
    const int xh = sl & 3, y = (sl >> 2) & 3, z = sl >> 4;               // z in {0, 1}; the second task sits at z + 2
    const int x = 2 * xh + ((y + z + colour) & 1);                        // brick origins are even in y and z
    const uint32_t offc = (uint32_t)(x + BX * ((y + 1) + (BY + 2) * (z + 1))) * ROWB + g8 * 16;
    const uint32_t offxf = XF_OFF + (uint32_t)(y + BY * z) * ROWB + g8 * 16;
    constexpr int DZ2 = 2 * PLANE_BYTES, DXF2 = 2 * BY * ROWB;            // the same offsets for the second task
    const uint32_t xm0 = x > 0 ? offc - ROWB : offxf, xm1 = x > 0 ? offc - ROWB + DZ2 : offxf + DXF2;
    const uint32_t xp0 = x < BX - 1 ? offc + ROWB : offxf + XF_BYTES, xp1 = x < BX - 1 ? offc + ROWB + DZ2 : offxf + XF_BYTES + DXF2;
    const uint32_t jmo = JM_OFF + (uint32_t)sl * 32;
    const uint32_t soff = (uint32_t)x + (uint32_t)L * ((uint32_t)y + (uint32_t)L * (uint32_t)z), LL2 = 2u * (uint32_t)L * (uint32_t)L;
    const uint32_t W4 = (uint32_t)p.W >> 2;
    uint4 *const spins4 = reinterpret_cast<uint4 *>(p.spins);
    uint4 *const flips4 = reinterpret_cast<uint4 *>(p.flips);

    // {first site of the brick, slab} comes from a table built once per state: decoding the brick index costs ~35
    // instructions per brick and thread; the entry of the NEXT brick is requested one iteration ahead
    uint2 org = __ldg(P.origin + min((int)blockIdx.x, total - 1));
    int k = 0;
    for (int bb = blockIdx.x; bb < total; bb += gridDim.x, k++) {
        const int s = k & 1;
        const uint32_t sb = sm0 + s * STAGE_BYTES, full = bars + 8 * s, empty = bars + 16 + 8 * s;
        const uint32_t i0 = org.x + soff, i1 = i0 + LL2;
        const uint32_t grp = org.y * 8u + (uint32_t)g8;
        org = __ldg(P.origin + min(bb + (int)gridDim.x, total - 1));
        uint32_t m[2][4], gg[2][4], h[2][4];
        bool slow[2], wslow[2];      // wslow: warp-uniform "some lane of the warp left the fast path" (only then can h be set)
        // both Philox chains in one basic block: their rounds interleave
        const cbp_env env = cbp_env_of(p, sbucket);
        const cbp_words<NW> rw0 = cbp_draw<NW>(p, env, i0, grp), rw1 = cbp_draw<NW>(p, env, i1, grp);
        slow[0] = cbp_task_hits<D, NW>(p, env, i0, grp, rw0, m[0], gg[0], h[0], &wslow[0]);
        slow[1] = cbp_task_hits<D, NW>(p, env, i1, grp, rw1, m[1], gg[1], h[1], &wslow[1]);
        if (k == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
        mbar_wait(full, (k >> 1) & 1);
#pragma unroll
        for (int j = 0; j < 2; j++) {
            // a0 = the z-1 neighbour's word (the lowest address this task reads in the planes): offsets are >= 0
            const uint32_t a0 = sb + offc - PLANE_BYTES, axm = sb + (j ? xm1 : xm0), axp = sb + (j ? xp1 : xp0), aj = sb + jmo;
            uint4 c, v[6], ja; uint2 jb;
            if (j == 0) {
                c = lds128<PLANE_BYTES>(a0);
                v[0] = lds128<0>(axp); v[1] = lds128<0>(axm);
                v[2] = lds128<PLANE_BYTES + BX * ROWB>(a0); v[3] = lds128<PLANE_BYTES - BX * ROWB>(a0);
                v[4] = lds128<2 * PLANE_BYTES>(a0); v[5] = lds128<0>(a0);
                ja = lds128<0>(aj); jb = lds64<16>(aj);
            } else {
                c = lds128<DZ2 + PLANE_BYTES>(a0);
                v[0] = lds128<0>(axp); v[1] = lds128<0>(axm);
                v[2] = lds128<DZ2 + PLANE_BYTES + BX * ROWB>(a0); v[3] = lds128<DZ2 + PLANE_BYTES - BX * ROWB>(a0);
                v[4] = lds128<DZ2 + 2 * PLANE_BYTES>(a0); v[5] = lds128<DZ2>(a0);
                ja = lds128<32 * 32>(aj); jb = lds64<32 * 32 + 16>(aj);
            }
            const uint32_t neg[6] = { ja.x, ja.y, ja.z, ja.w, jb.x, jb.y };
            uint32_t sc[4] = { c.x, c.y, c.z, c.w }, bp[4][2 * D], fl[4];
#pragma unroll
            for (int q = 0; q < 2 * D; q++) {
                bp[0][q] = lop3p<P_XOR3>(sc[0], v[q].x, neg[q]); bp[1][q] = lop3p<P_XOR3>(sc[1], v[q].y, neg[q]);
                bp[2][q] = lop3p<P_XOR3>(sc[2], v[q].z, neg[q]); bp[3][q] = lop3p<P_XOR3>(sc[3], v[q].w, neg[q]);
            }
#pragma unroll
            for (int w = 0; w < 4; w++) fl[w] = cbp_flip_planes<D>(bp[w], m[j][w], gg[j][w]);
            if (j == 1) mbar_arrive_after(empty, fl[0], p.one);   // this thread has read everything it needs from the stage (fl[0] depends on all of it)
            if (wslow[j]) {                              // a uniform branch, taken by ~13 % of the warps at β = 1
                asm volatile("" ::: "memory");           // (kept a branch: predicating it would cost every task 8 instructions)
                if (slow[j]) {
#pragma unroll
                    for (int w = 0; w < 4; w++) fl[w] |= h[j][w];   // a level-3 hit flips its lane whatever the bonds say
                }
            }
#pragma unroll
            for (int w = 0; w < 4; w++) sc[w] ^= fl[w];
            const uint32_t idx = (j ? i1 : i0) * W4 + grp;
            spins4[idx] = make_uint4(sc[0], sc[1], sc[2], sc[3]);
            if (flips4) flips4[idx] = make_uint4(fl[0], fl[1], fl[2], fl[3]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-sweep kernel: every half-sweep of a run in ONE launch.
//
// The per-colour launches above cost a launch boundary per half-sweep: blocks drain, the next grid's blocks run their
// prologue and wait for their first brick (measured: 1.5 of 16.3 µs per half-sweep at L = 64, R = 1024). Here the grid
// is persistent and co-resident (cooperative launch), block b owns bricks b, b + grid, ... of EVERY half-sweep, and a
// brick of half-sweep h is loaded as soon as the seven bricks it reads (itself and its six face neighbours) have
// completed half-sweep h - 1: done[brick] counts the half-sweeps completed on a brick since the state was created.
// The same condition orders the writes: a neighbour overwrites the halo a brick reads only in half-sweep h + 1, which
// it starts after this brick has published h. No deadlock: the oldest unfinished half-sweep always has a brick whose
// neighbours are complete, blocks walk their bricks in order, and every block is resident.
//
// Roles of a block: warps 0-7 consumers (as above); warp 8 lane 0 issuer (waits for the gate and for a free stage, then
// fence.proxy.async and the TMA loads); warp 9 lane 0 publisher (waits until the 256 consumers have stored a brick —
// mbarrier, release.cta/acquire.cta — then st.release.gpu of the brick's counter; cumulativity makes the consumers'
// stores visible before the counter); warp 10 lane 0 gatekeeper (polls the counters a brick depends on with relaxed
// loads, one acquire fence, then opens the gate in shared memory). Three lanes because each of the three costs an L2
// round trip or a gpu-scope fence per brick (measured together: as long as a brick's compute), and because a block
// that waits for a neighbour must still publish — the neighbour may in turn wait for THIS block's counter. Results are bit-identical to the per-colour launches: same
// Philox counters (site, group, sweep), same procedure.
constexpr int NFLOW = NCONS + 128;      // + one warpgroup of helpers: issuer, publisher and gatekeeper warps (the fourth idles)
constexpr int FLOW_REGS_HELPER = 32, FLOW_REGS_CONSUMER = 104;   // setmaxnreg: 2 blocks x 384 threads start with 80 each
constexpr int SMF_BARS = SM_BUCKET + CBP_BUCKETS * 8;      // full[2] | empty[2] | stored[4] | published
constexpr int SMEMF_BYTES = SMF_BARS + 128;

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *a)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *a)
{
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t *a, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_volatile(uint32_t a)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_volatile(uint32_t a, uint32_t v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory");
}

template <int NW, int MINB, bool PERGROUP, bool FLIPS>
__global__ void __launch_bounds__(NFLOW, MINB) k_checkerboard_flow(const __grid_constant__ cbf_params F)
{
    constexpr int D = 3;
    extern __shared__ __align__(1024) uint8_t smem[];
    const cbt_params &P = F.T;
    const cbp_params &p = P.p;
    const int t = threadIdx.x, L = p.L;
    const uint32_t sm0 = smem_u32(smem);
    const uint32_t bars = sm0 + SMF_BARS;                // full[s] +8s, empty[s] +16+8s, stored[i] +32+8i, published +64
    const uint32_t pub = bars + 64, gate = bars + 72;
    uint2 *sbucket = reinterpret_cast<uint2 *>(smem + SM_BUCKET);
    if (!PERGROUP) for (int k = t; k < CBP_BUCKETS; k += NFLOW) sbucket[k] = __ldg(p.bucket + k);
    if (t == 0) {
#pragma unroll
        for (int s = 0; s < NSTAGE; s++) { mbar_init(bars + 8 * s, 1); mbar_init(bars + 16 + 8 * s, NCONS); }
#pragma unroll
        for (int i = 0; i < 4; i++) mbar_init(bars + 32 + 8 * i, NCONS);
        sts_volatile(pub, 0u);
        sts_volatile(gate, 0u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int total = P.nbricks * P.nslab, grid = (int)gridDim.x;
    const int nk = (total - (int)blockIdx.x + grid - 1) / grid;          // bricks of this block per half-sweep (>= 1)
    const uint32_t nhalf = F.nhalf;

    if (t >= NCONS) {
        // the helpers need few registers: hand the rest of the warpgroup's share to the consumers
        if (MINB == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(FLOW_REGS_HELPER));
        if (t & 31) return;
        if (t == NCONS) {
            // ---------------- issuer: waits for the gate and the stage, then sends the brick's bulk copies ----------------
            uint32_t q = 0;
            const uint4 *ent = reinterpret_cast<const uint4 *>(F.bricks + blockIdx.x);
            uint4 e0 = __ldg(ent);
            for (uint32_t hs = 0; hs < nhalf; hs++) {
                const int colour = (int)((F.half0 + hs) & 1ull);
                for (int k = 0; k < nk; k++, q++) {
                    const int s = (int)(q & 1u);
                    const uint32_t sb = sm0 + s * STAGE_BYTES, full = bars + 8 * s, empty = bars + 16 + 8 * s;
                    if (hs > 0u) {
                        while ((int32_t)(lds_volatile(gate) - (q + 1u)) < 0) __nanosleep(500);
                        asm volatile("fence.acq_rel.cta;" ::: "memory");          // the gatekeeper's acquire, handed on
                        asm volatile("fence.proxy.async.global;" ::: "memory");   // the neighbours' generic-proxy stores, then TMA reads
                    }
                    if (q >= 2u) mbar_wait_sleep(empty, ((q >> 1) - 1u) & 1u);
                    if (q >= 3u) while ((int32_t)(lds_volatile(pub) - (q - 2u)) < 0) __nanosleep(500);   // the publisher is at most 3 bricks behind
                    const int slab = (int)e0.y, b = (int)e0.z;
                    const int x0 = (int)(e0.w & 1023u), y0 = (int)((e0.w >> 10) & 1023u), z0 = (int)(e0.w >> 20);
                    {
                        const int kn = k + 1 < nk ? k + 1 : 0;
                        e0 = __ldg(reinterpret_cast<const uint4 *>(F.bricks + ((int)blockIdx.x + kn * grid)));
                    }
                    mbar_expect_tx(full, TX_BYTES);
                    const bool wrap_lo = y0 == 0, wrap_hi = y0 + BY == L;
                    if (!wrap_lo && !wrap_hi) tma_load5(sb + PLANE_BYTES, &F.m_y6z, full, 0, slab, x0, y0 - 1, z0);   // four planes, one box
                    else {
#pragma unroll
                        for (int zz = 0; zz < BZ; zz++) {
                            const uint32_t dst = sb + (zz + 1) * PLANE_BYTES;
                            const int z = z0 + zz;
                            if (wrap_lo) tma_load5(dst, &P.m_y1, full, 0, slab, x0, L - 1, z);
                            if (wrap_hi) tma_load5(dst + 5 * BX * ROWB, &P.m_y1, full, 0, slab, x0, 0, z);
                            if (wrap_lo && wrap_hi) tma_load5(dst + BX * ROWB, &P.m_y4, full, 0, slab, x0, y0, z);
                            else if (wrap_lo) tma_load5(dst + BX * ROWB, &P.m_y5, full, 0, slab, x0, y0, z);
                            else tma_load5(dst, &P.m_y5, full, 0, slab, x0, y0 - 1, z);
                        }
                    }
                    tma_load5(sb + BX * ROWB, &P.m_y4, full, 0, slab, x0, y0, z0 == 0 ? L - 1 : z0 - 1);
                    tma_load5(sb + (BZ + 1) * PLANE_BYTES + BX * ROWB, &P.m_y4, full, 0, slab, x0, y0, z0 + BZ == L ? 0 : z0 + BZ);
                    tma_load5(sb + XF_OFF, &P.m_xf, full, 0, slab, x0 == 0 ? L - 1 : x0 - 1, y0, z0);
                    tma_load5(sb + XF_OFF + XF_BYTES, &P.m_xf, full, 0, slab, x0 + BX == L ? 0 : x0 + BX, y0, z0);
                    bulk_load(sb + JM_OFF, P.jbrick + ((size_t)colour * P.nbricks + b) * (NACT * 2), NACT * 32, full);
                }
            }
        } else if (t == NCONS + 32) {
            // ---------------- publisher ----------------
            uint32_t q = 0;
            for (uint32_t hs = 0; hs < nhalf; hs++)
                for (int k = 0; k < nk; k++, q++) {
                    mbar_wait_sleep(bars + 32 + 8 * (q & 3u), (q >> 2) & 1u);
                    st_release_gpu(F.done + ((int)blockIdx.x + k * grid), F.epoch0 + hs + 1u);
                    sts_volatile(pub, q + 1u);
                }
        } else if (t == NCONS + 64) {
            // ---------------- gatekeeper: opens the gate for bricks whose seven counters have arrived ----------------
            // Up to two bricks per round: their 14 relaxed loads are in flight together and ONE acquire fence covers
            // the bricks that passed (a poll is an L2 round trip, ~1600 cycles under load, the fence ~1400: per brick
            // they would take most of a brick's compute time; the counters are normally satisfied several bricks ahead).
            const uint32_t nq = nhalf * (uint32_t)nk;
            uint32_t gq = (uint32_t)nk, ghs = 1u; int gk = 0;
            while (gq < nq) {
                uint32_t v[2][7], need[2]; bool valid[2];
                {
                    uint32_t hs = ghs; int k = gk;
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        valid[j] = gq + (uint32_t)j < nq;
                        need[j] = F.epoch0 + hs;
                        const uint4 *ent = reinterpret_cast<const uint4 *>(F.bricks + ((int)blockIdx.x + k * grid));
                        const uint4 e1 = __ldg(ent + 1), e2 = __ldg(ent + 2);
                        const uint32_t ids[7] = { e1.x, e1.y, e1.z, e1.w, e2.x, e2.y, e2.z };
#pragma unroll
                        for (int i = 0; i < 7; i++) v[j][i] = valid[j] ? ld_relaxed_gpu(F.done + ids[i]) : need[j];
                        if (++k == nk) { k = 0; hs++; }
                    }
                }
                int n = 0;
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    int32_t worst = 0;
#pragma unroll
                    for (int i = 0; i < 7; i++) worst = min(worst, (int32_t)(v[j][i] - need[j]));
                    if (valid[j] && worst >= 0 && n == j) n = j + 1;
                }
                if (n == 0) { __nanosleep(1000); continue; }     // at the frontier: do not compete with the consumers for issue slots
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                gq += (uint32_t)n;
                for (int j = 0; j < n; j++) if (++gk == nk) { gk = 0; ghs++; }
                sts_volatile(gate, gq);
            }
        }
        return;
    }

    // ---------------- consumers: thread = (active-site slot s and s + 32, group g8) ----------------
    if (MINB == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(FLOW_REGS_CONSUMER));
    const int g8 = t & 7, sl = t >> 3;
    const int xh = sl & 3, y = (sl >> 2) & 3, z = sl >> 4;
    const uint32_t offxf = XF_OFF + (uint32_t)(y + BY * z) * ROWB + g8 * 16;
    constexpr int DZ2 = 2 * PLANE_BYTES, DXF2 = 2 * BY * ROWB;
    const uint32_t jmo = JM_OFF + (uint32_t)sl * 32;
    const uint32_t LL2 = 2u * (uint32_t)L * (uint32_t)L;
    const uint32_t W4 = (uint32_t)p.W >> 2;
    uint4 *const spins4 = reinterpret_cast<uint4 *>(p.spins);
    uint4 *const flips4 = reinterpret_cast<uint4 *>(p.flips);

    cbp_env env = cbp_env_of(p, sbucket);
    uint32_t cur_slab = 0xffffffffu;
    uint2 org = __ldg(reinterpret_cast<const uint2 *>(F.bricks + blockIdx.x));
    uint32_t q = 0;
    for (uint32_t hs = 0; hs < nhalf; hs++) {
        const uint64_t half = F.half0 + hs;
        const int colour = (int)(half & 1ull);
        env.t_lo = (uint32_t)(half >> 1); env.t_hi16 = (uint32_t)(half >> 33) << 16;
        const int x = 2 * xh + ((y + z + colour) & 1);                        // brick origins are even in y and z
        const uint32_t offc = (uint32_t)(x + BX * ((y + 1) + (BY + 2) * (z + 1))) * ROWB + g8 * 16;
        const uint32_t xm0 = x > 0 ? offc - ROWB : offxf, xm1 = x > 0 ? offc - ROWB + DZ2 : offxf + DXF2;
        const uint32_t xp0 = x < BX - 1 ? offc + ROWB : offxf + XF_BYTES, xp1 = x < BX - 1 ? offc + ROWB + DZ2 : offxf + XF_BYTES + DXF2;
        const uint32_t soff = (uint32_t)x + (uint32_t)L * ((uint32_t)y + (uint32_t)L * (uint32_t)z);
        for (int k = 0; k < nk; k++, q++) {
            const int s = (int)(q & 1u);
            const uint32_t sb = sm0 + s * STAGE_BYTES, full = bars + 8 * s, empty = bars + 16 + 8 * s;
            const uint32_t i0 = org.x + soff, i1 = i0 + LL2;
            const uint32_t grp = org.y * 8u + (uint32_t)g8;
            if (PERGROUP && org.y != cur_slab) {
                cur_slab = org.y;
                const cbp_group *G = F.groups + grp;
                env.tb0_0 = __ldg(&G->tb0_0); env.tb0_1 = __ldg(&G->tb0_1); env.tc0 = __ldg(&G->tc0);
                env.tbl = G->tbl; env.bucket = F.gbucket + (size_t)grp * CBP_BUCKETS;
            }
            {   // table entry of the next brick of this block (the first one again after the last of a half-sweep)
                const int kn = k + 1 < nk ? k + 1 : 0;
                org = __ldg(reinterpret_cast<const uint2 *>(F.bricks + ((int)blockIdx.x + kn * grid)));
            }
            uint32_t m[2][4], gg[2][4], h[2][4];
            bool slow[2], wslow[2];
            const cbp_words<NW> rw0 = cbp_draw<NW>(p, env, i0, grp), rw1 = cbp_draw<NW>(p, env, i1, grp);
            slow[0] = cbp_task_hits<D, NW>(p, env, i0, grp, rw0, m[0], gg[0], h[0], &wslow[0]);
            slow[1] = cbp_task_hits<D, NW>(p, env, i1, grp, rw1, m[1], gg[1], h[1], &wslow[1]);
            mbar_wait(full, (q >> 1) & 1u);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const uint32_t a0 = sb + offc - PLANE_BYTES, axm = sb + (j ? xm1 : xm0), axp = sb + (j ? xp1 : xp0), aj = sb + jmo;
                uint4 c, v[6], ja; uint2 jb;
                if (j == 0) {
                    c = lds128<PLANE_BYTES>(a0);
                    v[0] = lds128<0>(axp); v[1] = lds128<0>(axm);
                    v[2] = lds128<PLANE_BYTES + BX * ROWB>(a0); v[3] = lds128<PLANE_BYTES - BX * ROWB>(a0);
                    v[4] = lds128<2 * PLANE_BYTES>(a0); v[5] = lds128<0>(a0);
                    ja = lds128<0>(aj); jb = lds64<16>(aj);
                } else {
                    c = lds128<DZ2 + PLANE_BYTES>(a0);
                    v[0] = lds128<0>(axp); v[1] = lds128<0>(axm);
                    v[2] = lds128<DZ2 + PLANE_BYTES + BX * ROWB>(a0); v[3] = lds128<DZ2 + PLANE_BYTES - BX * ROWB>(a0);
                    v[4] = lds128<DZ2 + 2 * PLANE_BYTES>(a0); v[5] = lds128<DZ2>(a0);
                    ja = lds128<32 * 32>(aj); jb = lds64<32 * 32 + 16>(aj);
                }
                const uint32_t neg[6] = { ja.x, ja.y, ja.z, ja.w, jb.x, jb.y };
                uint32_t sc[4] = { c.x, c.y, c.z, c.w }, bp[4][2 * D], kc[4], tt[4], ns[4];
#pragma unroll
                for (int qq = 0; qq < 2 * D; qq++) {
                    bp[0][qq] = lop3p<P_XOR3>(sc[0], v[qq].x, neg[qq]); bp[1][qq] = lop3p<P_XOR3>(sc[1], v[qq].y, neg[qq]);
                    bp[2][qq] = lop3p<P_XOR3>(sc[2], v[qq].z, neg[qq]); bp[3][qq] = lop3p<P_XOR3>(sc[3], v[qq].w, neg[qq]);
                }
                // flip = kc | tt (cbp_flip_parts); the new word is formed straight from the two halves: sc ^ (kc | tt)
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    cbp_flip_parts(bp[w], m[j][w], gg[j][w], kc[w], tt[w]);
                    ns[w] = lop3p<0x1E>(sc[w], kc[w], tt[w]);
                }
                if (j == 1) mbar_arrive_after(empty, ns[0], p.one);   // this thread has read everything it needs from the stage (ns[0] depends on all of it)
                // a level-3 hit flips its lane whatever the bonds say. A uniform branch, taken by ~13 % of the warps at
                // β = 1, and kept a BRANCH (an out-of-line call): if-converted, its eight LOP3 would take issue slots in
                // every task, and integer issue is what bounds the kernel
                if (wslow[j]) {
                    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
                    const uint4 f = cbp_merge_level3(make_uint4(kc[0], kc[1], kc[2], kc[3]), make_uint4(tt[0], tt[1], tt[2], tt[3]),
                                                     slow[j] ? make_uint4(h[j][0], h[j][1], h[j][2], h[j][3]) : z4);
                    ns[0] = sc[0] ^ f.x; ns[1] = sc[1] ^ f.y; ns[2] = sc[2] ^ f.z; ns[3] = sc[3] ^ f.w;
                    kc[0] = f.x; kc[1] = f.y; kc[2] = f.z; kc[3] = f.w;      // (the flip mask, for FLIPS)
                }
                const uint32_t idx = (j ? i1 : i0) * W4 + grp;
                spins4[idx] = make_uint4(ns[0], ns[1], ns[2], ns[3]);
                if (FLIPS) flips4[idx] = make_uint4(kc[0] | tt[0], kc[1] | tt[1], kc[2] | tt[2], kc[3] | tt[3]);
            }
            mbar_arrive(bars + 32 + 8 * (q & 3u));           // both words of this thread are stored (release.cta)
        }
    }
}

// brick-ordered copy of the bond masks: jbrick[colour][brick][slot][8], slot = xh + 4 (y + 4 z)
__global__ void k_build_jbrick(const uint4 *__restrict__ jmask, uint4 *__restrict__ jbrick, int L, int nbx, int nby, int nbricks)
{
    const int id = blockIdx.x * blockDim.x + threadIdx.x;       // (colour, brick, slot)
    if (id >= 2 * nbricks * NACT) return;
    const int a = id % NACT, b = (id / NACT) % nbricks, colour = id / (NACT * nbricks);
    const int X = b % nbx, Y = (b / nbx) % nby, Z = b / (nbx * nby);
    const int xh = a & 3, y = (a >> 2) & 3, z = a >> 4;
    const int x = 2 * xh + ((y + z + colour) & 1);
    const size_t i = (size_t)(X * BX + x) + (size_t)L * ((size_t)(Y * BY + y) + (size_t)L * (size_t)(Z * BZ + z));
    jbrick[2 * (size_t)id] = jmask[2 * i];
    jbrick[2 * (size_t)id + 1] = jmask[2 * i + 1];
}

typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

rrrmc_status_t encode_map(encode_fn enc, CUtensorMap *m, void *base, int L, int W, int bx, int by, int bz)
{
    const cuuint64_t dims[5] = { 32, (cuuint64_t)(W / 32), (cuuint64_t)L, (cuuint64_t)L, (cuuint64_t)L };
    const cuuint64_t strides[4] = { 128, (cuuint64_t)W * 4, (cuuint64_t)W * 4 * L, (cuuint64_t)W * 4 * L * L };
    const cuuint32_t box[5] = { 32, 1, (cuuint32_t)bx, (cuuint32_t)by, (cuuint32_t)bz };
    const cuuint32_t es[5] = { 1, 1, 1, 1, 1 };
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { rrrmc_set_error("cuTensorMapEncodeTiled failed (%d) for box %dx%dx%d", (int)r, bx, by, bz); return RRRMC_ERR_CUDA; }
    return RRRMC_OK;
}

template <int NW, int MINB>
cudaError_t launch_one(const cbt_params &P, int colour, int sm_count, cudaStream_t st)
{
    static int configured = 0;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_checkerboard_tma<NW, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_checkerboard_tma<NW, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_checkerboard_tma<NW, MINB>, NCONS + 32, SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = occ < 1 ? 1 : occ;
    }
    const int total = P.nbricks * P.nslab;
    int grid = sm_count * configured;
    if (P.p.variant & 128) grid = 2;     // tests: many bricks per block
    if (grid > total) grid = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NCONS + 32); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (P.p.variant & 32) ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, k_checkerboard_tma<NW, MINB>, P, colour);
}
template <int MINB>
cudaError_t launch_nw(const cbt_params &P, int colour, int sm_count, cudaStream_t st)
{
    switch (P.p.NW) {
    case 1: return launch_one<1, MINB>(P, colour, sm_count, st);
    case 2: return launch_one<2, MINB>(P, colour, sm_count, st);
    case 4: return launch_one<4, MINB>(P, colour, sm_count, st);
    default: return launch_one<6, MINB>(P, colour, sm_count, st);
    }
}

} // namespace

bool checkerboard_tma_eligible(const rrrmc_state *s)
{
    const rrrmc_graph *g = s->g;
    if (!(g->kind == RRRMC_EA_PM1 && g->D == 3 && g->d_jmask && g->bipartite)) return false;
    if (s->W % 32 != 0 || g->L % BX != 0 || g->L % BY != 0 || g->L % BZ != 0) return false;
    const int64_t nbricks = (int64_t)(g->L / BX) * (g->L / BY) * (g->L / BZ);
    return nbricks * (s->W / 32) < ((int64_t)1 << 22) && g->N * s->W < ((int64_t)1 << 31);
}

void checkerboard_tma_free(rrrmc_state *s)
{
    if (s->tma) {
        cudaFree(s->tma->d_jbrick); cudaFree(s->tma->d_origin); cudaFree(s->tma->d_bricks); cudaFree(s->tma->d_done);
        cudaFree(s->tma->d_groups); cudaFree(s->tma->d_gbucket);
        delete s->tma; s->tma = nullptr;
    }
}

// Fills the TMA half of the launch parameters: tensor maps over the state's spin array (encoded once per state) and
// the brick-ordered bond masks (built once per state from the graph's jmask).
rrrmc_status_t checkerboard_tma_prepare(rrrmc_state *s, const cbp_params &p, cbt_params &P)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    const int L = g->L, W = (int)s->W;
    if (!s->tma) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        RR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
        if (!fn || qr != cudaDriverEntryPointSuccess) { rrrmc_set_error("cuTensorMapEncodeTiled is not available in this driver"); return RRRMC_ERR_CUDA; }
        cb_tma_store *c = new cb_tma_store();
        encode_fn enc = reinterpret_cast<encode_fn>(fn);
        rrrmc_status_t st = RRRMC_OK;
        if ((st = encode_map(enc, &c->m_y6, s->d_spins, L, W, BX, BY + 2, 1)) != RRRMC_OK ||
            (st = encode_map(enc, &c->m_y5, s->d_spins, L, W, BX, BY + 1, 1)) != RRRMC_OK ||
            (st = encode_map(enc, &c->m_y4, s->d_spins, L, W, BX, BY, 1)) != RRRMC_OK ||
            (st = encode_map(enc, &c->m_y1, s->d_spins, L, W, BX, 1, 1)) != RRRMC_OK ||
            (st = encode_map(enc, &c->m_xf, s->d_spins, L, W, 1, BY, BZ)) != RRRMC_OK ||
            (st = encode_map(enc, &c->m_y6z, s->d_spins, L, W, BX, BY + 2, BZ)) != RRRMC_OK) { delete c; return st; }
        c->nbx = L / BX; c->nby = L / BY; c->nbricks = c->nbx * c->nby * (L / BZ);
        const size_t n = (size_t)2 * c->nbricks * NACT;
        if (cudaMalloc(&c->d_jbrick, n * 32) != cudaSuccess) { delete c; rrrmc_set_error("cudaMalloc of the brick-ordered bond masks failed"); return RRRMC_ERR_CUDA; }
        k_build_jbrick<<<div_up((int64_t)n, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4 *>(g->d_jmask), c->d_jbrick, L, c->nbx, c->nby, c->nbricks);
        ctx->launches++;
        if (cudaGetLastError() != cudaSuccess) { cudaFree(c->d_jbrick); delete c; rrrmc_set_error("k_build_jbrick launch failed"); return RRRMC_ERR_CUDA; }
        // first site and slab of every brick of a launch, in launch order
        std::vector<uint2> org((size_t)c->nbricks * (W / 32));
        for (int slab = 0; slab < W / 32; slab++)
            for (int b = 0; b < c->nbricks; b++) {
                const int X = b % c->nbx, Y = (b / c->nbx) % c->nby, Z = b / (c->nbx * c->nby);
                org[(size_t)slab * c->nbricks + b] = make_uint2((uint32_t)(X * BX) + (uint32_t)L * ((uint32_t)(Y * BY) + (uint32_t)L * (uint32_t)(Z * BZ)), (uint32_t)slab);
            }
        // the multi-sweep kernel's table: position and face neighbours of every brick, and its progress counter.
        // Its launch order FOLDS the z layers (0, nbz-1, 1, nbz-2, ...): block b owns ranks b, b + grid, ... and all
        // blocks move through their lists at about the same pace, so a brick's neighbours should sit at about the
        // same rank — in natural order the periodic wrap would make the first bricks of a half-sweep wait for the
        // last bricks of the previous one. Folded, neighbouring bricks are at most three layers apart in rank.
        const int nbz = L / BZ, layer = c->nbx * c->nby;
        std::vector<cbf_brick> bt(org.size());
        auto rank = [&](int slab, int xx, int yy, int zz) {
            xx = (xx + c->nbx) % c->nbx; yy = (yy + c->nby) % c->nby; zz = (zz + nbz) % nbz;
            const int zr = zz < (nbz + 1) / 2 ? 2 * zz : 2 * (nbz - 1 - zz) + 1;
            return (uint32_t)(slab * c->nbricks + zr * layer + xx + c->nbx * yy);
        };
        for (int slab = 0; slab < W / 32; slab++)
            for (int b = 0; b < c->nbricks; b++) {
                const int X = b % c->nbx, Y = (b / c->nbx) % c->nby, Z = b / layer;
                cbf_brick &e = bt[rank(slab, X, Y, Z)];
                e.site0 = org[(size_t)slab * c->nbricks + b].x; e.slab = (uint32_t)slab; e.b = (uint32_t)b;
                e.xyz = (uint32_t)(X * BX) | (uint32_t)(Y * BY) << 10 | (uint32_t)(Z * BZ) << 20;
                e.nbr[0] = rank(slab, X + 1, Y, Z); e.nbr[1] = rank(slab, X - 1, Y, Z);
                e.nbr[2] = rank(slab, X, Y + 1, Z); e.nbr[3] = rank(slab, X, Y - 1, Z);
                e.nbr[4] = rank(slab, X, Y, Z + 1); e.nbr[5] = rank(slab, X, Y, Z - 1);
                e.nbr[6] = e.nbr[7] = rank(slab, X, Y, Z);
            }
        if (cudaMalloc(&c->d_origin, org.size() * sizeof(uint2)) != cudaSuccess ||
            cudaMalloc(&c->d_bricks, bt.size() * sizeof(cbf_brick)) != cudaSuccess ||
            cudaMalloc(&c->d_done, bt.size() * sizeof(uint32_t)) != cudaSuccess ||
            cudaMemcpyAsync(c->d_origin, org.data(), org.size() * sizeof(uint2), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemcpyAsync(c->d_bricks, bt.data(), bt.size() * sizeof(cbf_brick), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaMemsetAsync(c->d_done, 0, bt.size() * sizeof(uint32_t), ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
            cudaFree(c->d_jbrick); cudaFree(c->d_origin); cudaFree(c->d_bricks); cudaFree(c->d_done); delete c;
            rrrmc_set_error("upload of the brick table failed"); return RRRMC_ERR_CUDA;
        }
        c->epoch = 0;
        s->tma = c;
    }
    const cb_tma_store *c = s->tma;
    P.p = p;
    P.m_y6 = c->m_y6; P.m_y5 = c->m_y5; P.m_y4 = c->m_y4; P.m_y1 = c->m_y1; P.m_xf = c->m_xf;
    P.jbrick = c->d_jbrick; P.origin = c->d_origin;
    P.nbx = c->nbx; P.nby = c->nby; P.nbricks = c->nbricks; P.nslab = W / 32;
    P.inv_nbx = 1.0f / (float)c->nbx; P.inv_nby = 1.0f / (float)c->nby; P.inv_nbricks = 1.0f / (float)c->nbricks;
    return RRRMC_OK;
}

rrrmc_status_t launch_checkerboard_tma(rrrmc_ctx *ctx, cbt_params &P, int colour)
{
    const int mb = P.p.variant & 3;      // RRRMC_CB_VARIANT (tuning): resident blocks per SM
    cudaError_t e = mb == 1 ? launch_nw<1>(P, colour, ctx->sm_count, ctx->stream) : launch_nw<2>(P, colour, ctx->sm_count, ctx->stream);
    ctx->launches++;
    RR_CUDA(e);
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

namespace {

template <int NW, int MINB, bool PERGROUP, bool FLIPS>
cudaError_t flow_one(const cbf_params &F, int sm_count, cudaStream_t st)
{
    static int configured = 0;
    auto kern = k_checkerboard_flow<NW, MINB, PERGROUP, FLIPS>;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEMF_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NFLOW, SMEMF_BYTES);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        configured = occ;
    }
    const int total = F.T.nbricks * F.T.nslab;
    int grid = sm_count * configured;            // every block must be resident: the launch is cooperative
    if (F.T.p.variant & 128) grid = 2;           // tests: many bricks per block
    if (grid > total) grid = total;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NFLOW); cfg.dynamicSmemBytes = SMEMF_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, F);
}
template <int MINB, bool PERGROUP, bool FLIPS>
cudaError_t flow_nw(const cbf_params &F, int sm_count, cudaStream_t st)
{
    switch (F.T.p.NW) {
    case 1: return flow_one<1, MINB, PERGROUP, FLIPS>(F, sm_count, st);
    case 2: return flow_one<2, MINB, PERGROUP, FLIPS>(F, sm_count, st);
    case 4: return flow_one<4, MINB, PERGROUP, FLIPS>(F, sm_count, st);
    default: return flow_one<6, MINB, PERGROUP, FLIPS>(F, sm_count, st);
    }
}

} // namespace

rrrmc_status_t launch_checkerboard_flow(rrrmc_state *s, cbt_params &P, uint64_t sweep0, int64_t nsweeps,
                                        const cbp_group *groups, const uint2 *gbucket, int ngroups)
{
    rrrmc_ctx *ctx = s->g->ctx;
    cb_tma_store *c = s->tma;
    if (!c || !c->d_bricks) { rrrmc_set_error("launch_checkerboard_flow: the state has no brick table"); return RRRMC_ERR_STATE; }
    if (groups) {
        if (ngroups != (int)(s->W / 4)) { rrrmc_set_error("β ladder tables: expected %d groups, given %d", (int)(s->W / 4), ngroups); return RRRMC_ERR_ARG; }
        if (c->ngroups_alloc < ngroups) {
            if (!gbucket) { rrrmc_set_error("launch_checkerboard_flow: no device copy of the ladder tables"); return RRRMC_ERR_STATE; }
            cudaFree(c->d_groups); cudaFree(c->d_gbucket); c->d_groups = nullptr; c->d_gbucket = nullptr; c->ngroups_alloc = 0;
            RR_CUDA(cudaMalloc(&c->d_groups, sizeof(cbp_group) * ngroups));
            RR_CUDA(cudaMalloc(&c->d_gbucket, sizeof(uint2) * CBP_BUCKETS * ngroups));
            c->ngroups_alloc = ngroups;
        }
        if (gbucket) {
            RR_CUDA(cudaMemcpyAsync(c->d_groups, groups, sizeof(cbp_group) * ngroups, cudaMemcpyHostToDevice, ctx->stream));
            RR_CUDA(cudaMemcpyAsync(c->d_gbucket, gbucket, sizeof(uint2) * CBP_BUCKETS * ngroups, cudaMemcpyHostToDevice, ctx->stream));
            RR_CUDA(cudaStreamSynchronize(ctx->stream));   // caller buffers
            c->ladder_key.resize((size_t)ngroups * CBP_LEN);
            for (int k = 0; k < ngroups; k++) memcpy(c->ladder_key.data() + (size_t)k * CBP_LEN, groups[k].tbl, sizeof(uint32_t) * CBP_LEN);
        }
    }
    // A persistent grid whose blocks wait for each other must never share the device with a second one of its kind: two
    // half-resident grids (launched from two streams / contexts of this process) would wait for blocks that cannot become
    // resident. Launches are therefore chained device-wide: each waits for the previous one's completion event.
    static std::mutex flow_mu;
    static cudaEvent_t flow_done[64] = {};
    std::lock_guard<std::mutex> flow_lock(flow_mu);
    const int dev = ctx->device & 63;
    if (!flow_done[dev]) RR_CUDA(cudaEventCreateWithFlags(&flow_done[dev], cudaEventDisableTiming));
    else RR_CUDA(cudaStreamWaitEvent(ctx->stream, flow_done[dev], 0));
    cbf_params F;
    memset(&F, 0, sizeof F);
    F.T = P; F.m_y6z = c->m_y6z; F.bricks = c->d_bricks; F.done = c->d_done;
    F.groups = groups ? c->d_groups : nullptr; F.gbucket = groups ? c->d_gbucket : nullptr;
    const bool flips = P.p.flips != nullptr;
    const int mb = P.p.variant & 3;
    int64_t left = nsweeps;
    uint64_t sw = sweep0;
    while (left > 0) {
        const int64_t n = left < (1 << 22) ? left : (1 << 22);      // keeps the counters far from wrapping inside one launch
        F.epoch0 = c->epoch; F.nhalf = (uint32_t)(2 * n); F.half0 = 2 * sw;
        cudaError_t e;
        if (groups) e = flips ? flow_nw<2, true, true>(F, ctx->sm_count, ctx->stream) : flow_nw<2, true, false>(F, ctx->sm_count, ctx->stream);
        else if (mb == 1) e = flips ? flow_nw<1, false, true>(F, ctx->sm_count, ctx->stream) : flow_nw<1, false, false>(F, ctx->sm_count, ctx->stream);
        else e = flips ? flow_nw<2, false, true>(F, ctx->sm_count, ctx->stream) : flow_nw<2, false, false>(F, ctx->sm_count, ctx->stream);
        ctx->launches++;
        if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) {
            // the counters are only meaningful after a complete launch: start over
            cudaMemsetAsync(c->d_done, 0, sizeof(uint32_t) * (size_t)c->nbricks * (s->W / 32), ctx->stream); c->epoch = 0;
            if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported || e == cudaErrorLaunchOutOfResources) {
                // the grid cannot be made co-resident here (a shared or partitioned device): the caller falls back to
                // the per-colour launches of the same kernel family — another GPU path, same results
                c->flow_unavailable = true;
                rrrmc_set_error("the multi-sweep kernel needs a co-resident grid: %s", cudaGetErrorString(e));
                return RRRMC_ERR_UNSUPPORTED;
            }
            RR_CUDA(e);
            return RRRMC_ERR_CUDA;
        }
        c->epoch += F.nhalf;
        left -= n; sw += (uint64_t)n;
    }
    RR_CUDA(cudaEventRecord(flow_done[dev], ctx->stream));
    return RRRMC_OK;
}
