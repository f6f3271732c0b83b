// Dense kernels for GraphSKNormal on a replica batch (BASELINE config 4: N=4096 Gaussian couplings × 512 replicas).
//
//  * Local-field initialisation — the one dense contraction of the path, energy(X, C) of SK.jl:212-237:
//        lfields[r][i] = 2 σ_ri Σ_j J_ij σ_rj
//    as H = J·Sᵀ on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM). The spins are
//    exactly ±1, so the product is made *exact* instead of approximated in reduced-precision floating point: J is
//    quantised once to 40-bit fixed point and split into five signed 8-bit digit planes, each plane is an INT8 GEMM
//    with INT32 accumulation (no rounding, no order dependence), and the epilogue recombines the five accumulators
//    in int64. The only error is the 2^-P quantisation of J itself (|ΔH| <= N·2^-(P+1), ~1e-9 at N=4096), far inside
//    the 1e-6 relative tolerance of the contract.
//  * Lock-step Metropolis sweeps (new engine; the reference's per-flip update_cache! SK.jl:239-276 is the axpy):
//    all replicas attempt site i = 1..N in order; an accepted flip streams row i of J once per CTA and updates the
//    local fields of the CTA's replicas held in shared memory.
#include <cstdlib>
#include <algorithm>
#include <cmath>
#include <cstring>
#include "common.cuh"
#include "kernels.cuh"
#include "chain.cuh"
#include "philox.cuh"

constexpr int SKQ_SLICES = 5;          // 8-bit digit planes of the fixed-point couplings
constexpr int TC_M = 128, TC_N = 64, TC_KB = 128; // CTA tile: 128 replicas x 64 sites (x 5 digit planes), 128 bytes of K per stage

// ------------------------------------------------------------------------------------------------
// spins as int8 ±1, [R][Npad] (B operand of the GEMM; K-major)
// ------------------------------------------------------------------------------------------------
__global__ void k_spins_to_s8(const uint64_t *__restrict__ chunks, int64_t nchunks, int64_t R, int N, int Npad, int8_t *__restrict__ out)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= R * Npad) return;
    const int64_t r = tid / Npad; const int j = (int)(tid % Npad);
    out[tid] = j < N ? (int8_t)(2 * (int)((chunks[r * nchunks + (j >> 6)] >> (j & 63)) & 1ull) - 1) : (int8_t)0;
}

// ------------------------------------------------------------------------------------------------
// tcgen05 helpers (PTX ISA: tcgen05.alloc / mma / commit / ld; descriptors as in CUTLASS cute/arch/mma_sm100_desc.hpp)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// K-major operand tile with 128-byte swizzle: rows of 128 bytes, 8-row atoms of 1024 bytes (SBO), descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);          // start address
    d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset between 8-row atoms
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// kind::i8, S8 x S8 -> S32, both operands K-major, M x N tile
__host__ __device__ constexpr uint32_t umma_idesc_s8(int M, int N)
{
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(mbar), "r"(parity) : "memory");
}
// 16 bytes of row `row` at 16-byte chunk `c16` of a swizzled 128-byte-row tile
__device__ __forceinline__ uint32_t sw128_off(int row, int c16) { return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c16 ^ (row & 7)) << 4)); }

struct sk_tc_params {
    const int8_t *Jq;       // [SLICES][Npad][Npad] digit planes
    const int8_t *S8;       // [Rpad][Npad]
    double *lf;             // [R][N] out: local fields
    int N, Npad; int64_t R, Rpad;
    double scale;           // 2^-P
};

// Tiling: the M side of the MMA (128 TMEM lanes) is the REPLICAS, the N side 64 sites of one digit plane, five
// accumulators of 64 columns (TMEM's 512 columns allow no more). The kernel is bound by the operand bytes that cross
// L2 -> SM per MAC, and this orientation needs 128 + 5·64 = 448 bytes per K byte for 5·128·64 MACs where the other one
// (128 sites x 64 replicas) needs 704; each J tile is read by R/128 CTAs instead of R/64.
// Pipeline: a stage holds one 128-byte K chunk of the spin tile (16 KiB) and of the five digit-plane tiles (5 x 8 KiB);
// three stages in dynamic shared memory. All threads fill stage kc+2 with cp.async (16-byte chunks straight into the
// 128-byte-swizzled layout, four full 128-byte lines per warp instruction) AFTER the MMAs of chunk kc have been queued,
// so the tensor pipe always has the next chunk's instructions behind the running ones; a per-stage mbarrier armed by
// tcgen05.commit tells the producers when the MMAs have drained a stage.
constexpr int TC_STAGE_BYTES = (TC_M + SKQ_SLICES * TC_N) * TC_KB;   // 56 KiB
constexpr int TC_STAGES = 3;
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(g) : "memory");
}
__global__ void __launch_bounds__(128, 1) k_sk_fields_tc(sk_tc_params P)
{
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    __shared__ __align__(8) uint64_t mbar[TC_STAGES];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i0 = blockIdx.x * TC_N;                // first site of the tile
    const int64_t r0 = (int64_t)blockIdx.y * TC_M;   // first replica
    constexpr uint32_t TMEM_COLS = 512;              // 5 accumulators x 64 columns, rounded up to a power of two
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&tmem_base_s)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < TC_STAGES; st++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar[st])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = umma_idesc_s8(TC_M, TC_N);
    const int nkc = P.Npad / TC_KB;
    const int c16 = lane & 7, rsub = lane >> 3;      // this lane's 16-byte chunk and row offset inside a 4-row group

    auto fill = [&](int kc) {                        // all threads: stage kc % 3 <- K chunk kc
        uint8_t *stg = base + (size_t)(kc % TC_STAGES) * TC_STAGE_BYTES;
        const int8_t *A = P.S8 + (size_t)kc * TC_KB + c16 * 16;
        const uint32_t sA = smem_u32(stg);
#pragma unroll
        for (int j = 0; j < 8; j++) {                // 128 replica rows
            const int row = warp * 32 + j * 4 + rsub;
            cp_async16(sA + sw128_off(row, c16), A + (size_t)(r0 + row) * P.Npad);
        }
#pragma unroll
        for (int sl = 0; sl < SKQ_SLICES; sl++) {    // 64 site rows of each digit plane
            const int8_t *B = P.Jq + (size_t)sl * P.Npad * P.Npad + (size_t)kc * TC_KB + c16 * 16;
            const uint32_t sB = smem_u32(stg + (size_t)TC_M * TC_KB + (size_t)sl * TC_N * TC_KB);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int row = warp * 16 + j * 4 + rsub;
                cp_async16(sB + sw128_off(row, c16), B + (size_t)(i0 + row) * P.Npad);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto mma = [&](int kc) {                         // thread 0: 5 planes x 4 K-steps on stage kc % 3, then arm its mbarrier
        const uint32_t stg = smem_u32(base + (size_t)(kc % TC_STAGES) * TC_STAGE_BYTES);
        const uint64_t descA = umma_desc_k_sw128(stg);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int sl = 0; sl < SKQ_SLICES; sl++) {
            const uint64_t descB = umma_desc_k_sw128(stg + TC_M * TC_KB + sl * TC_N * TC_KB);
#pragma unroll
            for (int k = 0; k < TC_KB / 32; k++) {   // K = 32 bytes per instruction for 8-bit operands
                const uint32_t accumulate = (kc | k) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                    :: "r"(tmem_base + (uint32_t)(sl * TC_N)), "l"(descA + (uint64_t)(k * 32 >> 4)), "l"(descB + (uint64_t)(k * 32 >> 4)),
                       "r"(idesc), "r"(accumulate) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&mbar[kc % TC_STAGES])) : "memory");
    };

    fill(0);
    if (nkc > 1) fill(1);
    for (int kc = 0; kc < nkc; kc++) {
        if (kc + 1 < nkc) asm volatile("cp.async.wait_group 1;" ::: "memory");   // chunk kc has landed (this thread's part)
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) mma(kc);
        if (kc + 2 < nkc) {
            // stage (kc+2)%3 was last read by the MMAs of chunk kc-1 (running or done; chunk kc is queued behind them):
            // their commit is completion number (kc-1)/3 of that stage's barrier
            if (kc >= 1) mbar_wait(smem_u32(&mbar[(kc + 2) % TC_STAGES]), (uint32_t)(((kc - 1) / TC_STAGES) & 1));
            fill(kc + 2);
        }
    }
    // all MMAs done: the last commit of the last stage's barrier (commits complete in order)
    mbar_wait(smem_u32(&mbar[(nkc - 1) % TC_STAGES]), (uint32_t)(((nkc - 1) / TC_STAGES) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // epilogue: thread = TMEM lane = replica r0+tid; columns = sites. Recombine the five digit accumulators exactly.
    const int64_t r = r0 + tid;
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < TC_N; c0 += 16) {
        long long acc[16];
#pragma unroll
        for (int n = 0; n < 16; n++) acc[n] = 0;
#pragma unroll
        for (int s = SKQ_SLICES - 1; s >= 0; s--) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(lane_base + (uint32_t)(s * TC_N + c0)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int n = 0; n < 16; n++) acc[n] = acc[n] * 256 + (long long)(int32_t)v[n];
        }
        if (r < P.R) {
            const int4 sv = *reinterpret_cast<const int4 *>(P.S8 + (size_t)r * P.Npad + i0 + c0);   // the replica's 16 spins (±1, 0 in the padding)
            const int8_t *sb = reinterpret_cast<const int8_t *>(&sv);
#pragma unroll
            for (int n = 0; n < 16; n++) {
                const int i = i0 + c0 + n;
                if (i < P.N) {
                    const double H = (double)acc[n] * P.scale;           // Σ_j J_ij σ_rj
                    P.lf[(size_t)r * P.N + i] = 2.0 * (double)sb[n] * H;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "n"(TMEM_COLS) : "memory");
}

// CUDA-core path in the reference's summation order (bit-identical to energy() of SK.jl:218-231)
__global__ void k_sk_fields_ordered(const double *__restrict__ J, const uint64_t *__restrict__ chunks, int64_t nchunks, int64_t R, int N, double *__restrict__ lf)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= R * N) return;
    const int64_t r = tid / N; const int i = (int)(tid % N);
    const uint64_t *s = chunks + r * nchunks;
    const int si = (int)((s[i >> 6] >> (i & 63)) & 1ull);
    double a = 0.0;
    for (int j = 0; j < N; j++) {
        const int sj = (int)((s[j >> 6] >> (j & 63)) & 1ull);
        a = __dadd_rn(a, __dmul_rn((double)(1 - 2 * (si ^ sj)), J[(size_t)j * N + i]));
    }
    lf[tid] = 2 * a;
}
// E_r = -½ Σ_i lf_i / 2 summed in site order (SK.jl:232-236). A warp owns 32 replicas: it reads a 32-site tile of each
// with one coalesced 256-byte load, stages the tile in shared memory, and lane l then adds replica l's 32 values in site
// order — the reference's sequential sum, bit for bit, at streaming bandwidth (a thread walking its own row reads 8 bytes
// per 32-byte sector: 364 µs for 16.8 MB).
__global__ void __launch_bounds__(128) k_sk_energy_from_fields(const double *__restrict__ lf, int64_t R, int N, double *__restrict__ E)
{
    __shared__ double tile[4][32][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r0 = ((int64_t)blockIdx.x * 4 + wid) * 32;
    if (r0 >= R) return;
    double n = 0.0;
    for (int i0 = 0; i0 < N; i0 += 32) {
#pragma unroll 4
        for (int rp = 0; rp < 32; rp++)
            tile[wid][rp][lane] = (r0 + rp < R && i0 + lane < N) ? lf[(r0 + rp) * N + i0 + lane] : 0.0;
        __syncwarp();
        const int lim = min(32, N - i0);
        for (int k = 0; k < lim; k++) n = __dsub_rn(n, tile[wid][lane][k] / 2);
        __syncwarp();
    }
    if (r0 + lane < R) E[r0 + lane] = n / 2;
}

// ------------------------------------------------------------------------------------------------
// lock-step Metropolis sweeps: CTA = RPC replicas, their local fields in shared memory
// ------------------------------------------------------------------------------------------------
struct sk_ls_params {
    const double *J; uint64_t *chunks; int64_t nchunks; double *lf; double *E; long long *acc; const double *beta;
    int N; int64_t R; uint64_t seed, sweep0; int nsweeps;
};
template <int RPC>
__global__ void __launch_bounds__(512, 1) k_sk_lockstep(sk_ls_params P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int N = P.N, tid = threadIdx.x, nt = blockDim.x;
    double *lf = reinterpret_cast<double *>(smem_raw);                    // [RPC][N]
    uint32_t *sp = reinterpret_cast<uint32_t *>(lf + (size_t)RPC * N);     // [RPC][nw]
    const int nw = (N + 31) / 32;
    __shared__ int flag[RPC];
    __shared__ int snew[RPC];
    const int64_t rbase = (int64_t)blockIdx.x * RPC;
    for (int rp = 0; rp < RPC; rp++) {
        const int64_t r = rbase + rp;
        for (int j = tid; j < N; j += nt) lf[(size_t)rp * N + j] = r < P.R ? P.lf[r * N + j] : 0.0;
        for (int w = tid; w < nw; w += nt) {
            const uint64_t c = r < P.R ? P.chunks[r * P.nchunks + (w >> 1)] : 0ull;
            sp[rp * nw + w] = (uint32_t)(c >> ((w & 1) * 32));
        }
    }
    double E = 0.0, beta = 0.0; long long nacc = 0;
    if (tid < RPC && rbase + tid < P.R) { E = P.E[rbase + tid]; beta = P.beta[rbase + tid]; nacc = P.acc[rbase + tid]; }
    __syncthreads();
    for (int sw = 0; sw < P.nsweeps; sw++) {
        const uint64_t t = P.sweep0 + (uint64_t)sw;
        for (int i = 0; i < N; i++) {
            if (tid < RPC) {                                   // Metropolis decision, accept() of RRRMC.jl:39
                const int64_t r = rbase + tid;
                int ok = 0;
                if (r < P.R) {
                    const double dE = lf[(size_t)tid * N + i];                      // ΔE_i = +lfields[i], SK.jl:278-284
                    const double x = -beta * dE;
                    if (x >= 0) ok = 1;
                    else {
                        const philox_out u = philox4x32_10((uint32_t)i, (uint32_t)r, (uint32_t)t, (uint32_t)(t >> 32) ^ 0x534b4c53u,
                                                           (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
                        const double U = (double)((((uint64_t)u.y << 32) | u.x) >> 11) * 0x1.0p-53;
                        ok = U < exp(x);
                    }
                    if (ok) { E += dE; nacc++; }
                }
                flag[tid] = ok;
                snew[tid] = 1 ^ (int)((sp[tid * nw + (i >> 5)] >> (i & 31)) & 1u);
            }
            __syncthreads();
            bool any = false;
#pragma unroll
            for (int rp = 0; rp < RPC; rp++) any |= flag[rp] != 0;
            if (any) {                                         // update_cache!, SK.jl:252-265, for the accepted replicas
                const double *Ji = P.J + (size_t)i * N;
                for (int j = tid; j < N; j += nt) {
                    const double Jij = Ji[j];
#pragma unroll
                    for (int rp = 0; rp < RPC; rp++) {
                        if (!flag[rp]) continue;
                        const int sj = (int)((sp[rp * nw + (j >> 5)] >> (j & 31)) & 1u);
                        double *p = &lf[(size_t)rp * N + j];
                        if (j == i) *p = -*p;
                        else *p = __dadd_rn(*p, 4 * __dmul_rn((double)(1 - 2 * (snew[rp] ^ sj)), Jij));
                    }
                }
            }
            __syncthreads();
            if (tid < RPC && flag[tid]) sp[tid * nw + (i >> 5)] ^= 1u << (i & 31);
        }
    }
    __syncthreads();
    for (int rp = 0; rp < RPC; rp++) {
        const int64_t r = rbase + rp;
        if (r >= P.R) continue;
        for (int j = tid; j < N; j += nt) P.lf[r * N + j] = lf[(size_t)rp * N + j];
        for (int c = tid; c < (int)P.nchunks; c += nt) {
            const uint64_t lo = sp[rp * nw + 2 * c], hi = 2 * c + 1 < nw ? sp[rp * nw + 2 * c + 1] : 0u;
            P.chunks[r * P.nchunks + c] = lo | (hi << 32);
        }
    }
    if (tid < RPC && rbase + tid < P.R) { P.E[rbase + tid] = E; P.acc[rbase + tid] = nacc; }
}

// The same sweeps with the coupling rows staged by the TMA engine: row i+1 of J (N doubles, one 1-D bulk copy) lands in
// a second shared-memory buffer while the block decides and updates site i, so the L2/HBM latency of the row — which the
// kernel above pays in full between its two barriers of every site — is hidden behind the previous site's work.
// Identical arithmetic and draw stream (bit-identical results). Needs N even (16-byte bulk copies) and room for two rows.
__device__ __forceinline__ uint32_t sk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RPC>
__global__ void __launch_bounds__(512, 1) k_sk_lockstep_tma(sk_ls_params P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int N = P.N, tid = threadIdx.x, nt = blockDim.x;
    double *Jb = reinterpret_cast<double *>(smem_raw);                     // [2][N] coupling rows
    double *lf = Jb + 2 * (size_t)N;                                       // [RPC][N]
    uint32_t *sp = reinterpret_cast<uint32_t *>(lf + (size_t)RPC * N);     // [RPC][nw]
    const int nw = (N + 31) / 32;
    __shared__ int flag[RPC];
    __shared__ int snew[RPC];
    __shared__ __align__(8) uint64_t bar[2];
    const uint32_t rowbytes = (uint32_t)N * 8u;
    const int64_t rbase = (int64_t)blockIdx.x * RPC;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sk_smem_u32(&bar[0])) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sk_smem_u32(&bar[1])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int rp = 0; rp < RPC; rp++) {
        const int64_t r = rbase + rp;
        for (int j = tid; j < N; j += nt) lf[(size_t)rp * N + j] = r < P.R ? P.lf[r * N + j] : 0.0;
        for (int w = tid; w < nw; w += nt) {
            const uint64_t c = r < P.R ? P.chunks[r * P.nchunks + (w >> 1)] : 0ull;
            sp[rp * nw + w] = (uint32_t)(c >> ((w & 1) * 32));
        }
    }
    double E = 0.0, beta = 0.0; long long nacc = 0;
    if (tid < RPC && rbase + tid < P.R) { E = P.E[rbase + tid]; beta = P.beta[rbase + tid]; nacc = P.acc[rbase + tid]; }
    __syncthreads();
    auto fetch_row = [&](int row, int buf) {      // thread 0: row -> Jb[buf], completion on bar[buf]
        const uint32_t b = sk_smem_u32(&bar[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(rowbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(sk_smem_u32(Jb + (size_t)buf * N)), "l"(P.J + (size_t)row * N), "r"(rowbytes), "r"(b) : "memory");
    };
    if (tid == 0) fetch_row(0, 0);
    const long long nsteps = (long long)P.nsweeps * N;
    long long g = 0;
    for (int sw = 0; sw < P.nsweeps; sw++) {
        const uint64_t t = P.sweep0 + (uint64_t)sw;
        for (int i = 0; i < N; i++, g++) {
            // Jb[(g+1)&1] was last read in step g-1, which ended with a barrier: request the next row now
            if (tid == 0 && g + 1 < nsteps) fetch_row(i + 1 < N ? i + 1 : 0, (int)((g + 1) & 1));
            if (tid < RPC) {                                   // Metropolis decision, accept() of RRRMC.jl:39
                const int64_t r = rbase + tid;
                int ok = 0;
                if (r < P.R) {
                    const double dE = lf[(size_t)tid * N + i];                      // ΔE_i = +lfields[i], SK.jl:278-284
                    const double x = -beta * dE;
                    if (x >= 0) ok = 1;
                    else {
                        const philox_out u = philox4x32_10((uint32_t)i, (uint32_t)r, (uint32_t)t, (uint32_t)(t >> 32) ^ 0x534b4c53u,
                                                           (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
                        const double U = (double)((((uint64_t)u.y << 32) | u.x) >> 11) * 0x1.0p-53;
                        ok = U < exp(x);
                    }
                    if (ok) { E += dE; nacc++; }
                }
                flag[tid] = ok;
                snew[tid] = 1 ^ (int)((sp[tid * nw + (i >> 5)] >> (i & 31)) & 1u);
            }
            __syncthreads();
            {   // every thread observes the row's arrival (also when no replica accepted: the barrier's phases stay in step)
                const uint32_t b = sk_smem_u32(&bar[g & 1]), parity = (uint32_t)((g >> 1) & 1);
                asm volatile("{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" :: "r"(b), "r"(parity) : "memory");
            }
            bool any = false;
#pragma unroll
            for (int rp = 0; rp < RPC; rp++) any |= flag[rp] != 0;
            if (any) {                                         // update_cache!, SK.jl:252-265, for the accepted replicas
                // two sites per thread and iteration (16-byte accesses); 4·(±1·J) is formed as ±(4·J): scaling by four
                // and the sign are exact, so the sum is bit-identical to the reference's expression
                const double2 *Ji2 = reinterpret_cast<const double2 *>(Jb + (size_t)(g & 1) * N);
                for (int j2 = tid; j2 < N / 2; j2 += nt) {
                    const double2 Jv = Ji2[j2];
                    const double a0 = 4 * Jv.x, a1 = 4 * Jv.y;
                    const int j = 2 * j2;
#pragma unroll
                    for (int rp = 0; rp < RPC; rp++) {
                        if (!flag[rp]) continue;
                        const uint32_t w = sp[rp * nw + (j >> 5)] >> (j & 31);
                        const int s0 = snew[rp] ^ (int)(w & 1u), s1 = snew[rp] ^ (int)((w >> 1) & 1u);
                        double2 *p = reinterpret_cast<double2 *>(&lf[(size_t)rp * N + j]);
                        double2 v = *p;
                        v.x = j == i ? -v.x : __dadd_rn(v.x, s0 ? -a0 : a0);
                        v.y = j + 1 == i ? -v.y : __dadd_rn(v.y, s1 ? -a1 : a1);
                        *p = v;
                    }
                }
            }
            __syncthreads();
            if (tid < RPC && flag[tid]) sp[tid * nw + (i >> 5)] ^= 1u << (i & 31);
        }
    }
    __syncthreads();
    for (int rp = 0; rp < RPC; rp++) {
        const int64_t r = rbase + rp;
        if (r >= P.R) continue;
        for (int j = tid; j < N; j += nt) P.lf[r * N + j] = lf[(size_t)rp * N + j];
        for (int c = tid; c < (int)P.nchunks; c += nt) {
            const uint64_t lo = sp[rp * nw + 2 * c], hi = 2 * c + 1 < nw ? sp[rp * nw + 2 * c + 1] : 0u;
            P.chunks[r * P.nchunks + c] = lo | (hi << 32);
        }
    }
    if (tid < RPC && rbase + tid < P.R) { P.E[rbase + tid] = E; P.acc[rbase + tid] = nacc; }
}

// Register-resident variant (N even, N <= 4096): the local fields of the block's RPC replicas live in the REGISTERS of
// sixteen update warps (thread b owns the site pairs b, b+512, ... of every replica), so a site step moves no field
// through shared memory at all — only the 8N-byte coupling row is read from it.
//
// The registers hold u_j = σ_j·lf_j (σ = 2s - 1) instead of lf_j. update_cache! (SK.jl:252-265) adds 4·σ_i'·σ_j·J_ij to
// lf_j (σ_i' the new spin of the flipped site); multiplied by σ_j that is u_j += 4·σ_i'·J_ij — ONE multiplier per replica
// and step, the same for every j, so the update is a single DFMA per field with no per-element sign work:
//   u_j <- fma(J_ij, c, u_j),  c = ±4 for a replica that flipped, 0 for one that did not.
// It is bit-identical to the reference order: 4·J is exact, so the DFMA rounds once like lf + (±4J), and rounding to
// nearest is symmetric under the common sign σ_j. Site i itself needs nothing: J_ii = 0 (validated when the graph is
// created) leaves u_i alone, and lf_i -> -lf_i together with σ_i -> -σ_i is u_i -> u_i. ΔE_i = lf_i = σ_i·u_i.
//
// Three roles, one barrier per site:
//   warp 0, lanes < RPC   decide site i+1 of their replica while row i is being applied: u of site i+1 arrives through a
//                         2-slot mailbox (published by its owner one step earlier), the lane applies row i to that one
//                         value itself (same DFMA, same operands as the owner: the two copies are bit-identical), then
//                         accept() of RRRMC.jl:39 with the uniform that warp 17 prepared, and posts the multiplier c;
//   warps 1..16           the DFMAs above on their registers for every replica (branch-free: a site step runs this code
//                         once, so skipping a replica costs more in instruction fetch and issue than its 8 DFMAs), then
//                         the owner of site i+2 publishes its u for the next step;
//   warp 17               requests coupling row i+2 (cp.async.bulk, three row buffers), draws the Philox uniforms of
//                         site i+2 and waits for row i+1 before the barrier (so nobody else polls an mbarrier) — all
//                         off the deciders' dependency chain c(i) -> u(i+1) -> exp -> c(i+1), which bounds a site step
//                         together with the FP64 pipe (8 N RPC / 64 cycles).
// The acceptance test u < exp(x) is decided in single precision whenever that is certain: e = __expf(x) is within
// 2 + 1.173|x| ulp (CUDA math API) of exp(x), so for -32 <= x < 0 a uniform outside e·(1 ± 2^-15) has the same answer as the
// double-precision comparison (error budget < 7e-6, margin 3e-5); inside the margin (probability < 1e-4 per decision), or
// below -32 unless u is clearly larger than exp(-32), the lane evaluates the double-precision exp as the other kernels do.
// Fields, energies and configurations are bit-identical to the kernels above (same draw stream).
__device__ __forceinline__ double2 sk_lds_f64x2(uint32_t a) { double2 v; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ double sk_lds_f64(uint32_t a) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t sk_lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sk_sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sk_sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
struct sk_draw { double u; float lo, hi; };   // the uniform of a decision and its single-precision bracket u·(1 ∓ 2^-15)
template <int RPC>
__global__ void __launch_bounds__(576, 1) k_sk_lockstep_reg(sk_ls_params P)
{
    constexpr int NB = 512, SLOTS = 4, NBUF = 3;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int N = P.N, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, bt = tid - 32, half = N >> 1;
    double *Jb = reinterpret_cast<double *>(smem_raw);                     // [NBUF][N] coupling rows
    uint32_t *sp = reinterpret_cast<uint32_t *>(Jb + NBUF * (size_t)N);    // [RPC][nw] spins, kept by the deciders
    const int nw = (N + 31) / 32;
    __shared__ __align__(64) double cmul[2][RPC];  // multiplier of each replica for the step that reads the slot: ±4, or 0 (no flip)
    __shared__ __align__(64) double mail[2][RPC];                // u of the site decided in the step that reads the slot (before that step's row)
    __shared__ __align__(16) sk_draw ubuf[16][RPC]; // the uniform draws of a ring of 16 steps (slot = step mod 16)
    __shared__ __align__(8) uint64_t bar[NBUF];
    const uint32_t rowbytes = (uint32_t)N * 8u;
    const int64_t rbase = (int64_t)blockIdx.x * RPC;
    const bool bulk = warp >= 1 && warp <= 16, decider = warp == 0 && lane < RPC;   // warp 17: rows and draws
    if (tid == 0) {
        for (int b = 0; b < NBUF; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(sk_smem_u32(&bar[b])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    for (int rp = 0; rp < RPC; rp++) {
        const int64_t r = rbase + rp;
        for (int w = tid; w < nw; w += blockDim.x) {
            const uint64_t c = r < P.R ? P.chunks[r * P.nchunks + (w >> 1)] : 0ull;
            sp[rp * nw + w] = (uint32_t)(c >> ((w & 1) * 32));
        }
    }
    __syncthreads();
    auto fetch_row = [&](int row, int buf) {      // one lane: row -> Jb[buf], completion on bar[buf]
        const uint32_t b = sk_smem_u32(&bar[buf]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(rowbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(sk_smem_u32(Jb + (size_t)buf * N)), "l"(P.J + (size_t)row * N), "r"(rowbytes), "r"(b) : "memory");
    };
    auto wait_row = [&](int buf, uint32_t parity) {
        const uint32_t b = sk_smem_u32(&bar[buf]);
        asm volatile("{\n\t.reg .pred p;\nW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@!p bra W_%=;\n\t}" :: "r"(b), "r"(parity) : "memory");
    };
    constexpr int Q = 8, RING = 2 * Q;            // draws are made Q steps at a time, by Q·RPC lanes of warp 17, into a ring of 2Q steps
    // the draw of step st (site st mod N of sweep sweep0 + st div N) for replica rbase + rp, as in the kernels above
    auto draw = [&](long long st, int rp) {
        const int i = (int)(st % N); const uint64_t t = P.sweep0 + (uint64_t)(st / N);
        const philox_out u = philox4x32_10((uint32_t)i, (uint32_t)(rbase + rp), (uint32_t)t, (uint32_t)(t >> 32) ^ 0x534b4c53u,
                                           (uint32_t)P.seed, (uint32_t)(P.seed >> 32));
        sk_draw d;
        d.u = (double)((((uint64_t)u.y << 32) | u.x) >> 11) * 0x1.0p-53;
        const float uf = (float)d.u;
        d.lo = uf * (1.0f - 0x1.0p-15f); d.hi = uf * (1.0f + 0x1.0p-15f);
        return d;
    };
    const long long nsteps = (long long)P.nsweeps * N;
    const int s1 = 1 < N ? 1 : 0;
    const uint32_t a_J = sk_smem_u32(Jb), a_cmul = sk_smem_u32(&cmul[0][0]), a_mail = sk_smem_u32(&mail[0][0]);
    // One code path per role (the roles share nothing but the barrier — one before the first step, one per step, one after
    // the last — so each keeps only its own state in registers, with its shared-memory addresses as 32-bit values advanced
    // incrementally). Step g works on site i = g mod N: row i sits in buffer g % 3 — complete, warp 17 saw it arrive
    // before the last barrier. Slot g & 1 of cmul / mail belongs to step g.
    if (bulk) {
        double2 v[SLOTS][RPC];                     // u of the pairs k·512 + bt
#pragma unroll
        for (int k = 0; k < SLOTS; k++) {
            const int j2 = k * NB + bt;
#pragma unroll
            for (int rp = 0; rp < RPC; rp++) {
                const int64_t r = rbase + rp;
                v[k][rp] = make_double2(0.0, 0.0);
                if (j2 < half && r < P.R) {
                    const double2 f = *reinterpret_cast<const double2 *>(&P.lf[r * N + 2 * j2]);
                    const uint32_t s2 = (uint32_t)(P.chunks[r * P.nchunks + (j2 >> 5)] >> ((2 * j2) & 63));
                    v[k][rp] = make_double2((s2 & 1u) ? f.x : -f.x, (s2 & 2u) ? f.y : -f.y);
                }
            }
        }
        auto publish = [&](int site, uint32_t a) { // the owner of `site` posts its u for every replica at mailbox slot a
            const int p = site >> 1;
            if (bt == (p & (NB - 1))) {
                const int k0 = p / NB;
#pragma unroll
                for (int k = 0; k < SLOTS; k++)
                    if (k == k0) {
#pragma unroll
                        for (int rp = 0; rp < RPC; rp++) sk_sts_f64(a + rp * 8, (site & 1) ? v[k][rp].y : v[k][rp].x);
                    }
            }
        };
        if (nsteps > 0) publish(s1, a_mail);
        __syncthreads();
        // a dead slot (pair index >= N/2, only when N < 4096) reads pair bt of the row instead: its registers are never
        // published nor written back, so the loop needs no predicates
        uint32_t off[SLOTS];
#pragma unroll
        for (int k = 0; k < SLOTS; k++) off[k] = (uint32_t)((k * NB + bt < half ? k * NB + bt : (bt < half ? bt : 0)) * 16);
        int in2 = 2 < N ? 2 : 2 - N;                                               // the site of step g + 2
        uint32_t a_c = a_cmul, a_m = a_mail + RPC * 8, row = a_J;
        const uint32_t row_end = a_J + NBUF * rowbytes;
        for (long long g = 0; g < nsteps; g++) {
            double c[RPC];
            if (RPC == 4) {
                const double2 c01 = sk_lds_f64x2(a_c), c23 = sk_lds_f64x2(a_c + 16);
                c[0] = c01.x; c[1 % RPC] = c01.y; c[2 % RPC] = c23.x; c[3 % RPC] = c23.y;
            } else if (RPC == 2) {
                const double2 c01 = sk_lds_f64x2(a_c);
                c[0] = c01.x; c[1 % RPC] = c01.y;
            } else c[0] = sk_lds_f64(a_c);
            double2 Jv[SLOTS];
#pragma unroll
            for (int k = 0; k < SLOTS; k++) Jv[k] = sk_lds_f64x2(row + off[k]);
#pragma unroll
            for (int k = 0; k < SLOTS; k++) {
#pragma unroll
                for (int rp = 0; rp < RPC; rp++) {
                    v[k][rp].x = fma(Jv[k].x, c[rp], v[k][rp].x);
                    v[k][rp].y = fma(Jv[k].y, c[rp], v[k][rp].y);
                }
            }
            publish(in2, a_m);
            if (++in2 == N) in2 = 0;
            a_c ^= RPC * 8; a_m ^= RPC * 8;
            row += rowbytes; if (row == row_end) row = a_J;
            __syncthreads();
        }
        __syncthreads();                                                           // the deciders have flushed the final spins
#pragma unroll
        for (int k = 0; k < SLOTS; k++) {                                          // lf = σ·u
            const int j2 = k * NB + bt;
#pragma unroll
            for (int rp = 0; rp < RPC; rp++) {
                const int64_t r = rbase + rp;
                if (j2 < half && r < P.R) {
                    const uint32_t s2 = sp[rp * nw + (j2 >> 4)] >> ((2 * j2) & 31);
                    *reinterpret_cast<double2 *>(&P.lf[r * N + 2 * j2]) =
                        make_double2((s2 & 1u) ? v[k][rp].x : -v[k][rp].x, (s2 & 2u) ? v[k][rp].y : -v[k][rp].y);
                }
            }
        }
    } else if (warp == 0) {
        // (all 32 lanes run the loop and meet at the same barrier instructions; lanes >= RPC only keep the barrier count)
        const int64_t rmine = rbase + lane;
        const bool live = decider && rmine < P.R;
        double E = 0.0, beta = 0.0, cp = 0.0; long long nacc = 0;                  // cp: the multiplier of the current step
        if (live) { E = P.E[rmine]; beta = P.beta[rmine]; nacc = P.acc[rmine]; }
        const double nbeta = -beta;
        // accept() of RRRMC.jl:39 on x = -βΔE with the prepared draw at shared address ad (ex2.approx: 2 ulp, and
        // -32·log2(e) is far above the denormal range, so no range reduction is needed)
        auto accept = [&](double x, uint32_t ad) -> int {
            if (x >= 0) return 1;
            float lo, hi, e;
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(lo), "=f"(hi) : "r"(ad + 8) : "memory");
            const float xf = (float)x;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaxf(xf, -32.0f) * 1.4426950408889634f));
            if (xf >= -32.0f && hi < e) return 1;
            if (lo > e) return 0;
            return sk_lds_f64(ad) < exp(x);
        };
        // the replica's spins: the word that holds the site being decided stays in a register (nobody else reads the
        // shared copy before the end) and is written back when the sweep moves on to the next word
        const uint32_t a_sp = sk_smem_u32(sp + (size_t)(decider ? lane : 0) * nw);
        uint32_t w = decider ? sk_lds_u32(a_sp) : 0u, a_w = a_sp;
        if (decider && nsteps > 0) {
            const double f = live ? P.lf[rmine * N] : 0.0;
            const sk_draw d = draw(0, lane);
            ubuf[0][lane] = d;
            const int ok = live ? accept(-beta * f, sk_smem_u32(&ubuf[0][lane])) : 0;
            if (ok) { E += f; nacc++; }
            cp = ok ? ((w & 1u) ? -4.0 : 4.0) : 0.0;                               // 4·σ' of the new spin
            cmul[0][lane] = cp;
            w ^= (uint32_t)ok;
        }
        __syncthreads();
        uint32_t a_m = a_mail + lane * 8, a_c = a_cmul + lane * 8 + RPC * 8, row = a_J;
        const uint32_t a_u0 = sk_smem_u32(&ubuf[0][decider ? lane : 0]), a_uend = a_u0 + RING * RPC * 16, row_end = a_J + NBUF * rowbytes;
        uint32_t a_u = a_u0 + RPC * 16;                                            // the draw of step g + 1
        int in = s1;                                                               // the site of step g + 1
        for (long long g = 0; g + 1 < nsteps; g++) {
            if (decider) {
                if ((in & 31) == 0) { sk_sts_u32(a_w, w); a_w = a_sp + (uint32_t)(in >> 5) * 4; w = sk_lds_u32(a_w); }
                // the chain: mailbox and coupling -> u of site i+1 under row i -> x = -β·σ·u -> accept -> multiplier
                const double m = sk_lds_f64(a_m);
                const double a = sk_lds_f64(row + (uint32_t)in * 8);
                const uint32_t sj = (w >> (in & 31)) & 1u;
                const double bs = sj ? nbeta : beta;                                // -β·σ of site i+1
                const double c4 = sj ? -4.0 : 4.0, sd = sj ? 1.0 : -1.0;
                const double u = fma(a, cp, m);                                     // row i on u of site i+1 (N >= 2: in != i)
                const int ok = live ? accept(bs * u, a_u) : 0;
                cp = ok ? c4 : 0.0;
                sk_sts_f64(a_c, cp);
                if (ok) { E = fma(u, sd, E); nacc++; w ^= 1u << (in & 31); }        // ΔE = lfields[i+1] = σ·u, SK.jl:278-284
            }
            if (++in == N) in = 0;
            a_m ^= RPC * 8; a_c ^= RPC * 8;
            a_u += RPC * 16; if (a_u == a_uend) a_u = a_u0;
            row += rowbytes; if (row == row_end) row = a_J;
            __syncthreads();
        }
        if (decider) sk_sts_u32(a_w, w);
        if (nsteps > 0) __syncthreads();                                           // the last step: nothing left to decide
        __syncthreads();
        if (live) { P.E[rmine] = E; P.acc[rmine] = nacc; }
    } else {
        const int qs = lane / RPC, qr = lane % RPC;                                // lane -> (step of the batch, replica)
        if (nsteps > 0) {
            if (lane == 0) { fetch_row(0, 0); if (nsteps > 1) fetch_row(s1, 1); }
            if (qs < Q) ubuf[(1 + qs) % RING][qr] = draw(1 + qs, qr);              // steps 1..Q, decided during steps 0..Q-1
            if (lane == 0) wait_row(0, 0);
        }
        __syncthreads();
        int in2 = 2 < N ? 2 : 2 - N, rb1 = 1; uint32_t par1 = 0;                  // buffer of row g+1 and its mbarrier phase
        for (long long g = 0; g < nsteps; g++) {
            // buffer (g+2) % 3 was last read in step g-1, which ended with the barrier
            if (lane == 0 && g + 2 < nsteps) fetch_row(in2, rb1 + 1 == NBUF ? 0 : rb1 + 1);
            // every Q steps: the draws of steps g+Q+1 .. g+2Q (read during steps g+Q .. g+2Q-1; their ring slots were last
            // read during steps g-Q .. g-1)
            if ((g & (Q - 1)) == 0 && qs < Q && g + Q + 1 + qs < nsteps) ubuf[(g + Q + 1 + qs) % RING][qr] = draw(g + Q + 1 + qs, qr);
            if (lane == 0 && g + 1 < nsteps) wait_row(rb1, par1);
            if (++in2 == N) in2 = 0;
            if (++rb1 == NBUF) { rb1 = 0; par1 ^= 1u; }
            __syncthreads();
        }
        __syncthreads();
    }
    for (int rp = 0; rp < RPC; rp++) {
        const int64_t r = rbase + rp;
        if (r >= P.R) continue;
        for (int c = tid; c < (int)P.nchunks; c += blockDim.x) {
            const uint64_t lo = sp[rp * nw + 2 * c], hi = 2 * c + 1 < nw ? sp[rp * nw + 2 * c + 1] : 0u;
            P.chunks[r * P.nchunks + c] = lo | (hi << 32);
        }
    }
}


// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct sk_dense_store {
    double *lf = nullptr, *E = nullptr, *beta = nullptr; long long *acc = nullptr;
    int8_t *Jq = nullptr, *S8 = nullptr; int Npad = 0; int64_t Rpad = 0; double scale = 0; int P = 0;
    bool fields_valid = false;
};

void sk_dense_free(rrrmc_state *s)
{
    sk_dense_store *d = s->skd;
    if (!d) return;
    cudaFree(d->lf); cudaFree(d->E); cudaFree(d->beta); cudaFree(d->acc); cudaFree(d->Jq); cudaFree(d->S8);
    delete d;
    s->skd = nullptr;
}
void sk_dense_invalidate(rrrmc_state *s) { if (s->skd) s->skd->fields_valid = false; }

static rrrmc_status_t sk_dense_ensure(rrrmc_state *s)
{
    rrrmc_graph *g = s->g;
    if (g->kind != RRRMC_SK_F64) { rrrmc_set_error("the dense SK kernels need a GraphSKNormal (RRRMC_SK_F64) graph"); return RRRMC_ERR_UNSUPPORTED; }
    if (!s->skd) {
        sk_dense_store *d = new sk_dense_store();
        s->skd = d;
        RR_CUDA(cudaMalloc(&d->lf, sizeof(double) * s->R * g->N));
        RR_CUDA(cudaMalloc(&d->E, sizeof(double) * s->R));
        RR_CUDA(cudaMalloc(&d->beta, sizeof(double) * s->R));
        RR_CUDA(cudaMalloc(&d->acc, sizeof(long long) * s->R));
        RR_CUDA(cudaMemsetAsync(d->acc, 0, sizeof(long long) * s->R, g->ctx->stream));
    }
    return RRRMC_OK;
}
// quantise J once: J ≈ 2^-P Σ_s d_s 256^s with signed 8-bit digits d_s (balanced, so that Σ is exact in 40 bits)
static rrrmc_status_t sk_dense_quantise(rrrmc_state *s, const std::vector<double> &J)
{
    rrrmc_graph *g = s->g; sk_dense_store *d = s->skd;
    if (d->Jq) return RRRMC_OK;
    const int64_t N = g->N;
    d->Npad = (int)(((N + TC_KB - 1) / TC_KB) * TC_KB);
    d->Rpad = ((s->R + TC_M - 1) / TC_M) * TC_M;
    double mx = 0;
    for (double v : J) mx = std::max(mx, std::fabs(v));
    int e = 0; if (mx > 0) frexp(mx, &e);                 // mx < 2^e
    d->P = 38 - e; d->scale = ldexp(1.0, -d->P);           // |J·2^P| < 2^38: digits fit five signed bytes
    std::vector<int8_t> q((size_t)SKQ_SLICES * d->Npad * d->Npad, 0);
    for (int64_t i = 0; i < N; i++)
        for (int64_t j = 0; j < N; j++) {
            long long v = llrint(ldexp(J[i * N + j], d->P));
            for (int sl = 0; sl < SKQ_SLICES; sl++) {
                const long long dg = ((v + 128) & 255) - 128;  // balanced digit in [-128, 127]
                q[((size_t)sl * d->Npad + i) * d->Npad + j] = (int8_t)dg;
                v = (v - dg) / 256;
            }
        }
    RR_CUDA(cudaMalloc(&d->Jq, q.size()));
    RR_CUDA(cudaMemcpy(d->Jq, q.data(), q.size(), cudaMemcpyHostToDevice));
    RR_CUDA(cudaDeviceSynchronize());
    RR_CUDA(cudaMalloc(&d->S8, (size_t)d->Rpad * d->Npad));
    RR_CUDA(cudaMemsetAsync(d->S8, 0, (size_t)d->Rpad * d->Npad, s->g->ctx->stream)); // on the context stream: ordered before k_spins_to_s8
    return RRRMC_OK;
}

rrrmc_status_t sk_dense_fields_init(rrrmc_state *s, int use_tensor_cores, double *E_out, float *ms_out)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_TRY(sk_dense_ensure(s));
    RR_TRY(chain_sync_from_multispin(s));
    sk_dense_store *d = s->skd;
    const int N = (int)g->N;
    cudaEvent_t e0, e1;
    RR_CUDA(cudaEventCreate(&e0)); RR_CUDA(cudaEventCreate(&e1));
    if (use_tensor_cores) {
        RR_TRY(sk_dense_quantise(s, g->Jd));
        RR_CUDA(cudaEventRecord(e0, ctx->stream));
        k_spins_to_s8<<<div_up(s->R * d->Npad, 256), 256, 0, ctx->stream>>>(s->d_chunks, s->nchunks, s->R, N, d->Npad, d->S8);
        sk_tc_params P{ d->Jq, d->S8, d->lf, N, d->Npad, s->R, d->Rpad, d->scale };
        dim3 grid(d->Npad / TC_N, (unsigned)(d->Rpad / TC_M));
        const int tc_smem = TC_STAGES * TC_STAGE_BYTES + 1024;
        RR_CUDA(cudaFuncSetAttribute(k_sk_fields_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem));
        k_sk_fields_tc<<<grid, 128, tc_smem, ctx->stream>>>(P);
        ctx->launches += 2;
    } else {
        RR_CUDA(cudaEventRecord(e0, ctx->stream));
        k_sk_fields_ordered<<<div_up(s->R * N, 128), 128, 0, ctx->stream>>>(g->d_Jd, s->d_chunks, s->nchunks, s->R, N, d->lf);
        ctx->launches++;
    }
    RR_CUDA(cudaEventRecord(e1, ctx->stream));
    k_sk_energy_from_fields<<<div_up(s->R, 128), 128, 0, ctx->stream>>>(d->lf, s->R, N, d->E);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    if (E_out) RR_CUDA(cudaMemcpyAsync(E_out, d->E, sizeof(double) * s->R, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ms_out) RR_CUDA(cudaEventElapsedTime(ms_out, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    d->fields_valid = true;
    return RRRMC_OK;
}
rrrmc_status_t sk_dense_get_fields(rrrmc_state *s, double *lf_out)
{
    RR_ARG(s->skd && s->skd->fields_valid, "local fields are not initialised: call rrrmc_sk_fields_init first");
    RR_CUDA(cudaMemcpy(lf_out, s->skd->lf, sizeof(double) * s->R * s->g->N, cudaMemcpyDeviceToHost));
    return RRRMC_OK;
}

rrrmc_status_t sk_dense_sweeps(rrrmc_state *s, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                               double *E_out, int64_t *acc_out)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    RR_TRY(sk_dense_ensure(s));
    sk_dense_store *d = s->skd;
    if (!d->fields_valid) RR_TRY(sk_dense_fields_init(s, 1, nullptr, nullptr));
    RR_TRY(chain_sync_from_multispin(s));
    for (int64_t r = 0; r < s->R; r++) RR_ARG(std::isfinite(beta[r]), "β must be finite, given: %g", beta[r]);
    RR_CUDA(cudaMemcpyAsync(d->beta, beta, sizeof(double) * s->R, cudaMemcpyHostToDevice, ctx->stream));
    const int N = (int)g->N, nw = (N + 31) / 32;
    sk_ls_params P{ g->d_Jd, s->d_chunks, s->nchunks, d->lf, d->E, d->acc, d->beta, N, s->R, seed, sweep0, (int)nsweeps };
    int rpc = 4;
    while (rpc > 1 && (size_t)rpc * N * 8 + (size_t)rpc * nw * 4 > (size_t)200 * 1024) rpc >>= 1;
    const size_t smem = (size_t)rpc * N * 8 + (size_t)rpc * nw * 4;
    RR_ARG(smem <= (size_t)220 * 1024, "N = %d is too large for the lock-step SK kernel (local fields must fit shared memory)", N);
    const unsigned grid = div_up(s->R, rpc);
#define LS(RP) do { RR_CUDA(cudaFuncSetAttribute(k_sk_lockstep<RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
                    k_sk_lockstep<RP><<<grid, 512, smem, ctx->stream>>>(P); } while (0)
    // the TMA-staged kernel when two coupling rows fit beside the fields (RRRMC_SK_VARIANT=1 keeps the plain loads)
    const size_t smem_tma = smem + 2 * (size_t)N * 8;
    const char *skv = getenv("RRRMC_SK_VARIANT");
    const bool tma = N % 2 == 0 && smem_tma <= (size_t)225 * 1024 && !(skv && atoi(skv) == 1);
    const int variant = skv ? atoi(skv) : 0;   // RRRMC_SK_VARIANT: 0 best available, 1 plain loads, 2 TMA rows + fields in shared memory
    // fields in registers (N even, <= 4096): the default; one block per RPC replicas, RPC no larger than the SM count asks for
    const bool reg = N % 2 == 0 && N >= 2 && N <= 4096 && variant == 0;
#define LST(RP) do { RR_CUDA(cudaFuncSetAttribute(k_sk_lockstep_tma<RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_tma)); \
                     k_sk_lockstep_tma<RP><<<grid, 512, smem_tma, ctx->stream>>>(P); } while (0)
#define LSR(RP) do { const size_t sm = 3 * (size_t)N * 8 + (size_t)RP * nw * 4; \
                     RR_CUDA(cudaFuncSetAttribute(k_sk_lockstep_reg<RP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
                     k_sk_lockstep_reg<RP><<<div_up(s->R, RP), 576, sm, ctx->stream>>>(P); } while (0)
    if (reg) {
        const int nsm = ctx->sm_count > 0 ? ctx->sm_count : 148;
        if (s->R > 2 * (int64_t)nsm) LSR(4); else if (s->R > nsm) LSR(2); else LSR(1);
    }
    else if (tma) { if (rpc == 4) LST(4); else if (rpc == 2) LST(2); else LST(1); }
    else if (rpc == 4) LS(4); else if (rpc == 2) LS(2); else LS(1);
#undef LSR
#undef LS
#undef LST
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    s->ms_valid = false; s->chain_valid = true; s->chain_fields_valid = false; s->energy_valid = false;
    if (E_out) RR_CUDA(cudaMemcpyAsync(E_out, d->E, sizeof(double) * s->R, cudaMemcpyDeviceToHost, ctx->stream));
    if (acc_out) RR_CUDA(cudaMemcpyAsync(acc_out, d->acc, sizeof(long long) * s->R, cudaMemcpyDeviceToHost, ctx->stream));
    RR_CUDA(cudaStreamSynchronize(ctx->stream));
    return RRRMC_OK;
}
