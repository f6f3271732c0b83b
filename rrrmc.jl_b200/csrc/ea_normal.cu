// Checkerboard Metropolis half-sweep for continuous couplings — GraphEANormal (src/graphs/EA.jl:534-680) on the
// replica batch, the multispin layout of the ±J kernels (spins[N][W], one bit per replica).
//
// ΔE cannot be bit-sliced when the couplings are real numbers, so the arithmetic is per lane: a warp owns one active
// site and walks over four 32-replica words of it, lane l = replica 32w + l. The seven spin words and the 2D
// (neighbour, coupling) pairs of the site are warp-uniform loads (one transaction each, shared by 32 replicas; the
// couplings are shared by ALL replicas); every lane then accumulates
//     lf = 0;  for k = 1..2D:  lf -= J[x][k]·σx·σy      (slot order of energy(), EA.jl:590-603: products are exact,
//     ΔE = -2·lf                                          so lf is the reference's freshly initialised lfields[x]/2)
// in Float64 — bit for bit the value delta_energy() returns after energy() (EA.jl:665-672) — and applies
// accept(-βΔE) (RRRMC.jl:39): ΔE <= 0 flips, else u < exp(-βΔE) with a 53-bit uniform of its own Philox4x32-10 call,
// counter (sweep_hi<<16, site, replica, sweep_lo), key = seed. β is per replica (parallel-tempering ladders run as
// they are). The 32 decisions are collected with one ballot and lane 0 writes the word back.
// CPU restatement: oracle/rrrmc_oracle.c:orc_checkerboard_sweeps_f64 (same bits, tests/test_gpu_ea_normal.py).
// Bound: per-lane fp64 + Philox, ~5 warp-instructions per attempt — this is the generic path for real couplings, not
// the ±J headline kernel.
#include "common.cuh"
#include "philox.cuh"
#include "kernels.cuh"
#include "ea_normal.cuh"

template <int TWOD>
__global__ void __launch_bounds__(256) k_checkerboard_f64(const cbn_params p, int colour)
{
    const int lane = threadIdx.x & 31, Lh = p.L >> 1;
    const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); task < p.ntasks; task += nwarps) {
        const int64_t a = task / p.nwg; const int wg = (int)(task - a * p.nwg);
        // site a of this colour: half-row index xh, the other coordinates enumerated by `rest`
        const int64_t rest = a / Lh; const int xh = (int)(a - rest * Lh);
        int par = 0; { int64_t q = rest; for (int d = 1; d < p.D; d++) { par += (int)(q % p.L); q /= p.L; } }
        const int64_t i = 2 * xh + ((par + colour) & 1) + (int64_t)p.L * rest;
        int64_t nb[TWOD]; double Jk[TWOD];
#pragma unroll
        for (int k = 0; k < TWOD; k++) { nb[k] = (int64_t)__ldg(p.A + i * TWOD + k) * p.W; Jk[k] = __ldg(p.J + i * TWOD + k); }
        const int w1 = min(p.W, 4 * wg + 4);
        for (int w = 4 * wg; w < w1; w++) {
            const int64_t r = 32 * (int64_t)w + lane;
            const uint32_t c = p.spins[i * p.W + w];
            const uint32_t sx = (c >> lane) & 1u;
            double lf = 0.0;
#pragma unroll
            for (int k = 0; k < TWOD; k++) {
                const uint32_t sy = (p.spins[nb[k] + w] >> lane) & 1u;
                lf = (sx == sy) ? __dsub_rn(lf, Jk[k]) : __dadd_rn(lf, Jk[k]);      // lf -= J·σxσy, exactly
            }
            const double dE = -2.0 * lf;
            const double x = __dmul_rn(-(r < p.R ? p.beta[r] : 0.0), dE);
            bool flip = x >= 0.0;
            if (!flip) {
                const philox_out o = philox4x32_10(p.t_hi16, (uint32_t)i, (uint32_t)r, p.t_lo, p.k0, p.k1);
                const double u = (double)((((uint64_t)o.y << 32) | o.x) >> 11) * 0x1.0p-53;
                flip = u < exp(x);
            }
            const uint32_t mask = __ballot_sync(0xffffffffu, flip && r < p.R);
            if (lane == 0) {
                p.spins[i * p.W + w] = c ^ mask;
                if (p.flips) p.flips[i * p.W + w] = mask;
            }
        }
    }
}

rrrmc_status_t launch_checkerboard_f64(rrrmc_ctx *ctx, const cbn_params &p, int colour)
{
    const int warps_per_block = 8;
    int64_t blocks = (p.ntasks + warps_per_block - 1) / warps_per_block;
    const int64_t cap = (int64_t)ctx->sm_count * 64;                 // grid-stride beyond a few waves
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    switch (p.twoD) {
    case 2: k_checkerboard_f64<2><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, colour); break;
    case 4: k_checkerboard_f64<4><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, colour); break;
    case 6: k_checkerboard_f64<6><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, colour); break;
    case 8: k_checkerboard_f64<8><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, colour); break;
    default: rrrmc_set_error("checkerboard (continuous couplings): 2D = %d unsupported (D <= 4)", p.twoD); return RRRMC_ERR_UNSUPPORTED;
    }
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}
