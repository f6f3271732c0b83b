// Microbenchmark: issue rate of Philox4x32-10 per SM sub-partition under different ILP / occupancy (tuning aid).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rrrmc.jl_b200/csrc/philox.cuh"

template <int ILP>
__global__ void __launch_bounds__(256) k_philox(uint32_t *out, int ncall, uint32_t k0, uint32_t k1)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc[4] = {0, 0, 0, 0};
    for (int q = 0; q < ncall; q += ILP) {
#pragma unroll
        for (int j = 0; j < ILP; j++) {
            const philox_out r = philox4x32_10(q + j, tid, 7u, 11u, k0, k1);
            acc[0] |= r.x; acc[1] &= r.y; acc[2] |= r.z; acc[3] &= r.w;   // 1 LOP3-ish per word, like an op0 plane
        }
    }
    if ((acc[0] ^ acc[1] ^ acc[2] ^ acc[3]) == 0x12345u) out[tid] = acc[0];
}
// LOP3-only and IMAD-only loops to calibrate pipe rates
__global__ void __launch_bounds__(256) k_lop(uint32_t *out, int n, uint32_t a, uint32_t b)
{
    uint32_t x0 = threadIdx.x, x1 = a, x2 = b, x3 = a ^ b, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
    for (int i = 0; i < n; i++) {
        x0 = (x0 & x1) ^ x2; x1 = (x1 & x2) ^ x3; x2 = (x2 & x3) ^ x4; x3 = (x3 & x4) ^ x5;
        x4 = (x4 & x5) ^ x6; x5 = (x5 & x6) ^ x7; x6 = (x6 & x7) ^ x0; x7 = (x7 & x0) ^ x1;
    }
    if ((x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7) == 0x12345u) out[threadIdx.x] = x0;
}
__global__ void __launch_bounds__(256) k_mulwide(uint32_t *out, int n, uint32_t a, uint32_t b)
{
    uint32_t x0 = threadIdx.x, x1 = a, x2 = b, x3 = a ^ b, x4 = 5, x5 = 6, x6 = 7, x7 = 8;
    for (int i = 0; i < n; i++) {
        uint64_t p;
        p = (uint64_t)x0 * 0xD2511F53u; x0 = (uint32_t)p + (uint32_t)(p >> 32);
        p = (uint64_t)x1 * 0xD2511F53u; x1 = (uint32_t)p + (uint32_t)(p >> 32);
        p = (uint64_t)x2 * 0xD2511F53u; x2 = (uint32_t)p + (uint32_t)(p >> 32);
        p = (uint64_t)x3 * 0xD2511F53u; x3 = (uint32_t)p + (uint32_t)(p >> 32);
        p = (uint64_t)x4 * 0xD2511F53u; x4 = (uint32_t)p + (uint32_t)(p >> 32);
        p = (uint64_t)x5 * 0xD2511F53u; x5 = (uint32_t)p + (uint32_t)(p >> 32);
        p = (uint64_t)x6 * 0xD2511F53u; x6 = (uint32_t)p + (uint32_t)(p >> 32);
        p = (uint64_t)x7 * 0xD2511F53u; x7 = (uint32_t)p + (uint32_t)(p >> 32);
    }
    if ((x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7) == 0x12345u) out[threadIdx.x] = x0;
}

template <class F> float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main()
{
    uint32_t *d; cudaMalloc(&d, 1 << 24);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    const int ncall = 4096;
    for (int bps : {1, 2, 4, 6, 8}) {
        const int grid = sm * bps;
        const double warps_smsp = bps * 8 / 4.0;
        float m1 = timeit([&] { k_philox<1><<<grid, 256>>>(d, ncall, 1, 2); });
        float m2 = timeit([&] { k_philox<2><<<grid, 256>>>(d, ncall, 1, 2); });
        float m4 = timeit([&] { k_philox<4><<<grid, 256>>>(d, ncall, 1, 2); });
        // cycles per warp-level call per SMSP
        auto cpc = [&](float ms) { return ms * 1e-3 * clk / (ncall * warps_smsp); };
        printf("blocks/SM=%d warps/SMSP=%.0f  cycles per Philox call per SMSP: ILP1 %.1f  ILP2 %.1f  ILP4 %.1f\n", bps, warps_smsp, cpc(m1), cpc(m2), cpc(m4));
    }
    for (int bps : {2, 4, 8}) {
        const int grid = sm * bps; const int n = 20000;
        const double warps_smsp = bps * 8 / 4.0;
        float ml = timeit([&] { k_lop<<<grid, 256>>>(d, n, 3, 5); });
        float mm = timeit([&] { k_mulwide<<<grid, 256>>>(d, n, 3, 5); });
        printf("blocks/SM=%d  LOP3: %.2f cycles/warp-instr/SMSP   (IMAD.WIDE+IADD): %.2f cycles/pair/SMSP\n", bps,
               ml * 1e-3 * clk / (n * 8.0 * warps_smsp), mm * 1e-3 * clk / (n * 8.0 * warps_smsp));
    }
    return 0;
}
