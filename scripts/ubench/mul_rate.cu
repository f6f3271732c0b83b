// Microbenchmark: issue rates of the 32x32 multiply flavours Philox can be built from (tuning aid).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CH 8
#define BODY(...)                                                               \
    uint32_t x[CH];                                                             \
    for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 2654435761u + c + a;      \
    for (int i = 0; i < n; i++) {                                               \
        _Pragma("unroll") for (int c = 0; c < CH; c++) { __VA_ARGS__; }                \
    }                                                                           \
    uint32_t s = 0;                                                             \
    for (int c = 0; c < CH; c++) s ^= x[c];                                     \
    if (s == 0x12345u) out[threadIdx.x] = s;

__global__ void __launch_bounds__(256) k_wide_xor(uint32_t *out, int n, uint32_t a)
{ BODY(uint64_t p = (uint64_t)x[c] * 0xD2511F53u; x[c] = (uint32_t)p ^ (uint32_t)(p >> 32)) }
__global__ void __launch_bounds__(256) k_hi_lo_xor(uint32_t *out, int n, uint32_t a)
{ BODY(uint32_t h; uint32_t l; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(x[c]), "r"(0xD2511F53u)); asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(l) : "r"(x[c]), "r"(0xD2511F53u)); x[c] = h ^ l) }
__global__ void __launch_bounds__(256) k_hi(uint32_t *out, int n, uint32_t a)
{ BODY(x[c] = __umulhi(x[c], 0xD2511F53u) + 1u) }
__global__ void __launch_bounds__(256) k_hi_only(uint32_t *out, int n, uint32_t a)
{ BODY(asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(0xD2511F53u))) }
__global__ void __launch_bounds__(256) k_lo(uint32_t *out, int n, uint32_t a)
{ BODY(x[c] = x[c] * 0xD2511F53u + 12345u) }
__global__ void __launch_bounds__(256) k_wide_only(uint32_t *out, int n, uint32_t a)
{ BODY(uint32_t l, h; asm volatile("{.reg .u64 p; mul.wide.u32 p, %2, %3; mov.b64 {%0,%1}, p;}" : "=r"(l), "=r"(h) : "r"(x[c]), "r"(0xD2511F53u)); x[c] = l; if (h == 77u && i == n) x[c] ^= 1u) }
__global__ void __launch_bounds__(256) k_mad16(uint32_t *out, int n, uint32_t a)
{ BODY(uint32_t l = x[c] & 0xffffu; uint32_t h = x[c] >> 16; x[c] = l * 0x1F53u + h * 0xD251u) }
__global__ void __launch_bounds__(256) k_ffma(uint32_t *out, int n, uint32_t a)
{ BODY(float f = __uint_as_float(x[c]); f = f * 1.0001f + 0.5f; x[c] = __float_as_uint(f)) }
__global__ void __launch_bounds__(256) k_lop_and_lo(uint32_t *out, int n, uint32_t a)
{ BODY(x[c] = (x[c] * 0xD2511F53u) ^ (x[c] >> 3)) }

template <class F> float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main()
{
    uint32_t *d; cudaMalloc(&d, 1 << 20);
    int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3; const int n = 20000, bps = 4; const int grid = sm * bps; const double wps = bps * 8 / 4.0;
#define RUN(K) { float ms = timeit([&] { K<<<grid, 256>>>(d, n, 3); }); printf("%-14s %.2f cycles per chain-step per warp per SMSP\n", #K, ms * 1e-3 * clk / (n * (double)CH * wps)); }
    RUN(k_wide_xor) RUN(k_hi_lo_xor) RUN(k_hi) RUN(k_hi_only) RUN(k_lo) RUN(k_wide_only) RUN(k_mad16) RUN(k_ffma) RUN(k_lop_and_lo)
    return 0;
}
