// Philox4x32-10 counter-based RNG (Salmon, Moraes, Dror, Shaw, SC'11), host+device.
// Keys are kernel-uniform (seed only) so the per-round key schedule lives in uniform registers.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RRRMC_HD __host__ __device__ __forceinline__
#else
#define RRRMC_HD inline
#endif

struct philox_out { uint32_t x, y, z, w; };

RRRMC_HD philox_out philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    philox_out o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}

// Chain draw source: key = seed, counter = (n_lo, n_hi, chain_lo, tag ^ chain_hi); one call per draw.
struct chain_rng {
    uint64_t seed, chain, n; uint32_t tag;
    RRRMC_HD uint64_t u64()
    {
        philox_out o = philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)chain, tag ^ (uint32_t)(chain >> 32),
                                     (uint32_t)seed, (uint32_t)(seed >> 32));
        n++;
        return ((uint64_t)o.y << 32) | o.x;
    }
    RRRMC_HD double f64() { return (double)(u64() >> 11) * 0x1.0p-53; }  // rand(): [0,1), 53 bits
    RRRMC_HD int64_t range(int64_t nn)                                    // rand(1:n), unbiased (Lemire)
    {
        const uint64_t un = (uint64_t)nn;
        for (;;) {
            const uint64_t x = u64();
#ifdef __CUDA_ARCH__
            const uint64_t hi = __umul64hi(x, un), lo = x * un;
#else
            const unsigned __int128 m = (unsigned __int128)x * un;
            const uint64_t hi = (uint64_t)(m >> 64), lo = (uint64_t)m;
#endif
            if (lo < un) { const uint64_t t = (0 - un) % un; if (lo < t) continue; }
            return (int64_t)hi + 1;
        }
    }
};
