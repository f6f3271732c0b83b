// Microbenchmark: issue rate of LOP3 with three register sources, of SHF, and of LOP3 + IMAD.WIDE mixes per SM
// sub-partition, against the nominal 0.5 warp-instructions per cycle of the ALU pipe (tuning aid: what ALU-pipe
// utilisation can a LOP3-dominated kernel reach at all?).  nvcc -arch=sm_100a -O3 lop3_rate.cu -o lop3_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int LUT> __device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT)); return r;
}
// MODE 0: 16 independent LOP3 chains, three distinct register sources each (like the flip logic)
// MODE 1: LOP3 with two register sources + immediate-free constant (x ^ y)
// MODE 2: SHF (variable shift)
// MODE 3: 2 LOP3 : 1 IMAD.WIDE (Philox-like mix)
// MODE 4: IMAD.WIDE only
template <int MODE>
__global__ void __launch_bounds__(1024) k(uint32_t *out, int n, uint32_t a, long long *cyc)
{
    uint32_t x[16];
#pragma unroll
    for (int c = 0; c < 16; c++) x[c] = threadIdx.x * 2654435761u + c * a;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < n; i++) {
#pragma unroll
        for (int c = 0; c < 16; c++) {
            if (MODE == 0) x[c] = lop3<0x96>(x[c], x[(c + 5) & 15], x[(c + 11) & 15]);
            if (MODE == 1) x[c] = lop3<0x3c>(x[c], x[(c + 5) & 15], 0u);
            if (MODE == 2) { uint32_t r; asm volatile("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x[c]), "r"(x[(c + 5) & 15])); x[c] = r | 1u; }
            if (MODE == 3) {
                if (c % 3 == 2) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(0xD2511F53u)); x[c] = (uint32_t)(p >> 32); x[(c + 1) & 15] ^= (uint32_t)p; }
                else x[c] = lop3<0x96>(x[c], x[(c + 5) & 15], x[(c + 11) & 15]);
            }
            if (MODE == 4) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(x[c]), "r"(0xD2511F53u)); x[c] = (uint32_t)(p >> 32) + (uint32_t)p; }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < 16; c++) s ^= x[c];
    if (s == 0x12345u) out[threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE> void run(const char *name, int warps_per_smsp, double instr_per_iter)
{
    uint32_t *out; long long *cyc, h;
    cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int n = 20000, threads = 32 * 4 * warps_per_smsp;     // one block per SM, warps spread over the 4 sub-partitions
    k<MODE><<<148, threads>>>(out, 10, 3, cyc); cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, n, 3, cyc); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %d warps/SMSP: %.3f warp-instr/cycle/SMSP\n", name, warps_per_smsp, instr_per_iter * n * warps_per_smsp / (double)h);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    for (int w : {1, 2, 4, 8}) {
        if (w == 1) { run<0>("LOP3 3 register sources", 1, 16); run<1>("LOP3 2 register sources", 1, 16); run<2>("SHF variable", 1, 32); run<3>("2 LOP3 : 1 IMAD.WIDE (+1 LOP3)", 1, 16 + 5); run<4>("IMAD.WIDE + IADD", 1, 32); }
        if (w == 2) { run<0>("LOP3 3 register sources", 2, 16); run<1>("LOP3 2 register sources", 2, 16); run<2>("SHF variable", 2, 32); run<3>("2 LOP3 : 1 IMAD.WIDE (+1 LOP3)", 2, 16 + 5); run<4>("IMAD.WIDE + IADD", 2, 32); }
        if (w == 4) { run<0>("LOP3 3 register sources", 4, 16); run<1>("LOP3 2 register sources", 4, 16); run<2>("SHF variable", 4, 32); run<3>("2 LOP3 : 1 IMAD.WIDE (+1 LOP3)", 4, 16 + 5); run<4>("IMAD.WIDE + IADD", 4, 32); }
        if (w == 8) { run<0>("LOP3 3 register sources", 8, 16); run<1>("LOP3 2 register sources", 8, 16); run<2>("SHF variable", 8, 32); run<3>("2 LOP3 : 1 IMAD.WIDE (+1 LOP3)", 8, 16 + 5); run<4>("IMAD.WIDE + IADD", 8, 32); }
    }
    return 0;
}
