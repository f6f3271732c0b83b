// FP64 pipe rate and latency on one SM: independent DFMA chains per thread, CUDA-event timing, clock64 cycles.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dfma_rate scripts/ubench/dfma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void __launch_bounds__(1024) k(double *out, int iters, long long *cyc)
{
    double a[ILP];
    for (int k = 0; k < ILP; k++) a[k] = threadIdx.x * 1e-3 + k;
    const double b = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) a[k] = fma(a[k], b, c);
    }
    const long long t1 = clock64();
    double s = 0; for (int k = 0; k < ILP; k++) s += a[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int threads, int blocks, int iters)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * threads * blocks); cudaMalloc(&cyc, 8);
    k<ILP><<<blocks, threads>>>(out, iters, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<ILP><<<blocks, threads>>>(out, iters, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double n = (double)threads * ILP * iters;
    printf("ILP %d threads %4d blocks %3d: %.2f DFMA lanes/clk/SM (block 0), %.1f cycles per dependent DFMA round, %.2f TFLOP/s whole launch\n",
           ILP, threads, blocks, n / h, (double)h / iters, 2.0 * n * blocks / (ms * 1e-3) / 1e12);
    cudaFree(out); cudaFree(cyc);
}
int main()
{
    run<1>(32, 1, 100000);      // latency: one warp, one chain
    run<8>(32, 1, 100000);
    run<8>(128, 1, 100000);
    run<8>(512, 1, 100000);
    run<8>(1024, 1, 50000);
    run<8>(1024, 148, 50000);
    return 0;
}
