// Replay mode of the sequential samplers: k_chain_run fed a dumped typed draw stream instead of
// the Philox counter stream. Its own translation unit so that the two instantiations compile in parallel.
#include "chain_kernel.cuh"

void chain_launch_run_trace(const chain_params &P, unsigned grid, cudaStream_t st)
{
    k_chain_run<src_trace><<<grid, 32, 0, st>>>(P);
}
