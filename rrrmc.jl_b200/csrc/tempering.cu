// Parallel-tempering exchange for a β ladder laid over the 128-replica groups of a ±J GraphEA batch (the layout the
// checkerboard ladder kernel runs, ea_tma.cu): lane l of every group is one ladder, group g its rung at β_g.
// The reference has no tempering (it runs one standardMC per β, RRRMC.jl:81-127); BASELINE's north_star asks for
// "optional parallel-tempering swaps". Everything stays on the device: energies from k_energy_pm1, one decision per
// (pair of neighbouring groups, lane), and the accepted pairs exchange their CONFIGURATIONS (a masked swap of bit l
// between the two groups' words of every site) so that β stays constant inside a group, which the ladder kernel needs.
//
// Rule (the same as sharding.TemperingLadder.swap, classical action): pairs (g, g+1) with g ≡ round (mod 2);
//   ΔS = (β_g − β_{g+1})·(E_b − E_a), a = replica 128 g + l, b = replica 128 (g+1) + l; accept iff ΔS <= 0 or u < exp(−ΔS),
//   u = ((y:x) >> 11)·2^-53 from Philox4x32-10(counter = (round_lo, round_hi, g, l), key = seed).
// CPU restatement: oracle/rrrmc_oracle.c:orc_tempering_decide.
#include "kernels.cuh"
#include "philox.cuh"

namespace {

// one block per pair, one thread per lane; masks[g][4] = lanes of the pair (g, g+1) that exchange
__global__ void __launch_bounds__(128) k_pt_decide(const int *__restrict__ unsat, const double *__restrict__ beta_group, int G, int parity,
                                                   uint32_t k0, uint32_t k1, uint32_t r_lo, uint32_t r_hi,
                                                   uint32_t *__restrict__ masks, long long *__restrict__ accepted)
{
    const int g = 2 * (int)blockIdx.x + parity, l = threadIdx.x;
    if (g + 1 >= G) return;
    const int ua = unsat[128 * g + l], ub = unsat[128 * (g + 1) + l];
    // E = −D·N + 2·unsat: E_b − E_a = 2 (ub − ua), exact
    const double dS = __dmul_rn(__dsub_rn(beta_group[g], beta_group[g + 1]), 2.0 * (double)(ub - ua));
    bool acc = dS <= 0.0;
    if (!acc) {
        const philox_out o = philox4x32_10(r_lo, r_hi, (uint32_t)g, (uint32_t)l, k0, k1);
        const double u = (double)((((uint64_t)o.y << 32) | o.x) >> 11) * 0x1.0p-53;
        acc = u < exp(-dS);
    }
    const uint32_t word = __ballot_sync(0xffffffffu, acc);
    if ((l & 31) == 0) {
        masks[4 * g + (l >> 5)] = word;
        if (accepted && word) atomicAdd(reinterpret_cast<unsigned long long *>(accepted + g), (unsigned long long)__popc(word));
    }
}

// thread = (site, pair): exchanges the masked lanes of groups g and g+1
__global__ void __launch_bounds__(256) k_pt_exchange(uint4 *__restrict__ spins4, const uint4 *__restrict__ masks4, int64_t N, int G,
                                                     int parity, int npairs)
{
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= N * npairs) return;
    const int64_t i = id / npairs;
    const int g = 2 * (int)(id - i * npairs) + parity;
    const uint4 m = __ldg(masks4 + g);
    if ((m.x | m.y | m.z | m.w) == 0u) return;
    uint4 a = spins4[i * G + g], b = spins4[i * G + g + 1];
    const uint4 x = make_uint4((a.x ^ b.x) & m.x, (a.y ^ b.y) & m.y, (a.z ^ b.z) & m.z, (a.w ^ b.w) & m.w);
    a.x ^= x.x; a.y ^= x.y; a.z ^= x.z; a.w ^= x.w;
    b.x ^= x.x; b.y ^= x.y; b.z ^= x.z; b.w ^= x.w;
    spins4[i * G + g] = a; spins4[i * G + g + 1] = b;
}

} // namespace

// d_beta_group: [G] doubles on the device; d_masks: [4 G] words; d_accepted: [G] counters or nullptr
rrrmc_status_t launch_tempering_exchange(rrrmc_state *s, const double *d_beta_group, uint32_t *d_masks, long long *d_accepted,
                                         uint64_t seed, uint64_t round)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx;
    const int G = (int)(s->W / 4), parity = (int)(round & 1ull);
    const int npairs = (G - parity) / 2;
    if (npairs <= 0) return RRRMC_OK;
    RR_TRY(launch_energy_pm1(s, s->d_ibuf));
    k_pt_decide<<<npairs, 128, 0, ctx->stream>>>(s->d_ibuf, d_beta_group, G, parity, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                 (uint32_t)round, (uint32_t)(round >> 32), d_masks, d_accepted);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    k_pt_exchange<<<div_up(g->N * npairs, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint4 *>(s->d_spins), reinterpret_cast<const uint4 *>(d_masks),
                                                                       g->N, G, parity, npairs);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}
