// Launch parameters of the checkerboard Metropolis kernels (ea_multispin.cu), filled by api.cu.
#pragma once
#include "common.cuh"

constexpr int CB_MAXK = 32;

struct cb_params {
    uint32_t *spins;        // [N][W]
    uint32_t *flips;        // [N][W] or nullptr: per-attempt accept masks (for accepted counters)
    const uint8_t *jcode;   // [N]
    int L, Lh, W, G;        // lattice side, L/2, words per site, 128-replica groups per site
    uint32_t k0, k1;        // Philox key = seed
    uint32_t t_lo, t_hi16;  // sweep counter: low 32 bits, (high bits) << 16
    int K;                  // full bit planes (one Philox call per plane)
    int Ku;                 // leading planes whose threshold bit is class-independent (spin-independent part), <= K
    int Kz;                 // leading planes whose threshold bit is 0 for every class, <= Ku
    uint32_t rk[10][2];     // Philox round keys: key + r*(0x9E3779B9, 0xBB67AE85)
    int M;                  // merged planes after the full ones (one Philox call per four planes)
    float invG;             // 1/G
    int variant;            // occupancy variant (tuning)
    uint32_t zero;          // always 0 (opaque to ptxas: orders spin-dependent work after the first planes)
    uint8_t planeop[CB_MAXK]; // q < K+M. 0: threshold bit 0 for all classes, 1: bit 1 for all classes, 2: mixed
    uint32_t plane[CB_MAXK][3]; // plane[q][c-1] = all-ones iff bit (63-q) of thr64[c] is set
    uint32_t rem[3];        // bits [63-K .. 32-K] of thr64[c]
    uint32_t remM[3];       // bits [63-K-M .. 32-K-M] of thr64[c]
};

// "sparse" acceptance procedure: binomial counts by inverse CDF + uniform distinct positions
constexpr int CBS_T1 = 33;   // class 1 (ΔE=4): Bin(32, p1) per 32-lane word, entries k = 0..32
constexpr int CBS_TC = 129;  // classes 2..D: Bin(128, pc) per 128-lane task, entries k = 0..128
struct cbs_params {
    uint32_t *spins;        // [N][W]
    uint32_t *flips;        // [N][W] or nullptr
    const uint8_t *jcode;   // [N]
    const uint4 *jmask;     // [N][2]: whole-word sign masks of the six bonds of a site (api.cu)
    int L, Lh, W, G;
    int tpr;                // row-chunk kernel: threads per lattice row = (Lh/T)*G (set by the launcher)
    uint32_t t_lo, t_hi16;
    uint32_t rk[10][2];     // Philox round keys
    float invG;
    int Gshift;             // log2(G) when G is a power of two, else -1
    int variant;
    uint32_t tbl[CBS_T1 + 2 * CBS_TC]; // more than k lanes pass iff x > tbl[k]
};

// "poisson" acceptance procedure (ea_poisson.cu): Poisson hit counts by inverse CDF + uniform positions
// with replacement. tbl = TA[CBP_KA] | TB0[CBP_KR] | TB[CBP_KR] | TC[CBP_KR]: more than k hits iff x > T[k].
constexpr int CBP_KA = 64;
constexpr int CBP_KR = 32;
constexpr int CBP_LEN = CBP_KA + 3 * CBP_KR;
constexpr int CBP_BUCKETS = 1024;  // lookup on the top 10 bits of the level-1 count uniform
struct cbp_params {
    uint32_t *spins;        // [N][W]
    uint32_t *flips;        // [N][W] or nullptr
    const uint4 *jmask;     // [N][2]: whole-word sign masks of the six bonds of a site (api.cu)
    const uint2 *bucket;    // [CBP_BUCKETS] {T, a0}: for x in bucket e the level-1 count is a0 + (x > T); when two table
                            // entries fall into the bucket it holds {2^32-1, 64 + a0} (count >= a0: second tier)
    int L, Lh, W, G;
    int tpr;                // row-chunk kernel: threads per lattice row
    int NW;                 // static position words: 1, 2, 4 or 6
    int brick, sh_hbx, sh_by, bz; // brick mapping of blocks to sites (set by the launcher): log2(bx/2), log2(by), bz
    int nbx, nby, nbricks;  // bricks along x, y and in total (persistent kernel)
    float inv_nbx, inv_nby;
    uint32_t t_lo, t_hi16;
    uint32_t rk[10][2];     // Philox round keys
    float invG;
    int Gshift;             // log2(G) when G is a power of two, else -1
    int variant;
    uint32_t tb0_0, tb0_1;  // TB0[0], TB0[1]
    uint32_t tc0;           // TC[0]
    uint32_t one;           // always 1 (opaque to ptxas)
    uint32_t tbl[CBP_LEN];
};
