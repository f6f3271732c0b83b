// Launch parameters and per-state cache of the TMA-staged checkerboard kernel (ea_tma.cu).
#pragma once
#include <cuda.h>
#include "cb_params.cuh"

constexpr int CBT_BX = 8, CBT_BY = 4, CBT_BZ = 4;            // brick of sites handled per pipeline stage
constexpr int CBT_NACT = CBT_BX * CBT_BY * CBT_BZ / 2;       // active sites of one colour in a brick

struct alignas(64) cbt_params {
    CUtensorMap m_y6, m_y5, m_y4, m_y1;   // boxes (32 words, 1 slab, 8 x, {6,5,4,1} y, 1 z) over spins[z][y][x][slab][32]
    CUtensorMap m_xf;                     // box (32, 1, 1 x, 4 y, 4 z): an x face of a brick
    cbp_params p;                         // the poisson procedure's parameters (tables, Philox keys, spins, flips)
    const uint4 *jbrick;                  // [2 colours][nbricks][64 slots][2]: bond masks of the active sites, brick order
    const uint2 *origin;                  // [nslab * nbricks] {first site of the brick, slab} in launch order
    int nbx, nby, nbricks, nslab;         // bricks along x, y, per slab; 1024-replica slabs
    float inv_nbx, inv_nby, inv_nbricks;
};

struct cb_tma_store {
    CUtensorMap m_y6, m_y5, m_y4, m_y1, m_xf;
    uint4 *d_jbrick = nullptr;
    uint2 *d_origin = nullptr;
    int nbx = 0, nby = 0, nbricks = 0;
};

bool checkerboard_tma_eligible(const rrrmc_state *s);
void checkerboard_tma_free(rrrmc_state *s);
rrrmc_status_t checkerboard_tma_prepare(rrrmc_state *s, const cbp_params &p, cbt_params &P);
rrrmc_status_t launch_checkerboard_tma(rrrmc_ctx *ctx, cbt_params &P, int colour);
