// Launch parameters and per-state cache of the TMA-staged checkerboard kernel (ea_tma.cu).
#pragma once
#include <cuda.h>
#include <vector>
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

// One entry of the brick table (launch order: slab-major, then brick): where the brick sits and which bricks share a
// face with it. The multi-sweep kernel waits on the neighbours' progress counters before it loads a brick.
struct alignas(16) cbf_brick {
    uint32_t site0, slab;                 // first site of the brick (the consumers read these two words), 1024-replica slab
    uint32_t b, xyz;                      // brick index inside the slab; origin packed x0 | y0 << 10 | z0 << 20
    uint32_t nbr[8];                      // launch-order ids of the six face neighbours, then the brick itself (twice)
};

// Per-group count tables of a β ladder (one β per 128-replica group), multi-sweep kernel only.
struct alignas(16) cbp_group {
    uint32_t tbl[CBP_LEN];                // TA | TB0 | TB | TC of this group's β
    uint32_t tb0_0, tb0_1, tc0, pad;
};

// Launch parameters of the multi-sweep kernel: every half-sweep of a run in ONE launch, bricks ordered by progress
// counters instead of launch boundaries.
struct alignas(64) cbf_params {
    cbt_params T;
    CUtensorMap m_y6z;                    // box (32, 1, 8 x, 6 y, 4 z): the four planes of an interior brick with their y halo
    const cbf_brick *bricks;              // [nslab * nbricks]
    uint32_t *done;                       // [nslab * nbricks] half-sweeps completed on a brick, counted from the state's creation
    uint32_t epoch0;                      // value of every done[] entry when the launch starts
    uint32_t nhalf;                       // half-sweeps of this launch
    uint64_t half0;                       // index of the first one: sweep = half >> 1, colour = half & 1
    const cbp_group *groups;              // [groups of the batch] or nullptr (one β: the tables of T.p)
    const uint2 *gbucket;                 // [groups][CBP_BUCKETS] level-1 count lookups of the ladder
};

struct cb_tma_store {
    CUtensorMap m_y6, m_y5, m_y4, m_y1, m_xf, m_y6z;
    uint4 *d_jbrick = nullptr;
    uint2 *d_origin = nullptr;
    cbf_brick *d_bricks = nullptr;
    uint32_t *d_done = nullptr;
    uint32_t epoch = 0;                   // host mirror of done[]
    bool flow_unavailable = false;        // a cooperative launch failed for lack of co-residency: use the per-colour launches
    cbp_group *d_groups = nullptr;        // β ladder tables (uploaded per run)
    uint2 *d_gbucket = nullptr;
    int ngroups_alloc = 0;
    std::vector<uint32_t> ladder_key;     // the tables d_groups / d_gbucket were built from
    int nbx = 0, nby = 0, nbricks = 0;
};

bool checkerboard_tma_eligible(const rrrmc_state *s);
void checkerboard_tma_free(rrrmc_state *s);
rrrmc_status_t checkerboard_tma_prepare(rrrmc_state *s, const cbp_params &p, cbt_params &P);
rrrmc_status_t launch_checkerboard_tma(rrrmc_ctx *ctx, cbt_params &P, int colour);
// nsweeps whole sweeps starting with sweep counter sweep0 in one launch. groups: host tables of a β ladder (ngroups =
// W/4 entries) or nullptr; gbucket: their level-1 lookups, or nullptr when the state's device copy is current.
rrrmc_status_t launch_checkerboard_flow(rrrmc_state *s, cbt_params &P, uint64_t sweep0, int64_t nsweeps,
                                        const cbp_group *groups, const uint2 *gbucket, int ngroups);
