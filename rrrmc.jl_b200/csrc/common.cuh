// Shared declarations of the engine: handles, error plumbing, launch bookkeeping.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/rrrmc_b200.h"

void rrrmc_set_error(const char *fmt, ...);

#define RR_CUDA(call)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            rrrmc_set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return RRRMC_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)
#define RR_ARG(cond, ...)                                                                     \
    do { if (!(cond)) { rrrmc_set_error(__VA_ARGS__); return RRRMC_ERR_ARG; } } while (0)
#define RR_TRY(call) do { rrrmc_status_t s__ = (call); if (s__ != RRRMC_OK) return s__; } while (0)

struct rrrmc_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sm_count = 0;
    uint64_t launches = 0;
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    uint2 *d_cbp_bucket = nullptr;   // poisson checkerboard procedure: level-1 count lookup (cb_params.cuh)
    std::vector<uint32_t> cbp_bucket_key; // the TA table the lookup was built from
};

struct rrrmc_graph {
    rrrmc_ctx *ctx = nullptr;
    int kind = 0;
    int L = 0, D = 0, twoD = 0;
    int64_t N = 0;
    bool bipartite = false;          // even L >= 4: two-colour sweeps are valid
    std::vector<int32_t> A0;         // [N*twoD] 0-based neighbours, reference slot order
    std::vector<int32_t> uA0;        // unique neighbours per site (EA.jl:158)
    std::vector<int> nuA;
    std::vector<int64_t> Ji;         // [N*twoD] (PM1/INT)
    std::vector<double> Jd;          // [N*twoD] (F64)
    std::vector<double> allDE;
    // device
    uint32_t *d_jmask = nullptr;     // PM1 lattice: the six sign bits of d_jcode as whole-word masks, [N][8]
    uint8_t *d_jcode = nullptr;      // PM1 lattice: bit 2d = J(i -> i+e_d) < 0, bit 2d+1 = J(i-e_d -> i) < 0
    int32_t *d_A = nullptr;          // [N*twoD]
    int8_t *d_J8 = nullptr;          // [N*twoD] (PM1/INT)
    double *d_Jd = nullptr;          // [N*twoD] (EA F64); [N*N] (SK F64); [Nk*Nk] (QUANT inner SK F64)
    uint8_t *d_Jb = nullptr;         // [N*N] 0/1 (SK BIN); [Nk*Nk] (QUANT inner SK BIN)
    // SK / QT / QUANT (SK.jl, QT.jl)
    int64_t Nk = 0, M = 1;           // slice size and number of Trotter slices (SK: Nk = N, M = 1)
    int inner = 0;                   // QUANT: kind of the inner graph (RRRMC_SK_F64 / RRRMC_SK_BIN / RRRMC_EMPTY)
    double fourK = 0, Gamma = 0, beta = 0, sN = 1;
    int max_deg = 0;                 // upper bound of |neighbors(X, i)|
    bool nz_neighbors = false;       // GraphRRG family: neighbors() of the integer graph lists non-zero couplings only (RRG.jl:133)
};

struct rrrmc_state {
    rrrmc_graph *g = nullptr;
    int64_t R = 0;                   // replicas
    int64_t W = 0;                   // 32-replica words per site (R padded up)
    uint32_t *d_spins = nullptr;     // multispin layout: [N][W], site-major, replica-minor
    // scratch
    uint64_t *d_chunks = nullptr;    // [R][nchunks] staging in the reference BitVector layout
    int64_t nchunks = 0;
    int32_t *d_ibuf = nullptr;       // [max(N, W*32)] integer scratch (energies, ΔE)
    int64_t ibuf_len = 0;
    long long *d_acc = nullptr;      // [W*32] accepted counters
    uint32_t *d_flips = nullptr;     // [N][W] accept masks of the last sweep (count_accepted)
    uint32_t *d_mask = nullptr;      // [W] replica mask staging
    double *d_beta = nullptr;        // [W*32] per-replica β of the continuous-coupling checkerboard kernel
    double *d_pt_beta = nullptr;     // tempering exchange (tempering.cu): [W/4] β of the 128-replica groups,
    uint32_t *d_pt_masks = nullptr;  //   [W] exchange masks of the last round, [W/4] accepted exchanges per pair since the last read
    long long *d_pt_acc = nullptr;
    bool energy_valid = false;
    // chain layout (sequential samplers): d_chunks is the spin state, one BitVector per chain
    bool ms_valid = true;            // multispin copy is current
    bool chain_valid = false;        // d_chunks copy is current
    bool chain_fields_valid = false; // the chains' local-field caches match d_chunks
    // a GraphQuant batch whose replicas sit at different β (a parallel-tempering ladder): fourK is a function of β
    // (QT.jl:165), so every replica carries its own; empty = all replicas at the graph's β
    std::vector<double> q_beta, q_fourK; double *d_q_fourK = nullptr;
    struct chain_store *chain = nullptr;
    struct sk_dense_store *skd = nullptr; // dense GraphSKNormal kernels (sk_dense.cu)
    struct cb_tma_store *tma = nullptr;   // TMA-staged checkerboard kernel: tensor maps, brick-ordered bond masks (ea_tma.cu)
};

// a pair of timing events that cannot leak: created on demand, destroyed when the scope ends
struct event_pair {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t create()
    {
        cudaError_t e = cudaEventCreate(&e0);
        return e != cudaSuccess ? e : cudaEventCreate(&e1);
    }
    ~event_pair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

static inline unsigned div_up(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }
