// Warp-cooperative rrrMC (RRRMC.jl:149-219) and bklMC (RRRMC.jl:311-359) for GraphEA ±J lattices: one chain per warp,
// the whole chain state in shared memory, every step of a move spread over the lanes.
//
// The reference keeps the ΔE classes of DeltaE.jl:63-118 as ArraySets (ArraySets.jl:58-85): a move is ~20 dependent
// scattered accesses, which one GPU lane executes at ~4 µs per move (chain_ea.cu, the reference-order kernel). Here
//  * a site is ONE byte of shared memory: the three forward bond signs, its number u of unsatisfied bonds and its
//    spin. ΔE = 4(D − u), so the class (DeltaE.jl:108-118) is a function of the byte;
//  * a class is a BITMAP over the sites plus a count per 1024-site block. rand(1:t[k]) picks the p-th member in site
//    order: a warp prefix scan over the block counts, one over the 32 words of the block, and a find-nth-set-bit —
//    "per-replica class histograms, warp-shuffle prefix scans" (north_star). Any uniform pick among the members gives
//    the reference's chain law (DeltaE.jl:146-167); the TRAJECTORY for a given draw stream differs from the ArraySet
//    order, so this kernel is opt-in (rrrmc_opts_t.site_pick = RRRMC_PICK_RANK) and has its own CPU model,
//    oracle/rrrmc_oracle.c:orc_rank_rrrMC / orc_rank_bklMC, against which it is bit-exact;
//  * lanes 0..2D-1 re-file the neighbours of the moved site and lane 2D the site itself, in parallel; class c lives on
//    lane c-1: its size t, its weight T = t·f from the integer count, and the cumulative weight from ONE 8-lane scan
//    per move (the oracle's rk_scan) whose last element is z — the class pick is a ballot on r < cT;
//  * the counter-based draw stream (philox.cuh:chain_rng, draw n = Philox(n, chain)) is generated 32 draws at a time,
//    one per lane, and consumed by shuffle.
// Sampling instants, hook protocol (the kernel pauses after `quota` samples and resumes from the header) and the draw
// order are those of k_chain_ea / k_chain_run. Two instantiations: GraphEA lattices (neighbours and bond signs from the
// site index and the byte: shared memory only) and any ±J graph given by its adjacency table (GraphRRG, the benchmark
// family of the RRR paper: A and J8 through L1/L2). Eligibility: chain_warp_eligible().
#include <algorithm>
#include <cmath>
#include "chain.cuh"
#include "philox.cuh"

namespace {

constexpr int WK = 8;            // classes: 2·|allΔE| <= 8 (D <= 3)

struct warp_smem {
    uint8_t *st;                 // [N] site bytes: bits 0-2 forward bond is negative (d = 0, 1, 2), bits 3-5 u, bit 6 spin
    uint32_t *bm;                // [WK][NW32] class bitmaps
    int *cnt;                    // [WK][NB] members per 1024-site block
};

// class of a site (DeltaE.jl:108-118) from its degree K, its number u of unsatisfied bonds and its spin: ΔE = 2(K - 2u),
// allΔE = {0, 4, ..} (even K) or {2, 6, ..} (odd K), so the index of |ΔE| in allΔE is |K - 2u| >> 1
__device__ __forceinline__ int wclass(int K, int u, int sb)
{
    const int a = K - 2 * u, aa = (a < 0 ? -a : a) >> 1;
    const int up = a > 0 || (a == 0 && sb == 1);
    return aa + 1 + ((K >> 1) + 1) * up;
}

// 32 draws of the chain's stream at a time: lane l holds draw number base + l
struct warp_draws {
    uint64_t seed, chain, base; uint64_t mine; int used;
    __device__ __forceinline__ void fill(int lane)
    {
        const uint64_t n = base + (uint64_t)lane;
        const philox_out o = philox4x32_10((uint32_t)n, (uint32_t)(n >> 32), (uint32_t)chain, (uint32_t)(chain >> 32),
                                           (uint32_t)seed, (uint32_t)(seed >> 32));
        mine = ((uint64_t)o.y << 32) | o.x;
        used = 0;
    }
    __device__ __forceinline__ void init(uint64_t seed_, uint64_t chain_, uint64_t n, int lane) { seed = seed_; chain = chain_; base = n; fill(lane); }
    __device__ __forceinline__ uint64_t u64(int lane)
    {
        if (used == 32) { base += 32; fill(lane); }
        const uint64_t v = __shfl_sync(FULLMASK, mine, used);
        used++;
        return v;
    }
    __device__ __forceinline__ double f64(int lane) { return (double)(u64(lane) >> 11) * 0x1.0p-53; }
    __device__ __forceinline__ long long range(long long nn, int lane)   // rand(1:n), the procedure of chain_rng::range
    {
        const uint64_t un = (uint64_t)nn;
        for (;;) {
            const uint64_t x = u64(lane);
            const uint64_t hi = __umul64hi(x, un), lo = x * un;
            if (lo < un) { const uint64_t t = (0 - un) % un; if (lo < t) continue; }
            return (long long)hi + 1;
        }
    }
    __device__ long long consumed() const { return (long long)base + used; }
};

// inclusive prefix sum over the lanes of a value below 2^BITS: one ballot per bit plane (independent of each other, so
// the latency is one ballot instead of five dependent shuffles)
template <int BITS> __device__ __forceinline__ int warp_incl_scan(int v, int lane)
{
    const unsigned le = 0xffffffffu >> (31 - lane);
    int s = 0;
#pragma unroll
    for (int b = 0; b < BITS; b++) s += __popc(__ballot_sync(FULLMASK, (v >> b) & 1) & le) << b;
    return s;
}
// position of the p-th (1-based) set bit of x, p <= popc(x): binary search on the popcounts of the halves
__device__ __forceinline__ int nth_set_bit(uint32_t x, int p)
{
    int pos = 0;
#pragma unroll
    for (int w = 16; w >= 1; w >>= 1) {
        const int c = __popc(x & ((1u << w) - 1u));
        const bool hi = p > c;
        if (hi) { p -= c; x >>= w; pos += w; }
    }
    return pos;
}

// the p-th (1-based) member of a class in site order
__device__ __forceinline__ int rank_select(const uint32_t *bmk, const int *cntk, int NB, int NW32, int p, int lane)
{
    int base = 0, blk = 0;
    for (int b0 = 0; b0 < NB; b0 += 32) {
        const int c = b0 + lane < NB ? cntk[b0 + lane] : 0;
        const int incl = warp_incl_scan<11>(c, lane);              // a block holds at most 1024 members
        const unsigned m = __ballot_sync(FULLMASK, base + incl >= p);
        if (m) {
            const int l = __ffs(m) - 1;
            blk = b0 + l;
            p -= base + __shfl_sync(FULLMASK, incl - c, l);
            break;
        }
        base += __shfl_sync(FULLMASK, incl, 31);
    }
    const int w = blk * 32 + lane;
    const uint32_t x = w < NW32 ? bmk[w] : 0u;
    const int c = __popc(x), incl = warp_incl_scan<6>(c, lane);
    const unsigned m = __ballot_sync(FULLMASK, incl >= p);
    const int l = __ffs(m) - 1;
    p -= __shfl_sync(FULLMASK, incl - c, l);
    const uint32_t word = __shfl_sync(FULLMASK, x, l);
    return (blk * 32 + l) * 32 + nth_set_bit(word, p);
}

struct warp_hdr { double E, pdE; long long it, accepted, nextstep, skip; int pending, pmove; };

// LAT: a GraphEA lattice — neighbours and bond signs come from the site index and the byte's forward-bond bits (shared
// memory only); else any ±J graph of degree TWOD given by its adjacency table (GraphRRG: A and J8 through L1/L2)
template <int TWOD, bool LAT>
__global__ void __launch_bounds__(32) k_chain_warp(chain_params P)
{
    constexpr int D = TWOD / 2, LC = (TWOD >> 1) + 1, NK = 2 * LC;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x;
    const int64_t r = P.chain0 + blockIdx.x;
    chain_hdr &H = P.hdr[r];
    if (H.done) return;
    const int N = P.N, L = P.latL, NW32 = (N + 31) >> 5, NB = (NW32 + 31) >> 5;
    warp_smem S;
    S.cnt = reinterpret_cast<int *>(smem);
    S.bm = reinterpret_cast<uint32_t *>(S.cnt + WK * NB);
    S.st = reinterpret_cast<uint8_t *>(S.bm + (size_t)WK * NW32);
    uint64_t *chunks = P.chunks + r * P.nchunks;
    const double beta = P.beta[r];

    // ---- build the chain state from the configuration (every launch: the configuration is all that is kept) ----
    for (int i = lane; i < N; i += 32) {
        const uint8_t jc = LAT ? P.jcode[i] : (uint8_t)0;
        const int sb = (int)((chunks[i >> 6] >> (i & 63)) & 1ull);
        S.st[i] = (uint8_t)((jc & 1) | ((jc >> 1) & 2) | ((jc >> 2) & 4) | (sb << 6));
    }
    for (int k = lane; k < WK * NW32; k += 32) S.bm[k] = 0u;
    for (int k = lane; k < WK * NB; k += 32) S.cnt[k] = 0;
    __syncwarp();
    // u of every site; lane owns the sites of whole 32-site words, so the bitmaps need no atomics
    for (int w = lane; w < NW32; w += 32) {
        uint32_t wk[NK];
#pragma unroll
        for (int k = 0; k < NK; k++) wk[k] = 0u;
        for (int b = 0; b < 32; b++) {
            const int i = 32 * w + b;
            if (i >= N) break;
            const int bi = S.st[i], si = (bi >> 6) & 1;
            int u = 0;
            if (LAT) {
                int co[3], rem = i, stride = 1;
#pragma unroll
                for (int d = 0; d < D; d++) { co[d] = rem % L; rem /= L; }
#pragma unroll
                for (int d = 0; d < D; d++) {
                    const int up = i + ((co[d] + 1 == L ? 0 : co[d] + 1) - co[d]) * stride;
                    const int dn = i + ((co[d] == 0 ? L - 1 : co[d] - 1) - co[d]) * stride;
                    const int bu = S.st[up], bd = S.st[dn];
                    u += (si ^ ((bu >> 6) & 1)) ^ ((bi >> d) & 1);
                    u += (si ^ ((bd >> 6) & 1)) ^ ((bd >> d) & 1);
                    stride *= L;
                }
            } else {
#pragma unroll
                for (int q = 0; q < TWOD; q++) {
                    const int y = P.A[(int64_t)i * TWOD + q];
                    u += (si ^ ((S.st[y] >> 6) & 1)) ^ (P.J8[(int64_t)i * TWOD + q] < 0 ? 1 : 0);
                }
            }
            // (the other lanes only read bits 0-2 and 6 of this byte during the pass)
            S.st[i] = (uint8_t)(bi | (u << 3));
            const int k = wclass(TWOD, u, si);
#pragma unroll
            for (int kk = 0; kk < NK; kk++) if (kk + 1 == k) wk[kk] |= 1u << b;
        }
#pragma unroll
        for (int k = 0; k < NK; k++) {
            S.bm[k * NW32 + w] = wk[k];
            if (wk[k]) atomicAdd(&S.cnt[k * NB + (w >> 5)], __popc(wk[k]));
        }
    }
    __syncwarp();
    // class c + 1 lives on lane c: its size, its f (1 for the down half, exp(-β ΔE) for the up half, DeltaE.jl:83-95)
    // and the cumulative weight cT of classes 1..c+1
    int myt = 0;
    if (lane < NK) for (int b = 0; b < NB; b++) myt += S.cnt[lane * NB + b];
    const int myk = lane + 1;
    const double myf = (lane < NK && myk > LC) ? exp(-beta * P.DE[myk - LC - 1]) : 1.0;
    // inclusive scan over lanes 0..7 (steps 1, 2, 4): the oracle's rk_scan; -> z = the scan's last element
    auto scan8 = [&](double &x) {
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { const double y = __shfl_up_sync(FULLMASK, x, o); if (lane >= o) x = x + y; }
        return __shfl_sync(FULLMASK, x, 7);
    };
    double mycT = lane < NK ? (double)myt * myf : 0.0;
    double z = scan8(mycT);
    const bool pow2 = (L & (L - 1)) == 0;
    const int lsh = 31 - __clz(L);

    warp_hdr h;
    h.E = H.E; h.pdE = H.pdE; h.it = H.it; h.accepted = H.accepted; h.nextstep = H.nextstep; h.skip = H.skip;
    h.pending = H.pending; h.pmove = H.pmove;
    warp_draws src; src.init(P.seed, (uint64_t)r, (uint64_t)H.rng_n, lane);
    long long emitted = 0;
    const long long iters = P.iters, step = P.step;
    double *Es = P.Es;
    bool done = false;

    // class pick of rand_move (DeltaE.jl:146-167) on the cumulative weights, then the rank query: -> site, ΔE
    auto rand_move = [&](double &dE) {
        const double rr = src.f64(lane) * z;
        const unsigned hit = __ballot_sync(FULLMASK, lane < NK && rr < mycT);
        int k;
        if (hit) k = __ffs(hit);
        else k = 32 - __clz(__ballot_sync(FULLMASK, lane < NK && myt > 0));   // rounding: the last non-empty class
        dE = k <= LC ? -P.DE[k - 1] : P.DE[k - LC - 1];
        const int tk = __shfl_sync(FULLMASK, myt, k - 1);
        const int p = (int)src.range(tk, lane);
        return rank_select(S.bm + (size_t)(k - 1) * NW32, S.cnt + (k - 1) * NB, NB, NW32, p, lane);
    };
    // the re-filing of a flip of `move`, one entry per lane (lanes 0..2D-1 the neighbours, lane 2D the site); -> z'
    int ej = 0, ek0 = 0, ek1 = 0, ebyte = 0, mytp = 0; double mycTp = 0.0;
    auto plan = [&](int move) {
        const int bm_ = S.st[move], sm = (bm_ >> 6) & 1;
        int kk = 0;                                   // k0 | k1 << 4 of this lane's entry (0: no entry)
        if (lane < TWOD) {
            int neg;
            if (LAT) {
                const int d = lane >> 1, dir = lane & 1;
                int c, stride;
                if (pow2) { stride = 1 << (d * lsh); c = (move >> (d * lsh)) & (L - 1); }
                else {
                    int rem = move; stride = 1; c = 0;
                    for (int q = 0; q <= d; q++) { c = rem % L; rem /= L; if (q < d) stride *= L; }
                }
                const int cn = dir == 0 ? (c + 1 == L ? 0 : c + 1) : (c == 0 ? L - 1 : c - 1);
                ej = move + (cn - c) * stride;
                neg = dir == 0 ? (bm_ >> d) & 1 : (S.st[ej] >> d) & 1;
            } else {
                ej = P.A[(int64_t)move * TWOD + lane];
                neg = P.J8[(int64_t)move * TWOD + lane] < 0 ? 1 : 0;
            }
            const int by = S.st[ej], sy = (by >> 6) & 1, uy = (by >> 3) & 7;
            const int unsat = (sm ^ sy) ^ neg;
            const int u1 = uy + (unsat ? -1 : 1);
            ek0 = wclass(TWOD, uy, sy); ek1 = wclass(TWOD, u1, sy);
            ebyte = (by & ~(7 << 3)) | (u1 << 3);
            kk = ek0 | ek1 << 4;
        } else if (lane == TWOD) {
            const int um = (bm_ >> 3) & 7, u1 = TWOD - um;
            ej = move; ek0 = wclass(TWOD, um, sm); ek1 = wclass(TWOD, u1, sm ^ 1);
            ebyte = ((bm_ & 7) | (u1 << 3) | ((sm ^ 1) << 6));
            kk = ek0 | ek1 << 4;
        }
        // size of this lane's class after the flip: +1 per entry that arrives, -1 per entry that leaves
        int dl = 0;
#pragma unroll
        for (int e = 0; e <= TWOD; e++) {
            const int v = __shfl_sync(FULLMASK, kk, e);
            dl += ((v >> 4) == myk ? 1 : 0) - ((v & 15) == myk ? 1 : 0);
        }
        mytp = myt + dl;
        mycTp = lane < NK ? (double)mytp * myf : 0.0;
        return scan8(mycTp);
    };
    auto commit = [&](double zp) {
        if (lane <= TWOD) {
            S.st[ej] = (uint8_t)ebyte;
            const uint32_t bit = 1u << (ej & 31);
            atomicAnd(&S.bm[(size_t)(ek0 - 1) * NW32 + (ej >> 5)], ~bit);
            atomicOr(&S.bm[(size_t)(ek1 - 1) * NW32 + (ej >> 5)], bit);
            atomicAdd(&S.cnt[(ek0 - 1) * NB + (ej >> 10)], -1);
            atomicAdd(&S.cnt[(ek1 - 1) * NB + (ej >> 10)], 1);
        }
        myt = mytp; mycT = mycTp; z = zp;
        __syncwarp();
    };
#define EMIT_SAMPLE()                                                                        \
    do {                                                                                     \
        if (lane == 0 && Es && emitted < P.Es_rows) Es[emitted * P.R + (r - P.chain0)] = h.E; \
        emitted++;                                                                           \
    } while (0)

    if (P.sampler == CHAIN_RRR) { // RRRMC.jl:180-211 (always the staged form: nothing changes unless the move is accepted)
        long long to_sample = step - h.it % step;
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { done = true; break; }
                h.it++;
                if (--to_sample == 0) { to_sample = step; EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 1; break; } }
            }
            h.pending = 0;
            double dE0;
            const double z0 = z;
            const int move = rand_move(dE0);
            const double zp = plan(move);
            if (src.f64(lane) * zp < z0) { commit(zp); h.E += dE0; h.accepted++; }   // rand() < z/z' (RRRMC.jl:131-138) without the division
        }
    } else {                      // bklMC, RRRMC.jl:332-350
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { done = true; break; }
                // DeltaE.jl:141-144: floor(log1p(-rand()) / log1p(-z/N)); the two logarithms in ONE call, on two lanes
                const double us = src.f64(lane);
                const double ly = log1p(lane == 0 ? -us : -z / (double)N);
                h.skip = (long long)floor(__shfl_sync(FULLMASK, ly, 0) / __shfl_sync(FULLMASK, ly, 1));
                h.pmove = rand_move(h.pdE);
                h.pending = 1;
            }
            bool out = false, paused = false;
            while (h.it + h.skip + 1 >= h.nextstep) {
                if (h.pending == 2) h.pending = 1; // resuming right after the hook of this sample
                else { EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 2; paused = true; break; } }
                h.nextstep += step;
                if (h.nextstep > iters) { out = true; break; }
            }
            if (paused) break;
            if (out) { done = true; break; }
            const double zp = plan(h.pmove);
            commit(zp);
            h.it += h.skip + 1;
            h.E += h.pdE;
            h.accepted++;
            h.pending = 0;
        }
    }
#undef EMIT_SAMPLE
    // ---- write the configuration back ----
    __syncwarp();
    for (int64_t c = lane; c < P.nchunks; c += 32) {
        uint64_t v = 0;
        for (int b = 0; b < 64; b++) {
            const int64_t i = 64 * c + b;
            if (i < N) v |= (uint64_t)((S.st[i] >> 6) & 1) << b;
        }
        chunks[c] = v;
    }
    if (lane == 0) {
        H.rng_n = src.consumed();
        H.E = h.E; H.pdE = h.pdE; H.it = h.it; H.accepted = h.accepted; H.staged_its = h.it;
        H.nextstep = h.nextstep; H.skip = h.skip; H.pending = h.pending; H.pmove = h.pmove; H.z = z;
        H.built = 1;
        if (done) H.done = 1;
    }
}

size_t warp_smem_bytes(int N)
{
    const int NW32 = (N + 31) >> 5, NB = (NW32 + 31) >> 5;
    return 4 * (size_t)(WK * NB) + 4 * (size_t)WK * NW32 + (size_t)N + 16;
}

} // namespace

// lattice path: GraphEA ±J with L >= 3 (neighbours from the site index); graph path: any ±J GraphEA / GraphRRG whose rows
// hold 2..6 pairwise distinct neighbours and no zero coupling
static bool warp_is_lattice(const rrrmc_graph *g) { return g->d_jcode && g->L >= 3 && g->D >= 1 && g->D <= 3; }
bool chain_warp_eligible(const rrrmc_state *s, int sampler)
{
    const rrrmc_graph *g = s->g;
    if (!(sampler == CHAIN_RRR || sampler == CHAIN_BKL)) return false;
    if (g->kind != RRRMC_EA_PM1 || g->N >= ((int64_t)1 << 24)) return false;
    if (warp_smem_bytes((int)g->N) > 227 * 1024 - 1024) return false;
    if (warp_is_lattice(g)) return true;
    if (g->twoD < 2 || g->twoD > 6) return false;
    for (int64_t i = 0; i < g->N; i++)
        for (int k = 0; k < g->twoD; k++) {
            if (g->Ji[i * g->twoD + k] != 1 && g->Ji[i * g->twoD + k] != -1) return false;
            for (int q = 0; q < k; q++) if (g->A0[i * g->twoD + q] == g->A0[i * g->twoD + k]) return false;
        }
    return true;
}

rrrmc_status_t chain_warp_launch(rrrmc_state *s, const chain_params &P)
{
    const rrrmc_graph *g = s->g;
    cudaStream_t st = g->ctx->stream;
    const size_t sm = warp_smem_bytes(P.N);
    const bool lat = warp_is_lattice(g);
    static size_t configured[2][8] = {};
    auto go = [&](auto kern) -> cudaError_t {
        if (configured[lat][g->twoD] < sm) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return e;
            configured[lat][g->twoD] = sm;
        }
        kern<<<(unsigned)P.R, 32, sm, st>>>(P);
        return cudaGetLastError();
    };
    cudaError_t e;
    if (lat) e = g->D == 3 ? go(k_chain_warp<6, true>) : (g->D == 2 ? go(k_chain_warp<4, true>) : go(k_chain_warp<2, true>));
    else switch (g->twoD) {
        case 2: e = go(k_chain_warp<2, false>); break;
        case 3: e = go(k_chain_warp<3, false>); break;
        case 4: e = go(k_chain_warp<4, false>); break;
        case 5: e = go(k_chain_warp<5, false>); break;
        default: e = go(k_chain_warp<6, false>); break;
    }
    RR_CUDA(e);
    return RRRMC_OK;
}
