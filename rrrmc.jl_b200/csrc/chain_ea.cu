// GraphEA ±J fast path of the rejection-free samplers rrrMC (RRRMC.jl:149-219) and bklMC (RRRMC.jl:311-359).
//
// Same algorithm, same draw stream and same class-set order as the generic chain kernel (chain.cu:k_chain_run), which
// is the bit-for-bit restatement of DeltaE.jl:63-295 + ArraySets.jl:58-85 — trajectories are identical. What changes
// is the memory behaviour, which is what bounds a sequential chain on a GPU (every move is a chain of dependent,
// scattered accesses):
//  * compact per-chain state so that a 256-chain batch of L=32 lattices is L2 resident instead of streaming from
//    HBM: lfields as int8 (|lfields| <= 4D, EA.jl:214), class-set members and positions as uint16 (N < 65536), no
//    `lfields_last` (the one-level undo of EA.jl:231-241 restores integers exactly, so redoing the update is
//    equivalent) and no class array (the class is a function of lfields and the spin, DeltaE.jl:108-118);
//  * the loads of a move that do not depend on each other (neighbour list, couplings, the seven fields, spins and
//    set positions) are issued together, so a move costs a handful of L2 round trips instead of ~20.
// Eligibility (chain_ea_eligible): GraphEA{Int,(-1,1)} / GraphRRG{Int,(-1,1),K} with 2..6 pairwise distinct neighbours, N < 65536,
// Philox draw source. Everything else runs on the generic kernel.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include "chain.cuh"
#include "philox.cuh"

namespace {

constexpr int EA_MAXDEG = 6;   // 2D <= 6
constexpr int EA_MAXL = 4;     // |allΔE| = D+1 <= 4: classes 1..2L <= 8

__device__ __forceinline__ int sbit(const uint64_t *s, int i) { return (int)((s[i >> 6] >> (i & 63)) & 1ull); }

// lfields[x] = -2 σ_x Σ_k J_xk σ_k (EA.jl:201-215) for every chain and site
__global__ void k_ea_fields8(chain_params P)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.R * P.N) return;
    const int64_t r = tid / P.N; const int x = (int)(tid % P.N);
    const uint64_t *s = P.chunks + (P.chain0 + r) * P.nchunks;
    const int sx = 2 * sbit(s, x) - 1;
    int lf = 0;
    for (int k = 0; k < P.twoD; k++) {
        const int y = P.A[(int64_t)x * P.twoD + k];
        lf -= (int)P.J8[(int64_t)x * P.twoD + k] * sx * (2 * sbit(s, y) - 1);
    }
    P.ea_lf[(P.chain0 + r) * P.N + x] = (int8_t)(2 * lf);
}

// Draw source with look-ahead: the counter-based stream (philox.cuh:chain_rng — draw number n is Philox(n, chain))
// does not depend on the chain's state, so the three draws a move consumes are generated together at its start
// (instruction-level parallelism) instead of one after the other between dependent memory accesses.
struct ea_rng {
    chain_rng r; uint64_t q0, q1, q2; int cnt;
    __device__ void init(uint64_t seed, uint64_t chain, uint64_t n) { r.seed = seed; r.chain = chain; r.n = n; r.tag = 0; cnt = 0; }
    __device__ __forceinline__ void prefill() { while (cnt < 3) { const uint64_t v = r.u64(); if (cnt == 0) q0 = v; else if (cnt == 1) q1 = v; else q2 = v; cnt++; } }
    __device__ __forceinline__ uint64_t u64() { if (cnt == 0) return r.u64(); const uint64_t v = q0; q0 = q1; q1 = q2; cnt--; return v; }
    __device__ __forceinline__ double f64() { return (double)(u64() >> 11) * 0x1.0p-53; }
    __device__ __forceinline__ long long range(long long nn)   // rand(1:n), same procedure as chain_rng::range
    {
        const uint64_t un = (uint64_t)nn;
        for (;;) {
            const uint64_t x = u64();
            const uint64_t hi = __umul64hi(x, un), lo = x * un;
            if (lo < un) { const uint64_t t = (0 - un) % un; if (lo < t) continue; }
            return (long long)hi + 1;
        }
    }
    __device__ long long consumed() const { return (long long)r.n - cnt; }   // draws actually used
};

// per-chain view; the class weights T, sizes t live in shared memory (indexed by class at run time)
struct ea_chain {
    int N, L;
    const int32_t *A; const int8_t *J8;
    uint64_t *s; int8_t *lf; uint16_t *apos, *av;
    double ft[EA_MAXL], DE[EA_MAXL];
    double *T; int *t;
    double z;
};
// class of a site from its field and spin (DeltaE.jl:108-118 with ΔE = -lfields, EA.jl:274)
__device__ __forceinline__ int ea_class(int L, int lfv, int sb)
{
    const int dE = -lfv, a = (dE < 0 ? -dE : dE) >> 2;
    const int up = dE > 0 || (dE == 0 && sb == 1);
    return a + 1 + L * up;
}
__device__ __forceinline__ double ea_f(const ea_chain &c, int k) { return k > c.L ? c.ft[k - c.L - 1] : 1.0; }

__device__ void ea_build(ea_chain &c) // DeltaE.jl:74-104
{
    for (int k = 0; k <= 2 * c.L; k++) c.t[k] = 0;
    for (int i = 0; i < c.N; i++) {
        const int k = ea_class(c.L, c.lf[i], sbit(c.s, i));
        c.av[(int64_t)(k - 1) * c.N + c.t[k]] = (uint16_t)i;
        c.apos[i] = (uint16_t)c.t[k];
        c.t[k]++;
    }
    c.z = 0.0;
    for (int k = 1; k <= 2 * c.L; k++) { const double x = (double)c.t[k] * ea_f(c, k); c.z += x; c.T[k] = x; }
}
__device__ long long ea_rand_skip(const ea_chain &c, ea_rng &d) // DeltaE.jl:141-144
{
    return (long long)floor(log1p(-d.f64()) / log1p(-c.z / (double)c.N));
}
__device__ int ea_rand_move(const ea_chain &c, ea_rng &d, double &dE) // DeltaE.jl:146-167
{
    const int L = c.L;
    const double r = d.f64() * c.z;
    double cT = 0.0;
    int k = 1; bool broke = false;
    for (; k <= 2 * L; k++) { cT += c.T[k]; if (r < cT) { broke = true; break; } }
    if (!broke) k = 2 * L;
    if (!(r < cT)) while (c.T[k] == 0) k--;
    dE = k <= L ? -c.DE[k - 1] : c.DE[k - L - 1];
    const long long p = d.range(c.t[k]);
    return (int)c.av[(int64_t)(k - 1) * c.N + p - 1];
}

// One proposed flip of `move`: everything the class bookkeeping needs, gathered with independent loads.
// Entries 0..TWOD-1 are the neighbours in neighbors() order (EA.jl:292), entry TWOD is the moved site (last,
// DeltaE.jl:281-283); an entry is active when its class changes.
template <int TWOD> struct ea_plan {
    int site[TWOD + 1], k0[TWOD + 1], k1[TWOD + 1], pos[TWOD + 1], tail[TWOD + 1];
    bool act[TWOD + 1];
    int newlf[TWOD], lfm;
    double zp;
};
template <int TWOD>
__device__ __forceinline__ void ea_make_plan(const ea_chain &c, int move, ea_plan<TWOD> &pl)
{
    int J[TWOD], lfy[TWOD], sy[TWOD];
#pragma unroll
    for (int q = 0; q < TWOD; q++) { pl.site[q] = c.A[(int64_t)move * TWOD + q]; J[q] = c.J8[(int64_t)move * TWOD + q]; }
    const int lfm = c.lf[move], sm = sbit(c.s, move);
    pl.pos[TWOD] = c.apos[move];
#pragma unroll
    for (int q = 0; q < TWOD; q++) { lfy[q] = c.lf[pl.site[q]]; sy[q] = sbit(c.s, pl.site[q]); pl.pos[q] = c.apos[pl.site[q]]; }
    const int sx = sm ^ 1;       // the spin after the flip
    pl.lfm = lfm;
#pragma unroll
    for (int q = 0; q < TWOD; q++) {   // update rule EA.jl:248-259
        pl.newlf[q] = lfy[q] - 4 * (1 - 2 * (sx ^ sy[q])) * J[q];
        pl.k0[q] = ea_class(c.L, lfy[q], sy[q]); pl.k1[q] = ea_class(c.L, pl.newlf[q], sy[q]);
        pl.act[q] = pl.k0[q] != pl.k1[q];
    }
    // the moved site changes between the down and the up half (DeltaE.jl:224-226)
    pl.site[TWOD] = move; pl.k0[TWOD] = ea_class(c.L, lfm, sm);
    pl.k1[TWOD] = pl.k0[TWOD] > c.L ? pl.k0[TWOD] - c.L : pl.k0[TWOD] + c.L; pl.act[TWOD] = true;
    // speculative loads of the members that will fill the holes (ArraySets.jl:70-79): entry a deletes from class
    // k0[a] after `d` earlier deletions from that class, so — if nothing else interferes — its `last` is the
    // member d places before the current end. ea_commit validates each guess and reloads when it does not hold.
#pragma unroll
    for (int a = 0; a <= TWOD; a++) {
        int d = 0;
#pragma unroll
        for (int b = 0; b < a; b++) d += (pl.act[b] && pl.k0[b] == pl.k0[a]) ? 1 : 0;
        const int idx = c.t[pl.k0[a]] - 1 - d;
        pl.tail[a] = (pl.act[a] && idx >= 0) ? (int)c.av[(int64_t)(pl.k0[a] - 1) * c.N + idx] : -1;
    }
    // z' in the reference's summation order (DeltaE.jl:184-200, :248-283)
    double zp = c.z;
#pragma unroll
    for (int a = 0; a <= TWOD; a++)
        if (pl.act[a]) zp += ea_f(c, pl.k1[a]) - ea_f(c, pl.k0[a]);
    pl.zp = zp;
}
// flip the spin, update the fields (EA.jl:224-264) and move the planned sites between class sets (ArraySets.jl:58-85)
template <int TWOD>
__device__ __forceinline__ void ea_commit(ea_chain &c, int move, ea_plan<TWOD> &pl)
{
    c.s[move >> 6] ^= 1ull << (move & 63);
#pragma unroll
    for (int q = 0; q < TWOD; q++) c.lf[pl.site[q]] = (int8_t)pl.newlf[q];
    c.lf[move] = (int8_t)(-pl.lfm);
    int hole[TWOD + 1];
#pragma unroll
    for (int a = 0; a <= TWOD; a++) {
        hole[a] = -1;
        if (!pl.act[a]) continue;
        const int j = pl.site[a], k0 = pl.k0[a], k1 = pl.k1[a], p = pl.pos[a];
        const double f0 = ea_f(c, k0), f1 = ea_f(c, k1);
        c.T[k0] -= f0; c.T[k1] += f1;
        // delete!(ascache[k0], j): the last member fills the hole
        uint16_t *v0 = c.av + (int64_t)(k0 - 1) * c.N;
        const int idx = c.t[k0] - 1;
        // the speculative load is the truth iff no earlier entry pushed into this class and no earlier hole of this
        // class sits at idx; if the latest earlier operation on the class was a push, the pushed site is the last
        bool pushed = false, spoiled = false; int lastpush = -1;
#pragma unroll
        for (int b = 0; b < a; b++) {
            if (!pl.act[b]) continue;
            if (pl.k1[b] == k0) { pushed = true; lastpush = pl.site[b]; }
            if (pl.k0[b] == k0) { lastpush = -1; if (hole[b] == idx) spoiled = true; }
        }
        int last;
        if (!pushed && !spoiled) last = pl.tail[a];
        else if (lastpush >= 0) last = lastpush;
        else last = v0[idx];
        v0[p] = (uint16_t)last;
        c.apos[last] = (uint16_t)p;
        c.t[k0] = idx;
        hole[a] = p;
#pragma unroll
        for (int b = a + 1; b <= TWOD; b++) if (pl.site[b] == last) pl.pos[b] = p;   // its position was read before this move
        // push!(ascache[k1], j)
        const int e = c.t[k1];
        c.av[(int64_t)(k1 - 1) * c.N + e] = (uint16_t)j;
        c.apos[j] = (uint16_t)e;
        c.t[k1] = e + 1;
    }
    c.z = pl.zp;
}

// scalar part of chain_hdr kept in registers for the whole launch (the header's arrays are indexed at run time, which
// would otherwise push the whole struct — and every counter update — to local memory)
struct ea_hdr { double E, acc_rate, pdE; long long it, accepted, staged_its, nextstep, skip; int pending, pmove; };

template <int TWOD>
__global__ void __launch_bounds__(32) k_chain_ea(chain_params P)
{
    __shared__ double sT[2 * EA_MAXL + 1];
    __shared__ int st[2 * EA_MAXL + 1];
    if (threadIdx.x != 0) return;
    const int64_t r = P.chain0 + blockIdx.x;
    chain_hdr &H = P.hdr[r];
    if (H.done) return;
    ea_hdr h;
    h.E = H.E; h.acc_rate = H.acc_rate; h.pdE = H.pdE; h.it = H.it; h.accepted = H.accepted; h.staged_its = H.staged_its;
    h.nextstep = H.nextstep; h.skip = H.skip; h.pending = H.pending; h.pmove = H.pmove;
    const int N = P.N;
    const double beta = P.beta[r];
    ea_chain c;
    c.N = N; c.L = P.nDE; c.A = P.A; c.J8 = P.J8; c.T = sT; c.t = st;
    c.s = P.chunks + r * P.nchunks; c.lf = P.ea_lf + r * N; c.apos = P.ea_apos + r * N; c.av = P.ea_av + r * (int64_t)(2 * P.nDE) * N;
    for (int k = 0; k < c.L; k++) { c.DE[k] = P.DE[k]; c.ft[k] = exp(-beta * P.DE[k]); }
    for (int k = 0; k <= 2 * c.L; k++) { c.T[k] = H.T[k]; c.t[k] = H.t[k]; }
    c.z = H.z;
    if (!H.built) { ea_build(c); H.built = 1; }
    ea_rng src; src.init(P.seed, (uint64_t)r, (uint64_t)H.rng_n);
    long long emitted = 0;
    const long long iters = P.iters, step = P.step;
    double *Es = P.Es;
    ea_plan<TWOD> pl;
    bool done = false;
#define EMIT_SAMPLE()                                                         \
    do {                                                                      \
        if (Es && emitted < P.Es_rows) Es[emitted * P.R + (r - P.chain0)] = h.E; \
        emitted++;                                                            \
    } while (0)
    if (P.sampler == CHAIN_RRR) { // RRRMC.jl:180-211
        const double lambda = P.staged_thr_fact / (double)N;
        long long to_sample = step - h.it % step;   // iterations until `it % step == 0` (RRRMC.jl:104)
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { done = true; break; }
                h.it++;
                if (--to_sample == 0) { to_sample = step; EMIT_SAMPLE(); if (emitted >= P.quota) { h.pending = 1; break; } }
            }
            h.pending = 0;
            int acc = 0;
            double dE0;
            src.prefill();
            const double z = c.z;
            const int move = ea_rand_move(c, src, dE0);
            ea_make_plan<TWOD>(c, move, pl);
            if (h.acc_rate < P.staged_thr) {   // staged: nothing changes unless the move is accepted
                h.staged_its++;
                if (src.f64() < z / pl.zp) { ea_commit<TWOD>(c, move, pl); h.E += dE0; h.accepted++; acc = 1; }
            } else {                           // eager: apply, and apply again to undo on rejection
                ea_commit<TWOD>(c, move, pl);
                if (src.f64() < z / c.z) { h.E += dE0; h.accepted++; acc = 1; }
                else { ea_make_plan<TWOD>(c, move, pl); ea_commit<TWOD>(c, move, pl); }
            }
            h.acc_rate = h.acc_rate * (1 - lambda) + acc * lambda;
        }
    } else {                      // bklMC, RRRMC.jl:332-350
        for (;;) {
            if (!h.pending) {
                if (h.it >= iters) { done = true; break; }
                src.prefill();
                h.skip = ea_rand_skip(c, src);
                h.pmove = ea_rand_move(c, src, h.pdE);
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
            ea_make_plan<TWOD>(c, h.pmove, pl);
            ea_commit<TWOD>(c, h.pmove, pl);
            h.it += h.skip + 1;
            h.E += h.pdE;
            h.accepted++;
            h.pending = 0;
        }
    }
#undef EMIT_SAMPLE
    for (int k = 0; k <= 2 * c.L; k++) { H.T[k] = c.T[k]; H.t[k] = c.t[k]; }
    H.z = c.z;
    H.rng_n = src.consumed();
    H.E = h.E; H.acc_rate = h.acc_rate; H.pdE = h.pdE; H.it = h.it; H.accepted = h.accepted; H.staged_its = h.staged_its;
    H.nextstep = h.nextstep; H.skip = h.skip; H.pending = h.pending; H.pmove = h.pmove;
    if (done) H.done = 1;
}

} // namespace

bool chain_ea_eligible(const rrrmc_state *s, int sampler)
{
    const rrrmc_graph *g = s->g;
    if (getenv("RRRMC_CHAIN_GENERIC")) return false;   // tests: force the generic kernel
    if (!(sampler == CHAIN_RRR || sampler == CHAIN_BKL)) return false;
    if (g->kind != RRRMC_EA_PM1 || g->twoD < 2 || g->twoD > EA_MAXDEG || g->N >= 65536 || (int)g->allDE.size() > EA_MAXL) return false;
    for (int64_t i = 0; i < g->N; i++)                 // all neighbours distinct (L >= 3): uA == A (EA.jl:158)
        for (int k = 0; k + 1 < g->twoD; k++)
            if (g->A0[i * g->twoD + k] == g->A0[i * g->twoD + k + 1]) return false;
    return true;
}

rrrmc_status_t chain_ea_prepare(rrrmc_state *s, chain_params &P)
{
    rrrmc_graph *g = s->g; rrrmc_ctx *ctx = g->ctx; chain_store *c = s->chain;
    const size_t RN = (size_t)s->R * g->N;
    if (!c->ea_lf) {
        RR_CUDA(cudaMalloc(&c->ea_lf, RN));
        RR_CUDA(cudaMalloc(&c->ea_apos, RN * 2));
        RR_CUDA(cudaMalloc(&c->ea_av, RN * 2 * 2 * c->nDE));
    }
    P.fast = 1; P.ea_lf = c->ea_lf; P.ea_apos = c->ea_apos; P.ea_av = c->ea_av;
    k_ea_fields8<<<div_up(P.R * P.N, 256), 256, 0, ctx->stream>>>(P);
    ctx->launches++;
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}

rrrmc_status_t chain_ea_launch(rrrmc_state *s, const chain_params &P)
{
    cudaStream_t st = s->g->ctx->stream;
    if (P.twoD == 6) k_chain_ea<6><<<(unsigned)P.R, 32, 0, st>>>(P);
    else if (P.twoD == 5) k_chain_ea<5><<<(unsigned)P.R, 32, 0, st>>>(P);      // odd degrees: GraphRRG (RRG.jl), ΔE ∈ {2, 6, ..}
    else if (P.twoD == 4) k_chain_ea<4><<<(unsigned)P.R, 32, 0, st>>>(P);
    else if (P.twoD == 3) k_chain_ea<3><<<(unsigned)P.R, 32, 0, st>>>(P);
    else k_chain_ea<2><<<(unsigned)P.R, 32, 0, st>>>(P);
    RR_CUDA(cudaGetLastError());
    return RRRMC_OK;
}
