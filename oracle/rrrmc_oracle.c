/* rrrmc_oracle.c — CPU restatement of RRRMC.jl's single-spin-flip hot path (plain C11).
 *
 * TEST INFRASTRUCTURE ONLY (see rrrmc_oracle.h).  PARITY UNPINNED by reference golden
 * vectors (none exist; Julia is not runnable here) — pinned by reference invariants instead.
 *
 * Each block cites the reference file:line it restates.  Site indices are 1-based at this
 * API (like the reference); spin s_i is bit (i-1)&63 of chunk (i-1)>>6 (src/Common.jl:15-22,
 * src/Interface.jl:21-29), σ = 2s-1.
 */
#include "rrrmc_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ------------------------------------------------------------------------------------------
 * Config bits  (src/Interface.jl:21-43, src/Common.jl:15-22)
 * ---------------------------------------------------------------------------------------- */
static inline int cfg_get(const uint64_t *c, int64_t i) { return (int)((c[(i - 1) >> 6] >> ((i - 1) & 63)) & 1u); }
static inline void cfg_flip(uint64_t *c, int64_t i) { c[(i - 1) >> 6] ^= (uint64_t)1 << ((i - 1) & 63); }

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11; Random123 constants).  Shared *definition* with the
 * CUDA engine (which carries its own implementation); pinned by Random123's published KATs
 * in tests/test_oracle_philox.py.
 * ---------------------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void philox_src_next(orc_philox_src *s, uint32_t out[4])
{
    uint32_t ctr[4] = { (uint32_t)s->n, (uint32_t)(s->n >> 32), (uint32_t)s->chain, s->tag ^ (uint32_t)(s->chain >> 32) };
    uint32_t key[2] = { (uint32_t)s->seed, (uint32_t)(s->seed >> 32) };
    orc_philox4x32_10(ctr, key, out);
    s->n++;
}
uint64_t orc_philox_u64(orc_philox_src *s)
{
    uint32_t o[4]; philox_src_next(s, o);
    return ((uint64_t)o[1] << 32) | o[0];
}
/* uniform in [0,1) with 53 bits: same construction on the GPU side */
double orc_philox_f64(void *src)
{
    uint64_t x = orc_philox_u64((orc_philox_src *)src);
    return (double)(x >> 11) * 0x1.0p-53;
}
/* unbiased 1..n by multiply-shift with rejection (Lemire 2019); one Philox call per trial */
int64_t orc_philox_range(void *src, int64_t n)
{
    orc_philox_src *s = (orc_philox_src *)src;
    uint64_t un = (uint64_t)n;
    for (;;) {
        uint64_t x = orc_philox_u64(s);
        __uint128_t m = (__uint128_t)x * un;
        uint64_t lo = (uint64_t)m;
        if (lo < un) {
            uint64_t t = (0 - un) % un;
            if (lo < t) continue;
        }
        return (int64_t)(m >> 64) + 1;
    }
}

/* xoshiro256++ (Blackman & Vigna) — the generator family Julia >= 1.7 uses by default; used only for the
 * CPU-baseline timing in bench.py so that the baseline is not handicapped by a counter-based RNG. */
static inline uint64_t rotl64(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t xoshiro_next(orc_xoshiro_src *g)
{
    uint64_t *s = g->s;
    const uint64_t result = rotl64(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl64(s[3], 45);
    return result;
}
void orc_xoshiro_seed(orc_xoshiro_src *g, uint64_t seed)
{
    for (int k = 0; k < 4; k++) { /* splitmix64 */
        uint64_t z = (seed += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        g->s[k] = z ^ (z >> 31);
    }
}
double orc_xoshiro_f64(void *src) { return (double)(xoshiro_next((orc_xoshiro_src *)src) >> 11) * 0x1.0p-53; }
int64_t orc_xoshiro_range(void *src, int64_t n)
{
    const uint64_t un = (uint64_t)n;
    for (;;) {
        __uint128_t m = (__uint128_t)xoshiro_next((orc_xoshiro_src *)src) * un;
        uint64_t lo = (uint64_t)m;
        if (lo < un) { uint64_t t = (0 - un) % un; if (lo < t) continue; }
        return (int64_t)(m >> 64) + 1;
    }
}

/* ------------------------------------------------------------------------------------------
 * Draw traces
 * ---------------------------------------------------------------------------------------- */
orc_trace *orc_trace_new(void) { return (orc_trace *)calloc(1, sizeof(orc_trace)); }
void orc_trace_free(orc_trace *t) { if (!t) return; free(t->kind); free(t->ival); free(t->fval); free(t); }
static void trace_push(orc_trace *t, uint8_t kind, int64_t iv, double fv)
{
    if (t->len == t->cap) {
        t->cap = t->cap ? 2 * t->cap : 1024;
        t->kind = (uint8_t *)realloc(t->kind, (size_t)t->cap);
        t->ival = (int64_t *)realloc(t->ival, (size_t)t->cap * 8);
        t->fval = (double *)realloc(t->fval, (size_t)t->cap * 8);
    }
    t->kind[t->len] = kind; t->ival[t->len] = iv; t->fval[t->len] = fv; t->len++;
}
void orc_trace_load(orc_trace *t, int64_t len, const uint8_t *kind, const int64_t *ival, const double *fval)
{
    t->len = 0; t->pos = 0; t->error = 0;
    for (int64_t k = 0; k < len; k++) trace_push(t, kind[k], ival[k], fval[k]);
}
double orc_trace_rec_f64(void *p) { orc_trace *t = (orc_trace *)p; double v = t->inner.f64(t->inner.user); trace_push(t, 1, 0, v); return v; }
int64_t orc_trace_rec_range(void *p, int64_t n) { orc_trace *t = (orc_trace *)p; int64_t v = t->inner.range(t->inner.user, n); trace_push(t, 0, v, 0.0); return v; }
double orc_trace_play_f64(void *p)
{
    orc_trace *t = (orc_trace *)p;
    if (t->pos >= t->len || t->kind[t->pos] != 1) { t->error = 1; return 0.5; }
    return t->fval[t->pos++];
}
int64_t orc_trace_play_range(void *p, int64_t n)
{
    orc_trace *t = (orc_trace *)p;
    if (t->pos >= t->len || t->kind[t->pos] != 0 || t->ival[t->pos] < 1 || t->ival[t->pos] > n) { t->error = 1; return 1; }
    return t->ival[t->pos++];
}

/* ------------------------------------------------------------------------------------------
 * Graph object
 * ---------------------------------------------------------------------------------------- */
struct orc_graph {
    int kind;
    int64_t N;
    /* EA (EA.jl:138-169, 534-553) */
    int twoD;
    int64_t *A;      /* [N*twoD] 1-based, rows sorted ascending */
    int64_t *Ji;     /* [N*twoD] (EA_INT) */
    double *Jd;      /* [N*twoD] (EA_F64) ; [N*N] (SK_F64) */
    int64_t *uA;     /* unique neighbours per row (EA.jl:158 / :548) */
    int *nuA;
    /* LocalFields{ET} (Common.jl:27-36) */
    int64_t *lfi, *lfi_last;
    double *lfd, *lfd_last;
    int64_t move_last;
    /* allΔE */
    int nDE;
    double DE[64];
    /* SK */
    uint8_t *Jb; /* [N*N] 0/1 (SK_BIN) */
    double sN;
    int owns_J;
    /* QT / Quant */
    int64_t M, Nk;
    double fourK, beta, Gamma;
    orc_graph *X0;
    orc_graph **X1;
    uint64_t **C1;
};

int orc_kind(const orc_graph *g) { return g->kind; }
int64_t orc_getN(const orc_graph *g) { return g->N; }
double orc_quant_fourK(const orc_graph *g) { return g->kind == ORC_QUANT ? g->X0->fourK : g->fourK; }
orc_graph *orc_inner_graph(orc_graph *g) { return (g->kind == ORC_QUANT || g->kind == ORC_EA_DISCR) ? g->X0 : g; } /* Interface.jl:239-240 */

static int is_discr(const orc_graph *g) { return g->kind == ORC_EA_INT || g->kind == ORC_QT; }
static int is_double(const orc_graph *g) { return g->kind == ORC_QUANT || g->kind == ORC_EA_DISCR; }

/* gen_EA — src/graphs/EA.jl:24-43.  Column-major linear index, first coordinate fastest;
 * for every site and dimension one forward bond, recorded at both ends; rows sorted. */
static int cmp_i64(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y); }
int64_t orc_gen_EA(int64_t L, int D, int64_t *A)
{
    if (L < 2 || D < 1) return -1;
    int64_t N = 1; for (int d = 0; d < D; d++) N *= L;
    int twoD = 2 * D;
    int *cnt = (int *)calloc((size_t)N, sizeof(int));
    for (int64_t x0 = 0; x0 < N; x0++) {            /* CartesianIndices order == linear order */
        int64_t stride = 1, rem = x0;
        for (int d = 0; d < D; d++) {
            int64_t id = rem % L; rem /= L;
            int64_t id1 = (id + 1) % L;               /* mod1(i+1, L) in 0-based form */
            int64_t y0 = x0 + (id1 - id) * stride;
            A[x0 * twoD + cnt[x0]++] = y0 + 1;        /* push!(A[x], y) */
            A[y0 * twoD + cnt[y0]++] = x0 + 1;        /* push!(A[y], x) */
            stride *= L;
        }
    }
    for (int64_t x = 0; x < N; x++) qsort(A + x * twoD, (size_t)twoD, sizeof(int64_t), cmp_i64);
    free(cnt);
    return N;
}

/* gen_J — src/graphs/EA.jl:45-71.  One draw per bond with x<y in (x,k) order; mirrored into the
 * first still-empty slot of J[y]. `draws` supplies f() values in consumption order. */
int orc_gen_J_f64(int64_t N, int twoD, const int64_t *A, const double *draws, int64_t ndraws, double *J)
{
    const double m = -INFINITY; /* sentinel(Float64)=typemin */
    for (int64_t k = 0; k < N * twoD; k++) J[k] = m;
    int64_t nd = 0;
    for (int64_t x = 1; x <= N; x++)
        for (int k = 0; k < twoD; k++) {
            int64_t y = A[(x - 1) * twoD + k];
            if (x < y) {
                if (nd >= ndraws) return -1;
                double Jxy = draws[nd++];
                J[(x - 1) * twoD + k] = Jxy;
                int l = 0; while (l < twoD && J[(y - 1) * twoD + l] != m) l++;
                if (l == twoD) return -2;
                J[(y - 1) * twoD + l] = Jxy;
            }
        }
    for (int64_t k = 0; k < N * twoD; k++) if (J[k] == m) return -3;
    return (int)nd;
}

static void ea_common_init(orc_graph *g, int64_t N, int twoD, const int64_t *A)
{
    g->N = N; g->twoD = twoD;
    g->A = (int64_t *)malloc((size_t)(N * twoD) * 8); memcpy(g->A, A, (size_t)(N * twoD) * 8);
    g->uA = (int64_t *)malloc((size_t)(N * twoD) * 8);
    g->nuA = (int *)malloc((size_t)N * sizeof(int));
    for (int64_t x = 0; x < N; x++) {               /* unique of a sorted row */
        int n = 0;
        for (int k = 0; k < twoD; k++) {
            int64_t y = A[x * twoD + k];
            if (n == 0 || g->uA[x * twoD + n - 1] != y) g->uA[x * twoD + n++] = y;
        }
        g->nuA[x] = n;
    }
}

/* generic allΔE for GraphEA — src/graphs/EA.jl:295-309 (the (-1,1) special case :293 gives the same tuple) */
static int cmp_f64(const void *a, const void *b) { double x = *(const double *)a, y = *(const double *)b; return (x > y) - (x < y); }
static void ea_int_allDE(orc_graph *g, const int64_t *lev, int nlev)
{
    int64_t cap = 1 << 16, ne = 1, *es = (int64_t *)malloc((size_t)cap * 8), *nw = (int64_t *)malloc((size_t)cap * 8);
    es[0] = 0;
    for (int n = 0; n < g->twoD; n++) {
        int64_t nn = 0;
        for (int64_t a = 0; a < ne; a++)
            for (int l = 0; l < nlev; l++) { nw[nn++] = es[a] + lev[l]; nw[nn++] = es[a] - lev[l]; }
        qsort(nw, (size_t)nn, 8, cmp_i64);
        int64_t u = 0; for (int64_t a = 0; a < nn; a++) if (u == 0 || nw[u - 1] != nw[a]) nw[u++] = nw[a];
        memcpy(es, nw, (size_t)u * 8); ne = u;
    }
    int nd = 0; double tmp[4096];
    for (int64_t a = 0; a < ne; a++) tmp[nd++] = 2.0 * (double)llabs(es[a]);
    qsort(tmp, (size_t)nd, 8, cmp_f64);
    g->nDE = 0;
    for (int a = 0; a < nd; a++) if (g->nDE == 0 || g->DE[g->nDE - 1] != tmp[a]) g->DE[g->nDE++] = tmp[a];
    free(es); free(nw);
}

orc_graph *orc_ea_int_create(int64_t N, int twoD, const int64_t *A, const int64_t *J, const int64_t *lev, int nlev)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_EA_INT; ea_common_init(g, N, twoD, A);
    g->Ji = (int64_t *)malloc((size_t)(N * twoD) * 8); memcpy(g->Ji, J, (size_t)(N * twoD) * 8);
    g->lfi = (int64_t *)calloc((size_t)N, 8); g->lfi_last = (int64_t *)calloc((size_t)N, 8);
    g->owns_J = 1;
    ea_int_allDE(g, lev, nlev);
    return g;
}
static orc_graph *ea_f64_create_shared(int64_t N, int twoD, const int64_t *A, double *J, int own)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_EA_F64; ea_common_init(g, N, twoD, A);
    if (own) { g->Jd = (double *)malloc((size_t)(N * twoD) * 8); memcpy(g->Jd, J, (size_t)(N * twoD) * 8); }
    else g->Jd = J;
    g->owns_J = own;
    g->lfd = (double *)calloc((size_t)N, 8); g->lfd_last = (double *)calloc((size_t)N, 8);
    return g;
}
orc_graph *orc_ea_f64_create(int64_t N, int twoD, const int64_t *A, const double *J) { return ea_f64_create_shared(N, twoD, A, (double *)J, 1); }

/* GraphRRG{Int,LEV,K}(A, J) — RRG.jl:112-137: GraphEA's arithmetic on a K-regular adjacency (energy :165-190, update_cache!
 * :192-237, delta_energy :239-259), except that neighbors() lists only the entries with a non-zero coupling (:133, :261). */
orc_graph *orc_rrg_int_create(int64_t N, int K, const int64_t *A, const int64_t *J, const int64_t *lev, int nlev)
{
    orc_graph *g = orc_ea_int_create(N, K, A, J, lev, nlev);
    for (int64_t x = 0; x < N; x++) {
        int n = 0;
        for (int k = 0; k < K; k++) if (J[x * K + k] != 0) g->uA[x * K + n++] = A[x * K + k];
        g->nuA[x] = n;
    }
    return g;
}

/* GraphEANormalDiscretized{Int,LEV,twoD} <: DoubleGraph{DiscrGraph{Int},Float64} — EA.jl:311-344 with explicit continuous
 * couplings cJ (the constructor draws them with gen_J(randn)): every coupling is split by discretize (Common.jl:38-49:
 * nearest level, the first one wins ties) into a level dJ, which goes to the inner GraphEA{Int,LEV}, and a residual
 * rJ = cJ - dJ. The residual part has exactly the arithmetic of GraphEANormal on rJ (energy EA.jl:362-388 vs :584-611,
 * update_cache_residual! :452-487 vs :613-653, delta_energy_residual :489-497 vs :655-663), so it is held as one. */
static orc_graph *discretized_create(int64_t N, int twoD, const int64_t *A, const double *cJ, const int64_t *lev, int nlev, int rrg);
orc_graph *orc_ea_discretized_create(int64_t N, int twoD, const int64_t *A, const double *cJ, const int64_t *lev, int nlev)
{
    return discretized_create(N, twoD, A, cJ, lev, nlev, 0);
}
/* GraphRRGNormalDiscretized{Int,LEV,K} — RRG.jl:274-310: the same construction over GraphRRG; neighbors(X) is the whole
 * row A[i] (:499) while neighbors(inner_graph(X)) skips the zero levels (:133). */
orc_graph *orc_rrg_discretized_create(int64_t N, int K, const int64_t *A, const double *cJ, const int64_t *lev, int nlev)
{
    return discretized_create(N, K, A, cJ, lev, nlev, 1);
}
static orc_graph *discretized_create(int64_t N, int twoD, const int64_t *A, const double *cJ, const int64_t *lev, int nlev, int rrg)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_EA_DISCR; g->N = N; g->twoD = twoD; g->M = 1;
    int64_t *dJ = (int64_t *)malloc((size_t)(N * twoD) * 8);
    double *rJ = (double *)malloc((size_t)(N * twoD) * 8);
    for (int64_t a = 0; a < N * twoD; a++) {
        double x = cJ[a];
        int64_t d = lev[0]; double r = x - (double)d;
        for (int l = 1; l < nlev; l++) {
            double r1 = x - (double)lev[l];
            if (fabs(r1) < fabs(r)) { d = lev[l]; r = r1; }
        }
        dJ[a] = d; rJ[a] = r;
    }
    g->X0 = rrg ? orc_rrg_int_create(N, twoD, A, dJ, lev, nlev) : orc_ea_int_create(N, twoD, A, dJ, lev, nlev);
    g->X1 = (orc_graph **)calloc(1, sizeof(orc_graph *));
    g->X1[0] = orc_ea_f64_create(N, twoD, A, rJ);
    free(dJ); free(rJ);
    return g;
}

static orc_graph *sk_f64_create_shared(int64_t N, double *J, int own)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_SK_F64; g->N = N;
    if (own) { g->Jd = (double *)malloc((size_t)(N * N) * 8); memcpy(g->Jd, J, (size_t)(N * N) * 8); } else g->Jd = J;
    g->owns_J = own;
    g->lfd = (double *)calloc((size_t)N, 8); g->lfd_last = (double *)calloc((size_t)N, 8);
    return g;
}
orc_graph *orc_sk_f64_create(int64_t N, const double *J) { return sk_f64_create_shared(N, (double *)J, 1); }
static orc_graph *sk_bin_create_shared(int64_t N, uint8_t *J, int own)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_SK_BIN; g->N = N; g->sN = sqrt((double)N); /* SK.jl:47 √N */
    if (own) { g->Jb = (uint8_t *)malloc((size_t)(N * N)); memcpy(g->Jb, J, (size_t)(N * N)); } else g->Jb = J;
    g->owns_J = own;
    g->lfi = (int64_t *)calloc((size_t)N, 8); g->lfi_last = (int64_t *)calloc((size_t)N, 8);
    return g;
}
orc_graph *orc_sk_bin_create(int64_t N, const uint8_t *J) { return sk_bin_create_shared(N, (uint8_t *)J, 1); }

orc_graph *orc_qt_create(int64_t N, int64_t M, double fourK) /* QT.jl:42-54, allΔE :111 */
{
    if (M <= 2 || N % M != 0) return NULL;
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_QT; g->N = N; g->M = M; g->Nk = N / M; g->fourK = fourK;
    g->nDE = 2; g->DE[0] = 0.0; g->DE[1] = fourK;
    return g;
}
orc_graph *orc_empty_create(int64_t N)
{
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_EMPTY; g->N = N;
    return g;
}

/* round(x, digits=8) as Julia does it for Float64 (RoundNearest on x*10^8, then /10^8) — QT.jl:165 */
static double round8(double x) { return nearbyint(x * 1e8) / 1e8; }

orc_graph *orc_quant_create(int64_t Nk, int64_t M, double Gamma, double beta, int inner_kind,
                            const void *J_inner, int twoD, const int64_t *A_inner)
{
    if (!(Gamma >= 0) || M <= 2) return NULL;
    double fourK = round8(2.0 / beta * log(1.0 / tanh(beta * Gamma / (double)M))); /* QT.jl:165 */
    orc_graph *g = (orc_graph *)calloc(1, sizeof(orc_graph));
    g->kind = ORC_QUANT; g->N = Nk * M; g->M = M; g->Nk = Nk; g->beta = beta; g->Gamma = Gamma;
    g->X0 = orc_qt_create(g->N, M, fourK);
    g->X1 = (orc_graph **)calloc((size_t)M, sizeof(orc_graph *));
    g->C1 = (uint64_t **)calloc((size_t)M, sizeof(uint64_t *));
    void *Jshared = NULL;
    for (int64_t k = 0; k < M; k++) {
        switch (inner_kind) { /* QT.jl:139-145: M inner graphs sharing the constructor args, separate caches */
        case ORC_SK_BIN:
            if (k == 0) { g->X1[0] = orc_sk_bin_create(Nk, (const uint8_t *)J_inner); Jshared = g->X1[0]->Jb; }
            else g->X1[k] = sk_bin_create_shared(Nk, (uint8_t *)Jshared, 0);
            break;
        case ORC_SK_F64:
            if (k == 0) { g->X1[0] = orc_sk_f64_create(Nk, (const double *)J_inner); Jshared = g->X1[0]->Jd; }
            else g->X1[k] = sk_f64_create_shared(Nk, (double *)Jshared, 0);
            break;
        case ORC_EA_F64:
            if (k == 0) { g->X1[0] = orc_ea_f64_create(Nk, twoD, A_inner, (const double *)J_inner); Jshared = g->X1[0]->Jd; }
            else g->X1[k] = ea_f64_create_shared(Nk, twoD, A_inner, (double *)Jshared, 0);
            break;
        case ORC_EMPTY: g->X1[k] = orc_empty_create(Nk); break;
        default: return NULL;
        }
        g->C1[k] = (uint64_t *)calloc((size_t)((Nk + 63) / 64), 8); /* Config(Nk, init=false) QT.jl:144 */
    }
    return g;
}

void orc_graph_free(orc_graph *g)
{
    if (!g) return;
    if (g->kind == ORC_QUANT) {
        for (int64_t k = g->M - 1; k >= 0; k--) { orc_graph_free(g->X1[k]); free(g->C1[k]); }
        free(g->X1); free(g->C1); orc_graph_free(g->X0);
    }
    if (g->kind == ORC_EA_DISCR) { orc_graph_free(g->X1[0]); free(g->X1); orc_graph_free(g->X0); }
    free(g->A); free(g->uA); free(g->nuA);
    if (g->owns_J) { free(g->Ji); free(g->Jd); free(g->Jb); }
    free(g->lfi); free(g->lfi_last); free(g->lfd); free(g->lfd_last);
    free(g);
}

/* ------------------------------------------------------------------------------------------
 * energy  (also (re)initialises the cache — Interface.jl:103)
 * ---------------------------------------------------------------------------------------- */
static int64_t qt_energy0(const orc_graph *g, const uint64_t *s) /* QT.jl:68-82 */
{
    int64_t n = 0, M = g->M, Nk = g->Nk;
    for (int64_t i = 1; i <= Nk; i++) {
        int sj = cfg_get(s, i + (M - 1) * Nk);
        for (int64_t k = 1; k <= M; k++) {
            int sk = cfg_get(s, i + (k - 1) * Nk);
            n -= 1 - 2 * (sk ^ sj);
            sj = sk;
        }
    }
    return n;
}

double orc_energy(orc_graph *g, const uint64_t *s)
{
    int64_t N = g->N;
    switch (g->kind) {
    case ORC_EA_INT: { /* EA.jl:195-222 */
        int64_t n = 0;
        for (int64_t x = 1; x <= N; x++) {
            int64_t sx = 2 * cfg_get(s, x) - 1, lf = 0;
            for (int k = 0; k < g->twoD; k++) {
                int64_t y = g->A[(x - 1) * g->twoD + k];
                int64_t sy = 2 * cfg_get(s, y) - 1;
                lf -= g->Ji[(x - 1) * g->twoD + k] * sx * sy;
            }
            n += lf;
            g->lfi[x - 1] = 2 * lf;
        }
        g->move_last = 0;
        memset(g->lfi_last, 0, (size_t)N * 8);
        return (double)n / 2.0; /* n /= 2 ; discr(Int, n) */
    }
    case ORC_EA_F64: { /* EA.jl:584-611 */
        double E1 = 0.0;
        for (int64_t x = 1; x <= N; x++) {
            double sx = (double)(2 * cfg_get(s, x) - 1), lf = 0.0;
            for (int k = 0; k < g->twoD; k++) {
                int64_t y = g->A[(x - 1) * g->twoD + k];
                double sy = (double)(2 * cfg_get(s, y) - 1);
                lf -= g->Jd[(x - 1) * g->twoD + k] * sx * sy; /* (Jxy*σx)*σy */
            }
            E1 += lf;
            g->lfd[x - 1] = 2 * lf;
        }
        E1 /= 2;
        g->move_last = 0;
        memset(g->lfd_last, 0, (size_t)N * 8);
        return E1;
    }
    case ORC_SK_F64: { /* SK.jl:212-237 */
        double n = 0.0;
        for (int64_t i = 1; i <= N; i++) {
            const double *Ji = g->Jd + (i - 1) * N;
            int si = cfg_get(s, i);
            double lf = 0.0;
            for (int64_t j = 1; j <= N; j++) lf += (double)(1 - 2 * (si ^ cfg_get(s, j))) * Ji[j - 1];
            g->lfd[i - 1] = 2 * lf;
            n -= lf;
        }
        n /= 2;
        g->move_last = 0;
        memset(g->lfd_last, 0, (size_t)N * 8);
        return n;
    }
    case ORC_SK_BIN: { /* SK.jl:62-94 */
        int64_t sums = 0;
        for (int64_t i = 1; i <= N; i++) sums += cfg_get(s, i);
        int64_t n = -2 * sums;
        for (int64_t i = 1; i <= N; i++) {
            const uint8_t *Ji = g->Jb + (i - 1) * N;
            int64_t sc = 0;
            for (int64_t j = 1; j <= N; j++) sc += Ji[j - 1] ^ cfg_get(s, j);
            int64_t si = cfg_get(s, i);
            int64_t lf = -(2 * si - 1) * (N - 1 - 2 * sc);
            g->lfi[i - 1] = 2 * (-lf + 2 * si);
            n += lf;
        }
        n /= 2; /* @assert n % 2 == 0 ; n ÷= 2 */
        g->move_last = 0;
        memset(g->lfi_last, 0, (size_t)N * 8);
        return (double)n / g->sN;
    }
    case ORC_QT: /* QT.jl:84 */
        return (double)qt_energy0(g, s) * g->fourK / 4;
    case ORC_EMPTY:
        return 0.0;
    case ORC_EA_DISCR: { /* EA.jl:362-388: convert(Float64, E0 + E1) */
        double E0 = orc_energy(g->X0, s);
        double E1 = orc_energy(g->X1[0], s);
        return E0 + E1;
    }
    case ORC_QUANT: { /* QT.jl:185-199 */
        double E = orc_energy(g->X0, s);
        for (int64_t k = 1; k <= g->M; k++) {
            uint64_t *s1 = g->C1[k - 1];
            memset(s1, 0, (size_t)((g->Nk + 63) / 64) * 8);
            for (int64_t i = 1; i <= g->Nk; i++) if (cfg_get(s, (k - 1) * g->Nk + i)) cfg_flip(s1, i); /* copyto! */
            E += orc_energy(g->X1[k - 1], s1) / (double)g->M;
        }
        return E;
    }
    }
    return NAN;
}

/* ------------------------------------------------------------------------------------------
 * delta_energy  (called BEFORE the flip — Interface.jl:127-128)
 * ---------------------------------------------------------------------------------------- */
static void qt_neighbors(const orc_graph *g, int64_t i, int64_t *k1, int64_t *k2) /* QT.jl:105-108 */
{
    *k1 = i - g->Nk + g->N * (i <= g->Nk);
    *k2 = i + g->Nk - g->N * (i + g->Nk > g->N);
}

double orc_delta_energy_residual(orc_graph *g, const uint64_t *s, int64_t move) /* QT.jl:270-281 */
{
    (void)s;
    if (g->kind == ORC_EA_DISCR) return orc_delta_energy(g->X1[0], s, move); /* EA.jl:489-497: -lfields[move] */
    if (g->kind != ORC_QUANT) return 0.0;
    int64_t k = (move - 1) / g->Nk + 1, i = (move - 1) % g->Nk + 1;
    return orc_delta_energy(g->X1[k - 1], g->C1[k - 1], i) / (double)g->M;
}

double orc_delta_energy(orc_graph *g, const uint64_t *s, int64_t move)
{
    switch (g->kind) {
    case ORC_EA_INT: return (double)(-g->lfi[move - 1]);     /* EA.jl:266-275 */
    case ORC_EA_F64: return -g->lfd[move - 1];               /* EA.jl:655-663 */
    case ORC_SK_F64: return g->lfd[move - 1];                /* SK.jl:278-284 */
    case ORC_SK_BIN: return (double)g->lfi[move - 1] / g->sN; /* SK.jl:135-140 */
    case ORC_QT: {                                           /* QT.jl:86-103 */
        int64_t k1, k2; qt_neighbors(g, move, &k1, &k2);
        int sk = cfg_get(s, move), s1 = cfg_get(s, k1), s2 = cfg_get(s, k2);
        int d = (sk == s1) - (sk != s2);                     /* (sk ⊻ ~s1) - (sk ⊻ s2) on Bools */
        return (double)d * g->fourK;
    }
    case ORC_EMPTY: return 0.0;
    case ORC_QUANT: /* QT.jl:283-286 */
    case ORC_EA_DISCR: /* EA.jl:519-523 */
        return orc_delta_energy(g->X0, s, move) + orc_delta_energy_residual(g, s, move);
    }
    return NAN;
}

/* ------------------------------------------------------------------------------------------
 * update_cache!  (called AFTER the flip — Interface.jl:84-85) and spinflip! (Interface.jl:89-92)
 * ---------------------------------------------------------------------------------------- */
static void update_cache(orc_graph *g, uint64_t *s, int64_t move);

void orc_spinflip(orc_graph *g, uint64_t *s, int64_t move)
{
    cfg_flip(s, move);
    update_cache(g, s, move);
}

static void update_cache(orc_graph *g, uint64_t *s, int64_t move)
{
    int64_t N = g->N;
    switch (g->kind) {
    case ORC_EA_INT: { /* EA.jl:224-264 */
        const int64_t *U = g->uA + (move - 1) * g->twoD; int nU = g->nuA[move - 1];
        if (g->move_last == move) {
            for (int k = 0; k < nU; k++) { int64_t y = U[k] - 1, t = g->lfi[y]; g->lfi[y] = g->lfi_last[y]; g->lfi_last[y] = t; }
            g->lfi[move - 1] = -g->lfi[move - 1];
            g->lfi_last[move - 1] = -g->lfi_last[move - 1];
            return;
        }
        for (int k = 0; k < nU; k++) g->lfi_last[U[k] - 1] = g->lfi[U[k] - 1];
        int sx = cfg_get(s, move);
        for (int k = 0; k < g->twoD; k++) {
            int64_t y = g->A[(move - 1) * g->twoD + k];
            int64_t sxy = 1 - 2 * (sx ^ cfg_get(s, y));
            g->lfi[y - 1] = g->lfi[y - 1] - 4 * sxy * g->Ji[(move - 1) * g->twoD + k];
        }
        int64_t lfm = g->lfi[move - 1];
        g->lfi_last[move - 1] = lfm;
        g->lfi[move - 1] = -lfm;
        g->move_last = move;
        return;
    }
    case ORC_EA_F64: { /* EA.jl:613-653 */
        const int64_t *U = g->uA + (move - 1) * g->twoD; int nU = g->nuA[move - 1];
        if (g->move_last == move) {
            for (int k = 0; k < nU; k++) { int64_t y = U[k] - 1; double t = g->lfd[y]; g->lfd[y] = g->lfd_last[y]; g->lfd_last[y] = t; }
            g->lfd[move - 1] = -g->lfd[move - 1];
            g->lfd_last[move - 1] = -g->lfd_last[move - 1];
            return;
        }
        for (int k = 0; k < nU; k++) g->lfd_last[U[k] - 1] = g->lfd[U[k] - 1];
        int sx = cfg_get(s, move);
        for (int k = 0; k < g->twoD; k++) {
            int64_t y = g->A[(move - 1) * g->twoD + k];
            double f = (double)(4 * (1 - 2 * (sx ^ cfg_get(s, y)))); /* 4 * σxy */
            g->lfd[y - 1] -= f * g->Jd[(move - 1) * g->twoD + k];
        }
        double lfm = g->lfd[move - 1];
        g->lfd_last[move - 1] = lfm;
        g->lfd[move - 1] = -lfm;
        g->move_last = move;
        return;
    }
    case ORC_SK_F64: { /* SK.jl:239-276 */
        if (g->move_last == move) { double *t = g->lfd; g->lfd = g->lfd_last; g->lfd_last = t; return; }
        const double *Ji = g->Jd + (move - 1) * N;
        int si = cfg_get(s, move);
        double lfm = g->lfd[move - 1];
        for (int64_t j = 1; j <= N; j++) {
            double Js = (double)(1 - 2 * (si ^ cfg_get(s, j))) * Ji[j - 1];
            double lfj = g->lfd[j - 1];
            g->lfd_last[j - 1] = lfj;
            g->lfd[j - 1] = lfj + 4 * Js;
        }
        g->lfd_last[move - 1] = lfm;
        g->lfd[move - 1] = -lfm;
        g->move_last = move;
        return;
    }
    case ORC_SK_BIN: { /* SK.jl:96-133 */
        if (g->move_last == move) { int64_t *t = g->lfi; g->lfi = g->lfi_last; g->lfi_last = t; return; }
        const uint8_t *Ji = g->Jb + (move - 1) * N;
        int si = cfg_get(s, move);
        int64_t lfm = g->lfi[move - 1];
        for (int64_t j = 1; j <= N; j++) {
            int64_t Js = si ^ cfg_get(s, j) ^ Ji[j - 1];
            int64_t lfj = g->lfi[j - 1];
            g->lfi_last[j - 1] = lfj;
            g->lfi[j - 1] = lfj + 8 * Js - 4;
        }
        g->lfi_last[move - 1] = lfm;
        g->lfi[move - 1] = -lfm;
        g->move_last = move;
        return;
    }
    case ORC_QT: case ORC_EMPTY: return; /* Interface.jl:87 default: nothing */
    case ORC_QUANT: { /* QT.jl:172-183 */
        int64_t k = (move - 1) / g->Nk + 1, i = (move - 1) % g->Nk + 1;
        orc_spinflip(g->X1[k - 1], g->C1[k - 1], i);
        return;
    }
    case ORC_EA_DISCR: /* EA.jl:390-450. When the two caches disagree on move_last the reference updates them one
                        * after the other (:394-398); when they agree, its fused body performs the same two updates
                        * (both swaps when move_last == move, both ordinary updates otherwise). */
        update_cache(g->X0, s, move);
        update_cache(g->X1[0], s, move); /* update_cache_residual! :452-487 */
        return;
    }
}

/* neighbors — EA.jl:292,680 (uA); SK.jl:165,297 + Common.jl:78-92 (AllButOne); QT.jl:105-108; QT.jl:288-321 */
int orc_neighbors(const orc_graph *g, int64_t i, int64_t *out)
{
    switch (g->kind) {
    case ORC_EA_INT: case ORC_EA_F64: {
        int n = g->nuA[i - 1];
        for (int k = 0; k < n; k++) out[k] = g->uA[(i - 1) * g->twoD + k];
        return n;
    }
    case ORC_SK_F64: case ORC_SK_BIN: {
        int n = 0;
        for (int64_t j = 1; j <= g->N; j++) if (j != i) out[n++] = j;
        return n;
    }
    case ORC_QT: qt_neighbors(g, i, &out[0], &out[1]); return 2;
    case ORC_EMPTY: return 0;
    case ORC_EA_DISCR: return orc_neighbors(g->X1[0], i, out); /* EA.jl:525 (uA), RRG.jl:499 (A[i]): the distinct entries of the row */
    case ORC_QUANT: {
        qt_neighbors(g->X0, i, &out[0], &out[1]);
        int64_t k = (i - 1) / g->Nk + 1, j = (i - 1) % g->Nk + 1;
        int n = orc_neighbors(g->X1[k - 1], j, out + 2);
        for (int a = 0; a < n; a++) out[2 + a] += (k - 1) * g->Nk;
        return n + 2;
    }
    }
    return 0;
}

int orc_allDE(const orc_graph *g, double *out) /* Interface.jl:200-201,270 */
{
    const orc_graph *h = (g->kind == ORC_QUANT || g->kind == ORC_EA_DISCR) ? g->X0 : g;
    if (!is_discr(h)) return -1;
    for (int k = 0; k < h->nDE; k++) out[k] = h->DE[k];
    return h->nDE;
}

int64_t orc_get_lfields(const orc_graph *g, double *out)
{
    if (g->lfi) { for (int64_t i = 0; i < g->N; i++) out[i] = (double)g->lfi[i]; return g->N; }
    if (g->lfd) { for (int64_t i = 0; i < g->N; i++) out[i] = g->lfd[i]; return g->N; }
    return 0;
}

/* observables — QT.jl:113-121, 201-268 */
double orc_transverse_mag(orc_graph *g, const uint64_t *s, double beta)
{
    orc_graph *q = orc_inner_graph(g);
    double p = -(double)qt_energy0(q, s) / (double)q->N;
    double x = beta * q->fourK / 2;
    return cosh(x) - p * sinh(x);
}
double orc_Qenergy(orc_graph *g, const uint64_t *s)
{
    double E = -g->Gamma * orc_transverse_mag(g, s, g->beta);
    for (int64_t k = 0; k < g->M; k++) E += orc_energy(g->X1[k], g->C1[k]) / (double)g->N;
    return E;
}
void orc_Renergies(orc_graph *g, double *out) { for (int64_t k = 0; k < g->M; k++) out[k] = orc_energy(g->X1[k], g->C1[k]); }
void orc_overlaps(orc_graph *g, double *ovs)
{
    int64_t M = g->M, Nk = g->Nk;
    for (int64_t d = 0; d < M / 2; d++) ovs[d] = 0.0;
    for (int64_t k1 = 1; k1 <= M - 1; k1++)
        for (int64_t k2 = k1 + 1; k2 <= M; k2++) {
            int64_t sum = 0;
            for (int64_t i = 1; i <= Nk; i++) sum += cfg_get(g->C1[k1 - 1], i) ^ cfg_get(g->C1[k2 - 1], i);
            int64_t dl = k2 - k1 < M + k1 - k2 ? k2 - k1 : M + k1 - k2;
            ovs[dl - 1] += (double)(Nk - 2 * sum);
        }
    for (int64_t d = 1; d <= (M - 1) / 2; d++) ovs[d - 1] /= (double)(M * Nk);
    if (M % 2 == 0) ovs[M / 2 - 1] /= (double)(M * Nk) / 2;
}

/* ------------------------------------------------------------------------------------------
 * ArraySet — src/ArraySets.jl:19-85
 * ---------------------------------------------------------------------------------------- */
typedef struct { int64_t N, t; int64_t *v, *pos; } arrayset;
static void as_init(arrayset *a, int64_t N) { a->N = N; a->t = 0; a->v = (int64_t *)calloc((size_t)N + 1, 8); a->pos = (int64_t *)calloc((size_t)N + 1, 8); }
static void as_free(arrayset *a) { free(a->v); free(a->pos); }
static inline void as_push(arrayset *a, int64_t i) { a->t++; a->v[a->t] = i; a->pos[i] = a->t; }
static inline void as_delete(arrayset *a, int64_t i)
{
    int64_t p = a->pos[i];
    a->v[p] = a->v[a->t];
    a->pos[a->v[p]] = p;
    a->pos[i] = 0;
    a->t--;
}
static int as_check(const arrayset *a) /* ArraySets.jl:27-42 */
{
    if (a->t < 0 || a->t > a->N) return 1;
    int64_t c = 0;
    for (int64_t i = 1; i <= a->N; i++) {
        if (a->pos[i] == 0) continue;
        c++;
        if (a->pos[i] < 1 || a->pos[i] > a->t) return 2;
        if (a->v[a->pos[i]] != i) return 3;
    }
    if (c != a->t) return 4;
    for (int64_t i = 1; i <= a->t; i++) { if (a->v[i] == 0) return 5; if (a->pos[a->v[i]] != i) return 6; }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * DeltaECache (discrete) — src/DeltaE.jl:28-295
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int64_t N; int L;
    double DE[64], ft[64];
    double *T, *Tp; double z, zp;
    arrayset *as; int64_t *pos;
    int64_t (*staged)[3]; int64_t nstaged;
    int rrr;
} decache;

static int findk(const decache *c, double dE) /* DeltaE.jl:28-60: index of |ΔE| in ΔElist, 0 if absent */
{
    dE = fabs(dE);
    for (int k = 1; k <= c->L; k++) if (c->DE[k - 1] == dE) return k;
    return 0;
}
static inline double class_f(const decache *c, int k) { return k > c->L ? c->ft[k - c->L - 1] : 1.0; } /* DeltaE.jl:138-139 */

static decache *decache_new(orc_graph *X, const uint64_t *s, double beta, int rrr) /* DeltaE.jl:74-104 */
{
    decache *c = (decache *)calloc(1, sizeof(decache));
    c->N = X->N; c->L = orc_allDE(X, c->DE); c->rrr = rrr;
    int L = c->L;
    c->as = (arrayset *)calloc((size_t)(2 * L + 1), sizeof(arrayset));
    for (int k = 1; k <= 2 * L; k++) as_init(&c->as[k], c->N);
    c->pos = (int64_t *)calloc((size_t)c->N + 1, 8);
    for (int64_t i = 1; i <= c->N; i++) {
        double dE = orc_delta_energy(X, s, i);
        int aki = findk(c, dE);
        int upi = dE > 0 || (dE == 0 && cfg_get(s, i) == 1);
        int ki = aki + L * upi;
        c->pos[i] = ki;
        as_push(&c->as[ki], i);
    }
    c->staged = (int64_t(*)[3])malloc((size_t)(c->N + 1) * 3 * 8);
    for (int k = 0; k < L; k++) c->ft[k] = exp(-beta * c->DE[k]);
    c->T = (double *)calloc((size_t)(2 * L + 1), 8);
    c->z = 0.0;
    for (int k = 1; k <= 2 * L; k++) {
        double x = (double)c->as[k].t * class_f(c, k);
        c->z += x;
        c->T[k] = x;
    }
    c->Tp = rrr ? (double *)calloc((size_t)(2 * L + 1), 8) : c->T;
    c->zp = c->z;
    return c;
}
static void decache_free(decache *c)
{
    for (int k = 1; k <= 2 * c->L; k++) as_free(&c->as[k]);
    free(c->as); free(c->pos); free(c->staged);
    if (c->Tp != c->T) free(c->Tp);
    free(c->T); free(c);
}
static int decache_check(const decache *c) /* DeltaE.jl:120-136 */
{
    for (int k = 1; k <= 2 * c->L; k++) { int e = as_check(&c->as[k]); if (e) return 10 * k + e; }
    for (int64_t i = 1; i <= c->N; i++) {
        int64_t k = c->pos[i];
        if (k < 1 || k > 2 * c->L) return 1000;
        int64_t p = c->as[k].pos[i];
        if (p < 1 || p > c->as[k].t) return 1001;
        for (int k1 = 1; k1 <= 2 * c->L; k1++) if (k1 != k && c->as[k1].pos[i] != 0) return 1002;
    }
    return 0;
}

static int64_t d_rand_skip(const decache *c, orc_draws d) /* DeltaE.jl:141-144 */
{
    return (int64_t)floor(log1p(-d.f64(d.user)) / log1p(-c->z / (double)c->N));
}
static int64_t d_rand_move(const decache *c, orc_draws d, double *dE) /* DeltaE.jl:146-167 */
{
    int L = c->L;
    double r = d.f64(d.user) * c->z, cT = 0.0;
    int k = 0, broke = 0;
    for (k = 1; k <= 2 * L; k++) { cT += c->T[k]; if (r < cT) { broke = 1; break; } }
    if (!broke) k = 2 * L;                  /* `for outer k` leaves k at the last value */
    if (!(r < cT)) while (c->T[k] == 0) k--;
    *dE = k <= L ? -c->DE[k - 1] : c->DE[k - L - 1];
    const arrayset *a = &c->as[k];
    return a->v[d.range(d.user, a->t)];     /* ArraySets.jl:81-85 */
}
static void d_compute_staged(orc_graph *X, uint64_t *s, int64_t i, decache *c) /* DeltaE.jl:202-230 */
{
    int L = c->L; int64_t nb[64]; /* discrete graphs on this path have ≤ 2D (or 2) neighbours */
    orc_spinflip(X, s, i);
    c->nstaged = 0;
    int n = orc_neighbors(X, i, nb);
    for (int a = 0; a < n; a++) {
        int64_t j = nb[a], k0 = c->pos[j];
        double dE1 = orc_delta_energy(X, s, j);
        int ak1 = findk(c, dE1);
        int upj1 = dE1 > 0 || (dE1 == 0 && cfg_get(s, j) == 1);
        int64_t k1 = ak1 + L * upj1;
        if (k0 == k1) continue;
        c->staged[c->nstaged][0] = j; c->staged[c->nstaged][1] = k0; c->staged[c->nstaged][2] = k1; c->nstaged++;
    }
    int64_t k0 = c->pos[i], k1 = k0 - L * (2 * (k0 > L) - 1);
    c->staged[c->nstaged][0] = i; c->staged[c->nstaged][1] = k0; c->staged[c->nstaged][2] = k1; c->nstaged++;
    orc_spinflip(X, s, i);
}
static double d_reverse_probs(decache *c) /* DeltaE.jl:184-200 */
{
    double zp = c->z;
    if (c->Tp != c->T) memcpy(c->Tp, c->T, (size_t)(2 * c->L + 1) * 8);
    for (int64_t a = 0; a < c->nstaged; a++) {
        int k0 = (int)c->staged[a][1], k1 = (int)c->staged[a][2];
        double f0 = class_f(c, k0), f1 = class_f(c, k1);
        c->Tp[k0] -= f0;
        c->Tp[k1] += f1;
        zp += f1 - f0;
    }
    c->zp = zp;
    return zp;
}
static void d_apply_staged(decache *c) /* DeltaE.jl:169-182 */
{
    for (int64_t a = 0; a < c->nstaged; a++) {
        int64_t j = c->staged[a][0]; int k0 = (int)c->staged[a][1], k1 = (int)c->staged[a][2];
        as_delete(&c->as[k0], j);
        as_push(&c->as[k1], j);
        c->pos[j] = k1;
    }
    double *t = c->T; c->T = c->Tp; c->Tp = t; c->z = c->zp;
}
/* apply_move! for DiscrGraph / DoubleGraph{DiscrGraph} — DeltaE.jl:232-295 */
static double d_apply_move(orc_graph *X, uint64_t *s, int64_t move, decache *c)
{
    int L = c->L; int64_t nb[64];
    orc_spinflip(X, s, move);
    orc_graph *X0 = orc_inner_graph(X);
    double zp = c->z;
    int n = orc_neighbors(X0, move, nb);
    for (int a = 0; a <= n; a++) {
        int64_t j, k0, k1;
        if (a < n) {
            j = nb[a]; k0 = c->pos[j];
            double dE1 = orc_delta_energy(X0, s, j);
            int ak1 = findk(c, dE1);
            int upj1 = dE1 > 0 || (dE1 == 0 && cfg_get(s, j) == 1);
            k1 = ak1 + L * upj1;
            if (k0 == k1) continue;
        } else {
            j = move; k0 = c->pos[move]; k1 = k0 - L * (2 * (k0 > L) - 1);
        }
        double f0 = class_f(c, (int)k0), f1 = class_f(c, (int)k1);
        c->T[k0] -= f0;
        c->T[k1] += f1;
        zp += f1 - f0;
        as_delete(&c->as[k0], j);
        as_push(&c->as[k1], j);
        c->pos[j] = k1;
    }
    double cc = c->z / zp;
    c->z = zp;
    return cc;
}

int orc_check_discrete_cache(orc_graph *g, uint64_t *s, double beta, const int64_t *sites, int64_t nmoves)
{
    orc_graph *X0 = orc_inner_graph(g);
    orc_energy(g, s);
    decache *c = decache_new(X0, s, beta, 1);
    int e = decache_check(c);
    for (int64_t m = 0; m < nmoves && !e; m++) {
        d_apply_move(g, s, sites[m], c);
        e = decache_check(c);
        if (!e) { /* classes must equal a fresh classification */
            for (int64_t i = 1; i <= c->N && !e; i++) {
                double dE = orc_delta_energy(X0, s, i);
                int ki = findk(c, dE) + c->L * (dE > 0 || (dE == 0 && cfg_get(s, i) == 1));
                if (ki != c->pos[i]) e = 2000;
            }
            double z = 0; for (int k = 1; k <= 2 * c->L; k++) z += (double)c->as[k].t * class_f(c, k);
            if (fabs(z - c->z) > 1e-9 * (1 + fabs(z))) e = 2001;
        }
    }
    decache_free(c);
    return e;
}

/* ------------------------------------------------------------------------------------------
 * DynamicSampler — src/DynamicSamplers.jl:18-176 (tree walk instead of the tinds/tpos tables;
 * the reference keeps that equivalent form in comments, :182-196)
 * ---------------------------------------------------------------------------------------- */
typedef struct { double *v, *ps; double z; int64_t N, N2; int levs; int64_t trefresh; } dynsmp;

static void ds_add_path(dynsmp *d, int64_t i, double x) /* ps[k] += x along the path of element i (1-based) */
{
    int64_t k = 0, off = 1, u = d->levs > 0 ? (int64_t)1 << (d->levs - 1) : 0, i0 = i - 1;
    for (int lev = 1; lev <= d->levs; lev++) {
        if ((i0 & u) == 0) { d->ps[off + k] += x; k *= 2; }
        else k = 2 * k + 1;
        u >>= 1; off *= 2;
    }
}
static void ds_refresh(dynsmp *d) /* DynamicSamplers.jl:84-98 */
{
    double z = 0.0;
    for (int64_t i = 1; i <= d->N2; i++) z += d->v[i]; /* sum(v): Julia sums pairwise; differs by O(ε) only */
    d->z = z;
    memset(d->ps, 0, (size_t)(d->N2 + 1) * 8);
    for (int64_t i = 1; i <= d->N; i++) ds_add_path(d, i, d->v[i]);
    d->trefresh = 0;
}
static dynsmp *ds_new(int64_t N, const double *v) /* DynamicSamplers.jl:35-51 */
{
    dynsmp *d = (dynsmp *)calloc(1, sizeof(dynsmp));
    d->N = N; d->levs = 0; while (((int64_t)1 << d->levs) < N) d->levs++;
    d->N2 = (int64_t)1 << d->levs;
    d->v = (double *)calloc((size_t)d->N2 + 2, 8);
    d->ps = (double *)calloc((size_t)d->N2 + 2, 8);
    for (int64_t i = 1; i <= N; i++) d->v[i] = v[i - 1];
    ds_refresh(d);
    return d;
}
static void ds_free(dynsmp *d) { free(d->v); free(d->ps); free(d); }
static int64_t ds_getel(dynsmp *d, double x, int *err) /* DynamicSamplers.jl:130-152 */
{
    x *= d->z;
    int64_t k = 0, off = 1;
    for (int lev = 1; lev <= d->levs; lev++) {
        double p = d->ps[off + k];
        k *= 2;
        if (x > p) { x -= p; k += 1; }
        off *= 2;
    }
    if (k >= d->N || d->v[k + 1] == 0) {
        if (!(d->trefresh > 0)) { *err = 1; return 1; }
        ds_refresh(d);
        return ds_getel(d, x, err); /* sic: the reference recurses with the already-scaled residual x */
    }
    return k + 1;
}
static void ds_set(dynsmp *d, int64_t i, double x) /* DynamicSamplers.jl:159-176 */
{
    if (d->trefresh >= (d->N > 100 ? d->N : 100)) ds_refresh(d);
    d->trefresh++;
    double dd = x - d->v[i];
    d->v[i] = x;
    d->z += dd;
    ds_add_path(d, i, dd);
}

/* Test probes (tests/test_oracle_pins.py): the two container types exposed on their own, so that their state can be
 * compared with values derived by hand from the reference source. */
/* DynamicSampler(v) -> apply sets (1-based index, value) -> ps[1..N2-1], z; getel for each query x in [0,1) */
int orc_ds_probe(int64_t N, const double *v, int64_t nset, const int64_t *set_i, const double *set_x,
                 double *ps_out, double *z_out, int64_t nq, const double *xq, int64_t *el_out)
{
    dynsmp *d = ds_new(N, v);
    for (int64_t a = 0; a < nset; a++) ds_set(d, set_i[a], set_x[a]);
    for (int64_t k = 1; k < d->N2; k++) ps_out[k - 1] = d->ps[k];
    *z_out = d->z;
    int err = 0;
    for (int64_t q = 0; q < nq && !err; q++) el_out[q] = ds_getel(d, xq[q], &err);
    ds_free(d);
    return err;
}
/* ArraySet(N): op[a] > 0 push!(op[a]), op[a] < 0 delete!(-op[a]) -> v[1..t] in storage order, t */
int64_t orc_arrayset_probe(int64_t N, int64_t nops, const int64_t *op, int64_t *v_out)
{
    arrayset a; as_init(&a, N);
    for (int64_t k = 0; k < nops; k++) { if (op[k] > 0) as_push(&a, op[k]); else as_delete(&a, -op[k]); }
    const int64_t t = as_check(&a) ? -1 : a.t;
    for (int64_t k = 1; k <= a.t; k++) v_out[k - 1] = a.v[k];
    as_free(&a);
    return t;
}

/* ------------------------------------------------------------------------------------------
 * DeltaECacheCont — src/DeltaE.jl:297-410
 * ---------------------------------------------------------------------------------------- */
static inline double prior(double x) { return x > 0 ? exp(-x) : 1.0; } /* DeltaE.jl:297 */
typedef struct { dynsmp *ds; double *dEs; double beta; int64_t *sj; double *sdE, *sp; int64_t nstaged; int64_t *nb; } cocache;

static cocache *cocache_new(orc_graph *X, const uint64_t *s, double beta) /* DeltaE.jl:304-311 */
{
    cocache *c = (cocache *)calloc(1, sizeof(cocache));
    int64_t N = X->N;
    c->dEs = (double *)malloc((size_t)(N + 1) * 8);
    double *p = (double *)malloc((size_t)N * 8);
    for (int64_t i = 1; i <= N; i++) { c->dEs[i] = orc_delta_energy(X, s, i); p[i - 1] = prior(beta * c->dEs[i]); }
    c->ds = ds_new(N, p);
    free(p);
    c->beta = beta;
    c->sj = (int64_t *)malloc((size_t)(N + 1) * 8); c->sdE = (double *)malloc((size_t)(N + 1) * 8); c->sp = (double *)malloc((size_t)(N + 1) * 8);
    c->nb = (int64_t *)malloc((size_t)(N + 2) * 8);
    return c;
}
static void cocache_free(cocache *c) { ds_free(c->ds); free(c->dEs); free(c->sj); free(c->sdE); free(c->sp); free(c->nb); free(c); }
static int64_t c_rand_skip(const cocache *c, orc_draws d) /* DeltaE.jl:319-324 */
{
    double b = c->ds->z / (double)c->ds->N;
    if (b < DBL_MIN) b = DBL_MIN;
    if (b > 1.0) b = 1.0;
    return (int64_t)floor(log1p(-d.f64(d.user)) / log1p(-b));
}
static int64_t c_rand_move(cocache *c, orc_draws d, double *dE, int *err) /* DeltaE.jl:326-332 */
{
    int64_t move = ds_getel(c->ds, d.f64(d.user), err);
    *dE = c->dEs[move];
    return move;
}
static void c_compute_staged(orc_graph *X, uint64_t *s, int64_t i, cocache *c) /* DeltaE.jl:356-373 */
{
    orc_spinflip(X, s, i);
    c->nstaged = 0;
    double dE = orc_delta_energy(X, s, i);
    c->sj[0] = i; c->sdE[0] = dE; c->sp[0] = prior(c->beta * dE); c->nstaged = 1;
    int n = orc_neighbors(X, i, c->nb);
    for (int a = 0; a < n; a++) {
        int64_t j = c->nb[a];
        dE = orc_delta_energy(X, s, j);
        c->sj[c->nstaged] = j; c->sdE[c->nstaged] = dE; c->sp[c->nstaged] = prior(c->beta * dE); c->nstaged++;
    }
    orc_spinflip(X, s, i);
}
static double c_reverse_probs(cocache *c) /* DeltaE.jl:344-354 */
{
    double z = c->ds->z;
    for (int64_t a = 0; a < c->nstaged; a++) z += c->sp[a] - c->ds->v[c->sj[a]];
    if (z < DBL_MIN) z = DBL_MIN;
    if (z > (double)c->ds->N) z = (double)c->ds->N;
    return z;
}
static void c_apply_staged(cocache *c) /* DeltaE.jl:334-342 */
{
    for (int64_t a = 0; a < c->nstaged; a++) { c->dEs[c->sj[a]] = c->sdE[a]; ds_set(c->ds, c->sj[a], c->sp[a]); }
}
static double c_apply_move(orc_graph *X, uint64_t *s, int64_t move, cocache *c, int inner) /* DeltaE.jl:378-410 */
{
    orc_spinflip(X, s, move);
    orc_graph *X0 = inner ? orc_inner_graph(X) : X;
    double z = c->ds->z;
    double dE = orc_delta_energy(X0, s, move);
    c->dEs[move] = dE;
    ds_set(c->ds, move, prior(c->beta * dE));
    int n = orc_neighbors(X0, move, c->nb);
    for (int a = 0; a < n; a++) {
        int64_t j = c->nb[a];
        dE = orc_delta_energy(X0, s, j);
        c->dEs[j] = dE;
        ds_set(c->ds, j, prior(c->beta * dE));
    }
    return z / c->ds->z;
}

/* ------------------------------------------------------------------------------------------
 * Samplers — src/RRRMC.jl
 * ---------------------------------------------------------------------------------------- */
static inline int accept1(double x, orc_draws d) { return x >= 0 || d.f64(d.user) < exp(x); } /* RRRMC.jl:39 */
static inline int accept2(double c, double x, orc_draws d)                                    /* RRRMC.jl:40-44 */
{
    if (c >= 1 && x >= 0) return 1;
    double a = c * exp(x);
    return a >= 1 || d.f64(d.user) < a;
}
#define PUSH_SAMPLE()                                                              \
    do {                                                                           \
        if (res.nsamples < Es_cap && Es) Es[res.nsamples] = E;                     \
        res.nsamples++;                                                            \
    } while (0)

/* standardMC — RRRMC.jl:81-127 */
orc_result orc_standardMC(orc_graph *X, double beta, int64_t iters, int64_t step, uint64_t *s,
                          orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_result res = { 0, 0, 0, 0, 0 };
    int64_t N = X->N;
    double E = orc_energy(X, s);
    int64_t accepted = 0, it = 0;
    while (it < iters) {
        it++;
        if (it % step == 0) {
            PUSH_SAMPLE();
            if (hook && !hook(user, it, E, accepted)) break;
        }
        int64_t i = d.range(d.user, N);
        double dE = orc_delta_energy(X, s, i);
        if (!accept1(-beta * dE, d)) continue;
        orc_spinflip(X, s, i);
        E += dE;
        accepted++;
    }
    res.iters_done = it; res.accepted = accepted;
    return res;
}

/* rrrMC(::SingleGraph) RRRMC.jl:149-219 and rrrMC(::DoubleGraph) RRRMC.jl:221-290 */
orc_result orc_rrrMC(orc_graph *X, double beta, int64_t iters, int64_t step, uint64_t *s,
                     orc_draws d, double staged_thr, double staged_thr_fact,
                     orc_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_result res = { 0, 0, 0, 0, 0 };
    if (!isfinite(beta)) { res.status = -1; return res; }         /* ArgumentError :159/:230 */
    int dbl = is_double(X);
    orc_graph *X0 = orc_inner_graph(X);
    int discr = is_discr(X0);
    if (dbl && !discr) { res.status = -2; return res; }           /* not on this path */
    if (isnan(staged_thr)) staged_thr = dbl ? 0.5 : (discr ? 0.5 : 0.8); /* :163-165, :226 */
    int64_t N = X->N;
    double E = orc_energy(X, s);
    if (dbl && !isfinite(E)) { res.status = -3; return res; }     /* @assert isfinite(E) :238 */
    decache *dc = discr ? decache_new(X0, s, beta, 1) : NULL;     /* gen_ΔEcache :171/:240 */
    cocache *cc = discr ? NULL : cocache_new(X0, s, beta);
    double lambda = staged_thr_fact / (double)N;
    int64_t staged_its = 0, it = 0, accepted = 0;
    double acc_rate = 0.5;
    int err = 0;
    while (it < iters && !err) {
        it++;
        if (it % step == 0) {
            PUSH_SAMPLE();
            if (hook && !hook(user, it, E, accepted)) break;
        }
        int acc = 0;
        if (acc_rate < staged_thr) {
            staged_its++;
            /* step_rrr — RRRMC.jl:131-138 */
            double z, zp, dE0; int64_t move;
            if (discr) { z = dc->z; move = d_rand_move(dc, d, &dE0); d_compute_staged(X0, s, move, dc); zp = d_reverse_probs(dc); }
            else       { z = cc->ds->z; move = c_rand_move(cc, d, &dE0, &err); c_compute_staged(X0, s, move, cc); zp = c_reverse_probs(cc); }
            double c = z / zp;
            int ok; double dE1 = 0.0;
            if (dbl) { dE1 = orc_delta_energy_residual(X, s, move); ok = accept2(c, -beta * dE1, d); }
            else ok = d.f64(d.user) < c;
            if (ok) {
                orc_spinflip(X, s, move);
                if (discr) d_apply_staged(dc); else c_apply_staged(cc);
                E += dE0 + dE1;
                accepted++; acc = 1;
            }
        } else {
            double dE0, dE1 = 0.0; int64_t move;
            if (discr) move = d_rand_move(dc, d, &dE0); else move = c_rand_move(cc, d, &dE0, &err);
            if (dbl) dE1 = orc_delta_energy_residual(X, s, move);
            double c = discr ? d_apply_move(X, s, move, dc) : c_apply_move(X, s, move, cc, 1);
            int ok = dbl ? accept2(c, -beta * dE1, d) : (d.f64(d.user) < c);
            if (ok) { E += dE0 + dE1; accepted++; acc = 1; }
            else { if (discr) d_apply_move(X, s, move, dc); else c_apply_move(X, s, move, cc, 1); }
        }
        acc_rate = acc_rate * (1 - lambda) + acc * lambda;
    }
    if (dc) decache_free(dc);
    if (cc) cocache_free(cc);
    res.iters_done = it; res.accepted = accepted; res.staged_its = staged_its; res.status = err ? -4 : 0;
    return res;
}

/* bklMC — RRRMC.jl:311-359 (apply_step_bkl! :294-298) */
orc_result orc_bklMC(orc_graph *X, double beta, int64_t iters, int64_t step, uint64_t *s,
                     orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_result res = { 0, 0, 0, 0, 0 };
    int discr = is_discr(X);
    double E = orc_energy(X, s);
    decache *dc = discr ? decache_new(X, s, beta, 0) : NULL;
    cocache *cc = discr ? NULL : cocache_new(X, s, beta);
    int64_t it = 0, accepted = 0, nextstep = step;
    int err = 0;
    while (it < iters && !err) {
        int64_t skip = discr ? d_rand_skip(dc, d) : c_rand_skip(cc, d);
        double dE; int64_t move = discr ? d_rand_move(dc, d, &dE) : c_rand_move(cc, d, &dE, &err);
        int out = 0;
        while (it + skip + 1 >= nextstep) {
            PUSH_SAMPLE();
            if (hook && !hook(user, nextstep, E, accepted)) { out = 1; break; }
            nextstep += step;
            if (nextstep > iters) { out = 1; break; }
        }
        if (out) break;
        if (discr) d_apply_move(X, s, move, dc); else c_apply_move(X, s, move, cc, 0);
        it += skip + 1;
        E += dE;
        accepted++;
    }
    if (dc) decache_free(dc);
    if (cc) cocache_free(cc);
    res.iters_done = it; res.accepted = accepted; res.status = err ? -4 : 0;
    return res;
}

/* ------------------------------------------------------------------------------------------
 * wtmMC — RRRMC.jl:376-430: the rejection-free waiting-time method (Dall & Sibani) on src/WaitingTimes.jl.
 * Every spin holds the absolute time of its next flip in a mutable binary min-heap (DataStructures.jl's
 * MutableBinaryMinHeap, a third-party dependency that is not vendored in the reference: only the minimum and the
 * update! of a handle's value are used, so any correct min-heap gives the same trajectory — flip times are
 * continuous and ties have probability zero). τ_i = max(1, exp(βΔE_i)) (WaitingTimes.jl:15-16), waiting time
 * -τ·log1p(-rand()) (:17-21). Draw order: N draws for the initial heap in site order (:25-35); per move one draw for
 * the moved spin, then one per neighbour in neighbors() order (:39-51). `step` is a Float64 global-time interval,
 * divided by N (:394); the hook receives the sample index k (the reference passes the time k·step/N).
 * ---------------------------------------------------------------------------------------- */
typedef struct { int64_t N; double *v; int64_t *node, *pos; } theap;   /* v[h], node[h] in heap order; pos[site] = h */
static void th_swap(theap *H, int64_t a, int64_t b)
{
    double tv = H->v[a]; H->v[a] = H->v[b]; H->v[b] = tv;
    int64_t tn = H->node[a]; H->node[a] = H->node[b]; H->node[b] = tn;
    H->pos[H->node[a]] = a; H->pos[H->node[b]] = b;
}
static void th_up(theap *H, int64_t h) { while (h > 0) { int64_t q = (h - 1) / 2; if (!(H->v[h] < H->v[q])) break; th_swap(H, h, q); h = q; } }
static void th_down(theap *H, int64_t h, int64_t n)
{
    for (;;) {
        int64_t l = 2 * h + 1, r = l + 1, m = h;
        if (l < n && H->v[l] < H->v[m]) m = l;
        if (r < n && H->v[r] < H->v[m]) m = r;
        if (m == h) break;
        th_swap(H, h, m); h = m;
    }
}
static void th_update(theap *H, int64_t site, double val)
{
    int64_t h = H->pos[site];
    double old = H->v[h];
    H->v[h] = val;
    if (val < old) th_up(H, h); else th_down(H, h, H->N);
}
static inline double wt_tau(double beta, double dE) { double e = exp(beta * dE); return e > 1.0 ? e : 1.0; } /* max(1.0, exp(βΔE)) */
static inline double wt_gen(double tau, orc_draws d) { return -tau * log1p(-d.f64(d.user)); }

orc_result orc_wtmMC(orc_graph *X, double beta, int64_t samples, double step, uint64_t *s,
                     orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_result res = { 0, 0, 0, 0, 0 };
    int64_t N = X->N;
    double E = orc_energy(X, s);
    theap H; H.N = N;
    H.v = (double *)malloc((size_t)N * 8); H.node = (int64_t *)malloc((size_t)N * 8); H.pos = (int64_t *)malloc((size_t)N * 8);
    int64_t *nb = (int64_t *)malloc((size_t)(N + 2) * 8);
    double *taus = (double *)malloc((size_t)N * 8);
    for (int64_t i = 1; i <= N; i++) taus[i - 1] = wt_tau(beta, orc_delta_energy(X, s, i));   /* WaitingTimes.jl:28 */
    for (int64_t i = 0; i < N; i++) { H.v[i] = wt_gen(taus[i], d); H.node[i] = i; H.pos[i] = i; th_up(&H, i); } /* push! */
    free(taus);
    step /= (double)N;
    int64_t num_moves = 0;
    double tmax = step * (double)samples, t = 0.0, nextstep = step;
    while (t < tmax) {
        double tp = H.v[0]; int64_t move = H.node[0] + 1;                                      /* top_with_handle */
        int out = 0;
        while (tp >= nextstep) {
            PUSH_SAMPLE();
            if (hook && !hook(user, res.nsamples, E, num_moves)) { out = 1; break; }
            nextstep += step;
            if (nextstep > tmax + 1e-10) { out = 1; break; }
        }
        if (out) break;
        t = tp;
        double dE = orc_delta_energy(X, s, move);                                              /* update_heap! :39-51 */
        orc_spinflip(X, s, move);
        th_update(&H, move - 1, t + wt_gen(wt_tau(beta, -dE), d));
        int n = orc_neighbors(X, move, nb);
        for (int a = 0; a < n; a++) {
            int64_t j = nb[a];
            th_update(&H, j - 1, t + wt_gen(wt_tau(beta, orc_delta_energy(X, s, j)), d));
        }
        E += dE;
        num_moves++;
    }
    free(H.v); free(H.node); free(H.pos); free(nb);
    res.iters_done = num_moves; res.accepted = num_moves;
    return res;
}

/* ------------------------------------------------------------------------------------------
 * extremal_opt — RRRMC.jl:468-521 on the EOCache of DeltaE.jl:413-543 (τ-EO, Boettcher & Percus): the spins are
 * ranked by ΔE (classes in ascending ΔE, findks :413-422: K = 2L − has_zero classes, ΔE = 0 is ONE class here, unlike
 * DeltaECache), a rank i is drawn from the power law j^−τ (r = (1 − rand())·z, i = searchsortedfirst(fτ, r), :483-487),
 * the class holding rank i is found by walking the class sizes (:500-505) and a uniform member of it is flipped
 * unconditionally (:514); neighbours and the moved spin are re-classified (apply_move! :519-543). DiscrGraph only
 * (the generic EOCacheCont re-sorts all N spins per move, :545-635 — "very sub-optimal" in the reference's words).
 * fτ = cumsum(j^−τ) is an input: the host language computes it (Julia's cumsum is pairwise, its `^` is its own pow).
 * Draw order per move: one f64, one range(|class|).
 * ---------------------------------------------------------------------------------------- */
static int eo_findks(const decache *c, double dE, int has_zero) /* DeltaE.jl:413-422 */
{
    int ak = findk(c, dE);
    return dE >= 0 ? ak + c->L - has_zero : c->L + 1 - ak;
}
/* extremal_opt on EOCacheCont (DeltaE.jl:555-635): the graphs that are not DiscrGraph (GraphEANormal, GraphSKNormal, ...).
 * ΔEs[i] = delta_energy(X, C, i) for every spin, rank = sortperm(ΔEs) (:562-563); rand_move picks rank[i] with
 * i = searchsortedfirst(fτ, (1 - rand())·z) (:575-587) — ONE draw per move, no member draw; apply_move! flips, refreshes
 * ΔEs of the move and its neighbours and re-sorts (sortperm!(rank, ΔEs, initialized=true), :589-606). With continuous
 * couplings two spins never share a ΔE, so the sorted order is unique and rankshuffle! (:608-633, a shuffle inside groups
 * of equal ΔE that would consume further draws) does nothing; this restatement keeps equal values in their current
 * order (a stable insertion pass) and so differs from the reference only on ties. */
static orc_eo_result orc_extremal_opt_cont(orc_graph *X, const double *ftau, int64_t iters, int64_t step, uint64_t *s,
                                           uint64_t *Cmin, orc_draws d, orc_eo_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_eo_result res = { 0, 0, 0, 0.0, 0, 0.0 };
    const int64_t N = X->N, nch = (N + 63) / 64;
    double E = orc_energy(X, s), Emin = E;
    int64_t itmin = 0, it = 0;
    if (Cmin) memcpy(Cmin, s, (size_t)nch * 8);
    double *dEs = (double *)malloc((size_t)N * 8);
    int64_t *rank = (int64_t *)malloc((size_t)N * 8), *nb = (int64_t *)malloc((size_t)(N + 2) * 8);
    for (int64_t i = 0; i < N; i++) { dEs[i] = orc_delta_energy(X, s, i + 1); rank[i] = i; }
    for (int64_t p = 1; p < N; p++) {                     /* sortperm: stable insertion sort (ascending ΔE) */
        int64_t key = rank[p], q = p - 1; double kv = dEs[key];
        while (q >= 0 && dEs[rank[q]] > kv) { rank[q + 1] = rank[q]; q--; }
        rank[q + 1] = key;
    }
    const double z = ftau[N - 1];
    while (it < iters) {
        it++;
        if (it % step == 0) {
            if (res.nsamples < Es_cap && Es) Es[res.nsamples] = E;
            res.nsamples++;
            if (hook && !hook(user, it, E, Emin)) break;
        }
        const double r = (1 - d.f64(d.user)) * z;        /* rand_move, DeltaE.jl:575-587 */
        int64_t lo = 0, hi = N;
        while (lo < hi) { int64_t m = (lo + hi) >> 1; if (ftau[m] < r) lo = m + 1; else hi = m; }
        const int64_t i = lo + 1;
        if (i < 1 || i > N) { res.status = -4; break; }
        const int64_t move = rank[i - 1] + 1;
        const double dE = dEs[move - 1];
        orc_spinflip(X, s, move);                         /* apply_move!, DeltaE.jl:589-606 */
        dEs[move - 1] = orc_delta_energy(X, s, move);
        const int n = orc_neighbors(X, move, nb);
        for (int a = 0; a < n; a++) dEs[nb[a] - 1] = orc_delta_energy(X, s, nb[a]);
        for (int64_t p = 1; p < N; p++) {                 /* sortperm!(rank, ΔEs, initialized=true) */
            int64_t key = rank[p], q = p - 1; double kv = dEs[key];
            while (q >= 0 && dEs[rank[q]] > kv) { rank[q + 1] = rank[q]; q--; }
            rank[q + 1] = key;
        }
        E += dE;
        if (E < Emin) { Emin = E; itmin = it; if (Cmin) memcpy(Cmin, s, (size_t)nch * 8); }
    }
    free(dEs); free(rank); free(nb);
    res.iters_done = it; res.itmin = itmin; res.Emin = Emin; res.Efinal = E;
    return res;
}
orc_eo_result orc_extremal_opt(orc_graph *X, const double *ftau, int64_t iters, int64_t step, uint64_t *s,
                               uint64_t *Cmin, orc_draws d, orc_eo_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_eo_result res = { 0, 0, 0, 0.0, 0, 0.0 };
    if (!is_discr(X)) return orc_extremal_opt_cont(X, ftau, iters, step, s, Cmin, d, hook, user, Es, Es_cap);
    const int64_t N = X->N, nch = (N + 63) / 64;
    decache c0; memset(&c0, 0, sizeof c0);
    decache *c = &c0;
    c->N = N; c->L = orc_allDE(X, c->DE);
    const int L = c->L, has_zero = c->DE[0] == 0.0, K = 2 * L - has_zero;
    double E = orc_energy(X, s), Emin = E;
    int64_t itmin = 0, it = 0;
    if (Cmin) memcpy(Cmin, s, (size_t)nch * 8);
    arrayset *as = (arrayset *)calloc((size_t)K + 1, sizeof(arrayset));
    int64_t *pos = (int64_t *)calloc((size_t)N + 1, 8), nb[64];
    for (int k = 1; k <= K; k++) as_init(&as[k], N);
    for (int64_t i = 1; i <= N; i++) {                    /* EOCache ctor, DeltaE.jl:433-441 */
        int ki = eo_findks(c, orc_delta_energy(X, s, i), has_zero);
        pos[i] = ki; as_push(&as[ki], i);
    }
    const double z = ftau[N - 1];
    while (it < iters) {
        it++;
        if (it % step == 0) {
            if (res.nsamples < Es_cap && Es) Es[res.nsamples] = E;
            res.nsamples++;
            if (hook && !hook(user, it, E, Emin)) break;
        }
        /* rand_move, DeltaE.jl:480-517 */
        const double r = (1 - d.f64(d.user)) * z;
        int64_t lo = 0, hi = N;                          /* searchsortedfirst: first i (1-based) with fτ[i] >= r */
        while (lo < hi) { int64_t m = (lo + hi) >> 1; if (ftau[m] < r) lo = m + 1; else hi = m; }
        const int64_t i = lo + 1;
        if (i < 1 || i > N) { res.status = -4; break; }  /* @assert 1 ≤ i ≤ N */
        int k = 0; int64_t t = 0;
        while (i > t) { k++; t += as[k].t; }
        const double dE = k <= L ? -c->DE[L - k] : c->DE[k - L + has_zero - 1];
        const int64_t move = as[k].v[d.range(d.user, as[k].t)];
        /* apply_move!, DeltaE.jl:519-543 */
        orc_spinflip(X, s, move);
        const int n = orc_neighbors(X, move, nb);
        for (int a = 0; a <= n; a++) {
            const int64_t j = a < n ? nb[a] : move;
            const int k0 = (int)pos[j], k1 = eo_findks(c, orc_delta_energy(X, s, j), has_zero);
            if (k0 == k1) continue;
            as_delete(&as[k0], j); as_push(&as[k1], j); pos[j] = k1;
        }
        E += dE;
        if (E < Emin) { Emin = E; itmin = it; if (Cmin) memcpy(Cmin, s, (size_t)nch * 8); }
    }
    for (int k = 1; k <= K; k++) { if (!res.status && as_check(&as[k])) res.status = -5; as_free(&as[k]); }
    free(as); free(pos);
    res.iters_done = it; res.itmin = itmin; res.Emin = Emin; res.Efinal = E;
    return res;
}

/* ------------------------------------------------------------------------------------------
 * CPU model of the engine's checkerboard Metropolis (new-engine feature).
 * Deliberately scalar: per (site, replica) ΔE from the ±J definition (EA.jl:277-289 naive form),
 * acceptance = Metropolis (RRRMC.jl:39) with U drawn by the engine's per-task bit procedure:
 *   task=(site i, group g of 128 replicas, sweep t); call q of the task is
 *   Philox4x32-10(ctr=(q | (t>>32)<<16, i, g, t&0xffffffff), key=seed).
 *   Plane phase (q=0..K-1): bit b of word w of call q is the q-th most significant bit of U for
 *   replica 128g+32w+b; lanes are decided at the first bit where U differs from the 64-bit
 *   fixed-point threshold T_c.  Tail: the lanes still undecided after K planes are visited in
 *   ascending (w,b) order; the n-th one takes word n%4 of call K+n/4 as the next 32 bits of U
 *   and accepts iff it is < bits [63-K .. 32-K] of T_c.
 * ---------------------------------------------------------------------------------------- */
void orc_checkerboard_sweeps(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                             const uint64_t *thr, int K, int M, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                             int64_t *accepted)
{
    int64_t N = 1; for (int d = 0; d < D; d++) N *= L;
    int64_t W = R / 32, G = (R + 127) / 128;
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int64_t sw = 0; sw < nsweeps; sw++) {
        uint64_t t = sweep0 + (uint64_t)sw;
        for (int colour = 0; colour < 2; colour++)
            for (int64_t i = 0; i < N; i++) {
                int64_t co[3] = { 0, 0, 0 }, rem = i, par = 0;
                for (int d = 0; d < D; d++) { co[d] = rem % L; rem /= L; par += co[d]; }
                if ((par & 1) != colour) continue;
                /* neighbour sites and couplings */
                int64_t nbr[6]; int Jn[6]; int64_t stride = 1;
                for (int d = 0; d < D; d++) {
                    int64_t up = i + (((co[d] + 1) % L) - co[d]) * stride;
                    int64_t dn = i + (((co[d] + L - 1) % L) - co[d]) * stride;
                    nbr[2 * d] = up; Jn[2 * d] = Jfwd[i * D + d];
                    nbr[2 * d + 1] = dn; Jn[2 * d + 1] = Jfwd[dn * D + d];
                    stride *= L;
                }
                for (int64_t g = 0; g < G; g++) {
                    int cls[128]; /* 0: ΔE≤0 (always flip); c≥1: ΔE = 4c */
                    int und[128];
                    for (int l = 0; l < 128; l++) {
                        int64_t r = 128 * g + l; cls[l] = -1; und[l] = 0;
                        if (r >= R) continue;
                        int64_t w = r >> 5; int b = (int)(r & 31);
                        int sc = (spins[i * W + w] >> b) & 1, acc = 0;
                        for (int k = 0; k < 2 * D; k++) {
                            int sk = (spins[nbr[k] * W + w] >> b) & 1;
                            acc += Jn[k] * (2 * sc - 1) * (2 * sk - 1);
                        }
                        int dE = 2 * acc;
                        cls[l] = dE <= 0 ? 0 : dE / 4;
                        und[l] = cls[l] > 0;
                    }
                    int flip[128]; for (int l = 0; l < 128; l++) flip[l] = (cls[l] == 0);
                    uint32_t ctr[4], out[4];
                    ctr[1] = (uint32_t)i; ctr[2] = (uint32_t)g; ctr[3] = (uint32_t)t;
                    for (int q = 0; q < K; q++) {
                        int any = 0; for (int l = 0; l < 128; l++) any |= und[l];
                        if (!any) break;
                        ctr[0] = (uint32_t)q | ((uint32_t)(t >> 32) << 16);
                        orc_philox4x32_10(ctr, key, out);
                        for (int l = 0; l < 128; l++) {
                            if (!und[l]) continue;
                            int ub = (out[l >> 5] >> (l & 31)) & 1;
                            int tb = (int)((thr[cls[l] - 1] >> (63 - q)) & 1);
                            if (ub != tb) { und[l] = 0; flip[l] = (ub < tb); }
                        }
                    }
                    /* merged planes: still-undecided lanes are sparse, so the four words are overlaid on one:
                     * at bit position b the lowest word with an undecided lane is "merged" and reads bit b of
                     * word j%4 of call K+j/4 as bit K+j of its U; lanes shadowed at their position keep K bits. */
                    int merged[128];
                    for (int l = 0; l < 128; l++) merged[l] = 0;
                    if (M > 0) {
                        for (int b = 0; b < 32; b++)
                            for (int w = 0; w < 4; w++)
                                if (und[32 * w + b]) { merged[32 * w + b] = 1; break; }
                        for (int j = 0; j < M; j++) {
                            if ((j & 3) == 0) { ctr[0] = (uint32_t)(K + j / 4) | ((uint32_t)(t >> 32) << 16); orc_philox4x32_10(ctr, key, out); }
                            for (int l = 0; l < 128; l++) {
                                if (!und[l] || !merged[l]) continue;
                                int ub = (out[j & 3] >> (l & 31)) & 1;
                                int tb = (int)((thr[cls[l] - 1] >> (63 - K - j)) & 1);
                                if (ub != tb) { und[l] = 0; flip[l] = (ub < tb); }
                            }
                        }
                    }
                    /* tail: the n-th lane still undecided takes 32 fresh bits against its next 32 threshold bits */
                    int n = 0, call0 = K + (M + 3) / 4;
                    for (int l = 0; l < 128; l++) {
                        if (!und[l]) continue;
                        if ((n & 3) == 0) { ctr[0] = (uint32_t)(call0 + n / 4) | ((uint32_t)(t >> 32) << 16); orc_philox4x32_10(ctr, key, out); }
                        uint32_t V = out[n & 3];
                        int used = merged[l] ? K + M : K;
                        uint32_t rem32 = (uint32_t)((used ? (thr[cls[l] - 1] << used) : thr[cls[l] - 1]) >> 32);
                        flip[l] = V < rem32;
                        n++;
                    }
                    for (int l = 0; l < 128; l++) {
                        int64_t r = 128 * g + l;
                        if (r >= R || !flip[l]) continue;
                        spins[i * W + (r >> 5)] ^= (uint32_t)1 << (r & 31);
                        if (accepted) accepted[r]++;
                    }
                }
            }
    }
}

/* ---- "sparse" acceptance procedure ---------------------------------------------------------------------------
 * Random words of a task: call 0 = (A0..A3), call 1 = (B0..B3). S = B0 | B1<<32.
 *   S bits [0,40): eight 5-bit "static" slots; word w of class 1 takes slots 2w and 2w+1 for its first two draws
 *   S bits [40,61): the first three 7-bit slots of the task's overflow stream
 *   calls 2,3,...: words (0,1) of each form a 64-bit window holding nine more 7-bit overflow slots (LSB first)
 * The overflow stream serves, in this order: the 3rd, 4th... draws (and redraws) of class-1 words 0..3, then the
 * draws of class 2, then of class 3. */
typedef struct { uint32_t ctr[4], key[2]; uint64_t win; int left; uint32_t call; uint32_t hi16; } orc_slots;
static uint32_t orc_slot_next(orc_slots *s)
{
    if (s->left == 0) {
        uint32_t out[4];
        s->ctr[0] = s->call | s->hi16;
        orc_philox4x32_10(s->ctr, s->key, out);
        s->win = (uint64_t)out[0] | ((uint64_t)out[1] << 32);
        s->left = 9; s->call++;
    }
    uint32_t v = (uint32_t)(s->win & 127u);
    s->win >>= 7; s->left--;
    return v;
}

void orc_cb_sparse_tables(const uint64_t *thr, int D, uint32_t *tbl)
{
    for (int c = 1; c <= D; c++) {
        int n = c == 1 ? 32 : 128;
        uint32_t *T = c == 1 ? tbl : tbl + ORC_CB_T1 + (c - 2) * ORC_CB_TC;
        long double p = (long double)thr[c - 1] / 18446744073709551616.0L, q = 1.0L - p;
        long double pk = powl(q, (long double)n), cdf = 0.0L;
        for (int k = 0; k <= n; k++) {
            cdf += pk;
            /* "more than k lanes pass" iff x > T[k], x uniform on 32 bits: P = 1 - round(CDF·2^32)/2^32 */
            long double v = rintl(cdf * 4294967296.0L);
            T[k] = (k == n || v >= 4294967296.0L) ? 0xffffffffu : (v < 1.0L ? 0u : (uint32_t)(v - 1.0L));
            pk = q > 0.0L ? pk * (long double)(n - k) / (long double)(k + 1) * (p / q) : (k + 1 == n ? 1.0L : 0.0L);
        }
    }
}

void orc_checkerboard_sweeps_sparse(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                                    const uint32_t *tbl, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                                    int64_t *accepted)
{
    int64_t N = 1; for (int d = 0; d < D; d++) N *= L;
    int64_t W = R / 32, G = (R + 127) / 128;
    for (int64_t sw = 0; sw < nsweeps; sw++) {
        uint64_t t = sweep0 + (uint64_t)sw;
        for (int colour = 0; colour < 2; colour++)
            for (int64_t i = 0; i < N; i++) {
                int64_t co[3] = { 0, 0, 0 }, rem = i, par = 0;
                for (int d = 0; d < D; d++) { co[d] = rem % L; rem /= L; par += co[d]; }
                if ((par & 1) != colour) continue;
                int64_t nbr[6]; int Jn[6]; int64_t stride = 1;
                for (int d = 0; d < D; d++) {
                    int64_t up = i + (((co[d] + 1) % L) - co[d]) * stride;
                    int64_t dn = i + (((co[d] + L - 1) % L) - co[d]) * stride;
                    nbr[2 * d] = up; Jn[2 * d] = Jfwd[i * D + d];
                    nbr[2 * d + 1] = dn; Jn[2 * d + 1] = Jfwd[dn * D + d];
                    stride *= L;
                }
                for (int64_t g = 0; g < G; g++) {
                    /* pass[c][l] = 1: lane l passes the filter of class c (ΔE = 4c) in this attempt */
                    unsigned char pass[4][128];
                    memset(pass, 0, sizeof pass);
                    orc_slots st;
                    uint32_t A[4], B[4];
                    st.key[0] = (uint32_t)seed; st.key[1] = (uint32_t)(seed >> 32);
                    st.ctr[1] = (uint32_t)i; st.ctr[2] = (uint32_t)g; st.ctr[3] = (uint32_t)t;
                    st.hi16 = (uint32_t)(t >> 32) << 16;
                    st.ctr[0] = 0u | st.hi16; orc_philox4x32_10(st.ctr, st.key, A);
                    st.ctr[0] = 1u | st.hi16; orc_philox4x32_10(st.ctr, st.key, B);
                    uint64_t S = (uint64_t)B[0] | ((uint64_t)B[1] << 32);
                    st.win = S >> 40; st.left = 3; st.call = 2;
                    /* class 1: one binomial count per 32-lane word, uniform = word w of call 0 */
                    for (int w = 0; w < 4; w++) {
                        int s = 0, j = 0;
                        while (s < 32 && A[w] > tbl[s]) {
                            uint32_t pos = j < 2 ? (uint32_t)(S >> (5 * (2 * w + j))) & 31u : orc_slot_next(&st) & 31u;
                            j++;
                            if (pass[1][32 * w + pos]) continue; /* duplicate: redraw */
                            pass[1][32 * w + pos] = 1; s++;
                        }
                    }
                    /* classes 2..D: one count per task, uniform = word c of call 1 */
                    for (int c = 2; c <= D; c++) {
                        const uint32_t *T = tbl + ORC_CB_T1 + (c - 2) * ORC_CB_TC;
                        int s = 0;
                        while (s < 128 && B[c] > T[s]) {
                            uint32_t pos = orc_slot_next(&st);
                            if (pass[c][pos]) continue;
                            pass[c][pos] = 1; s++;
                        }
                    }
                    for (int l = 0; l < 128; l++) {
                        int64_t r = 128 * g + l;
                        if (r >= R) continue;
                        int64_t w = r >> 5; int b = (int)(r & 31);
                        int sc = (spins[i * W + w] >> b) & 1, acc = 0;
                        for (int k = 0; k < 2 * D; k++) {
                            int sk = (spins[nbr[k] * W + w] >> b) & 1;
                            acc += Jn[k] * (2 * sc - 1) * (2 * sk - 1);
                        }
                        int dE = 2 * acc;
                        int flip = dE <= 0 ? 1 : pass[dE / 4][l];
                        if (!flip) continue;
                        spins[i * W + w] ^= (uint32_t)1 << b;
                        if (accepted) accepted[r]++;
                    }
                }
            }
    }
}

/* ---- "poisson" acceptance procedure ------------------------------------------------------------------------------
 * Each lane of a task carries D independent Poisson hit processes: "level l" hits arrive with rate
 * lam_l - lam_{l+1}, lam_c = -log(1 - p_c), p_c = exp(-β·4c) (lam_{D+1} = 0). A lane of class c (ΔE = 4c) flips iff
 * it received a hit of level >= c: probability 1 - exp(-lam_c) = p_c, independently across lanes — exactly accept()
 * of RRRMC.jl:39. Hits are sampled per task (128 lanes): a Poisson COUNT per level by inverse CDF on a 32-bit
 * uniform, then one uniform 7-bit POSITION per hit, WITH replacement (two hits on one lane are harmless, so nothing
 * is ever redrawn and the common case has no data-dependent control flow).
 * Random words of a task (Philox call q has ctr = (q | (t>>32)<<16, site, group, t&0xffffffff), key = seed):
 *   call 0 = (X0, X1, P[0], P[1]); when NW > 2, call 1 = (P[2..5]).
 *   X0: count a of level-1 hits: a > k iff X0 > TA[k].
 *   X1: counts (b, c) of level-2 and level-3 hits: if X1 <= TC[0] then c = 0 and b > k iff X1 > TB0[k] (TB0 = the
 *       CDF of b rescaled to [0, TC[0]]); else c > k iff X1 > TC[k] and b > k iff Y > TB[k], Y a fresh uniform.
 *   static position slots: slot j = byte j&3 of P[j>>2], low 7 bits. Slots 0..NS-1 (NS = 4·NW-1) serve the first NS
 *       level-1 hits; the last slot (byte 3 of P[NW-1]) serves the first level-2 hit.
 *   overflow stream (rare): calls CO, CO+1, ... with CO = 1 if NW <= 2 else 2. Word 0 of call CO is Y; words 1..3 of
 *       every overflow call hold twelve byte slots (low 7 bits each, LSB byte first). The stream serves, in order, the
 *       level-1 hits NS.., the level-2 hits 1.., then all level-3 hits.
 * tbl = TA[64] | TB0[32] | TB[32] | TC[32]; each table ends with the value that stops its scan. */
typedef struct { uint32_t ctr[4], key[2], hi16, call, w[4], Y; int used, loaded; } orc_pstream;
static void orc_pstream_fetch(orc_pstream *s)
{
    s->ctr[0] = s->call | s->hi16;
    orc_philox4x32_10(s->ctr, s->key, s->w);
    if (!s->loaded) s->Y = s->w[0];
    s->loaded = 1; s->used = 0; s->call++;
}
static uint32_t orc_pstream_slot(orc_pstream *s)
{
    if (!s->loaded || s->used == 12) orc_pstream_fetch(s);
    uint32_t v = (s->w[1 + s->used / 4] >> (8 * (s->used % 4))) & 127u;
    s->used++;
    return v;
}
static uint32_t orc_pstream_Y(orc_pstream *s)
{
    if (!s->loaded) orc_pstream_fetch(s);
    return s->Y;
}

static void orc_poisson_table(long double mu, long double scale, uint32_t last, uint32_t *T, int n)
{
    long double pk = expl(-mu), cdf = 0.0L;
    for (int k = 0; k < n; k++) {
        cdf += pk;
        long double v = rintl(cdf * scale);              /* count > k iff x > T[k] */
        T[k] = (k == n - 1 || v > (long double)last) ? last : (v < 1.0L ? 0u : (uint32_t)(v - 1.0L));
        pk = pk * mu / (long double)(k + 1);
    }
}
void orc_cb_poisson_tables(const uint64_t *thr, int D, uint32_t *tbl)
{
    long double lam[5] = { 0, 0, 0, 0, 0 };
    for (int c = 1; c <= D; c++) lam[c] = -log1pl(-(long double)thr[c - 1] / 18446744073709551616.0L);
    uint32_t *TA = tbl, *TB0 = TA + ORC_CBP_KA, *TB = TB0 + ORC_CBP_KR, *TC = TB + ORC_CBP_KR;
    orc_poisson_table(128.0L * (lam[1] - lam[2]), 4294967296.0L, 0xffffffffu, TA, ORC_CBP_KA);
    orc_poisson_table(128.0L * (lam[2] - lam[3]), 4294967296.0L, 0xffffffffu, TB, ORC_CBP_KR);
    orc_poisson_table(128.0L * lam[3], 4294967296.0L, 0xffffffffu, TC, ORC_CBP_KR);
    orc_poisson_table(128.0L * (lam[2] - lam[3]), (long double)TC[0] + 1.0L, TC[0], TB0, ORC_CBP_KR);
}

/* tbl_stride = 0: one β, every 128-replica group reads tbl; tbl_stride = ORC_CBP_KA + 3 ORC_CBP_KR: a β ladder, group g
   reads tbl + g * tbl_stride (same Philox counters, same procedure, its own count tables) */
static void orc_checkerboard_sweeps_poisson_impl(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                                                 const uint32_t *tbl, int64_t tbl_stride, int NW, uint64_t seed, uint64_t sweep0,
                                                 int64_t nsweeps, int64_t *accepted)
{
    int64_t N = 1; for (int d = 0; d < D; d++) N *= L;
    int64_t W = R / 32, G = (R + 127) / 128;
    const int NS = 4 * NW - 1;
    for (int64_t sw = 0; sw < nsweeps; sw++) {
        uint64_t t = sweep0 + (uint64_t)sw;
        for (int colour = 0; colour < 2; colour++)
            for (int64_t i = 0; i < N; i++) {
                int64_t co[3] = { 0, 0, 0 }, rem = i, par = 0;
                for (int d = 0; d < D; d++) { co[d] = rem % L; rem /= L; par += co[d]; }
                if ((par & 1) != colour) continue;
                int64_t nbr[6]; int Jn[6]; int64_t stride = 1;
                for (int d = 0; d < D; d++) {
                    int64_t up = i + (((co[d] + 1) % L) - co[d]) * stride;
                    int64_t dn = i + (((co[d] + L - 1) % L) - co[d]) * stride;
                    nbr[2 * d] = up; Jn[2 * d] = Jfwd[i * D + d];
                    nbr[2 * d + 1] = dn; Jn[2 * d + 1] = Jfwd[dn * D + d];
                    stride *= L;
                }
                for (int64_t g = 0; g < G; g++) {
                    const uint32_t *TA = tbl + g * tbl_stride, *TB0 = TA + ORC_CBP_KA, *TB = TB0 + ORC_CBP_KR, *TC = TB + ORC_CBP_KR;
                    int lvl[128];   /* highest level of a hit on the lane, 0 = none */
                    memset(lvl, 0, sizeof lvl);
                    orc_pstream st;
                    uint32_t X[4], P[6] = { 0, 0, 0, 0, 0, 0 };
                    memset(&st, 0, sizeof st);
                    st.key[0] = (uint32_t)seed; st.key[1] = (uint32_t)(seed >> 32);
                    st.ctr[1] = (uint32_t)i; st.ctr[2] = (uint32_t)g; st.ctr[3] = (uint32_t)t;
                    st.hi16 = (uint32_t)(t >> 32) << 16;
                    st.ctr[0] = 0u | st.hi16; orc_philox4x32_10(st.ctr, st.key, X);
                    P[0] = X[2]; P[1] = X[3];
                    if (NW > 2) { st.ctr[0] = 1u | st.hi16; orc_philox4x32_10(st.ctr, st.key, P + 2); }
                    st.call = NW > 2 ? 2u : 1u;
#define ORC_STATIC_SLOT(j) ((P[(j) >> 2] >> (8 * ((j) & 3))) & 127u)
                    int a = 0, b = 0, c = 0;
                    while (X[0] > TA[a]) a++;
                    for (int j = 0; j < a; j++) {
                        uint32_t pos = j < NS ? ORC_STATIC_SLOT(j) : orc_pstream_slot(&st);
                        if (lvl[pos] < 1) lvl[pos] = 1;
                    }
                    if (X[1] <= TC[0]) { while (X[1] > TB0[b]) b++; }
                    else {
                        while (X[1] > TC[c]) c++;
                        uint32_t Y = orc_pstream_Y(&st);
                        while (Y > TB[b]) b++;
                    }
                    for (int j = 0; j < b; j++) {
                        uint32_t pos = j == 0 ? ORC_STATIC_SLOT(NS) : orc_pstream_slot(&st);
                        if (lvl[pos] < 2) lvl[pos] = 2;
                    }
                    for (int j = 0; j < c; j++) lvl[orc_pstream_slot(&st)] = 3;
#undef ORC_STATIC_SLOT
                    for (int l = 0; l < 128; l++) {
                        int64_t r = 128 * g + l;
                        if (r >= R) continue;
                        int64_t w = r >> 5; int bb = (int)(r & 31);
                        int sc = (spins[i * W + w] >> bb) & 1, acc = 0;
                        for (int k = 0; k < 2 * D; k++) {
                            int sk = (spins[nbr[k] * W + w] >> bb) & 1;
                            acc += Jn[k] * (2 * sc - 1) * (2 * sk - 1);
                        }
                        int dE = 2 * acc;
                        int flip = dE <= 0 ? 1 : (lvl[l] >= dE / 4);
                        if (!flip) continue;
                        spins[i * W + w] ^= (uint32_t)1 << bb;
                        if (accepted) accepted[r]++;
                    }
                }
            }
    }
}

void orc_checkerboard_sweeps_poisson(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                                     const uint32_t *tbl, int NW, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                                     int64_t *accepted)
{
    orc_checkerboard_sweeps_poisson_impl(L, D, R, spins, Jfwd, tbl, 0, NW, seed, sweep0, nsweeps, accepted);
}
/* β ladder: tbls[G][ORC_CBP_KA + 3 ORC_CBP_KR], one table set per 128-replica group */
void orc_checkerboard_sweeps_poisson_ladder(int L, int D, int64_t R, uint32_t *spins, const int8_t *Jfwd,
                                            const uint32_t *tbls, int NW, uint64_t seed, uint64_t sweep0, int64_t nsweeps,
                                            int64_t *accepted)
{
    orc_checkerboard_sweeps_poisson_impl(L, D, R, spins, Jfwd, tbls, ORC_CBP_KA + 3 * ORC_CBP_KR, NW, seed, sweep0, nsweeps, accepted);
}

/* ------------------------------------------------------------------------------------------
 * Checkerboard Metropolis for continuous couplings (GraphEANormal, EA.jl:534-680), CPU model of
 * rrrmc.jl_b200/csrc/ea_normal.cu. Same two-colour schedule as the ±J kernels; per (site, replica):
 *   lf = 0; for k = 1..2D: lf -= J[x][k]·σx·σy  (the slot order of energy(), EA.jl:590-603); ΔE = -2·lf
 *   accept(-βΔE): ΔE <= 0 flips, else u < exp(-βΔE) (RRRMC.jl:39), u = 53-bit uniform from
 *   Philox4x32-10(counter = (t_hi<<16, site, replica, t_lo), key = seed): (y:x) >> 11 · 2^-53.
 * A, J: the reference layout (1-based neighbours, slot-aligned couplings). beta: per replica.
 * ---------------------------------------------------------------------------------------- */
void orc_checkerboard_sweeps_f64(int L, int D, int64_t R, uint32_t *spins, const int64_t *A, const double *J,
                                 const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps, int64_t *accepted)
{
    int64_t N = 1; for (int d = 0; d < D; d++) N *= L;
    const int64_t W = (R + 31) / 32; const int twoD = 2 * D;
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int64_t sw = 0; sw < nsweeps; sw++) {
        const uint64_t t = sweep0 + (uint64_t)sw;
        for (int colour = 0; colour < 2; colour++)
            for (int64_t i = 0; i < N; i++) {
                int64_t rem = i, par = 0;
                for (int d = 0; d < D; d++) { par += rem % L; rem /= L; }
                if ((par & 1) != colour) continue;
                for (int64_t r = 0; r < R; r++) {
                    const int64_t w = r >> 5; const int bb = (int)(r & 31);
                    const int sx = (spins[i * W + w] >> bb) & 1;
                    double lf = 0.0;
                    for (int k = 0; k < twoD; k++) {
                        const int64_t y = A[i * twoD + k] - 1;
                        const int sy = (spins[y * W + w] >> bb) & 1;
                        lf -= J[i * twoD + k] * (double)((2 * sx - 1) * (2 * sy - 1));
                    }
                    const double dE = -2.0 * lf, x = -beta[r] * dE;
                    int flip = x >= 0;
                    if (!flip) {
                        uint32_t ctr[4] = { (uint32_t)(t >> 32) << 16, (uint32_t)i, (uint32_t)r, (uint32_t)t }, o[4];
                        orc_philox4x32_10(ctr, key, o);
                        const double u = (double)((((uint64_t)o[1] << 32) | o[0]) >> 11) * 0x1.0p-53;
                        flip = u < exp(x);
                    }
                    if (!flip) continue;
                    spins[i * W + w] ^= (uint32_t)1 << bb;
                    if (accepted) accepted[r]++;
                }
            }
    }
}

/* ------------------------------------------------------------------------------------------
 * Lock-step Metropolis sweeps on GraphSKNormal, CPU model of rrrmc.jl_b200/csrc/sk_dense.cu:k_sk_lockstep*.
 * Every sweep visits the sites in order 1..N (the same site for every replica of the batch); per (site, replica):
 *   ΔE = lfields[i]                                   (delta_energy, SK.jl:278-284)
 *   accept(-βΔE): ΔE <= 0 flips, else u < exp(-βΔE)   (RRRMC.jl:39), u = 53-bit uniform from
 *     Philox4x32-10(counter = (site, replica, t_lo, t_hi ^ "SKLS"), key = seed): (y:x) >> 11 · 2^-53
 *   on a flip: E += ΔE, s[i] ^= 1, then update_cache! (SK.jl:252-265): lfields[j] += 4·(1 - 2(s_i ⊻ s_j))·J[i][j] for
 *   j != i with s_i the NEW spin, lfields[i] = -lfields[i].
 * chunks: [R][nchunks] packed spins (bit j of a replica = site j), lf: [R][N] fields, E / acc / beta: per replica.
 * ---------------------------------------------------------------------------------------- */
void orc_sk_lockstep_sweeps(int N, int64_t R, const double *J, uint64_t *chunks, int64_t nchunks, double *lf, double *E,
                            int64_t *acc, const double *beta, uint64_t seed, uint64_t sweep0, int64_t nsweeps)
{
    const uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    for (int64_t r = 0; r < R; r++) {
        uint64_t *s = chunks + r * nchunks; double *f = lf + r * (int64_t)N;
        for (int64_t sw = 0; sw < nsweeps; sw++) {
            const uint64_t t = sweep0 + (uint64_t)sw;
            for (int i = 0; i < N; i++) {
                const double dE = f[i], x = -beta[r] * dE;
                int flip = x >= 0;
                if (!flip) {
                    uint32_t ctr[4] = { (uint32_t)i, (uint32_t)r, (uint32_t)t, (uint32_t)(t >> 32) ^ 0x534b4c53u }, o[4];
                    orc_philox4x32_10(ctr, key, o);
                    const double u = (double)((((uint64_t)o[1] << 32) | o[0]) >> 11) * 0x1.0p-53;
                    flip = u < exp(x);
                }
                if (!flip) continue;
                E[r] += dE; acc[r]++;
                s[i >> 6] ^= 1ull << (i & 63);
                const int si = (int)((s[i >> 6] >> (i & 63)) & 1ull);
                const double *Ji = J + (int64_t)i * N;
                for (int j = 0; j < N; j++) {
                    if (j == i) continue;
                    const int sj = (int)((s[j >> 6] >> (j & 63)) & 1ull);
                    f[j] = f[j] + 4 * ((double)(1 - 2 * (si ^ sj)) * Ji[j]);
                }
                f[i] = -f[i];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Parallel-tempering exchange decisions, CPU model of rrrmc.jl_b200/csrc/tempering.cu:k_pt_decide.
 * A β ladder lies over the 128-replica groups of a batch: lane l of every group is one ladder, group g its rung.
 * Pairs (g, g+1) with g ≡ round (mod 2): a = replica 128 g + l, b = replica 128 (g+1) + l,
 *   ΔS = (β_g − β_{g+1})·(E_b − E_a); accept iff ΔS <= 0 or u < exp(−ΔS),
 *   u = ((y:x) >> 11)·2^-53 from Philox4x32-10(counter = (round_lo, round_hi, g, l), key = seed).
 * E[R]: energies (integers for ±J lattices, so E_b − E_a is exact). swap[(G-1)*128] = 1 where the pair exchanges.
 * The reference has no tempering (RRRMC.jl:81-127 runs one β); this is the rule of sharding.TemperingLadder.swap.
 * ---------------------------------------------------------------------------------------- */
void orc_tempering_decide(int64_t G, const double *beta_group, const double *E, uint64_t seed, uint64_t round, uint8_t *swap)
{
    memset(swap, 0, (size_t)((G - 1) * 128));
    for (int64_t g = (int64_t)(round & 1u); g + 1 < G; g += 2)
        for (int l = 0; l < 128; l++) {
            double dS = (beta_group[g] - beta_group[g + 1]) * (E[128 * (g + 1) + l] - E[128 * g + l]);
            int acc = dS <= 0.0;
            if (!acc) {
                uint32_t ctr[4] = { (uint32_t)round, (uint32_t)(round >> 32), (uint32_t)g, (uint32_t)l };
                uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) }, o[4];
                orc_philox4x32_10(ctr, key, o);
                double u = (double)((((uint64_t)o[1] << 32) | o[0]) >> 11) * 0x1.0p-53;
                acc = u < exp(-dS);
            }
            swap[g * 128 + l] = (uint8_t)acc;
        }
}

/* ------------------------------------------------------------------------------------------
 * Rank-select rrrMC / bklMC for GraphEA ±J: CPU model of rrrmc.jl_b200/csrc/chain_warp.cu.
 *
 * The same Markov chains as rrrMC (RRRMC.jl:149-219) and bklMC (RRRMC.jl:311-359) on the ΔE classes of DeltaE.jl:63-118,
 * with two implementation choices changed so that a warp can run one chain out of shared memory:
 *  (1) the member of class k that rand(1:t[k]) picks is the p-th member IN SITE ORDER (a rank query on the class
 *      bitmap) instead of the p-th entry of the ArraySet (ArraySets.jl:58-85, insertion order with hole filling).
 *      Either is a uniform pick among the t[k] members (DeltaE.jl:146-167 only needs that), so the chains have the
 *      same law; the trajectories for a given draw stream differ from the reference-order kernels.
 *  (2) the class weights are T[k] = t[k]·f(k) recomputed from the integer counts, their cumulative sums come from a
 *      fixed 8-lane scan (rk_scan; missing classes count 0) and z is its last element, instead of the reference's
 *      running sums (DeltaE.jl:184-200, 248-283), whose rounding depends on the order in which neighbours are re-filed.
 * Everything else is the reference's: class of a site from ΔE and its spin (DeltaE.jl:108-118), f(k) = 1 for the
 * down half, exp(-β·ΔE) for the up half (:83-95), the class scan of rand_move with its fallback (:146-167), the
 * acceptance rand() < z/z' (RRRMC.jl:131-138; evaluated as rand()·z' < z), rand_skip (DeltaE.jl:141-144), the sampling instants of the two drivers.
 * Draw order: rrrMC: rand() (class), rand(1:t[k]) (member), rand() (accept). bklMC: rand() (skip), rand(), rand(1:t[k]).
 * X: an ORC_EA_INT graph with couplings ±1 (any degree 2D <= 6, allΔE = 0, 4, .., 4D or the odd-degree set).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    orc_graph *X; const uint64_t *s; int64_t N; int L, twoD;
    int *u;          /* unsatisfied bonds of a site */
    int *cls;        /* class 1..2L */
    int64_t t[17]; double f[17], z;
} rk_cache;
static inline int rk_spin(const uint64_t *s, int64_t i) { return (int)((s[i >> 6] >> (i & 63)) & 1u); }
static int rk_class(const rk_cache *c, int u, int sb)
{
    /* ΔE = 2·(satisfied − unsatisfied) = 2·(2D − 2u); class index from |ΔE| through allΔE (DeltaE.jl:108-118) */
    double dE = 2.0 * (double)(c->twoD - 2 * u);
    double a = dE < 0 ? -dE : dE;
    int k = 0; while (k < c->L && c->X->DE[k] != a) k++;
    int up = dE > 0 || (dE == 0 && sb == 1);
    return k + 1 + c->L * up;
}
/* cumulative class weights cT[1..8] by the 8-lane Hillis-Steele scan the warp runs (steps 1, 2, 4): the association of
   every partial sum is fixed by the scan, cT[8] is the tree ((T1+T2)+(T3+T4))+((T5+T6)+(T7+T8)); returns z = cT[8] */
static double rk_scan(const rk_cache *c, const int64_t *t, double *cT)
{
    for (int k = 1; k <= 8; k++) cT[k] = k <= 2 * c->L ? (double)t[k] * c->f[k] : 0.0;
    for (int o = 1; o < 8; o <<= 1)
        for (int k = 8; k > o; k--) cT[k] = cT[k] + cT[k - o];
    return cT[8];
}
static double rk_z(const rk_cache *c, const int64_t *t) { double cT[9]; return rk_scan(c, t, cT); }
static rk_cache *rk_new(orc_graph *X, const uint64_t *s, double beta)
{
    rk_cache *c = (rk_cache *)calloc(1, sizeof *c);
    c->X = X; c->s = s; c->N = X->N; c->L = X->nDE; c->twoD = X->twoD;
    c->u = (int *)malloc((size_t)c->N * sizeof(int)); c->cls = (int *)malloc((size_t)c->N * sizeof(int));
    for (int k = 1; k <= 2 * c->L; k++) c->f[k] = k > c->L ? exp(-beta * X->DE[k - c->L - 1]) : 1.0;
    for (int64_t i = 0; i < c->N; i++) {
        int u = 0, si = rk_spin(s, i);
        for (int q = 0; q < c->twoD; q++) {
            int64_t y = X->A[i * c->twoD + q] - 1;
            int neg = X->Ji[i * c->twoD + q] < 0;
            u += (si ^ rk_spin(s, y)) ^ neg;
        }
        c->u[i] = u; c->cls[i] = rk_class(c, u, si); c->t[c->cls[i]]++;
    }
    c->z = rk_z(c, c->t);
    return c;
}
static void rk_free(rk_cache *c) { free(c->u); free(c->cls); free(c); }
static int64_t rk_rand_move(rk_cache *c, orc_draws d, double *dE)   /* DeltaE.jl:146-167, member = rank in site order */
{
    int L = c->L;
    double cT[9];
    rk_scan(c, c->t, cT);
    double r = d.f64(d.user) * c->z;
    int k = 1, broke = 0;
    for (; k <= 2 * L; k++) if (r < cT[k]) { broke = 1; break; }
    if (!broke) { k = 2 * L; while (c->t[k] == 0) k--; }
    *dE = k <= L ? -c->X->DE[k - 1] : c->X->DE[k - L - 1];
    int64_t p = d.range(d.user, c->t[k]);
    for (int64_t i = 0; i < c->N; i++) if (c->cls[i] == k && --p == 0) return i;
    return -1;
}
/* counts after flipping `move` (nothing is modified); returns z' */
static double rk_plan(const rk_cache *c, int64_t move, int64_t *tp)
{
    memcpy(tp, c->t, sizeof c->t);
    int sm = rk_spin(c->s, move);
    for (int q = 0; q < c->twoD; q++) {
        int64_t y = c->X->A[move * c->twoD + q] - 1;
        int neg = c->X->Ji[move * c->twoD + q] < 0, sy = rk_spin(c->s, y);
        int unsat = (sm ^ sy) ^ neg;
        int k1 = rk_class(c, c->u[y] + (unsat ? -1 : 1), sy);
        tp[c->cls[y]]--; tp[k1]++;
    }
    tp[c->cls[move]]--; tp[rk_class(c, c->twoD - c->u[move], sm ^ 1)]++;
    return rk_z(c, tp);
}
static void rk_commit(rk_cache *c, uint64_t *s, int64_t move, const int64_t *tp, double zp)
{
    int sm = rk_spin(s, move);
    for (int q = 0; q < c->twoD; q++) {
        int64_t y = c->X->A[move * c->twoD + q] - 1;
        int neg = c->X->Ji[move * c->twoD + q] < 0, sy = rk_spin(s, y);
        int unsat = (sm ^ sy) ^ neg;
        c->u[y] += unsat ? -1 : 1;
        c->cls[y] = rk_class(c, c->u[y], sy);
    }
    s[move >> 6] ^= (uint64_t)1 << (move & 63);
    c->u[move] = c->twoD - c->u[move];
    c->cls[move] = rk_class(c, c->u[move], sm ^ 1);
    memcpy(c->t, tp, sizeof c->t);
    c->z = zp;
}
orc_result orc_rank_rrrMC(orc_graph *X, double beta, int64_t iters, int64_t step, uint64_t *s,
                          orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_result res = { 0, 0, 0, 0, 0 };
    if (!isfinite(beta) || X->kind != ORC_EA_INT) { res.status = -1; return res; }
    double E = orc_energy(X, s);
    rk_cache *c = rk_new(X, s, beta);
    int64_t it = 0, accepted = 0, tp[17];
    while (it < iters) {
        it++;
        if (it % step == 0) {
            PUSH_SAMPLE();
            if (hook && !hook(user, it, E, accepted)) break;
        }
        double dE0, z = c->z;
        int64_t move = rk_rand_move(c, d, &dE0);
        double zp = rk_plan(c, move, tp);
        if (d.f64(d.user) * zp < z) { rk_commit(c, s, move, tp, zp); E += dE0; accepted++; }   /* rand() < z/z' (RRRMC.jl:131-138), as a product */
    }
    rk_free(c);
    res.iters_done = it; res.accepted = accepted; res.staged_its = it;
    return res;
}
orc_result orc_rank_bklMC(orc_graph *X, double beta, int64_t iters, int64_t step, uint64_t *s,
                          orc_draws d, orc_hook hook, void *user, double *Es, int64_t Es_cap)
{
    orc_result res = { 0, 0, 0, 0, 0 };
    if (!isfinite(beta) || X->kind != ORC_EA_INT) { res.status = -1; return res; }
    double E = orc_energy(X, s);
    rk_cache *c = rk_new(X, s, beta);
    int64_t it = 0, accepted = 0, nextstep = step, tp[17];
    while (it < iters) {
        int64_t skip = (int64_t)floor(log1p(-d.f64(d.user)) / log1p(-c->z / (double)c->N));   /* DeltaE.jl:141-144 */
        double dE; int64_t move = rk_rand_move(c, d, &dE);
        int out = 0;
        while (it + skip + 1 >= nextstep) {
            PUSH_SAMPLE();
            if (hook && !hook(user, nextstep, E, accepted)) { out = 1; break; }
            nextstep += step;
            if (nextstep > iters) { out = 1; break; }
        }
        if (out) break;
        double zp = rk_plan(c, move, tp);
        rk_commit(c, s, move, tp, zp);
        it += skip + 1;
        E += dE;
        accepted++;
    }
    rk_free(c);
    res.iters_done = it; res.accepted = accepted;
    return res;
}

/* ------------------------------------------------------------------------------------------
 * DFloat64 — src/DFloats.jl:11-62: "a Real type which is actually an integer in disguise", the value x is held as the
 * Int64 round(x·10^5) and every operation the graphs use is exact integer arithmetic.
 *   convert(DFloat64, x::Real) = round(Int64, x·dfact)   (:24, Julia's round = nearest, ties to even)
 *   Float64(x::DFloat64) = d2i(x) / dfact                 (:28)
 *   x ± y, -x: integer ±; Integer·x: integer product; x / Integer: integer quotient ÷ (truncating) (:30-39)
 * orc_dfloat_ea_energy restates energy(::GraphEA{DFloat64}) (EA.jl:195-222) with these operations on real-valued
 * couplings J (each converted once, like the graph constructor does, EA.jl:191): the engine's fractional-level graphs
 * (an integer-level graph in units of gcd/10^5) are tested against it.
 * ---------------------------------------------------------------------------------------- */
#define ORC_DFACT 100000
int64_t orc_dfloat_from_f64(double x) { return (int64_t)nearbyint(x * (double)ORC_DFACT); }   /* round-to-nearest-even mode */
double orc_dfloat_to_f64(int64_t d) { return (double)d / (double)ORC_DFACT; }
int64_t orc_dfloat_add(int64_t a, int64_t b) { return a + b; }
int64_t orc_dfloat_sub(int64_t a, int64_t b) { return a - b; }
int64_t orc_dfloat_mul_int(int64_t k, int64_t a) { return k * a; }
int64_t orc_dfloat_div_int(int64_t a, int64_t k) { return a / k; }                         /* ÷: truncation toward zero */
/* A: [N*twoD] 1-based neighbours (0 = no neighbour: ragged rows), J: [N*twoD] real couplings; -> Float64(energy) and,
   when lf2 != NULL, the local fields discr(ET, 2 lf) as DFloat64 integers (EA.jl:214) */
double orc_dfloat_ea_energy(int64_t N, int twoD, const int64_t *A, const double *J, const uint64_t *s, int64_t *lf2)
{
    int64_t n = 0;
    for (int64_t x = 0; x < N; x++) {
        int64_t sx = 2 * (int64_t)((s[x >> 6] >> (x & 63)) & 1u) - 1, lf = 0;
        for (int k = 0; k < twoD; k++) {
            int64_t y = A[x * twoD + k] - 1;
            if (y < 0) continue;
            int64_t sy = 2 * (int64_t)((s[y >> 6] >> (y & 63)) & 1u) - 1;
            lf = orc_dfloat_sub(lf, orc_dfloat_mul_int(sx * sy, orc_dfloat_from_f64(J[x * twoD + k])));
        }
        n = orc_dfloat_add(n, lf);
        if (lf2) lf2[x] = orc_dfloat_mul_int(2, lf);
    }
    return orc_dfloat_to_f64(orc_dfloat_div_int(n, 2));
}
