/*
 * Plain-C restatement of the reference's GRF sampler + per-length accumulation.
 *
 * TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  Only tests/, smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Follows (reference paths):
 *   efficient_graph_gp_sparse/random_walk_samplers_sparse/sparse_sampler.py:36-54
 *       the walk loop: accumulate (start, cur) += load; deg == 0 -> stop
 *       without drawing; draw u, stop if u < p_halt; draw k in [0, deg);
 *       load *= deg * w / (1 - p_halt)
 *   efficient_graph_gp/random_walk_samplers/sampler.py:163-184
 *       the two other load rules (assignment / ablation)
 *   sparse_sampler.py:117-130   COO -> CSR with sorted columns, "/ W" == "* (1/W)"
 *   sampler.py:196-201          "value / W"
 *
 * Per (step, start, node) the loads are added in walk order to a double that
 * starts at 0.0 -- the summation order of the reference's defaultdict(float).
 *
 * Draw sources: 0 = replay of a recorded trace (trace_u / trace_k indexed
 * [walk_id * L + step], walk_id = start * W + w), 1 = the native stream of the
 * CUDA walker: Philox4x32-10 (Salmon et al., SC'11), counter (walk_lo, walk_hi,
 * step / 2, 0), key (seed_lo, seed_hi); one block serves two steps: words (0, 1)
 * at the even step, (2, 3) at the odd one; first word < floor(p * 2^32) halts,
 * the neighbour index is (second word * deg) >> 32.
 *
 * Parity: pinned against reference-generated fixtures via
 * tests/test_oracle_golden.py (C path checked against the numpy path and the
 * golden vectors in tests/test_oracle_c.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t node;
    int32_t seq;
    double load;
} Rec;

typedef struct {
    int64_t n_local;
    int32_t L;
    int64_t visits;
    int64_t *nnz;      /* [L] */
    int64_t *cap;      /* [L] */
    int64_t **indptr;  /* [L][n_local + 1] */
    int32_t **indices; /* [L][nnz] */
    double **data;     /* [L][nnz] */
} GrfOracleResult;

static void philox_round(uint32_t c[4], const uint32_t k[2]) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    c[0] = n0;
    c[1] = (uint32_t)p1;
    c[2] = n2;
    c[3] = (uint32_t)p0;
}

void grf_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

static int rec_cmp(const void *a, const void *b) {
    const Rec *x = (const Rec *)a, *y = (const Rec *)b;
    if (x->node != y->node) return x->node < y->node ? -1 : 1;
    return x->seq < y->seq ? -1 : (x->seq > y->seq);
}

static void push_entry(GrfOracleResult *r, int step, int32_t col, double val) {
    if (r->nnz[step] == r->cap[step]) {
        int64_t nc = r->cap[step] ? r->cap[step] * 2 : 1024;
        r->indices[step] = (int32_t *)realloc(r->indices[step], (size_t)nc * sizeof(int32_t));
        r->data[step] = (double *)realloc(r->data[step], (size_t)nc * sizeof(double));
        r->cap[step] = nc;
    }
    r->indices[step][r->nnz[step]] = col;
    r->data[step][r->nnz[step]] = val;
    r->nnz[step]++;
}

GrfOracleResult *grf_oracle_build(int64_t n_nodes, const int32_t *indptr, const int32_t *indices, const double *data,
                                  int64_t start_lo, int64_t start_hi, int32_t W, int32_t L, double p_halt,
                                  int32_t draw_mode, uint64_t seed, const double *trace_u, const int32_t *trace_k,
                                  int32_t load_mode, int32_t scale_mode) {
    (void)n_nodes;
    GrfOracleResult *r = (GrfOracleResult *)calloc(1, sizeof(*r));
    int64_t n_local = start_hi - start_lo;
    r->n_local = n_local;
    r->L = L;
    r->nnz = (int64_t *)calloc((size_t)L, sizeof(int64_t));
    r->cap = (int64_t *)calloc((size_t)L, sizeof(int64_t));
    r->indptr = (int64_t **)calloc((size_t)L, sizeof(int64_t *));
    r->indices = (int32_t **)calloc((size_t)L, sizeof(int32_t *));
    r->data = (double **)calloc((size_t)L, sizeof(double *));
    for (int s = 0; s < L; ++s) r->indptr[s] = (int64_t *)calloc((size_t)n_local + 1, sizeof(int64_t));

    Rec *recs = (Rec *)malloc((size_t)L * (size_t)W * sizeof(Rec));
    int32_t *cnt = (int32_t *)malloc((size_t)L * sizeof(int32_t));
    const double one_minus_p = 1.0 - p_halt;
    double thr = floor(p_halt * 4294967296.0);
    if (thr < 0) thr = 0;
    if (thr > 4294967296.0) thr = 4294967296.0;
    const uint64_t halt_thr = (uint64_t)thr;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    const double recip = 1.0 / (double)W;

    for (int64_t start = start_lo; start < start_hi; ++start) {
        memset(cnt, 0, (size_t)L * sizeof(int32_t));
        for (int32_t w = 0; w < W; ++w) {
            const uint64_t walk_id = (uint64_t)start * (uint64_t)W + (uint64_t)w;
            int32_t cur = (int32_t)start;
            double load = 1.0;
            for (int32_t step = 0; step < L; ++step) {
                Rec *rec = &recs[(size_t)step * W + cnt[step]];
                rec->node = cur;
                rec->seq = cnt[step];
                rec->load = load;
                cnt[step]++;
                r->visits++;
                const int32_t s = indptr[cur];
                const int32_t deg = indptr[cur + 1] - s;
                if (deg == 0) break;
                int32_t k;
                if (draw_mode == 0) {
                    const double u = trace_u[walk_id * (uint64_t)L + (uint64_t)step];
                    if (u < p_halt) break;
                    k = trace_k[walk_id * (uint64_t)L + (uint64_t)step];
                } else {
                    const uint32_t ctr[4] = {(uint32_t)walk_id, (uint32_t)(walk_id >> 32), (uint32_t)(step >> 1), 0u};
                    uint32_t x[4];
                    grf_oracle_philox(ctr, key, x);
                    const uint32_t xh = x[2 * (step & 1)], xk = x[2 * (step & 1) + 1];
                    if ((uint64_t)xh < halt_thr) break;
                    k = (int32_t)(((uint64_t)xk * (uint64_t)(uint32_t)deg) >> 32);
                }
                const double wgt = data[s + k];
                /* volatile keeps gcc from contracting into an FMA on any -march */
                volatile double scaled = (double)deg * wgt;
                scaled = scaled / one_minus_p;
                if (load_mode == 0) {
                    volatile double nl = load * scaled;
                    load = nl;
                } else if (load_mode == 1) {
                    load = scaled;
                } else {
                    load = wgt;
                }
                cur = indices[s + k];
            }
        }
        const int64_t row = start - start_lo;
        for (int32_t step = 0; step < L; ++step) {
            Rec *base = &recs[(size_t)step * W];
            qsort(base, (size_t)cnt[step], sizeof(Rec), rec_cmp);
            int32_t i = 0;
            while (i < cnt[step]) {
                volatile double sum = 0.0;
                int32_t j = i;
                while (j < cnt[step] && base[j].node == base[i].node) {
                    sum = sum + base[j].load;
                    ++j;
                }
                double v = sum;
                v = scale_mode == 0 ? v * recip : v / (double)W;
                push_entry(r, step, base[i].node, v);
                i = j;
            }
            r->indptr[step][row + 1] = r->nnz[step];
        }
    }
    free(recs);
    free(cnt);
    return r;
}

int64_t grf_oracle_nnz(const GrfOracleResult *r, int32_t step) { return r->nnz[step]; }
int64_t grf_oracle_visits(const GrfOracleResult *r) { return r->visits; }

void grf_oracle_copy(const GrfOracleResult *r, int32_t step, int64_t *indptr, int32_t *indices, double *data) {
    memcpy(indptr, r->indptr[step], (size_t)(r->n_local + 1) * sizeof(int64_t));
    memcpy(indices, r->indices[step], (size_t)r->nnz[step] * sizeof(int32_t));
    memcpy(data, r->data[step], (size_t)r->nnz[step] * sizeof(double));
}

void grf_oracle_free(GrfOracleResult *r) {
    if (!r) return;
    for (int s = 0; s < r->L; ++s) {
        free(r->indptr[s]);
        free(r->indices[s]);
        free(r->data[s]);
    }
    free(r->indptr);
    free(r->indices);
    free(r->data);
    free(r->nnz);
    free(r->cap);
    free(r);
}
