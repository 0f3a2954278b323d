// Normalized Laplacian of a CSR adjacency on the device: D^-1/2 (D - A) D^-1/2.
//
// Replaces efficient_graph_gp_sparse/utils_sparse/graph_utils.py:5-30 (two scipy SpGEMMs against
// diagonal matrices; 3.9 s on the host for the 4 M-node / 67 M-edge graph, against 0.04 s for the
// whole Phi build).  Bit-identical to the reference for canonical CSR input (sorted columns, no
// duplicates):
//   deg_i   = sum_j A_ij in the order scipy's A.sum(axis=1) uses: np.add.reduceat over the row, i.e.
//             a[0] + pairwise_sum(a[1:]) with numpy's pairwise summation (8 accumulators up to 128
//             elements, recursive halving above) -- irrelevant for unit weights, needed for the
//             last bit with weighted graphs
//   dis_i   = deg_i > 0 ? 1 / sqrt(deg_i) : 0          (np.sqrt, then 1.0 / x; inf -> 0)
//   L_ij    = deg_i - A_ii on the diagonal, -A_ij elsewhere; exact zeros dropped (csr binop)
//   out_ij  = (dis_i * L_ij) * dis_j, in that order; exact zeros dropped at either product
//             (scipy's SpGEMM) -- so zero-degree rows come out empty
// Two passes (count, fill) of one warp per row around a scan; streaming, HBM-bound.  Rows in which nothing is
// dropped (the usual case) are filled without any cross-lane dependency, so hub rows pipeline.

#include "grf_common.cuh"

namespace grf {

// numpy's pairwise summation (numpy/_core/src/umath/loops_utils.h.src, DOUBLE_pairwise_sum), restated.
// Blocks of up to 128 elements are summed with 8 accumulators; longer runs are halved recursively
// (left half rounded down to a multiple of 8) and the two halves added.  The recursion is unrolled
// onto an explicit stack: a hub row of a power-law graph has 10^5 neighbours, and a recursive device
// function of that depth overruns the default 1 KB call stack (illegal memory access at 4 M nodes).
__device__ __forceinline__ double numpy_pairwise_leaf(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int64_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

__device__ double numpy_pairwise_sum(const double *a, int64_t n) {
    // post-order walk of the halving tree; the current root-to-node path has one node per level, so
    // the per-level arrays are the whole stack (depth <= log2(2^31 / 128) + 1 < 32)
    constexpr int kDepth = 32;
    int64_t right_off[kDepth], right_n[kDepth];
    double left_sum[kDepth];
    bool left_done[kDepth];
    int level = 0;
    int64_t off = 0, len = n;
    for (;;) {
        while (len > 128) {  // descend along left children, remembering the right ones
            int64_t n2 = len / 2;
            n2 -= n2 % 8;
            right_off[level] = off + n2;
            right_n[level] = len - n2;
            left_done[level] = false;
            len = n2;
            ++level;
        }
        double ret = numpy_pairwise_leaf(a + off, len);
        for (;;) {  // ascend
            if (level == 0) return ret;
            const int parent = level - 1;
            if (!left_done[parent]) {  // that was the parent's left child: its right child is next
                left_sum[parent] = ret;
                left_done[parent] = true;
                off = right_off[parent];
                len = right_n[parent];
                break;
            }
            ret = __dadd_rn(left_sum[parent], ret);  // right child done: parent = left + right
            level = parent;
        }
    }
}

// warp per row.  The summation order is scipy's (see the header comment) -- which only matters when the sum is
// inexact: if every weight of the row is an integer and sum |a_ij| <= 2^53, every partial sum in ANY order is an
// exactly representable integer, so the lanes add their strided share and a shuffle tree finishes (unit-weight
// graphs: every row).  Otherwise lane 0 walks numpy's pairwise tree.  (One thread per row, always in numpy's order:
// 10 ms for the 175 303-neighbour hub row of the config-4 graph, 30x the time of the rest of the graph.)
__global__ void __launch_bounds__(256) lap_degree_kernel(const int32_t *__restrict__ row_ptr,
                                                         const double *__restrict__ val, int64_t n,
                                                         double *__restrict__ deg, double *__restrict__ dis) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const int32_t b = row_ptr[r], e = row_ptr[r + 1];
        double s = 0.0, sa = 0.0;
        bool whole = true;
        for (int32_t i = b + lane; i < e; i += 32) {
            const double a = val[i];
            s += a;
            sa += fabs(a);
            whole = whole && (a == rint(a));  // false for NaN; infinities fail the bound below
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            sa += __shfl_xor_sync(0xffffffffu, sa, d);
        }
        const bool exact = __all_sync(0xffffffffu, whole) && sa <= 9007199254740992.0;
        if (lane == 0) {
            if (!exact) s = e - b <= 1 ? (e > b ? val[b] : 0.0) : __dadd_rn(val[b], numpy_pairwise_sum(val + b + 1, e - b - 1));
            deg[r] = s;
            double d = __ddiv_rn(1.0, __dsqrt_rn(s));
            if (isinf(d)) d = 0.0;  // d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0
            dis[r] = d;
        }
    }
}

constexpr int kLapUnroll = 8;  // independent 32-entry batches of a row in flight per warp

// value of output entry (r, c) given the Laplacian value l = L_rc; keep = survives both products
__device__ __forceinline__ double lap_scale(double dis_r, double l, double dis_c, bool &keep) {
    const double x = __dmul_rn(dis_r, l);
    const double y = __dmul_rn(x, dis_c);
    keep = (l != 0.0) && (x != 0.0) && (y != 0.0);
    return y;
}

// out_cnt[r] = entries of output row r.  Lanes count their strided share of the row (no cross-lane dependency
// between the batches: a hub row's loads pipeline), one reduction at the end.
__global__ void __launch_bounds__(256) lap_count_kernel(const int32_t *__restrict__ row_ptr,
                                                        const int32_t *__restrict__ col, const double *__restrict__ val,
                                                        const double *__restrict__ deg, const double *__restrict__ dis,
                                                        int64_t n, int32_t *__restrict__ out_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const int32_t b = row_ptr[r], e = row_ptr[r + 1];
        const double dr = dis[r], dg = deg[r];
        int cnt = 0;
        double a_diag = 0.0;  // a stored self-loop (at most one: canonical CSR)
        // kLapUnroll batches in flight: the loads of a batch depend on each other (col -> dis[col]), so one
        // batch at a time made the 175 303-neighbour hub row of the config-4 graph a chain of 5 500 round trips
        // (6.6 ms for this kernel, all of it that one warp)
        for (int32_t i0 = b + lane; i0 < e; i0 += 32 * kLapUnroll) {
            int32_t c[kLapUnroll];
            double a[kLapUnroll], dc[kLapUnroll];
#pragma unroll
            for (int u = 0; u < kLapUnroll; ++u) {
                const int32_t i = i0 + 32 * u;
                c[u] = i < e ? col[i] : -1;
                a[u] = i < e ? val[i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < kLapUnroll; ++u) dc[u] = (c[u] >= 0 && c[u] != (int32_t)r) ? dis[c[u]] : 0.0;
#pragma unroll
            for (int u = 0; u < kLapUnroll; ++u) {
                if (c[u] < 0) continue;
                if (c[u] == (int32_t)r) {
                    a_diag = a[u];
                } else {
                    bool keep;
                    lap_scale(dr, -a[u], dc[u], keep);
                    cnt += keep;
                }
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            a_diag += __shfl_xor_sync(0xffffffffu, a_diag, d);  // all other lanes hold 0.0
        }
        if (lane == 0) {
            bool keep_d;
            lap_scale(dr, __dsub_rn(dg, a_diag), dr, keep_d);
            out_cnt[r] = cnt + (keep_d ? 1 : 0);
        }
    }
}

// write the output rows at out_ptr[r]
__global__ void __launch_bounds__(256) lap_fill_kernel(const int32_t *__restrict__ row_ptr,
                                                       const int32_t *__restrict__ col, const double *__restrict__ val,
                                                       const double *__restrict__ deg, const double *__restrict__ dis,
                                                       int64_t n, const int32_t *__restrict__ out_ptr,
                                                       int32_t *__restrict__ out_col, double *__restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const int32_t b = row_ptr[r], e = row_ptr[r + 1];
        const double dr = dis[r], dg = deg[r];
        int32_t out = out_ptr[r];
        if (e > b && out_ptr[r + 1] - out == e - b + 1) {
            // every stored entry survives, so does the diagonal, and there is no self-loop (the row could not
            // have e - b + 1 entries otherwise): positions follow from the index alone -- entry i goes to
            // i - b, one further if its column lies beyond the diagonal -- and the batches are independent
            for (int32_t i0 = b + lane; i0 < e; i0 += 32 * kLapUnroll) {
                int32_t c[kLapUnroll];
                double a[kLapUnroll], dc[kLapUnroll];
#pragma unroll
                for (int u = 0; u < kLapUnroll; ++u) {
                    const int32_t i = i0 + 32 * u;
                    c[u] = i < e ? col[i] : -1;
                    a[u] = i < e ? val[i] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kLapUnroll; ++u) dc[u] = c[u] >= 0 ? dis[c[u]] : 0.0;
#pragma unroll
                for (int u = 0; u < kLapUnroll; ++u) {
                    if (c[u] < 0) continue;
                    const int32_t i = i0 + 32 * u;
                    bool keep;
                    const double v = lap_scale(dr, -a[u], dc[u], keep);
                    const int32_t pos = out + (i - b) + (c[u] > (int32_t)r ? 1 : 0);
                    out_col[pos] = c[u];
                    out_val[pos] = v;
                    // the diagonal sits before the first column beyond r, or after the last entry
                    const bool first_beyond = c[u] > (int32_t)r && (i == b || col[i - 1] < (int32_t)r);
                    const bool last_below = i == e - 1 && c[u] < (int32_t)r;
                    if (first_beyond || last_below) {
                        bool keep_d;
                        const int32_t pd = out + (i - b) + (last_below ? 1 : 0);
                        out_col[pd] = (int32_t)r;
                        out_val[pd] = lap_scale(dr, dg, dr, keep_d);
                    }
                }
            }
            continue;
        }
        // general path: zeros dropped / a stored self-loop; positions by ballot, batch after batch.
        // the diagonal entry goes before the first stored column > r; if every stored column is < r
        // (or the row is empty) one more, all-padding batch emits it
        bool diag_done = false;
        for (int32_t base = b; base < e || !diag_done; base += 32) {
            const int32_t i = base + lane;
            int32_t c = 0x7fffffff;
            double a = 0.0;
            if (i < e) {
                c = col[i];
                a = val[i];
            }
            // does the diagonal have to be emitted inside / before this batch?
            const unsigned gt = __ballot_sync(0xffffffffu, c > (int32_t)r);   // stored cols beyond r (or padding)
            const unsigned eq = __ballot_sync(0xffffffffu, c == (int32_t)r);  // a stored self-loop
            bool emit_diag = false;
            int diag_lane = 32;
            double a_diag = 0.0;
            if (!diag_done && (gt | eq)) {
                emit_diag = true;
                diag_lane = eq ? __ffs(eq) - 1 : __ffs(gt) - 1;  // position in this batch where it is inserted
                if (eq) a_diag = __shfl_sync(0xffffffffu, a, diag_lane);
                diag_done = true;
            }
            // ordinary entries (the self-loop lane is replaced by the diagonal)
            bool keep = false;
            double v = 0.0;
            if (i < e && c != (int32_t)r) v = lap_scale(dr, -a, dis[c], keep);
            bool keep_d = false;
            double vd = 0.0;
            if (emit_diag) vd = lap_scale(dr, __dsub_rn(dg, a_diag), dr, keep_d);
            const unsigned kept = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                int pos = __popc(kept & ((1u << lane) - 1u));
                if (keep_d && lane >= diag_lane) ++pos;  // the diagonal sits before this lane's entry
                out_col[out + pos] = c;
                out_val[out + pos] = v;
            }
            if (keep_d && lane == 0) {
                const int pos = __popc(kept & ((1u << diag_lane) - 1u));
                out_col[out + pos] = (int32_t)r;
                out_val[out + pos] = vd;
            }
            out += __popc(kept) + (keep_d ? 1 : 0);
        }
    }
}

static inline int lap_grid(int64_t n, int per_block) {
    int64_t g = (n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)kSmCount * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace grf

using namespace grf;

extern "C" int grf_laplacian_count(const GrfGraph *adj, double *deg, double *dis, int32_t *out_cnt, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, out_cnt);
    GRF_REQUIRE(adj, "grf_laplacian_count: null graph");
    GRF_REQUIRE(adj->n_nodes >= 0 && adj->n_nodes < (1ll << 31) && adj->nnz < (1ll << 31) - adj->n_nodes,
                "grf_laplacian_count: graph exceeds int32 index range");
    if (adj->n_nodes == 0) return GRF_OK;
    GRF_REQUIRE(adj->row_ptr && deg && dis && out_cnt, "grf_laplacian_count: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    lap_degree_kernel<<<lap_grid(adj->n_nodes, 8), 256, 0, st>>>(adj->row_ptr, adj->val, adj->n_nodes, deg, dis);
    GRF_CUDA_OK(cudaGetLastError());
    lap_count_kernel<<<lap_grid(adj->n_nodes, 8), 256, 0, st>>>(adj->row_ptr, adj->col_idx, adj->val, deg, dis,
                                                                adj->n_nodes, out_cnt);
    return check_cuda(cudaGetLastError(), "laplacian count launch");
}

extern "C" int grf_laplacian_fill(const GrfGraph *adj, const double *deg, const double *dis, const int32_t *out_ptr,
                                  int32_t *out_col, double *out_val, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, out_ptr);
    GRF_REQUIRE(adj, "grf_laplacian_fill: null graph");
    if (adj->n_nodes == 0) return GRF_OK;
    GRF_REQUIRE(adj->row_ptr && deg && dis && out_ptr, "grf_laplacian_fill: null buffer");
    lap_fill_kernel<<<lap_grid(adj->n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
        adj->row_ptr, adj->col_idx, adj->val, deg, dis, adj->n_nodes, out_ptr, out_col, out_val);
    return check_cuda(cudaGetLastError(), "laplacian fill launch");
}
