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
// Two passes (count, fill) of one warp per row around a scan; streaming, HBM-bound.

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

// thread per row; the summation order is scipy's (see the header comment)
__global__ void __launch_bounds__(256) lap_degree_kernel(const int32_t *__restrict__ row_ptr,
                                                         const double *__restrict__ val, int64_t n,
                                                         double *__restrict__ deg, double *__restrict__ dis) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t b = row_ptr[r], e = row_ptr[r + 1];
        double s = 0.0;
        if (e > b) s = e - b == 1 ? val[b] : __dadd_rn(val[b], numpy_pairwise_sum(val + b + 1, e - b - 1));
        deg[r] = s;
        double d = __ddiv_rn(1.0, __dsqrt_rn(s));
        if (isinf(d)) d = 0.0;  // d_inv_sqrt[np.isinf(d_inv_sqrt)] = 0
        dis[r] = d;
    }
}

// value of output entry (r, c) given the Laplacian value l = L_rc; keep = survives both products
__device__ __forceinline__ double lap_scale(double dis_r, double l, double dis_c, bool &keep) {
    const double x = __dmul_rn(dis_r, l);
    const double y = __dmul_rn(x, dis_c);
    keep = (l != 0.0) && (x != 0.0) && (y != 0.0);
    return y;
}

// FILL = false: out_cnt[r] = entries of output row r;  FILL = true: write them at out_ptr[r]
template <bool FILL>
__global__ void __launch_bounds__(256) lap_rows_kernel(const int32_t *__restrict__ row_ptr,
                                                       const int32_t *__restrict__ col, const double *__restrict__ val,
                                                       const double *__restrict__ deg, const double *__restrict__ dis,
                                                       int64_t n, int32_t *__restrict__ out_cnt,
                                                       const int32_t *__restrict__ out_ptr,
                                                       int32_t *__restrict__ out_col, double *__restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const int32_t b = row_ptr[r], e = row_ptr[r + 1];
        const double dr = dis[r], dg = deg[r];
        int32_t out = FILL ? out_ptr[r] : 0;
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
            if (FILL) {
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
            }
            out += __popc(kept) + (keep_d ? 1 : 0);
        }
        if (!FILL && lane == 0) out_cnt[r] = out;
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
    lap_degree_kernel<<<lap_grid(adj->n_nodes, 256), 256, 0, st>>>(adj->row_ptr, adj->val, adj->n_nodes, deg, dis);
    GRF_CUDA_OK(cudaGetLastError());
    lap_rows_kernel<false><<<lap_grid(adj->n_nodes, 8), 256, 0, st>>>(adj->row_ptr, adj->col_idx, adj->val, deg, dis,
                                                                       adj->n_nodes, out_cnt, nullptr, nullptr,
                                                                       nullptr);
    return check_cuda(cudaGetLastError(), "laplacian count launch");
}

extern "C" int grf_laplacian_fill(const GrfGraph *adj, const double *deg, const double *dis, const int32_t *out_ptr,
                                  int32_t *out_col, double *out_val, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, out_ptr);
    GRF_REQUIRE(adj, "grf_laplacian_fill: null graph");
    if (adj->n_nodes == 0) return GRF_OK;
    GRF_REQUIRE(adj->row_ptr && deg && dis && out_ptr, "grf_laplacian_fill: null buffer");
    lap_rows_kernel<true><<<lap_grid(adj->n_nodes, 8), 256, 0, (cudaStream_t)stream>>>(
        adj->row_ptr, adj->col_idx, adj->val, deg, dis, adj->n_nodes, nullptr, out_ptr, out_col, out_val);
    return check_cuda(cudaGetLastError(), "laplacian fill launch");
}
