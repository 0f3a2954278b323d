// Union rows: merging the L per-length segments of every Phi (or Phi^T) row.
//
// On lattices, rings and any graph whose walks revisit nodes, the per-length
// matrices M_l of one row overlap heavily (config 2: sum_l nnz(M_l) = 6.36 M but
// the union pattern of Phi = sum_l f_l M_l has 2.76 M entries).  The matvec is
// bound by one L1 wavefront per *gathered* entry, so for a CG solve -- hundreds
// of products with the same modulator f -- Phi_f is materialised once on the
// union pattern and the products run on 2.3x fewer entries and bytes.  f stays a
// matvec-time quantity: re-materialising after an optimiser step is one
// streaming pass, and the gradient still uses the per-length blocks.
//
// Layout (built once per Phi, static):
//   merged values  mval[nnz]   the row's entries re-ordered by (col, length)
//   union headers  uhdr[nU]    {col, mask}: bit l of mask set <=> M_l[row, col] stored;
//                              the values of a union entry are contiguous in mval, ascending l
//   uptr[n_rows + 1]           union entries of row r: [uptr[r], uptr[r+1])
// Per modulator:  ent_f[nU] = {col, sum_{l in mask} f[l] * mval[..]} -- a plain CSR
// that the ordinary spmm kernel multiplies with L = 1, f = [1].
//
// All kernels are streaming passes (HBM-bound); the merge itself is done by rank:
// an entry's position in the (col, length) order of its row is its own index in
// its segment plus, for every other segment, a binary search -- no sorting network,
// any row length.

#include "grf_common.cuh"

namespace grf {

// position of `key` = (col, step) among the entries of segment [b, e) (sorted by col, unique)
__device__ __forceinline__ int32_t lower_bound_col(const GrfEntry *__restrict__ ent, int32_t b, int32_t e,
                                                   uint32_t col, bool strict) {
    // number of entries with col' < col (strict) or col' <= col
    int32_t lo = b, hi = e;
    while (lo < hi) {
        const int32_t mid = (lo + hi) >> 1;
        const uint32_t c = (uint32_t)ent[mid].col & kColMask;
        const bool left = strict ? (c < col) : (c <= col);
        if (left)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo - b;
}

// Work is handed out in TASKS, not rows: a task is a row, or -- for a row longer than the caller's chunk
// size -- one chunk of consecutive entries of its flat run (task_row[k], entries [task_b[k], task_e[k])).
// The hub columns of a power-law Phi^T hold 10^5..10^6 entries; with one warp per row the longest one was a
// serial chain of 2*10^4 iterations of dependent binary searches (config 4: 294 ms for the union build and
// 44 ms per materialisation against 40 ms for the whole Phi build).

// writes mkey (col << 5 | step) and mval at the merged position; one warp per task
__global__ void __launch_bounds__(256) union_rank_kernel(const int32_t *__restrict__ ptr,
                                                         const GrfEntry *__restrict__ ent,
                                                         const int32_t *__restrict__ task_row,
                                                         const int32_t *__restrict__ task_b,
                                                         const int32_t *__restrict__ task_e, int64_t n_tasks,
                                                         int32_t L, uint32_t *__restrict__ mkey,
                                                         float *__restrict__ mval) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < n_tasks; k += nwarps) {
        const int32_t *rp = ptr + (int64_t)task_row[k] * L;
        const int32_t row_b = rp[0];
        const int32_t tb = task_b[k], te = task_e[k];
        for (int32_t i = tb + lane; i < te; i += 32) {
            int s = 0;
            while (s + 1 < L && i >= rp[s + 1]) ++s;  // the segment (walk length) entry i belongs to
            const GrfEntry en = ent[i];
            const uint32_t col = (uint32_t)en.col & kColMask;
            int32_t rank = i - rp[s];
            for (int s2 = 0; s2 < L; ++s2) {
                if (s2 == s) continue;
                // entries of an earlier length with the same col come first, of a later length after
                rank += lower_bound_col(ent, rp[s2], rp[s2 + 1], col, /*strict=*/s2 > s);
            }
            mkey[row_b + rank] = (col << 5) | (uint32_t)s;
            mval[row_b + rank] = en.val;
        }
    }
}

// distinct columns among the merged entries of each task (a column's first entry decides the task)
__global__ void __launch_bounds__(256) union_count_kernel(const int32_t *__restrict__ ptr,
                                                          const uint32_t *__restrict__ mkey,
                                                          const int32_t *__restrict__ task_row,
                                                          const int32_t *__restrict__ task_b,
                                                          const int32_t *__restrict__ task_e, int64_t n_tasks,
                                                          int32_t L, int32_t *__restrict__ task_cnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < n_tasks; k += nwarps) {
        const int32_t row_b = ptr[(int64_t)task_row[k] * L];
        const int32_t b = task_b[k], e = task_e[k];
        int cnt = 0;
        for (int32_t i = b + lane; i < e; i += 32) cnt += (i == row_b) || ((mkey[i - 1] >> 5) != (mkey[i] >> 5));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0) task_cnt[k] = cnt;
    }
}

// headers: {col, mask}; one warp per task, heads ranked with ballots.  task_v0[k] = merged position of the
// task's first head (where the values of its first union entry start).
__global__ void __launch_bounds__(256) union_fill_kernel(const int32_t *__restrict__ ptr,
                                                         const uint32_t *__restrict__ mkey,
                                                         const int32_t *__restrict__ task_row,
                                                         const int32_t *__restrict__ task_b,
                                                         const int32_t *__restrict__ task_e, int64_t n_tasks,
                                                         int32_t L, const int32_t *__restrict__ task_u0,
                                                         int2 *__restrict__ uhdr, int32_t *__restrict__ task_v0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < n_tasks; k += nwarps) {
        const int64_t r = task_row[k];
        const int32_t row_b = ptr[r * L], row_e = ptr[(r + 1) * L];
        const int32_t b = task_b[k], e = task_e[k];
        int32_t out = task_u0[k];
        bool first_seen = false;
        for (int32_t base = b; base < e; base += 32) {
            const int32_t i = base + lane;
            bool head = false;
            uint32_t col = 0;
            if (i < e) {
                col = mkey[i] >> 5;
                head = (i == row_b) || ((mkey[i - 1] >> 5) != col);
            }
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            if (head) {
                uint32_t mask = 0;  // the group may run past the end of the task (never past the row)
                for (int32_t q = i; q < row_e && (mkey[q] >> 5) == col; ++q) mask |= 1u << (mkey[q] & 31u);
                uhdr[out + __popc(heads & ((1u << lane) - 1u))] = make_int2((int)col, (int)mask);
            }
            if (!first_seen && heads) {
                if (lane == __ffs(heads) - 1) task_v0[k] = i;
                first_seen = true;
            }
            out += __popc(heads);
        }
        if (!first_seen && lane == 0) task_v0[k] = e;
    }
}

// ent_f[u] = {col, sum_{l in mask} f[l] * mval[..]}; one warp per task: union entries
// [task_u0[k], task_u0[k+1]), their values from merged position task_v0[k] on
__global__ void __launch_bounds__(256) union_materialize_kernel(const int32_t *__restrict__ task_u0,
                                                                const int32_t *__restrict__ task_v0,
                                                                int64_t n_tasks, const int2 *__restrict__ uhdr,
                                                                const float *__restrict__ mval,
                                                                const float *__restrict__ f, int32_t L,
                                                                GrfEntry *__restrict__ ent_f) {
    __shared__ float fs[kMaxSteps];
    if (threadIdx.x < kMaxSteps) fs[threadIdx.x] = threadIdx.x < L ? f[threadIdx.x] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < n_tasks; k += nwarps) {
        const int32_t ub = task_u0[k], ue = task_u0[k + 1];
        int32_t voff = task_v0[k];
        for (int32_t base = ub; base < ue; base += 32) {
            const int32_t u = base + lane;
            int2 h = make_int2(0, 0);
            if (u < ue) h = uhdr[u];
            const int n = __popc((unsigned)h.y);
            int incl = n;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            if (u < ue) {
                int32_t q = voff + incl - n;
                float acc = 0.f;
                unsigned m = (unsigned)h.y;
                while (m) {
                    const int l = __ffs(m) - 1;
                    m &= m - 1;
                    acc = fmaf(fs[l], mval[q++], acc);
                }
                GrfEntry o;
                o.col = h.x;  // length bits 0: the merged matrix is multiplied with L = 1
                o.val = acc;
                ent_f[u] = o;
            }
            voff += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

static inline int warp_grid(int64_t n_tasks) {
    int64_t g = (n_tasks + 7) / 8;
    const int64_t cap = (int64_t)kSmCount * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace grf

using namespace grf;

extern "C" int grf_union_rank(const int32_t *blk_ptr, const GrfEntry *entries, int32_t n_steps,
                              const int32_t *task_row, const int32_t *task_b, const int32_t *task_e, int64_t n_tasks,
                              uint32_t *mkey, float *mval, int32_t *task_cnt, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, task_cnt);
    GRF_REQUIRE(n_tasks >= 0 && n_steps >= 1 && n_steps <= kMaxSteps, "grf_union_rank: bad shape");
    if (n_tasks == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && task_row && task_b && task_e && mkey && mval && task_cnt, "grf_union_rank: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    union_rank_kernel<<<warp_grid(n_tasks), 256, 0, st>>>(blk_ptr, entries, task_row, task_b, task_e, n_tasks, n_steps,
                                                         mkey, mval);
    GRF_CUDA_OK(cudaGetLastError());
    union_count_kernel<<<warp_grid(n_tasks), 256, 0, st>>>(blk_ptr, mkey, task_row, task_b, task_e, n_tasks, n_steps,
                                                          task_cnt);
    return check_cuda(cudaGetLastError(), "union_rank/count launch");
}

extern "C" int grf_union_fill(const int32_t *blk_ptr, const uint32_t *mkey, int32_t n_steps, const int32_t *task_row,
                              const int32_t *task_b, const int32_t *task_e, int64_t n_tasks, const int32_t *task_u0,
                              int32_t *uhdr, int32_t *task_v0, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, task_u0);
    GRF_REQUIRE(n_tasks >= 0 && n_steps >= 1, "grf_union_fill: bad shape");
    if (n_tasks == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && mkey && task_row && task_b && task_e && task_u0 && uhdr && task_v0,
                "grf_union_fill: null buffer");
    union_fill_kernel<<<warp_grid(n_tasks), 256, 0, (cudaStream_t)stream>>>(
        blk_ptr, mkey, task_row, task_b, task_e, n_tasks, n_steps, task_u0, (int2 *)uhdr, task_v0);
    return check_cuda(cudaGetLastError(), "union_fill_kernel launch");
}

extern "C" int grf_union_materialize(const int32_t *task_u0, const int32_t *task_v0, int64_t n_tasks,
                                     const int32_t *uhdr, const float *mval, const float *f, int32_t n_steps,
                                     GrfEntry *entries_f, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, task_u0);
    GRF_REQUIRE(n_tasks >= 0 && n_steps >= 1 && n_steps <= kMaxSteps, "grf_union_materialize: bad shape");
    if (n_tasks == 0) return GRF_OK;
    GRF_REQUIRE(task_u0 && task_v0 && f, "grf_union_materialize: null buffer");
    union_materialize_kernel<<<warp_grid(n_tasks), 256, 0, (cudaStream_t)stream>>>(
        task_u0, task_v0, n_tasks, (const int2 *)uhdr, mval, f, n_steps, entries_f);
    return check_cuda(cudaGetLastError(), "union_materialize_kernel launch");
}
