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

// one warp per row; writes mkey (col << 5 | step) and mval at the merged position
__global__ void __launch_bounds__(256) union_rank_kernel(const int32_t *__restrict__ ptr,
                                                         const GrfEntry *__restrict__ ent, int64_t n_rows, int32_t L,
                                                         uint32_t *__restrict__ mkey, float *__restrict__ mval) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t *rp = ptr + r * L;
        const int32_t row_b = rp[0];
        for (int s = 0; s < L; ++s) {
            const int32_t b = rp[s], e = rp[s + 1];
            for (int32_t i = b + lane; i < e; i += 32) {
                const GrfEntry en = ent[i];
                const uint32_t col = (uint32_t)en.col & kColMask;
                int32_t rank = i - b;
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
}

__global__ void __launch_bounds__(256) union_count_kernel(const int32_t *__restrict__ ptr,
                                                          const uint32_t *__restrict__ mkey, int64_t n_rows,
                                                          int32_t L, int32_t *__restrict__ ucnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t b = ptr[r * L], e = ptr[(r + 1) * L];
        int cnt = 0;
        for (int32_t i = b + lane; i < e; i += 32) cnt += (i == b) || ((mkey[i - 1] >> 5) != (mkey[i] >> 5));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0) ucnt[r] = cnt;
    }
}

// headers: {col, mask}; one warp per row, heads ranked with ballots
__global__ void __launch_bounds__(256) union_fill_kernel(const int32_t *__restrict__ ptr,
                                                         const uint32_t *__restrict__ mkey, int64_t n_rows, int32_t L,
                                                         const int32_t *__restrict__ uptr, int2 *__restrict__ uhdr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t b = ptr[r * L], e = ptr[(r + 1) * L];
        int32_t out = uptr[r];
        for (int32_t base = b; base < e; base += 32) {
            const int32_t i = base + lane;
            bool head = false;
            uint32_t col = 0;
            if (i < e) {
                col = mkey[i] >> 5;
                head = (i == b) || ((mkey[i - 1] >> 5) != col);
            }
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            if (head) {
                uint32_t mask = 0;
                for (int32_t q = i; q < e && (mkey[q] >> 5) == col; ++q) mask |= 1u << (mkey[q] & 31u);
                uhdr[out + __popc(heads & ((1u << lane) - 1u))] = make_int2((int)col, (int)mask);
            }
            out += __popc(heads);
        }
    }
}

// ent_f[u] = {col, sum_{l in mask} f[l] * mval[..]}; one warp per row
__global__ void __launch_bounds__(256) union_materialize_kernel(const int32_t *__restrict__ ptr,
                                                                const int32_t *__restrict__ uptr,
                                                                const int2 *__restrict__ uhdr,
                                                                const float *__restrict__ mval,
                                                                const float *__restrict__ f, int64_t n_rows,
                                                                int32_t L, GrfEntry *__restrict__ ent_f) {
    __shared__ float fs[kMaxSteps];
    if (threadIdx.x < kMaxSteps) fs[threadIdx.x] = threadIdx.x < L ? f[threadIdx.x] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t ub = uptr[r], ue = uptr[r + 1];
        int32_t voff = ptr[r * L];  // the row's values start where its per-length entries start
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

static inline int warp_grid(int64_t n_rows) {
    int64_t g = (n_rows + 7) / 8;
    const int64_t cap = (int64_t)kSmCount * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace grf

using namespace grf;

extern "C" int grf_union_rank(const int32_t *blk_ptr, const GrfEntry *entries, int64_t n_rows, int32_t n_steps,
                              uint32_t *mkey, float *mval, int32_t *ucnt, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, ucnt);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1 && n_steps <= kMaxSteps, "grf_union_rank: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && mkey && mval && ucnt, "grf_union_rank: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    union_rank_kernel<<<warp_grid(n_rows), 256, 0, st>>>(blk_ptr, entries, n_rows, n_steps, mkey, mval);
    GRF_CUDA_OK(cudaGetLastError());
    union_count_kernel<<<warp_grid(n_rows), 256, 0, st>>>(blk_ptr, mkey, n_rows, n_steps, ucnt);
    return check_cuda(cudaGetLastError(), "union_rank/count launch");
}

extern "C" int grf_union_fill(const int32_t *blk_ptr, const uint32_t *mkey, int64_t n_rows, int32_t n_steps,
                              const int32_t *uptr, int32_t *uhdr, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, uptr);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1, "grf_union_fill: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && mkey && uptr && uhdr, "grf_union_fill: null buffer");
    union_fill_kernel<<<warp_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(blk_ptr, mkey, n_rows, n_steps, uptr,
                                                                          (int2 *)uhdr);
    return check_cuda(cudaGetLastError(), "union_fill_kernel launch");
}

extern "C" int grf_union_materialize(const int32_t *blk_ptr, const int32_t *uptr, const int32_t *uhdr,
                                     const float *mval, const float *f, int64_t n_rows, int32_t n_steps,
                                     GrfEntry *entries_f, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, uptr);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1 && n_steps <= kMaxSteps, "grf_union_materialize: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && uptr && f, "grf_union_materialize: null buffer");
    union_materialize_kernel<<<warp_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(
        blk_ptr, uptr, (const int2 *)uhdr, mval, f, n_rows, n_steps, entries_f);
    return check_cuda(cudaGetLastError(), "union_materialize_kernel launch");
}
