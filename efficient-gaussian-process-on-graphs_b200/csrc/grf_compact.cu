// Staging -> CSR: the array half of K2 and the one-off Phi^T build (K4).
//
// Replaces  sparse_sampler.py:117-130  (keys -> rows/cols/vals -> csr_matrix / W)
//           sampler.py:188-203          (dict -> dense, value / W)
//           graph_preprocessor.py:117-139 (float32 values for the matvec)
//           sparse_lo.py:23-25          (.t().to_sparse_csr(), per forward in the reference)
//
// All of these are HBM-bound streaming passes: every kernel reads and writes
// each byte once with warp-contiguous accesses; the only irregular traffic is
// the column-keyed scatter of the transpose (int32 atomics for the counts and
// the slot claim).

#include "grf_common.cuh"

namespace grf {

// ---------------------------------------------------------------------------
// exclusive scan of int32 counts (three short kernels: tile sums, scan of the
// tile sums by one CTA, tile-local scan + offset)
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int64_t scan_fetch(const int32_t *cnt, int64_t j, int64_t n_rows, int32_t n_steps,
                                              int32_t order) {
    if (order == GRF_ORDER_ROW_MAJOR) return cnt[j];
    const int64_t s = j / n_rows;
    const int64_t r = j - s * n_rows;
    return cnt[r * n_steps + s];
}

__device__ __forceinline__ int64_t block_reduce_sum(int64_t v, int64_t *sh) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    int64_t total = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) total += sh[i];
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const int32_t *cnt, int64_t n, int64_t n_rows,
                                                               int32_t n_steps, int32_t order, int64_t *tile_sums) {
    __shared__ int64_t sh[32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
    int64_t v = 0;
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t j = base + (int64_t)i * kScanThreads + threadIdx.x;
        if (j < n) v += scan_fetch(cnt, j, n_rows, n_steps, order);
    }
    const int64_t total = block_reduce_sum(v, sh);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// also leaves the grand total in *total (the caller sizes its output from it with one 8-byte read)
__global__ void __launch_bounds__(1024) scan_of_tile_sums(int64_t *tile_sums, int64_t n_tiles, int64_t *total) {
    __shared__ int64_t sh[32];
    __shared__ int64_t carry_sh;
    if (threadIdx.x == 0) carry_sh = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_tiles; base += blockDim.x) {
        const int64_t j = base + threadIdx.x;
        const int64_t v = j < n_tiles ? tile_sums[j] : 0;
        // inclusive scan across the CTA
        int64_t incl = v;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t nb = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += nb;
        }
        if (lane == 31) sh[warp] = incl;
        __syncthreads();
        int64_t before = 0, all = 0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
            if (i < warp) before += sh[i];
            all += sh[i];
        }
        const int64_t carry = carry_sh;
        if (j < n_tiles) tile_sums[j] = carry + before + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) carry_sh = carry + all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_sh;
}

template <typename OutT>
__global__ void __launch_bounds__(kScanThreads) scan_apply(const int32_t *cnt, int64_t n, int64_t n_rows,
                                                           int32_t n_steps, int32_t order, const int64_t *tile_sums,
                                                           OutT *offsets) {
    __shared__ int64_t sh[32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int64_t local[kScanItems];
    int64_t mine = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t j = base + i;
        local[i] = j < n ? scan_fetch(cnt, j, n_rows, n_steps, order) : 0;
        mine += local[i];
    }
    // exclusive scan of `mine` across the CTA
    int64_t incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t nb = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += nb;
    }
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    int64_t before = 0;
    for (int i = 0; i < warp; ++i) before += sh[i];
    int64_t run = tile_sums[blockIdx.x] + before + incl - mine;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        const int64_t j = base + i;
        if (j < n) offsets[j] = (OutT)run;
        run += local[i];
        if (j == n - 1) offsets[n] = (OutT)run;
    }
}

// ---------------------------------------------------------------------------
// staging -> reference layout (step-major CSR, float64) / matvec layout
// ---------------------------------------------------------------------------
__device__ __forceinline__ double scale_sum(double sum, int32_t scale_mode, double recip, double w) {
    return scale_mode == GRF_SCALE_MUL_RECIP ? __dmul_rn(sum, recip) : __ddiv_rn(sum, w);
}

// one warp per row
__global__ void __launch_bounds__(256) compact_steps_kernel(const int32_t *stage_col, const double *stage_sum,
                                                            const int32_t *row_cnt, const int64_t *off_sm,
                                                            int64_t n_rows, int32_t L, int64_t stride, int32_t W,
                                                            int32_t scale_mode, int32_t *out_col, double *out_val) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double recip = 1.0 / (double)W;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        int64_t src = r * stride;
        for (int s = 0; s < L; ++s) {
            const int c = row_cnt[r * L + s];
            const int64_t dst = off_sm[(int64_t)s * n_rows + r];
            for (int i = lane; i < c; i += 32) {
                out_col[dst + i] = stage_col[src + i];
                out_val[dst + i] = scale_sum(stage_sum[src + i], scale_mode, recip, (double)W);
            }
            src += c;
        }
    }
}

__global__ void __launch_bounds__(256) compact_blocks_kernel(const int32_t *stage_col, const double *stage_sum,
                                                             const int32_t *blk_ptr, int64_t n_rows, int32_t L,
                                                             int64_t stride, int32_t W, int32_t scale_mode,
                                                             GrfEntry *entries) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double recip = 1.0 / (double)W;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int64_t src = r * stride;
        const int32_t dst = blk_ptr[r * L];
        const int32_t c = blk_ptr[(r + 1) * L] - dst;  // the row's segments are contiguous on both sides
        for (int i = lane; i < c; i += 32) {
            int step = 0;
            while (step + 1 < L && dst + i >= blk_ptr[r * L + step + 1]) ++step;
            GrfEntry e;
            e.col = pack_col(stage_col[src + i], step);
            e.val = (float)scale_sum(stage_sum[src + i], scale_mode, recip, (double)W);
            entries[dst + i] = e;
        }
    }
}

// staging already holds finished entries (GrfWalkCfg.stage_entries): one warp moves a row's run
__global__ void __launch_bounds__(256) compact_entries_kernel(const int2 *__restrict__ stage,
                                                              const int32_t *__restrict__ blk_ptr, int64_t n_rows,
                                                              int32_t L, int64_t stride, int2 *__restrict__ entries) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int2 *src = stage + r * stride;
        const int32_t dst = __ldg(blk_ptr + r * L);
        const int32_t c = __ldg(blk_ptr + (r + 1) * L) - dst;
        for (int i0 = 0; i0 < c; i0 += 128) {  // four independent 8-byte loads per lane in flight
            int2 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 32 + lane;
                if (i < c) v[u] = __ldcs(src + i);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 32 + lane;
                if (i < c) entries[dst + i] = v[u];
            }
        }
    }
}

__global__ void __launch_bounds__(256) count_from_steps_kernel(const int64_t *off_sm, int64_t n_rows, int32_t L,
                                                               int32_t *row_cnt) {
    const int64_t n = n_rows * L;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = j / n_rows, r = j - s * n_rows;
        row_cnt[r * L + s] = (int32_t)(off_sm[j + 1] - off_sm[j]);
    }
}

__global__ void __launch_bounds__(256) blocks_from_steps_kernel(const int64_t *off_sm, const int32_t *col,
                                                                const double *val, const int32_t *blk_ptr,
                                                                int64_t n_rows, int32_t L, GrfEntry *entries) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        for (int s = 0; s < L; ++s) {
            const int64_t src = off_sm[(int64_t)s * n_rows + r];
            const int c = (int)(off_sm[(int64_t)s * n_rows + r + 1] - src);
            const int32_t dst = blk_ptr[r * L + s];
            for (int i = lane; i < c; i += 32) {
                GrfEntry e;
                e.col = pack_col(col[src + i], s);
                e.val = (float)val[src + i];  // torch .float(): round to nearest even
                entries[dst + i] = e;
            }
        }
    }
}

// Row statistics of one side of the Phi blocks in one pass over the row pointers:
// census[0] = rows longer than `threshold`, census[1] = chunks of `threshold` entries those rows
// split into, census[2] = non-empty rows.  Decides (without a host round trip per quantity)
// whether the matvec needs the long-row split and whether a non-empty-column list pays off.
__global__ void __launch_bounds__(256) row_census_kernel(const int32_t *__restrict__ ptr, int64_t n_rows, int32_t L,
                                                         int32_t threshold, int32_t *__restrict__ census) {
    int32_t n_long = 0, n_chunks = 0, n_full = 0;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t len = __ldg(ptr + (r + 1) * L) - __ldg(ptr + r * L);
        n_full += len > 0;
        if (len > threshold) {
            n_long += 1;
            n_chunks += (len + threshold - 1) / threshold;
        }
    }
    n_long = __reduce_add_sync(0xffffffffu, n_long);
    n_chunks = __reduce_add_sync(0xffffffffu, n_chunks);
    n_full = __reduce_add_sync(0xffffffffu, n_full);
    if ((threadIdx.x & 31) == 0) {
        if (n_long) atomicAdd(census + 0, n_long);
        if (n_chunks) atomicAdd(census + 1, n_chunks);
        if (n_full) atomicAdd(census + 2, n_full);
    }
}

// ids of the non-empty rows, ascending, without a host round trip: flags -> scan -> scatter
__global__ void __launch_bounds__(256) row_flags_kernel(const int32_t *__restrict__ ptr, int64_t n_rows, int32_t L,
                                                        int32_t *__restrict__ flags) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
        flags[r] = __ldg(ptr + (r + 1) * L) > __ldg(ptr + r * L);
}

__global__ void __launch_bounds__(256) scatter_ids_kernel(const int32_t *__restrict__ flags,
                                                          const int32_t *__restrict__ pos, int64_t n_rows,
                                                          int32_t *__restrict__ ids) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (int64_t)gridDim.x * blockDim.x)
        if (flags[r]) ids[pos[r]] = (int32_t)r;
}

}  // namespace grf

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
using namespace grf;

static inline int grid_for_warps(int64_t n_rows, int threads) {
    const int64_t warps_per_cta = threads / 32;
    int64_t g = (n_rows + warps_per_cta - 1) / warps_per_cta;
    const int64_t cap = (int64_t)kSmCount * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" int64_t grf_scan_workspace_bytes(int64_t n_items) {
    const int64_t tiles = (n_items + kScanTile - 1) / kScanTile;
    return (tiles + 2) * (int64_t)sizeof(int64_t);  // [grand total][tile sums]
}

extern "C" int grf_scan_counts(const int32_t *row_cnt, int64_t n_rows, int32_t n_steps, int32_t order, void *offsets,
                               int32_t out_is_i64, void *workspace, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, offsets);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1, "grf_scan_counts: bad shape");
    GRF_REQUIRE(order == GRF_ORDER_ROW_MAJOR || order == GRF_ORDER_STEP_MAJOR, "grf_scan_counts: bad order");
    GRF_REQUIRE(offsets && workspace, "grf_scan_counts: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_rows * n_steps;
    if (n == 0) {
        GRF_CUDA_OK(cudaMemsetAsync(offsets, 0, out_is_i64 ? 8 : 4, st));
        GRF_CUDA_OK(cudaMemsetAsync(workspace, 0, sizeof(int64_t), st));
        return GRF_OK;
    }
    GRF_REQUIRE(row_cnt, "grf_scan_counts: null counts");
    const int64_t tiles = (n + kScanTile - 1) / kScanTile;
    GRF_REQUIRE(tiles < (1ll << 31), "grf_scan_counts: too many items");
    int64_t *total = (int64_t *)workspace;
    int64_t *tile_sums = total + 1;
    scan_tile_sums<<<(int)tiles, kScanThreads, 0, st>>>(row_cnt, n, n_rows, n_steps, order, tile_sums);
    scan_of_tile_sums<<<1, 1024, 0, st>>>(tile_sums, tiles, total);
    if (out_is_i64)
        scan_apply<int64_t><<<(int)tiles, kScanThreads, 0, st>>>(row_cnt, n, n_rows, n_steps, order, tile_sums,
                                                                 (int64_t *)offsets);
    else
        scan_apply<int32_t><<<(int)tiles, kScanThreads, 0, st>>>(row_cnt, n, n_rows, n_steps, order, tile_sums,
                                                                 (int32_t *)offsets);
    return check_cuda(cudaGetLastError(), "scan kernels launch");
}

extern "C" int grf_compact_steps(const int32_t *stage_col, const double *stage_sum, const int32_t *row_cnt,
                                 const int64_t *offsets_step_major, int64_t n_rows, int32_t n_steps,
                                 int64_t stage_stride, int32_t walks_per_node, int32_t scale_mode, int32_t *out_col,
                                 double *out_val, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, row_cnt);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1 && walks_per_node >= 1, "grf_compact_steps: bad shape");
    GRF_REQUIRE(scale_mode == GRF_SCALE_MUL_RECIP || scale_mode == GRF_SCALE_DIV, "grf_compact_steps: bad scale_mode");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(stage_col && stage_sum && row_cnt && offsets_step_major && out_col && out_val,
                "grf_compact_steps: null buffer");
    compact_steps_kernel<<<grid_for_warps(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(
        stage_col, stage_sum, row_cnt, offsets_step_major, n_rows, n_steps, stage_stride, walks_per_node, scale_mode,
        out_col, out_val);
    return check_cuda(cudaGetLastError(), "compact_steps_kernel launch");
}

extern "C" int grf_compact_blocks(const int32_t *stage_col, const double *stage_sum, const int32_t *row_cnt,
                                  const int32_t *blk_ptr, int64_t n_rows, int32_t n_steps, int64_t stage_stride,
                                  int32_t walks_per_node, int32_t scale_mode, GrfEntry *entries, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, blk_ptr);
    (void)row_cnt;
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1 && walks_per_node >= 1, "grf_compact_blocks: bad shape");
    GRF_REQUIRE(scale_mode == GRF_SCALE_MUL_RECIP || scale_mode == GRF_SCALE_DIV,
                "grf_compact_blocks: bad scale_mode");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(stage_col && stage_sum && blk_ptr && entries, "grf_compact_blocks: null buffer");
    compact_blocks_kernel<<<grid_for_warps(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(
        stage_col, stage_sum, blk_ptr, n_rows, n_steps, stage_stride, walks_per_node, scale_mode, entries);
    return check_cuda(cudaGetLastError(), "compact_blocks_kernel launch");
}

extern "C" int grf_compact_entries(const GrfEntry *stage_entries, const int32_t *blk_ptr, int64_t n_rows,
                                   int32_t n_steps, int64_t stage_stride, GrfEntry *entries, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, blk_ptr);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1 && stage_stride >= 1, "grf_compact_entries: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(stage_entries && blk_ptr && entries, "grf_compact_entries: null buffer");
    compact_entries_kernel<<<grid_for_warps(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(
        (const int2 *)stage_entries, blk_ptr, n_rows, n_steps, stage_stride, (int2 *)entries);
    return check_cuda(cudaGetLastError(), "compact_entries_kernel launch");
}

extern "C" int grf_count_from_steps(const int64_t *offsets_step_major, int64_t n_rows, int32_t n_steps,
                                    int32_t *row_cnt, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, row_cnt);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1, "grf_count_from_steps: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(offsets_step_major && row_cnt, "grf_count_from_steps: null buffer");
    const int64_t n = n_rows * n_steps;
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)kSmCount * 32) g = (int64_t)kSmCount * 32;
    count_from_steps_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(offsets_step_major, n_rows, n_steps, row_cnt);
    return check_cuda(cudaGetLastError(), "count_from_steps_kernel launch");
}

extern "C" int grf_row_census(const int32_t *blk_ptr, int64_t n_rows, int32_t n_steps, int32_t threshold,
                              int32_t *census, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, census);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1 && threshold >= 1, "grf_row_census: bad shape");
    GRF_REQUIRE(census, "grf_row_census: null census");
    cudaStream_t st = (cudaStream_t)stream;
    GRF_CUDA_OK(cudaMemsetAsync(census, 0, 3 * sizeof(int32_t), st));
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr, "grf_row_census: null blk_ptr");
    int64_t g = (n_rows + 255) / 256;
    if (g > (int64_t)kSmCount * 8) g = (int64_t)kSmCount * 8;
    row_census_kernel<<<(int)g, 256, 0, st>>>(blk_ptr, n_rows, n_steps, threshold, census);
    return check_cuda(cudaGetLastError(), "row_census_kernel launch");
}

namespace grf {
// Chunk table of the long rows (GrfLongRows), built without a host round trip: the counts come from
// grf_row_census.  One packed 64-bit ticket hands a long row its slot in `rows` AND its first chunk id, so chunk
// ids grow with the slot (chunk_ptr ascending) whatever order the rows arrive in; the table's order differs from
// run to run, the sums do not (every row adds its own chunks in order).
__global__ void __launch_bounds__(256) long_rows_claim_kernel(const int32_t *__restrict__ ptr, int64_t n_rows,
                                                              int32_t L, int32_t threshold, int32_t chunk,
                                                              unsigned long long *__restrict__ ticket,
                                                              int32_t *__restrict__ rows,
                                                              int32_t *__restrict__ chunk_ptr, int32_t n_long,
                                                              int32_t n_chunks) {
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows;
         r += (int64_t)gridDim.x * blockDim.x) {
        const int32_t len = ptr[(r + 1) * L] - ptr[r * L];
        if (len > threshold) {
            const unsigned long long nch = (unsigned long long)((len + chunk - 1) / chunk);
            const unsigned long long old = atomicAdd(ticket, (1ull << 32) | nch);
            const int32_t slot = (int32_t)(old >> 32);
            if (slot < n_long) {
                rows[slot] = (int32_t)r;
                chunk_ptr[slot] = (int32_t)(old & 0xffffffffull);
            }
        }
        if (r == 0) chunk_ptr[n_long] = n_chunks;
    }
}

// bounds[k] = entry range of chunk k; first[k] = the first X row it gathers (its issue key)
__global__ void __launch_bounds__(256) long_rows_chunks_kernel(const int32_t *__restrict__ ptr, int32_t L,
                                                               int32_t chunk, const GrfEntry *__restrict__ ent,
                                                               const int32_t *__restrict__ rows,
                                                               const int32_t *__restrict__ chunk_ptr, int32_t n_long,
                                                               int32_t n_chunks, int2 *__restrict__ bounds,
                                                               int32_t *__restrict__ first) {
    for (int32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_chunks; k += gridDim.x * blockDim.x) {
        int32_t lo = 0, hi = n_long;  // last slot whose first chunk is <= k
        while (hi - lo > 1) {
            const int32_t mid = (lo + hi) >> 1;
            if (__ldg(chunk_ptr + mid) <= k) lo = mid; else hi = mid;
        }
        const int64_t r = rows[lo];
        const int32_t b = ptr[r * L] + (k - __ldg(chunk_ptr + lo)) * chunk;
        const int32_t e = min(b + chunk, ptr[(r + 1) * L]);
        bounds[k] = make_int2(b, e);
        if (first) first[k] = (int32_t)((uint32_t)ent[b].col & kColMask);
    }
}
}  // namespace grf

extern "C" int grf_long_rows_build(const int32_t *blk_ptr, const GrfEntry *entries, int64_t n_rows, int32_t n_steps,
                                   int32_t threshold, int32_t chunk, int32_t n_long, int32_t n_chunks,
                                   unsigned long long *ticket,
                                   int32_t *rows, int32_t *chunk_ptr, int32_t *bounds, int32_t *first,
                                   void *stream) {
    GRF_ON_STREAM_DEVICE(stream, rows);
    using namespace grf;
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1 && threshold >= 1 && chunk >= 1 && chunk <= threshold && n_long >= 0 &&
                    n_chunks >= n_long,
                "grf_long_rows_build: bad shape");
    if (n_long == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && ticket && rows && chunk_ptr && bounds, "grf_long_rows_build: null buffer");
    GRF_REQUIRE(!first || entries, "grf_long_rows_build: issue keys need the entries");
    cudaStream_t st = (cudaStream_t)stream;
    GRF_CUDA_OK(cudaMemsetAsync(ticket, 0, sizeof(unsigned long long), st));
    int64_t g = (n_rows + 255) / 256;
    if (g > (int64_t)kSmCount * 16) g = (int64_t)kSmCount * 16;
    long_rows_claim_kernel<<<(int)g, 256, 0, st>>>(blk_ptr, n_rows, n_steps, threshold, chunk, ticket, rows, chunk_ptr,
                                                   n_long, n_chunks);
    GRF_CUDA_OK(cudaGetLastError());
    int64_t gc = ((int64_t)n_chunks + 255) / 256;
    if (gc > (int64_t)kSmCount * 16) gc = (int64_t)kSmCount * 16;
    long_rows_chunks_kernel<<<(int)gc, 256, 0, st>>>(blk_ptr, n_steps, chunk, entries, rows, chunk_ptr, n_long,
                                                     n_chunks, (int2 *)bounds, first);
    return check_cuda(cudaGetLastError(), "long_rows kernels launch");
}

extern "C" int grf_blocks_from_steps(const int64_t *offsets_step_major, const int32_t *col, const double *val,
                                     const int32_t *blk_ptr, int64_t n_rows, int32_t n_steps, GrfEntry *entries,
                                     void *stream) {
    GRF_ON_STREAM_DEVICE(stream, blk_ptr);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1, "grf_blocks_from_steps: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(offsets_step_major && blk_ptr, "grf_blocks_from_steps: null buffer");
    blocks_from_steps_kernel<<<grid_for_warps(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(
        offsets_step_major, col, val, blk_ptr, n_rows, n_steps, entries);
    return check_cuda(cudaGetLastError(), "blocks_from_steps_kernel launch");
}

extern "C" int grf_nonempty_rows(const int32_t *blk_ptr, int64_t n_rows, int32_t n_steps, int32_t *flags,
                                 int32_t *pos, void *scan_workspace, int32_t *ids, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, flags);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1, "grf_nonempty_rows: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && flags && pos && scan_workspace && ids, "grf_nonempty_rows: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = (n_rows + 255) / 256;
    if (g > (int64_t)kSmCount * 16) g = (int64_t)kSmCount * 16;
    row_flags_kernel<<<(int)g, 256, 0, st>>>(blk_ptr, n_rows, n_steps, flags);
    GRF_CUDA_OK(cudaGetLastError());
    const int rc = grf_scan_counts(flags, n_rows, 1, GRF_ORDER_ROW_MAJOR, pos, 0, scan_workspace, stream);
    if (rc != GRF_OK) return rc;
    scatter_ids_kernel<<<(int)g, 256, 0, st>>>(flags, pos, n_rows, ids);
    return check_cuda(cudaGetLastError(), "nonempty_rows kernels launch");
}
