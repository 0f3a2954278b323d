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

// ---------------------------------------------------------------------------
// Phi^T blocks
// ---------------------------------------------------------------------------
// A row's L per-length segments are contiguous and every entry carries its length in the top bits
// of `col`, so a warp takes the whole row as one flat run: ~2 rounds of 32 entries per row at
// config 2 instead of one (mostly short) round per length -- the kernels are bound by the latency
// of the atomic round trips, i.e. by the number of rounds.
__global__ void __launch_bounds__(256) transpose_count_kernel(const int32_t *blk_ptr, const GrfEntry *entries,
                                                              int64_t n_rows, int32_t L, int32_t *tcnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t b = blk_ptr[r * L], e = blk_ptr[(r + 1) * L];
        for (int32_t i = b + lane; i < e; i += 32) {
            const int32_t packed = entries[i].col;
            atomicAdd(&tcnt[(int64_t)entry_col(packed) * L + entry_step(packed)], 1);
        }
    }
}

__global__ void __launch_bounds__(256) transpose_fill_kernel(const int32_t *blk_ptr, const GrfEntry *entries,
                                                             int64_t n_rows, int32_t L, int32_t *cursor,
                                                             GrfEntry *tentries) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t b = blk_ptr[r * L], e = blk_ptr[(r + 1) * L];
        for (int32_t i = b + lane; i < e; i += 32) {
            const GrfEntry src = entries[i];
            const int s = entry_step(src.col);
            const int32_t slot = atomicAdd(&cursor[(int64_t)entry_col(src.col) * L + s], 1);
            GrfEntry dst;
            dst.col = pack_col((int32_t)r, s);
            dst.val = src.val;
            tentries[slot] = dst;
        }
    }
}

// Slots inside one (column, length) segment were claimed in arbitrary order;
// order them by row so that the layout (and the fp32 summation order of
// Phi^T V) is deterministic.  Three tiers by segment length:
//   <= 32   one thread per segment: keys (row << 5 | slot) sorted by a register
//           sorting network (rows are unique and < 2^27, so the key is 32 bits),
//           values re-read through the sorted slot index;
//   <= 256  one warp per segment, register bitonic sort with shuffles;
//   longer  (hub columns) one CTA per segment, bitonic network in global memory.
// (v1 used a per-thread insertion sort in global memory: 420 us at config 2.)
constexpr int kTinySeg = 8;
constexpr int kShortSeg = 32;
constexpr int kMidSeg = 256;

template <int C>
__device__ __forceinline__ void sort_segment_regs(GrfEntry *seg, int len) {
    uint32_t key[C];
    const uint32_t step_bits = (uint32_t)seg[0].col & ~kColMask;
#pragma unroll
    for (int i = 0; i < C; ++i)
        key[i] = i < len ? ((((uint32_t)seg[i].col & kColMask) << 5) | (uint32_t)i) : 0xffffffffu;
    sort_network_u32<C>(key);
    float val[C];
#pragma unroll
    for (int i = 0; i < C; ++i) val[i] = i < len ? seg[key[i] & 31u].val : 0.f;
#pragma unroll
    for (int i = 0; i < C; ++i) {
        if (i < len) {
            GrfEntry e;
            e.col = (int32_t)(step_bits | (key[i] >> 5));
            e.val = val[i];
            seg[i] = e;
        }
    }
}

// lists: [0] = number of mid segments, [1] = number of long segments, then the two lists
//
// A warp takes 32 consecutive segments (one per lane).  Their entries are one contiguous stretch of
// tentries: it is copied to shared memory with coalesced 8-byte loads, every lane sorts its own
// segment there, and the stretch is written back coalesced.  (One thread per segment straight on
// global memory -- 32 lanes striding through 32 different segments -- took 66 us at config 2.)
constexpr int kSortStage = 1024;  // entries a warp stages (32 segments of <= 32 entries always fit)

__global__ void __launch_bounds__(128) transpose_sort_short_kernel(const int32_t *tblk_ptr, int64_t n_segs,
                                                                   GrfEntry *tentries, int32_t *counts,
                                                                   int32_t *mid_list, int32_t *long_list) {
    __shared__ int2 stage_all[4][kSortStage];
    const int lane = threadIdx.x & 31;
    int2 *stage = stage_all[threadIdx.x >> 5];
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int2 *ent2 = reinterpret_cast<int2 *>(tentries);
    for (int64_t g0 = warp0 * 32; g0 < n_segs; g0 += nwarps * 32) {
        const int64_t g = g0 + lane;
        int32_t b = 0, len = 0;
        if (g < n_segs) {
            b = tblk_ptr[g];
            len = tblk_ptr[g + 1] - b;
        }
        // longer segments go to the warp / CTA tiers; their entries are staged and written back untouched
        if (len > kMidSeg)
            long_list[-atomicAdd(&counts[1], 1)] = (int32_t)g;  // grows downwards from the end
        else if (len > kShortSeg)
            mid_list[atomicAdd(&counts[0], 1)] = (int32_t)g;   // grows upwards
        const bool mine = len > 1 && len <= kShortSeg;
        if (!__any_sync(0xffffffffu, mine)) continue;
        const int32_t base = __shfl_sync(0xffffffffu, b, 0);
        const int last = (int)min((int64_t)31, n_segs - 1 - g0);
        const int32_t total = __shfl_sync(0xffffffffu, b + len, last) - base;
        const bool staged = total <= kSortStage;
        GrfEntry *seg = tentries + b;
        if (staged) {
            for (int32_t i = lane; i < total; i += 32) stage[i] = ent2[base + i];
            __syncwarp();
            seg = reinterpret_cast<GrfEntry *>(stage) + (b - base);
        }
        if (mine) {
            if (len <= kTinySeg)
                sort_segment_regs<kTinySeg>(seg, len);
            else if (len <= 16)
                sort_segment_regs<16>(seg, len);
            else
                sort_segment_regs<kShortSeg>(seg, len);
        }
        if (staged) {
            __syncwarp();
            for (int32_t i = lane; i < total; i += 32) ent2[base + i] = stage[i];
            __syncwarp();
        }
    }
}

template <int KPL>
__device__ __forceinline__ void sort_segment_warp(GrfEntry *seg, int len, int lane) {
    unsigned long long key[KPL];
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const int i = r * 32 + lane;
        unsigned long long kk = ~0ull;
        if (i < len) {
            const GrfEntry e = seg[i];
            kk = ((unsigned long long)(uint32_t)e.col << 32) | (unsigned long long)(uint32_t)__float_as_int(e.val);
        }
        key[r] = kk;
    }
    __syncwarp();
    warp_bitonic_sort<unsigned long long, KPL>(key, lane);
#pragma unroll
    for (int r = 0; r < KPL; ++r) {
        const int i = lane * KPL + r;
        if (i < len) {
            GrfEntry e;
            e.col = (int32_t)(uint32_t)(key[r] >> 32);
            e.val = __int_as_float((int)(uint32_t)key[r]);
            seg[i] = e;
        }
    }
}

__global__ void __launch_bounds__(128) transpose_sort_mid_kernel(const int32_t *tblk_ptr, GrfEntry *tentries,
                                                                 const int32_t *counts, const int32_t *mid_list) {
    const int lane = threadIdx.x & 31;
    const int n_mid = counts[0];
    const int warp0 = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)(((int64_t)gridDim.x * blockDim.x) >> 5);
    for (int li = warp0; li < n_mid; li += nwarps) {
        const int32_t g = mid_list[li];
        const int32_t b = tblk_ptr[g];
        const int len = tblk_ptr[g + 1] - b;
        GrfEntry *seg = tentries + b;
        if (len <= 64)
            sort_segment_warp<2>(seg, len, lane);
        else if (len <= 128)
            sort_segment_warp<4>(seg, len, lane);
        else
            sort_segment_warp<8>(seg, len, lane);
        __syncwarp();
    }
}

// One CTA per long segment.  Segments of up to kSmemSortCap entries (128 KB) are staged in shared
// memory, sorted there and written back; only longer ones run the network on global memory.  (On an
// R-MAT graph of 2^20 nodes 71 k segments hold 60 % of all entries, the largest 83 k entries: the
// global-memory network took 26.9 ms for them.)
// Three launches by size class, (256, 1024] with 256 threads, (1024, 4096] with 512, the rest with
// 1024: a small segment spends its time in the ~50 barriers of the network, which cost less with
// fewer warps, and the small classes leave room for several CTAs per SM.
constexpr int kSmemSortCap = 16384;

// Bitonic network in its all-ascending form (mirror step, then half-cleaners): every comparator
// moves the smaller key to the lower index, so the virtual +inf padding at [len, np2) never moves
// and comparators that touch it are simply skipped.  All threads of the CTA call these together.
__device__ __forceinline__ void sort_compare_swap(GrfEntry *seg, uint32_t lo, uint32_t hi) {
    const GrfEntry a = seg[lo], bb = seg[hi];
    if (a.col > bb.col) {
        seg[lo] = bb;
        seg[hi] = a;
    }
}

__device__ __forceinline__ void sort_mirror_stage(GrfEntry *seg, uint32_t len, uint32_t np2, uint32_t k) {
    const uint32_t half = k >> 1;
    for (uint32_t c = threadIdx.x; c < np2 / 2; c += blockDim.x) {
        const uint32_t blk = c / half, off = c - blk * half;
        const uint32_t lo = blk * k + off, hi = blk * k + k - 1 - off;
        if (hi < len) sort_compare_swap(seg, lo, hi);
    }
    __syncthreads();
}

__device__ __forceinline__ void sort_clean_stages(GrfEntry *seg, uint32_t len, uint32_t np2, uint32_t j_from,
                                                  uint32_t j_to) {
    for (uint32_t j = j_from; j >= j_to && j > 0; j >>= 1) {
        for (uint32_t c = threadIdx.x; c < np2 / 2; c += blockDim.x) {
            const uint32_t lo = ((c & ~(j - 1)) << 1) | (c & (j - 1));
            const uint32_t hi = lo + j;
            if (hi < len) sort_compare_swap(seg, lo, hi);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void sort_copy(GrfEntry *dst, const GrfEntry *src, uint32_t n) {
    const int2 *s2 = reinterpret_cast<const int2 *>(src);
    int2 *d2 = reinterpret_cast<int2 *>(dst);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) d2[i] = s2[i];
    __syncthreads();
}

template <int kThreads>
__global__ void __launch_bounds__(kThreads) transpose_sort_long_kernel(const int32_t *tblk_ptr, GrfEntry *tentries,
                                                                       const int32_t *counts,
                                                                       const int32_t *long_list, uint32_t len_above,
                                                                       uint32_t len_upto, uint32_t smem_cap) {
    extern __shared__ __align__(16) unsigned char sort_smem[];
    GrfEntry *sbuf = reinterpret_cast<GrfEntry *>(sort_smem);
    const int32_t n_long = counts[1];
    for (int32_t li = blockIdx.x; li < n_long; li += gridDim.x) {
        const int32_t g = long_list[-li];
        const int32_t b = tblk_ptr[g];
        const uint32_t len = (uint32_t)(tblk_ptr[g + 1] - b);
        if (len <= len_above || len > len_upto) continue;  // another launch's size class (uniform over the CTA)
        GrfEntry *seg = tentries + b;
        if (len <= smem_cap) {
            // the whole segment fits: one trip through shared memory
            const uint32_t np2 = next_pow2(len);
            sort_copy(sbuf, seg, len);
            for (uint32_t k = 2; k <= np2; k <<= 1) {
                sort_mirror_stage(sbuf, len, np2, k);
                sort_clean_stages(sbuf, len, np2, k >> 2, 1);
            }
            sort_copy(seg, sbuf, len);
            continue;
        }
        // Longer than the staging area (smem_cap = T entries, a power of two): every stage whose
        // comparators stay inside a T-aligned tile runs on the tile in shared memory; only the
        // mirror steps and the half-cleaners at distance >= T touch global memory.  For the 83 k-entry
        // hub segment of the R-MAT test graph that is 6 global stages instead of 153.
        const uint32_t T = smem_cap;
        const uint32_t np2 = next_pow2(len);
        for (uint32_t t0 = 0; t0 < len; t0 += T) {  // tiles sorted on their own (k = 2 .. T)
            const uint32_t tl = min(T, len - t0);
            sort_copy(sbuf, seg + t0, tl);
            for (uint32_t k = 2; k <= T; k <<= 1) {
                sort_mirror_stage(sbuf, tl, T, k);
                sort_clean_stages(sbuf, tl, T, k >> 2, 1);
            }
            sort_copy(seg + t0, sbuf, tl);
        }
        for (uint32_t k = 2 * T; k <= np2; k <<= 1) {
            sort_mirror_stage(seg, len, np2, k);
            sort_clean_stages(seg, len, np2, k >> 2, T);  // distances k/4 .. T on global memory
            for (uint32_t t0 = 0; t0 < len; t0 += T) {      // distances T/2 .. 1 inside the tiles
                const uint32_t tl = min(T, len - t0);
                sort_copy(sbuf, seg + t0, tl);
                sort_clean_stages(sbuf, tl, T, T >> 1, 1);
                sort_copy(seg + t0, sbuf, tl);
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
    GRF_ON_STREAM_DEVICE(stream);
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
    GRF_ON_STREAM_DEVICE(stream);
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
    GRF_ON_STREAM_DEVICE(stream);
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

extern "C" int grf_count_from_steps(const int64_t *offsets_step_major, int64_t n_rows, int32_t n_steps,
                                    int32_t *row_cnt, void *stream) {
    GRF_ON_STREAM_DEVICE(stream);
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
    GRF_ON_STREAM_DEVICE(stream);
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

extern "C" int grf_blocks_from_steps(const int64_t *offsets_step_major, const int32_t *col, const double *val,
                                     const int32_t *blk_ptr, int64_t n_rows, int32_t n_steps, GrfEntry *entries,
                                     void *stream) {
    GRF_ON_STREAM_DEVICE(stream);
    GRF_REQUIRE(n_rows >= 0 && n_steps >= 1, "grf_blocks_from_steps: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(offsets_step_major && blk_ptr, "grf_blocks_from_steps: null buffer");
    blocks_from_steps_kernel<<<grid_for_warps(n_rows, 256), 256, 0, (cudaStream_t)stream>>>(
        offsets_step_major, col, val, blk_ptr, n_rows, n_steps, entries);
    return check_cuda(cudaGetLastError(), "blocks_from_steps_kernel launch");
}

extern "C" int grf_nonempty_rows(const int32_t *blk_ptr, int64_t n_rows, int32_t n_steps, int32_t *flags,
                                 int32_t *pos, void *scan_workspace, int32_t *ids, void *stream) {
    GRF_ON_STREAM_DEVICE(stream);
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

static int64_t transpose_cursor_bytes(int64_t n_cols, int32_t n_steps) {
    return ((n_cols * n_steps + 2) * (int64_t)sizeof(int32_t) + 15) / 16 * 16;
}

extern "C" int64_t grf_transpose_workspace_bytes(int64_t n_cols, int32_t n_steps) {
    // [segment counts, later the fill cursor / sort work lists: n_cols*L + 2 ints][scan workspace][census: 8 ints]
    return transpose_cursor_bytes(n_cols, n_steps) + grf_scan_workspace_bytes(n_cols * n_steps) + 32;
}

extern "C" int grf_transpose_offsets(const int32_t *blk_ptr, const GrfEntry *entries, int64_t n_rows, int64_t n_cols,
                                     int32_t n_steps, const int32_t *col_counts, int32_t *tblk_ptr, void *workspace,
                                     int32_t census_threshold, int32_t *census_host, void *stream) {
    GRF_ON_STREAM_DEVICE(stream);
    GRF_REQUIRE(n_rows >= 0 && n_cols >= 0 && n_steps >= 1, "grf_transpose_offsets: bad shape");
    GRF_REQUIRE(tblk_ptr && workspace, "grf_transpose_offsets: null buffer");
    GRF_REQUIRE(!census_host || census_threshold >= 1, "grf_transpose_offsets: bad census threshold");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *tcnt = (int32_t *)workspace;
    char *scan_ws = (char *)workspace + transpose_cursor_bytes(n_cols, n_steps);
    int32_t *census = (int32_t *)(scan_ws + grf_scan_workspace_bytes(n_cols * n_steps));
    if (col_counts) {
        tcnt = const_cast<int32_t *>(col_counts);  // the walker counted while it emitted the entries
    } else if (n_cols > 0) {
        GRF_CUDA_OK(cudaMemsetAsync(tcnt, 0, (size_t)n_cols * n_steps * sizeof(int32_t), st));
        if (n_rows > 0) {
            GRF_REQUIRE(blk_ptr, "grf_transpose_offsets: null blk_ptr");
            transpose_count_kernel<<<grid_for_warps(n_rows, 256), 256, 0, st>>>(blk_ptr, entries, n_rows, n_steps,
                                                                                 tcnt);
            GRF_CUDA_OK(cudaGetLastError());
        }
    }
    int rc = grf_scan_counts(tcnt, n_cols, n_steps, GRF_ORDER_ROW_MAJOR, tblk_ptr, 0, scan_ws, stream);
    if (rc != GRF_OK) return rc;
    if (census_host) {
        // row statistics of both sides while the fill / sort kernels that follow keep the GPU busy: the
        // host reads them from pinned memory after an event recorded behind this call
        rc = grf_row_census(blk_ptr, n_rows, n_steps, census_threshold, census, stream);
        if (rc != GRF_OK) return rc;
        rc = grf_row_census(tblk_ptr, n_cols, n_steps, census_threshold, census + 3, stream);
        if (rc != GRF_OK) return rc;
        GRF_CUDA_OK(cudaMemcpyAsync(census_host, census, 6 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    return GRF_OK;
}

extern "C" int grf_transpose_fill(const int32_t *blk_ptr, const GrfEntry *entries, int64_t n_rows, int64_t n_cols,
                                  int32_t n_steps, const int32_t *tblk_ptr, int32_t *cursor, GrfEntry *tentries,
                                  void *stream) {
    GRF_ON_STREAM_DEVICE(stream);
    GRF_REQUIRE(n_rows >= 0 && n_cols >= 0 && n_steps >= 1, "grf_transpose_fill: bad shape");
    if (n_cols == 0 || n_rows == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && tblk_ptr && cursor, "grf_transpose_fill: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n_segs = n_cols * n_steps;
    GRF_CUDA_OK(cudaMemcpyAsync(cursor, tblk_ptr, (size_t)n_segs * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
    transpose_fill_kernel<<<grid_for_warps(n_rows, 256), 256, 0, st>>>(blk_ptr, entries, n_rows, n_steps, cursor,
                                                                        tentries);
    GRF_CUDA_OK(cudaGetLastError());
    // deterministic order inside every segment; `cursor` is dead now and is reused for the
    // work lists: [n_mid, n_long | mid list growing up ... long list growing down]; at most
    // n_segs segments are listed in total, which is what the n_segs + 2 ints of scratch hold
    GRF_CUDA_OK(cudaMemsetAsync(cursor, 0, 2 * sizeof(int32_t), st));
    if (n_segs >= 2) {
        int64_t g = (n_segs + 127) / 128;
        if (g > (int64_t)kSmCount * 64) g = (int64_t)kSmCount * 64;
        int32_t *mid_list = cursor + 2;
        int32_t *long_list = cursor + 2 + (n_segs - 1);
        transpose_sort_short_kernel<<<(int)g, 128, 0, st>>>(tblk_ptr, n_segs, tentries, cursor, mid_list, long_list);
        GRF_CUDA_OK(cudaGetLastError());
        transpose_sort_mid_kernel<<<kSmCount * 8, 128, 0, st>>>(tblk_ptr, tentries, cursor, mid_list);
        GRF_CUDA_OK(cudaGetLastError());
        const size_t big_smem = (size_t)kSmemSortCap * sizeof(GrfEntry);
        GRF_CUDA_OK(cudaFuncSetAttribute(transpose_sort_long_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)big_smem));
        transpose_sort_long_kernel<256><<<kSmCount * 8, 256, 1024 * sizeof(GrfEntry), st>>>(
            tblk_ptr, tentries, cursor, long_list, (uint32_t)kMidSeg, 1024u, 1024u);
        GRF_CUDA_OK(cudaGetLastError());
        transpose_sort_long_kernel<512><<<kSmCount * 4, 512, 4096 * sizeof(GrfEntry), st>>>(
            tblk_ptr, tentries, cursor, long_list, 1024u, 4096u, 4096u);
        GRF_CUDA_OK(cudaGetLastError());
        transpose_sort_long_kernel<1024><<<kSmCount, 1024, big_smem, st>>>(
            tblk_ptr, tentries, cursor, long_list, 4096u, 0xffffffffu, (uint32_t)kSmemSortCap);
    }
    return check_cuda(cudaGetLastError(), "transpose kernels launch");
}
