// Phi^T blocks by a stable LSD radix sort (K4 in SURVEY.md 2a).
//
// Replaces sparse_lo.py:23-25 (`.t().to_sparse_csr()`, redone by the reference on every forward)
// with one build per Phi.  Round 1 scattered the entries through one atomic cursor per (column,
// length) segment and then sorted every segment by row: at config 4 (521 M entries, hub columns
// with 6*10^5 entries) that took 65 ms -- 26 ms of returning atomics at 2 % issue utilisation and
// 32 ms of segment sorts whose tail is one CTA per hub column -- twice the walker.  Sorting the
// entries by the key  column * L + length  with a STABLE least-significant-digit radix sort needs
// no atomics on global memory, no per-segment pass and no size classes: the input is ordered by
// row, so every (column, length) segment comes out ordered by row -- the fixed summation order
// of Phi^T V -- and hub columns cost the same per entry as any other.
//
// Passes: ceil(key bits / digit bits) with digit width <= 11 (config 2: 19 bits -> 2 passes;
// config 4: 25 bits -> 3).  Per pass, three launches:
//   radix_hist     per-CTA digit histogram of the CTA's contiguous span of tiles (shared-memory
//                  atomics), written digit-major: hist[digit][cta]
//   scan           exclusive scan of hist (reuses the grf_scan_counts kernels)
//   radix_scatter  the CTA walks its tiles in order; inside a tile a warp owns 32*K consecutive
//                  items and ranks them digit by digit with match.any + a warp-private counter
//                  table in shared memory, the tables are scanned across the 8 warps, and every
//                  item goes to  base[digit] + warp offset + rank  -- stable by construction.
// Items are (key u32, payload {length << 27 | local row, value}) = 12 B; the last pass writes the
// payload only, straight into the Phi^T entries.  HBM-bound streaming: per entry 8 + 12 (keys +
// payload out of the row pass) + passes * (4 + 12 + 12) bytes.

#include "grf_common.cuh"

namespace grf {

constexpr int kRadixThreads = 256;
constexpr int kRadixWarps = kRadixThreads / 32;
constexpr int kRadixItems = 16;                                  // items per thread and tile
constexpr int kRadixTile = kRadixThreads * kRadixItems;          // 4096
constexpr int kRadixMaxBits = 10;                                // 1024 bins: 8 warp tables of u16 = 16 KB

// keys[i] = column * L + length, payload[i] = {length << 27 | row - row0, value} for the entries of
// rows [0, n_rows) of this block (one warp per row; i counts from the block's first entry)
__global__ void __launch_bounds__(256) transpose_keys_kernel(const int32_t *__restrict__ blk_ptr,
                                                             const int2 *__restrict__ entries, int64_t n_rows,
                                                             int32_t L, int32_t entry_lo,
                                                             uint32_t *__restrict__ keys, int2 *__restrict__ payload) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t b = __ldg(blk_ptr + r * L), e = __ldg(blk_ptr + (r + 1) * L);
        for (int32_t i = b + lane; i < e; i += 32) {
            const int2 en = __ldg(entries + i);
            const uint32_t col = (uint32_t)en.x & kColMask, len = (uint32_t)en.x >> kStepShift;
            keys[i - entry_lo] = col * (uint32_t)L + len;
            payload[i - entry_lo] = make_int2((int32_t)((len << kStepShift) | (uint32_t)r), en.y);
        }
    }
}

__global__ void __launch_bounds__(kRadixThreads) radix_hist_kernel(const uint32_t *__restrict__ keys, int64_t n,
                                                                   int32_t shift, int32_t nbins,
                                                                   int32_t tiles_per_cta,
                                                                   int32_t *__restrict__ hist) {
    extern __shared__ uint32_t sh_hist[];
    for (int d = threadIdx.x; d < nbins; d += kRadixThreads) sh_hist[d] = 0u;
    __syncthreads();
    const uint32_t mask = (uint32_t)nbins - 1u;
    const int64_t begin = (int64_t)blockIdx.x * tiles_per_cta * kRadixTile;
    const int64_t end = min(n, begin + (int64_t)tiles_per_cta * kRadixTile);
    for (int64_t i0 = begin + threadIdx.x; i0 < end; i0 += (int64_t)kRadixThreads * 4) {
        uint32_t k[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (int64_t)u * kRadixThreads;
            k[u] = i < end ? __ldg(keys + i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (int64_t)u * kRadixThreads;
            if (i < end) atomicAdd(&sh_hist[(k[u] >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < nbins; d += kRadixThreads) hist[(int64_t)d * gridDim.x + blockIdx.x] = (int32_t)sh_hist[d];
}

// ---- bulk-async staging (cp.async.bulk = SASS UBLKCP, completion on an mbarrier) ------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// kLast: the payload goes to its final place and the keys are dropped.
//
// Per tile of 4096 items (a CTA walks its contiguous span of tiles in order):
//   stage     the tile's keys (16 KB) and payloads (32 KB) arrive in shared memory by bulk-async copies
//             that one thread issued a tile ahead -- keys double-buffered (the next tile's keys land while
//             this tile is ranked), payloads single-buffered (needed only at write-out; re-armed as soon as
//             the previous write-out has finished) -- each completing on its own mbarrier;
//   rank      every warp ranks its 512 consecutive items digit by digit: match.any finds the lanes with
//             the same digit, the lowest of them bumps the warp's private counter;
//   scan      the warp tables are scanned across the 8 warps and the tile totals across the digits: every
//             item now has its place in the tile's digit-sorted order;
//   exchange  the item's index goes to that place (a 2-byte permutation, not the 12-byte item);
//   write-out consecutive threads take consecutive places, fetch key and payload through the permutation
//             from the staged tile and store them: a store instruction covers a few output runs.
// Stable by construction.  History (ncu, 137 M entries): v1 wrote every item from the thread that loaded it
// -- 32 lanes x 32 digit streams per store, 1.7 GB working set against a 256 MB TLB reach: 7 % issue
// utilisation, 58 k cycles per tile.  v2 ordered the tile in shared memory first: 2.1 ms per pass, but 58 %
// of the stall samples were long-scoreboard waits on the tile's own (synchronous) loads.  v3 = this.
template <bool kLast>
__global__ void __launch_bounds__(kRadixThreads, 2) radix_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                         const int2 *__restrict__ pay_in,
                                                                         int64_t n, int32_t shift, int32_t nbins,
                                                                         int32_t tiles_per_cta,
                                                                         const int32_t *__restrict__ offsets,
                                                                         uint32_t *__restrict__ keys_out,
                                                                         int2 *__restrict__ pay_out) {
    extern __shared__ __align__(128) unsigned char sh_bytes[];
    // [payload stage 32 KB][key stage 2 x 16 KB][permutation 8 KB][gbase][dstart][warp tables]
    int2 *spay = reinterpret_cast<int2 *>(sh_bytes);
    uint32_t *skey0 = reinterpret_cast<uint32_t *>(spay + kRadixTile);
    unsigned short *xsrc = reinterpret_cast<unsigned short *>(skey0 + 2 * kRadixTile);
    uint32_t *gbase = reinterpret_cast<uint32_t *>(xsrc + kRadixTile);  // global slot of the tile's first item of each digit
    uint32_t *dstart = gbase + nbins;                                   // first place of each digit in the tile's order
    unsigned short *cnt = reinterpret_cast<unsigned short *>(dstart + nbins);  // [warps][nbins]
    __shared__ uint32_t warp_sums[kRadixWarps];
    __shared__ __align__(8) uint64_t bar_key[2], bar_pay;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t mask = (uint32_t)nbins - 1u;
    const unsigned lt = (1u << lane) - 1u;
    // digits owned by this thread in the scans: dpt consecutive ones
    const int dpt = nbins >= kRadixThreads ? nbins / kRadixThreads : 1;
    const int d0 = threadIdx.x * dpt;
    const bool owner = d0 < nbins;
    const int64_t begin = (int64_t)blockIdx.x * tiles_per_cta * kRadixTile;
    const int64_t end = min(n, begin + (int64_t)tiles_per_cta * kRadixTile);
    if (begin >= end) return;
    for (int d = threadIdx.x; d < nbins; d += kRadixThreads) gbase[d] = (uint32_t)offsets[(int64_t)d * gridDim.x + blockIdx.x];
    // bytes of a tile's copies, rounded up to the 16-byte granule of cp.async.bulk (the buffers are padded to it)
    auto key_bytes = [&](int64_t tile) { return (uint32_t)(((min((int64_t)kRadixTile, end - tile) * 4) + 15) & ~15ll); };
    auto pay_bytes = [&](int64_t tile) { return (uint32_t)(((min((int64_t)kRadixTile, end - tile) * 8) + 15) & ~15ll); };
    if (threadIdx.x == 0) {
        mbar_init(&bar_key[0], 1);
        mbar_init(&bar_key[1], 1);
        mbar_init(&bar_pay, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(&bar_key[0], key_bytes(begin));
        bulk_g2s(skey0, keys_in + begin, key_bytes(begin), &bar_key[0]);
        mbar_expect_tx(&bar_pay, pay_bytes(begin));
        bulk_g2s(spay, pay_in + begin, pay_bytes(begin), &bar_pay);
        if (begin + kRadixTile < end) {
            mbar_expect_tx(&bar_key[1], key_bytes(begin + kRadixTile));
            bulk_g2s(skey0 + kRadixTile, keys_in + begin + kRadixTile, key_bytes(begin + kRadixTile), &bar_key[1]);
        }
    }
    unsigned short *mycnt = cnt + w * nbins;
    uint32_t it = 0;
    for (int64_t tile = begin; tile < end; tile += kRadixTile, ++it) {
        const uint32_t *skey = skey0 + (it & 1u) * kRadixTile;
        {   // clear the warp tables (warps * nbins u16)
            uint32_t *z = reinterpret_cast<uint32_t *>(cnt);
            for (int i = threadIdx.x; i < nbins * kRadixWarps / 2; i += kRadixThreads) z[i] = 0u;
        }
        __syncthreads();  // (also: the mbarrier inits are visible before the first wait)
        mbar_wait(&bar_key[it & 1u], (it >> 1) & 1u);
        const int n_tile = (int)min((int64_t)kRadixTile, end - tile);
        unsigned short rank[kRadixItems];
        const int lbase = w * (32 * kRadixItems) + lane;  // this lane's first item in the tile
#pragma unroll
        for (int k = 0; k < kRadixItems; ++k) {
            const int li = lbase + k * 32;
            const bool valid = li < n_tile;
            const uint32_t d = valid ? (skey[li] >> shift) & mask : 0u;
            // lanes past the end get a private pseudo-digit so that they match nobody
            const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : (0x80000000u | (uint32_t)lane));
            unsigned short prev = 0;
            if (valid) prev = mycnt[d];
            __syncwarp();
            if (valid && (peers & lt) == 0u) mycnt[d] = (unsigned short)(prev + __popc(peers));  // lowest peer updates
            __syncwarp();
            rank[k] = (unsigned short)(prev + __popc(peers & lt));
        }
        __syncthreads();
        // per digit: exclusive scan of the warp tables across the warps (left in place) and the tile total
        uint32_t tot[4];  // dpt <= 4 (1024 bins / 256 threads)
        uint32_t mine = 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            tot[j] = 0u;
            if (j < dpt && owner) {
                const int d = d0 + j;
                uint32_t run = 0u;
#pragma unroll
                for (int ww = 0; ww < kRadixWarps; ++ww) {
                    const uint32_t c = cnt[ww * nbins + d];
                    cnt[ww * nbins + d] = (unsigned short)run;
                    run += c;
                }
                tot[j] = run;
                mine += run;
            }
        }
        // exclusive scan of the tile totals over the digits (thread order == digit order)
        uint32_t incl = mine;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const uint32_t o = __shfl_up_sync(0xffffffffu, incl, dd);
            if (lane >= dd) incl += o;
        }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        uint32_t before = 0u;
#pragma unroll
        for (int ww = 0; ww < kRadixWarps; ++ww) before += ww < w ? warp_sums[ww] : 0u;
        uint32_t run = before + incl - mine;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < dpt && owner) {
                dstart[d0 + j] = run;
                run += tot[j];
            }
        }
        __syncthreads();
        // exchange: the item's index goes to its place in the tile's digit-sorted order
#pragma unroll
        for (int k = 0; k < kRadixItems; ++k) {
            const int li = lbase + k * 32;
            if (li < n_tile) {
                const uint32_t d = (skey[li] >> shift) & mask;
                xsrc[dstart[d] + mycnt[d] + rank[k]] = (unsigned short)li;
            }
        }
        __syncthreads();
        mbar_wait(&bar_pay, it & 1u);
#pragma unroll 4
        for (int i = threadIdx.x; i < n_tile; i += kRadixThreads) {
            const int src = xsrc[i];
            const uint32_t kk = skey[src];
            const uint32_t d = (kk >> shift) & mask;
            const uint32_t pos = gbase[d] + ((uint32_t)i - dstart[d]);
            if (!kLast) keys_out[pos] = kk;
            pay_out[pos] = spay[src];
        }
        __syncthreads();  // the staged tile has been consumed: its buffers may be refilled
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j < dpt && owner) gbase[d0 + j] += tot[j];
        if (threadIdx.x == 0) {
            const int64_t next = tile + kRadixTile, next2 = tile + 2 * (int64_t)kRadixTile;
            if (next < end) {
                mbar_expect_tx(&bar_pay, pay_bytes(next));
                bulk_g2s(spay, pay_in + next, pay_bytes(next), &bar_pay);
            }
            if (next2 < end) {  // into the key buffer this tile just released
                mbar_expect_tx(&bar_key[it & 1u], key_bytes(next2));
                bulk_g2s(skey0 + (it & 1u) * kRadixTile, keys_in + next2, key_bytes(next2), &bar_key[it & 1u]);
            }
        }
        // the next tile clears `cnt` and rewrites dstart / the permutation only after its own barriers;
        // gbase[d] is touched by its owner thread alone
    }
}

struct RadixPlan {
    int passes;
    int bits[4];
    int grid;
    int tiles_per_cta;
};

static RadixPlan radix_plan(int64_t n_keys_space, int64_t n) {
    RadixPlan p;
    const int key_bits = bit_width((uint64_t)(n_keys_space > 1 ? n_keys_space - 1 : 1));
    p.passes = (key_bits + kRadixMaxBits - 1) / kRadixMaxBits;
    if (p.passes < 1) p.passes = 1;
    // spread the bits evenly over the passes (narrow digits keep the per-tile scan of the warp tables short)
    const int base = key_bits / p.passes, extra = key_bits % p.passes;
    // the wider digits go last: the last pass writes payloads only (8-byte items), so its shorter digit runs
    // still fill whole sectors; a 9-bit first pass wrote 32-byte key runs at arbitrary offsets (ncu, config 4:
    // 8.3 GB written for 6.3 GB of items, 6.7 ms against 4.1 ms for the 8-bit second pass)
    // (widest first: config 4 18.9 ms, config 2 0.298 ms; widest last: 17.2 ms, 0.272 ms)
    for (int i = 0; i < p.passes; ++i) p.bits[i] = base + (i >= p.passes - extra ? 1 : 0);
    for (int i = 0; i < p.passes; ++i)
        if (p.bits[i] < 1) p.bits[i] = 1;
    const int64_t tiles = (n + kRadixTile - 1) / kRadixTile;
    int64_t grid = (int64_t)kSmCount * 2;  // two resident CTAs per SM (84 .. 96 KB of shared memory each)
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    p.grid = (int)grid;
    p.tiles_per_cta = (int)((tiles + grid - 1) / grid);
    return p;
}

static inline int64_t align16(int64_t x) { return (x + 15) / 16 * 16; }

}  // namespace grf

using namespace grf;

static int64_t transpose_counts_bytes(int64_t n_cols, int32_t n_steps) {
    return align16((n_cols * n_steps + 2) * (int64_t)sizeof(int32_t));
}

// workspace layout: [segment counts][scan workspace for them][census: 8 ints]
//                   [keys A][keys B][payload A][payload B][hist / offsets][scan workspace for hist]
extern "C" int64_t grf_transpose_workspace_bytes(int64_t n_cols, int32_t n_steps, int64_t nnz) {
    const RadixPlan rp = radix_plan(n_cols * n_steps, nnz);
    const int64_t hist_items = (int64_t)(1 << kRadixMaxBits) * rp.grid;
    int64_t b = transpose_counts_bytes(n_cols, n_steps) + align16(grf_scan_workspace_bytes(n_cols * n_steps)) + 32;
    b += 2 * align16(nnz * 4);
    b += (rp.passes >= 2 ? 2 : 1) * align16(nnz * 8);
    b += 2 * align16((hist_items + 1) * 4) + align16(grf_scan_workspace_bytes(hist_items));
    return b;
}

namespace grf {
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
}  // namespace grf

static inline int grid_rows(int64_t n_rows) {
    int64_t g = (n_rows + 7) / 8;
    const int64_t cap = (int64_t)kSmCount * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

extern "C" int grf_transpose_offsets(const int32_t *blk_ptr, const GrfEntry *entries, int64_t n_rows, int64_t n_cols,
                                     int32_t n_steps, const int32_t *col_counts, int32_t *tblk_ptr, void *workspace,
                                     int32_t census_threshold, int32_t *census_host, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, tblk_ptr);
    GRF_REQUIRE(n_rows >= 0 && n_cols >= 0 && n_steps >= 1, "grf_transpose_offsets: bad shape");
    GRF_REQUIRE(tblk_ptr && workspace, "grf_transpose_offsets: null buffer");
    GRF_REQUIRE(!census_host || census_threshold >= 1, "grf_transpose_offsets: bad census threshold");
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *tcnt = (int32_t *)workspace;
    char *scan_ws = (char *)workspace + transpose_counts_bytes(n_cols, n_steps);
    int32_t *census = (int32_t *)(scan_ws + align16(grf_scan_workspace_bytes(n_cols * n_steps)));
    if (col_counts) {
        tcnt = const_cast<int32_t *>(col_counts);  // the walker counted while it emitted the entries
    } else if (n_cols > 0) {
        GRF_CUDA_OK(cudaMemsetAsync(tcnt, 0, (size_t)n_cols * n_steps * sizeof(int32_t), st));
        if (n_rows > 0) {
            GRF_REQUIRE(blk_ptr, "grf_transpose_offsets: null blk_ptr");
            transpose_count_kernel<<<grid_rows(n_rows), 256, 0, st>>>(blk_ptr, entries, n_rows, n_steps, tcnt);
            GRF_CUDA_OK(cudaGetLastError());
        }
    }
    int rc = grf_scan_counts(tcnt, n_cols, n_steps, GRF_ORDER_ROW_MAJOR, tblk_ptr, 0, scan_ws, stream);
    if (rc != GRF_OK) return rc;
    if (census_host) {
        // row statistics of both sides while the sort that follows keeps the GPU busy: the host reads
        // them from pinned memory after an event recorded behind this call
        rc = grf_row_census(blk_ptr, n_rows, n_steps, census_threshold, census, stream);
        if (rc != GRF_OK) return rc;
        rc = grf_row_census(tblk_ptr, n_cols, n_steps, census_threshold, census + 3, stream);
        if (rc != GRF_OK) return rc;
        GRF_CUDA_OK(cudaMemcpyAsync(census_host, census, 6 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    }
    return GRF_OK;
}

extern "C" int grf_transpose_fill(const int32_t *blk_ptr, const GrfEntry *entries, int64_t n_rows, int64_t n_cols,
                                  int32_t n_steps, int64_t entry_lo, int64_t nnz, void *workspace,
                                  GrfEntry *tentries, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, workspace);
    GRF_REQUIRE(n_rows >= 0 && n_cols >= 0 && n_steps >= 1 && nnz >= 0 && entry_lo >= 0,
                "grf_transpose_fill: bad shape");
    GRF_REQUIRE(entry_lo + nnz < (1ll << 31), "grf_transpose_fill: entry range exceeds the 32-bit offsets");
    GRF_REQUIRE(n_cols * n_steps < (1ll << 32), "grf_transpose_fill: (column, length) keys exceed 32 bits");
    if (n_cols == 0 || n_rows == 0 || nnz == 0) return GRF_OK;
    GRF_REQUIRE(blk_ptr && entries && workspace && tentries, "grf_transpose_fill: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const RadixPlan rp = radix_plan(n_cols * n_steps, nnz);
    char *p = (char *)workspace + transpose_counts_bytes(n_cols, n_steps) +
              align16(grf_scan_workspace_bytes(n_cols * n_steps)) + 32;
    uint32_t *keys[2];
    keys[0] = (uint32_t *)p;
    p += align16(nnz * 4);
    keys[1] = (uint32_t *)p;
    p += align16(nnz * 4);
    int2 *pay[2];
    pay[0] = (int2 *)p;
    p += align16(nnz * 8);
    pay[1] = pay[0];
    if (rp.passes >= 2) {  // ping-pong; a single pass goes straight from pay[0] to the Phi^T entries
        pay[1] = (int2 *)p;
        p += align16(nnz * 8);
    }
    const int64_t hist_cap = (int64_t)(1 << kRadixMaxBits) * rp.grid;
    int32_t *hist = (int32_t *)p;
    p += align16((hist_cap + 1) * 4);
    int32_t *offs = (int32_t *)p;
    p += align16((hist_cap + 1) * 4);
    void *scan_ws = p;

    {
        const int max_smem = kRadixTile * (8 + 2 * 4 + 2) + (1 << kRadixMaxBits) * (8 + 2 * kRadixWarps);
        GRF_CUDA_OK(cudaFuncSetAttribute(radix_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        GRF_CUDA_OK(cudaFuncSetAttribute(radix_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    }
    transpose_keys_kernel<<<grid_rows(n_rows), 256, 0, st>>>(blk_ptr, (const int2 *)entries, n_rows, n_steps,
                                                             (int32_t)entry_lo, keys[0], pay[0]);
    GRF_CUDA_OK(cudaGetLastError());
    int shift = 0;
    int cur = 0;
    for (int pass = 0; pass < rp.passes; ++pass) {
        const int nbins = 1 << rp.bits[pass];
        const bool last = pass + 1 == rp.passes;
        radix_hist_kernel<<<rp.grid, kRadixThreads, nbins * sizeof(uint32_t), st>>>(keys[cur], nnz, shift, nbins,
                                                                                     rp.tiles_per_cta, hist);
        GRF_CUDA_OK(cudaGetLastError());
        const int rc = grf_scan_counts(hist, (int64_t)nbins * rp.grid, 1, GRF_ORDER_ROW_MAJOR, offs, 0, scan_ws, stream);
        if (rc != GRF_OK) return rc;
        const size_t smem = (size_t)kRadixTile * (8 + 2 * 4 + 2) + (size_t)nbins * 2 * sizeof(uint32_t) +
                            (size_t)nbins * kRadixWarps * sizeof(unsigned short);
        if (last) {
            radix_scatter_kernel<true><<<rp.grid, kRadixThreads, smem, st>>>(
                keys[cur], pay[cur], nnz, shift, nbins, rp.tiles_per_cta, offs, nullptr, (int2 *)tentries);
        } else {
            radix_scatter_kernel<false><<<rp.grid, kRadixThreads, smem, st>>>(
                keys[cur], pay[cur], nnz, shift, nbins, rp.tiles_per_cta, offs, keys[cur ^ 1], pay[cur ^ 1]);
        }
        GRF_CUDA_OK(cudaGetLastError());
        shift += rp.bits[pass];
        cur ^= 1;
    }
    return GRF_OK;
}
