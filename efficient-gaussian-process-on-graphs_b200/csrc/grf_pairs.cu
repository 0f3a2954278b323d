// Line-pair layout of the merged Phi_f (and Phi_f^T) for t = 16 right-hand sides.
//
// With 16 float32 columns a row of X is 64 bytes -- half a 128-byte line -- and the matvec is bound by L1
// data-stage wavefronts: one per gathered row, whatever its width (ncu on config 2: 77 % of the wavefront peak at
// 1.2 wavefronts per entry slot, DRAM at 14 %).  On lattices, rings and any node numbering with locality the
// columns of a Phi row come in runs (a 4-hop diamond of a 316-wide grid is 9 runs of 1..9 consecutive ids), so
// columns 2p and 2p+1 are usually both present: stored as ONE entry {p, w_even, w_odd} they cost ONE wavefront
// for the whole line X[2p : 2p+2, :].  Four lanes own a row of Phi and take 32 bytes of the line each (256-bit
// loads): lanes 0-1 the halves of X[2p, :] with weight w_even, lanes 2-3 those of X[2p+1, :] with w_odd; the two
// row halves are added once per row.
//
// Replaces, like grf_matvec.cu, the 2L cuSPARSE SpMMs of utils_sparse/sparse_lo.py:16-18 -- here for the CG
// steady state (fixed modulator) on graphs where the pairing pays; the caller (engine.MatvecPlan) checks the
// pair count against the union count and falls back to grf_phi_matvec otherwise (power-law graphs: 0.3 % of
// the entries have their partner).
//
// Built from the union layout (grf_union.cu): pptr / pidx once per Phi, the entries once per modulator.

#include "grf_common.cuh"

namespace grf {

// heads: union entries whose line (col >> 1) differs from their predecessor's in the row
__global__ void __launch_bounds__(256) pair_count_kernel(const int32_t *__restrict__ uptr,
                                                         const int2 *__restrict__ uent, int64_t n_rows,
                                                         int32_t *__restrict__ pcnt) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t b = uptr[r], e = uptr[r + 1];
        int cnt = 0;
        for (int32_t i = b + lane; i < e; i += 32)
            cnt += (i == b) || ((((uint32_t)uent[i - 1].x & kColMask) >> 1) != (((uint32_t)uent[i].x & kColMask) >> 1));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
        if (lane == 0) pcnt[r] = cnt;
    }
}

// pidx[u] = pair slot of union entry u
__global__ void __launch_bounds__(256) pair_index_kernel(const int32_t *__restrict__ uptr,
                                                         const int2 *__restrict__ uent, int64_t n_rows,
                                                         const int32_t *__restrict__ pptr,
                                                         int32_t *__restrict__ pidx) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_rows; r += nwarps) {
        const int32_t b = uptr[r], e = uptr[r + 1];
        int32_t out = pptr[r];
        for (int32_t base = b; base < e; base += 32) {
            const int32_t i = base + lane;
            bool head = false;
            if (i < e)
                head = (i == b) ||
                       ((((uint32_t)uent[i - 1].x & kColMask) >> 1) != (((uint32_t)uent[i].x & kColMask) >> 1));
            const unsigned heads = __ballot_sync(0xffffffffu, head);
            // slot = heads at or before this lane, minus one
            if (i < e) pidx[i] = out + __popc(heads & ((2u << lane) - 1u)) - 1;
            out += __popc(heads);
        }
    }
}

// pent zeroed by the caller; every union entry writes its line index and its half
__global__ void __launch_bounds__(256) pair_scatter_kernel(const int2 *__restrict__ uent,
                                                           const int32_t *__restrict__ pidx, int64_t n_union,
                                                           int32_t *__restrict__ pent) {
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_union;
         u += (int64_t)gridDim.x * blockDim.x) {
        const int2 en = uent[u];
        const uint32_t col = (uint32_t)en.x & kColMask;
        int32_t *dst = pent + (int64_t)pidx[u] * 4;
        dst[0] = (int32_t)(col >> 1);
        dst[1 + (col & 1u)] = en.y;
    }
}

// Y[k, 0:16] = sum over the pair entries of row k of  w_even * X[2p, :] + w_odd * X[2p + 1, :]
// X: [n_x][16] float32 (leading dimension exactly 16: a pair of rows is one 128-byte line).
//
// Four lanes per row, eight rows per warp -- the shape of spmm_blocks_kernel -- but every lane takes 32 bytes of
// the line with ONE 256-bit load (sm_100: LDG.E.256): lanes 0-1 the two halves of X[2p, :] with the even weight,
// lanes 2-3 those of X[2p + 1, :] with the odd one.  A gather instruction of the warp therefore touches 8 lines
// for 8 pair entries: one L1 wavefront per PAIR, where the 64-byte rows of the plain layout cost one per entry.
// Lane `sub` holds, per round and register q, entry base + 2q + (sub & 1) -- its line offset and the weight of ITS
// half (sub >> 1) -- so slot m of the round reaches every lane with two width-4 shuffles whose source lane is
// (half << 1) | (m & 1).  Slots beyond the end of a row issue no gather (predicated: a padded gather would cost
// the wavefront the layout is there to save); the rows of a warp run the rounds of the longest of them.
struct F8 {
    float v[8];
};
// 32 bytes of X, or zeros when there is nothing to gather (off < 0).  The predicate lives inside the asm block:
// written as `if (off >= 0) x = load(...)` nvcc re-uses one register group for all the gathers of a batch and
// serialises load -> FMA -> load (the three first versions of this kernel all took 42 us per half, whatever their
// shape; the same trap is described in spmm_row).
__device__ __forceinline__ F8 ldg256_if(const float *base, int32_t off) {
    F8 r;
#pragma unroll
    for (int c = 0; c < 8; ++c) r.v[c] = 0.f;
    asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %9, 0;\n\t"
        "@p ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t}"
        : "+f"(r.v[0]), "+f"(r.v[1]), "+f"(r.v[2]), "+f"(r.v[3]), "+f"(r.v[4]), "+f"(r.v[5]), "+f"(r.v[6]), "+f"(r.v[7])
        : "l"(base + (off < 0 ? 0 : off)), "r"(off));
    return r;
}

constexpr int kPairRegs = 8;                    // entries per lane and round
constexpr int kPairRound = 2 * kPairRegs;       // entry slots per row and round
constexpr int kPairBatch = 8;                   // gathers in flight per lane (32 bytes each)
__global__ void __launch_bounds__(256, 2) spmm_pairs_kernel(const int32_t *__restrict__ pptr,
                                                            const int4 *__restrict__ pent,
                                                            const int32_t *__restrict__ row_ids, int64_t n_tasks,
                                                            int64_t row_lo, int64_t n_rows,
                                                            const float *__restrict__ X, int64_t n_x,
                                                            float *__restrict__ Y, int64_t ldy) {
    const int lane = threadIdx.x & 31;
    const int sub = lane & 3, g = lane >> 2, half = sub >> 1;
    const int c0 = (sub & 1) * 8;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t n_iters = (n_tasks + 7) / 8;
    // bounds of iteration `it` for this lane's group
    auto bounds = [&](int64_t it, int32_t &b, int32_t &e, bool &mine) {
        b = e = 0;
        mine = false;
        const int64_t k = it * 8 + g;
        if (it < n_iters && k < n_tasks) {
            const int64_t row = row_ids ? (int64_t)__ldg(row_ids + k) - row_lo : k;
            if (row >= 0 && row < n_rows) {
                mine = true;
                b = __ldg(pptr + row);
                e = __ldg(pptr + row + 1);
            }
        }
    };
    // this lane's share of the round that starts at entry `base` of [b, e): offsets of the X row of this lane's
    // half (in floats; -1 = no entry, or a row of X that does not exist) and the weights of that half
    auto fetch = [&](int32_t b, int32_t e, int32_t base, int32_t (&off)[kPairRegs], float (&wt)[kPairRegs]) {
#pragma unroll
        for (int q = 0; q < kPairRegs; ++q) {
            const int32_t idx = b + base + q * 2 + (sub & 1);
            off[q] = -1;
            wt[q] = 0.f;
            if (idx < e) {
                const int4 en = __ldg(pent + idx);
                const int64_t xrow = 2 * (int64_t)en.x + half;
                if (xrow < n_x) off[q] = (int32_t)(xrow * 16);
                wt[q] = __int_as_float(half ? en.z : en.y);
            }
        }
    };
    int32_t b, e, nb, ne;
    bool mine, nmine;
    int64_t it = warp0;
    bounds(it, b, e, mine);
    int32_t off[kPairRegs];
    float wt[kPairRegs];
    fetch(b, e, 0, off, wt);
    for (; it < n_iters; it += nwarps) {
        bounds(it + nwarps, nb, ne, nmine);  // the next iteration's rows, while this one's gathers run
        const int32_t len_max = __reduce_max_sync(0xffffffffu, e - b);
        float acc[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = 0.f;
        int32_t base = 0;
        do {
#pragma unroll
            for (int m0 = 0; m0 < kPairRound; m0 += kPairBatch) {
                F8 x[kPairBatch];
                float w[kPairBatch];
#pragma unroll
                for (int m = 0; m < kPairBatch; ++m) {
                    const int q = (m0 + m) >> 1, j = (m0 + m) & 1;
                    const int32_t o = __shfl_sync(0xffffffffu, off[q], (half << 1) | j, 4);
                    w[m] = __shfl_sync(0xffffffffu, wt[q], (half << 1) | j, 4);
                    x[m] = ldg256_if(X + c0, o);
                }
                if (m0 + kPairBatch == kPairRound) {
                    // every slot of this round has been handed out: the registers take the next round (or the
                    // first round of the next iteration) while the last gathers are in flight
                    if (base + kPairRound < len_max)
                        fetch(b, e, base + kPairRound, off, wt);
                    else
                        fetch(nb, ne, 0, off, wt);
                }
#pragma unroll
                for (int m = 0; m < kPairBatch; ++m) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) acc[c] = fmaf(w[m], x[m].v[c], acc[c]);
                }
            }
            base += kPairRound;
        } while (base < len_max);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
        if (mine && half == 0) {
            float4 *dst = reinterpret_cast<float4 *>(Y + (it * 8 + g) * ldy + c0);
            dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
            dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
        b = nb;
        e = ne;
        mine = nmine;
    }
}

static inline int pairs_warp_grid(int64_t n_rows) {
    int64_t g = (n_rows + 7) / 8;
    const int64_t cap = (int64_t)kSmCount * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace grf

using namespace grf;

extern "C" int grf_pairs_count(const int32_t *uptr, const GrfEntry *uent, int64_t n_rows, int32_t *pcnt,
                               void *stream) {
    GRF_ON_STREAM_DEVICE(stream, pcnt);
    GRF_REQUIRE(n_rows >= 0, "grf_pairs_count: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(uptr && pcnt, "grf_pairs_count: null buffer");
    pair_count_kernel<<<pairs_warp_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(uptr, (const int2 *)uent, n_rows,
                                                                                 pcnt);
    return check_cuda(cudaGetLastError(), "pair_count_kernel launch");
}

extern "C" int grf_pairs_index(const int32_t *uptr, const GrfEntry *uent, int64_t n_rows, const int32_t *pptr,
                               int32_t *pidx, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, pptr);
    GRF_REQUIRE(n_rows >= 0, "grf_pairs_index: bad shape");
    if (n_rows == 0) return GRF_OK;
    GRF_REQUIRE(uptr && pptr && pidx, "grf_pairs_index: null buffer");
    pair_index_kernel<<<pairs_warp_grid(n_rows), 256, 0, (cudaStream_t)stream>>>(uptr, (const int2 *)uent, n_rows,
                                                                                 pptr, pidx);
    return check_cuda(cudaGetLastError(), "pair_index_kernel launch");
}

extern "C" int grf_pairs_fill(const GrfEntry *uent, const int32_t *pidx, int64_t n_union, int64_t n_pairs,
                              int32_t *pent, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, pent);
    GRF_REQUIRE(n_union >= 0 && n_pairs >= 0, "grf_pairs_fill: bad shape");
    if (n_pairs == 0) return GRF_OK;
    GRF_REQUIRE(uent && pidx && pent, "grf_pairs_fill: null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    GRF_CUDA_OK(cudaMemsetAsync(pent, 0, (size_t)n_pairs * 16, st));
    int64_t g = (n_union + 255) / 256;
    if (g > (int64_t)kSmCount * 32) g = (int64_t)kSmCount * 32;
    if (g < 1) g = 1;
    pair_scatter_kernel<<<(int)g, 256, 0, st>>>((const int2 *)uent, pidx, n_union, pent);
    return check_cuda(cudaGetLastError(), "pair_scatter_kernel launch");
}

extern "C" int grf_pairs_spmm(const int32_t *pptr, const int32_t *pent, const int32_t *row_ids, int64_t n_tasks,
                              int64_t row_lo, int64_t n_rows, const float *x, int64_t n_x, float *y, int64_t ldy,
                              void *stream) {
    GRF_ON_STREAM_DEVICE(stream, y);
    GRF_REQUIRE(n_tasks >= 0 && n_rows >= 0 && n_x >= 0, "grf_pairs_spmm: bad shape");
    GRF_REQUIRE(row_ids || n_tasks == n_rows, "grf_pairs_spmm: n_tasks must equal n_rows when row_ids is NULL");
    if (n_tasks == 0) return GRF_OK;
    GRF_REQUIRE(pptr && x && y && ldy >= 16, "grf_pairs_spmm: null buffer or ldy < 16");
    GRF_REQUIRE(((uintptr_t)x & 127) == 0 && ((uintptr_t)y & 15) == 0 && ldy % 4 == 0,
                "grf_pairs_spmm: X must be 128-byte aligned (a pair of rows is one line), Y rows 16-byte aligned");
    GRF_REQUIRE(n_x * 16 < (1ll << 31), "grf_pairs_spmm: X beyond 2^27 rows (32-bit element offsets)");
    int64_t g = (n_tasks + 63) / 64;  // 8 warps x 8 rows per CTA
    const int64_t cap = (int64_t)kSmCount * 2;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    spmm_pairs_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(pptr, (const int4 *)pent, row_ids, n_tasks, row_lo,
                                                                n_rows, x, n_x, y, ldy);
    return check_cuda(cudaGetLastError(), "spmm_pairs_kernel launch");
}

// Both halves of out = Phi_f[x1] (Phi_f^T V) on pair entries in one call (one host round trip per CG product):
// which & 1: U = Phi_f^T V (all n_cols rows of the transposed side);  which & 2: out = Phi_f[x1] U.
extern "C" int grf_pairs_matvec(const int32_t *tpptr, const int32_t *tpent, const int32_t *pptr, const int32_t *pent,
                                const int32_t *x1, int64_t n1, int64_t row_lo, int64_t n_rows, int64_t n_cols,
                                const float *v, float *u, float *out, int64_t ldo, int32_t which, void *stream) {
    if (which & 1) {
        const int rc = grf_pairs_spmm(tpptr, tpent, nullptr, n_cols, 0, n_cols, v, n_rows, u, 16, stream);
        if (rc != GRF_OK) return rc;
    }
    if (which & 2) return grf_pairs_spmm(pptr, pent, x1, n1, row_lo, n_rows, u, n_cols, out, ldo, stream);
    return GRF_OK;
}
