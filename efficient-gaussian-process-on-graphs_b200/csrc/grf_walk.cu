// GRF walker + per-start-node merge (K1 + the dict half of K2 in SURVEY.md 2a).
//
// Replaces the reference's CPython walk loop
//   efficient_graph_gp_sparse/random_walk_samplers_sparse/sparse_sampler.py:36-54
//   efficient_graph_gp/random_walk_samplers/sampler.py:40-59, :163-184
// and its defaultdict accumulation / merge (:42, :110-114).
//
// One *group* (a warp, or a whole CTA when W is large) owns one start node at a
// time.  Phase 1: every lane runs walks w = lane, lane+G, ... of that node:
// CSR row_ptr pair -> halting draw -> neighbour draw -> (col, val) gather, the
// load update deg*w/(1-p) applied in registers in float64, and the visit
// (node, load) of every length >= 1 parked in shared memory.  Phase 2: per
// length, the <= W visits are bitonic-sorted by (node, walk) -- in registers
// with warp shuffles when a warp owns the node, in shared memory when a CTA
// does -- and, per distinct node, the loads are added *in walk order* into a double
// that starts at 0.0 -- exactly the order in which the reference's
// defaultdict(float) accumulates them, so the sums are bit-identical when both
// sides see the same draws.  The merged (column-sorted) per-length segments go
// to the row's staging region; grf_compact_* turns staging into CSR.
//
// Draw sources: Philox4x32-10 keyed by the seed with counter (walk id, step / 2)
// -- one block serves two steps: words (0, 1) halt / pick the even step, words
// (2, 3) the odd one; independent of how start nodes are sharded over GPUs --
// or a replayed trace of the reference's own PCG64 draws.
//
// Roofline: per executed walk-step the algorithmic traffic is row_ptr pair 8 B
// + col 4 B + val 8 B + staging write 12 B/(merged entry); three dependent
// random 32-B sectors per step => latency / sector bound, not bandwidth bound.

#include "grf_common.cuh"

namespace grf {

struct WalkParams {
    const int32_t *row_ptr;
    const int32_t *col_idx;
    const double *val;
    const double *scaled_val;
    int64_t start_lo;
    int64_t n_local;
    int32_t W, L, Wp, wbits;
    double p_halt, one_minus_p;
    unsigned long long halt_thr;
    uint32_t k0, k1;
    int32_t draw_mode, load_mode;
    const double *trace_u;
    const int32_t *trace_k;
    int64_t stride;
    int32_t *stage_col;
    double *stage_sum;
    int32_t *row_cnt;
    unsigned long long *visits;
    uint32_t group_bytes, loads_bytes, nodes_bytes, sorted_bytes;
    int32_t *col_counts;  // optional [n_nodes][L]: entries per (column, length), for the Phi^T offsets
};

template <bool kBlock>
__device__ __forceinline__ void group_sync() {
    if (kBlock)
        __syncthreads();
    else
        __syncwarp();
}

// exclusive scan of one int per thread over the group; `total` = group sum
template <bool kBlock>
__device__ __forceinline__ int group_excl_scan(int v, int *scratch, int &total) {
    const int lane = threadIdx.x & 31;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += n;
    }
    if (!kBlock) {
        total = __shfl_sync(0xffffffffu, incl, 31);
        return incl - v;
    }
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    int before = 0, all = 0;
    for (int i = 0; i < nwarps; ++i) {
        const int s = scratch[i];
        if (i < warp) before += s;
        all += s;
    }
    __syncthreads();
    total = all;
    return before + incl - v;
}

// KPL = keys per lane of the warp-per-node variant (sort width 32*KPL >= W); 0 for the
// CTA-per-node variant, which sorts in shared memory.
#ifndef GRF_WALK_MINBLOCKS
#define GRF_WALK_MINBLOCKS 5  // register cap for the warp-per-node variant (walk-steps/s: grid config 2 | R-MAT 2^20 nodes)
// 1 or 4: 85 registers 97 G | 37 G;  5: 74 registers 102 G | 45 G;  6 or 7: 58-60 registers 102 G | 30 G
#endif
// kFast: Philox draws, cumulative load, per-edge factors precomputed (the production setting) --
// strips the per-step mode tests; the generic instantiation serves replay / ablation / sequential.
template <bool kBlock, typename KeyT, int KPL, bool kFast>
__global__ void __launch_bounds__(kBlock ? 256 : 128, kBlock ? 1 : GRF_WALK_MINBLOCKS) walk_merge_kernel(const WalkParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int GS = kBlock ? (int)blockDim.x : 32;
    const int tg = kBlock ? (int)threadIdx.x : (int)(threadIdx.x & 31);
    const int groups_per_cta = kBlock ? 1 : (int)(blockDim.x >> 5);
    const int group_in_cta = kBlock ? 0 : (int)(threadIdx.x >> 5);

    unsigned char *base = smem_raw + (size_t)group_in_cta * p.group_bytes;
    double *loads = reinterpret_cast<double *>(base);                                  // [(L-1)][W]
    int32_t *nodes = reinterpret_cast<int32_t *>(base + p.loads_bytes);                // [(L-1)][W]
    double *sorted_loads = reinterpret_cast<double *>(base + p.loads_bytes + p.nodes_bytes);  // [32*KPL], warp variant
    KeyT *keys = reinterpret_cast<KeyT *>(base + p.loads_bytes + p.nodes_bytes + p.sorted_bytes);  // [Wp | 32*KPL]
    __shared__ int scan_scratch[32];

    const int W = p.W, L = p.L, Wp = p.Wp, wbits = p.wbits;
    const KeyT KEY_MAX = ~(KeyT)0;
    const KeyT wmask = ((KeyT)1 << wbits) - 1;
    const int chunk = (Wp + GS - 1) / GS;
    unsigned long long my_visits = 0;

    for (int64_t row = (int64_t)blockIdx.x * groups_per_cta + group_in_cta; row < p.n_local;
         row += (int64_t)gridDim.x * groups_per_cta) {
        const int64_t start = p.start_lo + row;

        // ---------------- phase 1: the walks -------------------------------
        // a walk that stops at length s leaves no visit at lengths > s: mark every slot empty first
        // (16-byte stores; per-walk tail loops cost 6 % of the kernel's instructions under ncu)
        {
            int4 *n4 = reinterpret_cast<int4 *>(nodes);
            const int n_vec = (int)(p.nodes_bytes >> 4);
            for (int i = tg; i < n_vec; i += GS) n4[i] = make_int4(-1, -1, -1, -1);
        }
        group_sync<kBlock>();
        for (int w = tg; w < W; w += GS) {
            const unsigned long long walk_id = (unsigned long long)start * (unsigned long long)W + (unsigned)w;
            int32_t cur = (int32_t)start;
            double load = 1.0;
            ++my_visits;  // the length-0 visit (start, 1.0)
            int step = 0;
            // one transition: false = the walk ends before a visit at length step + 1
            auto advance = [&](const uint32_t xh, const uint32_t xk) -> bool {
                const int32_t rs = __ldg(p.row_ptr + cur);
                const int32_t re = __ldg(p.row_ptr + cur + 1);
                const int32_t deg = re - rs;
                if (deg == 0) return false;  // dead end: stop without drawing (sparse_sampler.py:47)
                int32_t k;
                if (!kFast && p.draw_mode == GRF_DRAW_REPLAY) {
                    const unsigned long long ti = walk_id * (unsigned)L + (unsigned)step;
                    if (__ldg(p.trace_u + ti) < p.p_halt) return false;
                    k = __ldg(p.trace_k + ti);
                } else {
                    if ((unsigned long long)xh < p.halt_thr) return false;
                    k = (int32_t)__umulhi(xk, (uint32_t)deg);
                }
                const int64_t e = (int64_t)rs + k;
                const int32_t nxt = __ldg(p.col_idx + e);
                // load *= degree * weight / (1 - p_halt), evaluated left to right in float64 with
                // no contraction (sparse_sampler.py:54).  (deg * w) / (1 - p) depends on the edge
                // only, so grf_edge_scale may have computed it once per edge (same roundings).
                if (!kFast && p.load_mode == GRF_LOAD_ABLATION) {
                    load = __ldg(p.val + e);
                } else {
                    const double scaled = (kFast || p.scaled_val)
                                              ? __ldg(p.scaled_val + e)
                                              : __ddiv_rn(__dmul_rn((double)deg, __ldg(p.val + e)), p.one_minus_p);
                    load = (kFast || p.load_mode == GRF_LOAD_CUMULATIVE) ? __dmul_rn(load, scaled) : scaled;
                }
                cur = nxt;
                nodes[step * W + w] = cur;  // the visit at length step+1
                loads[step * W + w] = load;
                ++my_visits;
                ++step;
                return true;
            };
            if (!kFast && p.draw_mode == GRF_DRAW_REPLAY) {
                while (step < L - 1 && advance(0u, 0u)) {
                }
            } else {
                // one Philox block serves two consecutive steps: words (0,1) the even one, (2,3) the odd one
                while (step < L - 1) {
                    uint32_t x[4];
                    philox4x32_10((uint32_t)walk_id, (uint32_t)(walk_id >> 32), (uint32_t)(step >> 1), 0u, p.k0, p.k1,
                                  x);
                    if (!advance(x[0], x[1])) break;
                    if (step >= L - 1 || !advance(x[2], x[3])) break;
                }
            }
        }

        int32_t *out_col = p.stage_col + row * p.stride;
        double *out_sum = p.stage_sum + row * p.stride;
        if (tg == 0) {
            out_col[0] = (int32_t)start;  // M_0 = I: W visits of load 1.0, summed exactly
            out_sum[0] = (double)W;
            p.row_cnt[row * L] = 1;
            if (p.col_counts) atomicAdd(p.col_counts + start * L, 1);
        }
        int off = 1;
        group_sync<kBlock>();

        // ---------------- phase 2: merge the visits of each length ----------
        for (int si = 0; si < L - 1; ++si) {
            if constexpr (!kBlock) {
                // warp-per-node: sort in registers, park the sorted keys in shared memory
                // for the run walks below
                const int lane = tg;
                KeyT key[KPL];
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    const int w = r * 32 + lane;
                    KeyT kk = KEY_MAX;
                    if (w < W) {
                        const int32_t nd = nodes[si * W + w];
                        if (nd >= 0) kk = ((KeyT)(uint32_t)nd << wbits) | (KeyT)w;
                    }
                    key[r] = kk;
                }
                warp_bitonic_sort<KeyT, KPL>(key, lane);
                // park the loads in sorted order; a run (= one output entry) is then a contiguous
                // stretch of that stream.  Heads write their node and where their run starts; one lane
                // per run then adds the run's loads in walk order (the reference's summation order).
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    const KeyT kk = key[r];
                    if (kk != KEY_MAX) sorted_loads[lane * KPL + r] = loads[si * W + (int)(kk & wmask)];
                }
                const KeyT prev_last = __shfl_up_sync(0xffffffffu, key[KPL - 1], 1);
                bool head[KPL];
                int counts = 0;  // heads in the low half, valid elements in the high half: one scan for both
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    const KeyT kk = key[r];
                    const KeyT pv = r == 0 ? prev_last : key[r > 0 ? r - 1 : 0];
                    const bool first = (lane == 0 && r == 0);
                    head[r] = (kk != KEY_MAX) && (first || (pv >> wbits) != (kk >> wbits));
                    counts += (int)head[r] + ((int)(kk != KEY_MAX) << 16);
                }
                int totals;
                int rank = group_excl_scan<false>(counts, scan_scratch, totals) & 0xffff;
                const int total = totals & 0xffff, n_valid = totals >> 16;
                int32_t *run_start = reinterpret_cast<int32_t *>(keys);  // [total + 1]
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    if (head[r]) {
                        run_start[rank] = lane * KPL + r;
                        const int32_t node = (int32_t)(key[r] >> wbits);
                        out_col[off + rank] = node;
                        if (p.col_counts) atomicAdd(p.col_counts + (int64_t)node * L + si + 1, 1);
                        ++rank;
                    }
                }
                if (lane == 0) run_start[total] = n_valid;
                __syncwarp();
                for (int j = lane; j < total; j += 32) {
                    const int qb = run_start[j], qe = run_start[j + 1];
                    double sum = 0.0;
                    for (int q = qb; q < qe; ++q) sum = __dadd_rn(sum, sorted_loads[q]);
                    out_sum[off + j] = sum;
                }
                if (lane == 0) p.row_cnt[row * L + si + 1] = total;
                off += total;
                __syncwarp();
            } else {
                for (int i = tg; i < Wp; i += GS) {
                    KeyT key = KEY_MAX;
                    if (i < W) {
                        const int32_t nd = nodes[si * W + i];
                        if (nd >= 0) key = ((KeyT)(uint32_t)nd << wbits) | (KeyT)i;
                    }
                    keys[i] = key;
                }
                __syncthreads();
                for (uint32_t k = 2; k <= (uint32_t)Wp; k <<= 1) {
                    for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                        for (uint32_t c = tg; c < (uint32_t)Wp / 2; c += GS) {
                            const uint32_t idx = ((c & ~(j - 1)) << 1) | (c & (j - 1));
                            const uint32_t ixj = idx | j;
                            const KeyT a = keys[idx], b = keys[ixj];
                            const bool up = (idx & k) == 0;
                            if ((a > b) == up) {
                                keys[idx] = b;
                                keys[ixj] = a;
                            }
                        }
                        __syncthreads();
                    }
                }
                // distinct nodes = run heads of the sorted keys; thread tg owns a
                // contiguous chunk so that ranks follow column order
                const int lo = tg * chunk;
                const int hi = min(Wp, lo + chunk);
                int heads = 0;
                for (int i = lo; i < hi; ++i) {
                    const KeyT key = keys[i];
                    if (key == KEY_MAX) break;
                    heads += (i == 0) || ((keys[i - 1] >> wbits) != (key >> wbits));
                }
                int total;
                int rank = group_excl_scan<true>(heads, scan_scratch, total);
                for (int i = lo; i < hi; ++i) {
                    const KeyT key = keys[i];
                    if (key == KEY_MAX) break;
                    const KeyT node = key >> wbits;
                    if ((i == 0) || ((keys[i - 1] >> wbits) != node)) {
                        double sum = 0.0;
                        for (int q = i; q < Wp; ++q) {
                            const KeyT kq = keys[q];
                            if ((kq >> wbits) != node) break;
                            sum = __dadd_rn(sum, loads[si * W + (int)(kq & wmask)]);
                        }
                        out_col[off + rank] = (int32_t)node;
                        out_sum[off + rank] = sum;
                        if (p.col_counts) atomicAdd(p.col_counts + (int64_t)node * L + si + 1, 1);
                        ++rank;
                    }
                }
                if (tg == 0) p.row_cnt[row * L + si + 1] = total;
                off += total;
                __syncthreads();
            }
        }
    }

    if (p.visits != nullptr) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) my_visits += __shfl_xor_sync(0xffffffffu, my_visits, d);
        if ((threadIdx.x & 31) == 0 && my_visits) atomicAdd(p.visits, my_visits);
    }
}

template <bool kBlock, typename KeyT, int KPL, bool kFast = false>
static int launch_walk(const WalkParams &p, size_t smem, int threads, int grid, cudaStream_t stream) {
    auto kern = walk_merge_kernel<kBlock, KeyT, KPL, kFast>;
    GRF_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, threads, smem, stream>>>(p);
    return check_cuda(cudaGetLastError(), "walk_merge_kernel launch");
}

}  // namespace grf

extern "C" int64_t grf_walk_stage_stride(int32_t walks_per_node, int32_t max_walk_length) {
    if (walks_per_node < 1 || max_walk_length < 1) return 0;
    return 1 + (int64_t)(max_walk_length - 1) * (int64_t)walks_per_node;
}

extern "C" int grf_walk(const GrfGraph *graph, const GrfWalkCfg *cfg, int64_t stage_stride, int32_t *stage_col,
                        double *stage_sum, int32_t *row_cnt, unsigned long long *visits_out, void *stream) {
    GRF_ON_STREAM_DEVICE(stream);
    using namespace grf;
    GRF_REQUIRE(graph && cfg, "grf_walk: null graph/cfg");
    GRF_REQUIRE(graph->n_nodes >= 0 && graph->n_nodes < (1ll << 31), "grf_walk: n_nodes %lld out of int32 range",
                (long long)graph->n_nodes);
    GRF_REQUIRE(graph->nnz >= 0 && graph->nnz < (1ll << 31), "grf_walk: nnz %lld out of int32 range",
                (long long)graph->nnz);
    GRF_REQUIRE(cfg->walks_per_node >= 1, "grf_walk: walks_per_node must be >= 1");
    GRF_REQUIRE(cfg->max_walk_length >= 1, "grf_walk: max_walk_length must be >= 1");
    GRF_REQUIRE(cfg->p_halt >= 0.0 && cfg->p_halt <= 1.0, "grf_walk: p_halt must be in [0, 1]");
    GRF_REQUIRE(cfg->start_lo >= 0 && cfg->start_lo <= cfg->start_hi && cfg->start_hi <= graph->n_nodes,
                "grf_walk: start range [%lld, %lld) outside [0, %lld)", (long long)cfg->start_lo,
                (long long)cfg->start_hi, (long long)graph->n_nodes);
    GRF_REQUIRE(cfg->draw_mode == GRF_DRAW_PHILOX || cfg->draw_mode == GRF_DRAW_REPLAY, "grf_walk: bad draw_mode");
    GRF_REQUIRE(cfg->load_mode >= GRF_LOAD_CUMULATIVE && cfg->load_mode <= GRF_LOAD_ABLATION,
                "grf_walk: bad load_mode");
    GRF_REQUIRE(cfg->draw_mode != GRF_DRAW_REPLAY || (cfg->trace_u && cfg->trace_k),
                "grf_walk: replay mode needs trace_u and trace_k");
    GRF_REQUIRE(stage_stride >= grf_walk_stage_stride(cfg->walks_per_node, cfg->max_walk_length),
                "grf_walk: stage_stride too small");
    const int64_t n_local = cfg->start_hi - cfg->start_lo;
    if (n_local == 0) return GRF_OK;
    GRF_REQUIRE(graph->row_ptr && stage_col && stage_sum && row_cnt, "grf_walk: null buffer");
    GRF_REQUIRE(graph->nnz == 0 || (graph->col_idx && graph->val), "grf_walk: null edge arrays");

    WalkParams p;
    p.row_ptr = graph->row_ptr;
    p.col_idx = graph->col_idx;
    p.val = graph->val;
    p.scaled_val = cfg->scaled_val;
    p.start_lo = cfg->start_lo;
    p.n_local = n_local;
    p.W = cfg->walks_per_node;
    p.L = cfg->max_walk_length;
    p.Wp = (int32_t)next_pow2((uint32_t)p.W);
    p.wbits = bit_width((uint64_t)p.Wp - 1);
    p.p_halt = cfg->p_halt;
    p.one_minus_p = 1.0 - cfg->p_halt;
    double thr = floor(cfg->p_halt * 4294967296.0);
    thr = thr < 0.0 ? 0.0 : (thr > 4294967296.0 ? 4294967296.0 : thr);
    p.halt_thr = (unsigned long long)thr;
    p.k0 = (uint32_t)cfg->seed;
    p.k1 = (uint32_t)(cfg->seed >> 32);
    p.draw_mode = cfg->draw_mode;
    p.load_mode = cfg->load_mode;
    p.trace_u = cfg->trace_u;
    p.trace_k = cfg->trace_k;
    p.stride = stage_stride;
    p.stage_col = stage_col;
    p.stage_sum = stage_sum;
    p.row_cnt = row_cnt;
    p.visits = visits_out;
    p.col_counts = cfg->col_counts;

    const int node_bits = bit_width((uint64_t)(graph->n_nodes > 0 ? graph->n_nodes - 1 : 0));
    const bool key32 = node_bits + p.wbits <= 31;
    const size_t key_size = key32 ? 4 : 8;
    const size_t steps = (size_t)(p.L - 1);
    p.loads_bytes = (uint32_t)((steps * p.W * 8 + 15) & ~(size_t)15);  // keeps `nodes` 16-byte aligned
    p.nodes_bytes = (uint32_t)((steps * p.W * 4 + 15) & ~(size_t)15);  // whole int4s: the kernel clears it with 16-byte stores
    const bool warp_variant = p.W <= 256;
    const int kpl = p.Wp <= 32 ? 1 : p.Wp / 32;  // warp variant sorts 32*kpl keys
    const size_t n_keys = warp_variant ? (size_t)32 * kpl : (size_t)p.Wp;
    p.sorted_bytes = warp_variant ? (uint32_t)(n_keys * 8) : 0u;
    // + 4: the warp variant reuses the key area as run_start[runs + 1] (runs <= n_keys)
    const size_t gb = ((size_t)p.loads_bytes + p.nodes_bytes + p.sorted_bytes + n_keys * key_size + 4 + 15) & ~(size_t)15;
    p.group_bytes = (uint32_t)gb;
    const size_t kMaxSmem = 227 * 1024 - 256;
    GRF_REQUIRE((uint64_t)p.W * (uint64_t)p.L < (1ull << 31), "grf_walk: W*L too large");

    cudaStream_t st = (cudaStream_t)stream;
    if (warp_variant) {
        const int warps = 4;
        const int64_t want = (n_local + warps - 1) / warps;
        const int grid = (int)(want < (int64_t)kSmCount * 64 ? want : (int64_t)kSmCount * 64);
        const size_t smem = warps * gb;
        const int threads = warps * 32;
        const bool fast = p.draw_mode == GRF_DRAW_PHILOX && p.load_mode == GRF_LOAD_CUMULATIVE && p.scaled_val;
#define GRF_WALK_CASE(K)                                                                              \
    case K:                                                                                           \
        if (fast)                                                                                     \
            return key32 ? launch_walk<false, uint32_t, K, true>(p, smem, threads, grid, st)         \
                         : launch_walk<false, unsigned long long, K, true>(p, smem, threads, grid, st); \
        return key32 ? launch_walk<false, uint32_t, K>(p, smem, threads, grid, st)                   \
                     : launch_walk<false, unsigned long long, K>(p, smem, threads, grid, st)
        switch (kpl) {
            GRF_WALK_CASE(1);
            GRF_WALK_CASE(2);
            GRF_WALK_CASE(4);
            default:
                GRF_WALK_CASE(8);
        }
#undef GRF_WALK_CASE
    }
    if (gb > kMaxSmem)
        return fail(GRF_ERR_UNSUPPORTED,
                    "grf_walk: W=%d, L=%d needs %zu B of shared memory per start node (max %zu)", p.W, p.L, gb,
                    kMaxSmem);
    const int64_t want = n_local;
    const int grid = (int)(want < (int64_t)kSmCount * 32 ? want : (int64_t)kSmCount * 32);
    return key32 ? launch_walk<true, uint32_t, 0>(p, gb, 256, grid, st)
                 : launch_walk<true, unsigned long long, 0>(p, gb, 256, grid, st);
}

namespace grf {
__global__ void __launch_bounds__(256) edge_scale_kernel(const int32_t *__restrict__ row_ptr,
                                                         const double *__restrict__ val, int64_t n_nodes,
                                                         double one_minus_p, double *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n_nodes; r += nwarps) {
        const int32_t b = row_ptr[r], e = row_ptr[r + 1];
        const double deg = (double)(e - b);
        for (int32_t i = b + lane; i < e; i += 32) out[i] = __ddiv_rn(__dmul_rn(deg, val[i]), one_minus_p);
    }
}
}  // namespace grf

extern "C" int grf_edge_scale(const GrfGraph *graph, double p_halt, double *scaled_val, void *stream) {
    GRF_ON_STREAM_DEVICE(stream);
    using namespace grf;
    GRF_REQUIRE(graph, "grf_edge_scale: null graph");
    GRF_REQUIRE(p_halt >= 0.0 && p_halt <= 1.0, "grf_edge_scale: p_halt must be in [0, 1]");
    if (graph->n_nodes == 0 || graph->nnz == 0) return GRF_OK;
    GRF_REQUIRE(graph->row_ptr && graph->val && scaled_val, "grf_edge_scale: null buffer");
    int64_t g = (graph->n_nodes + 7) / 8;
    if (g > (int64_t)kSmCount * 32) g = (int64_t)kSmCount * 32;
    edge_scale_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(graph->row_ptr, graph->val, graph->n_nodes,
                                                               1.0 - p_halt, scaled_val);
    return check_cuda(cudaGetLastError(), "edge_scale_kernel launch");
}
