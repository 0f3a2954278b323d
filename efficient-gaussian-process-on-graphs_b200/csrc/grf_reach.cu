// Which row shards can reach which column: owner bitmasks propagated along the walk graph.
//
// A row-sharded Phi(Phi^T V) only has to exchange the columns that more than one shard touches
// (grf_b200/sharding.py).  Shard g touches column v only if some start node it owns reaches v in at
// most L - 1 hops, so a superset of the touched columns follows from the graph alone -- once per
// (graph, sharding, L), with no collective and no dependence on the random draws.  Bit g of
// mask[v] = "a start node of shard g reaches v"; one round ORs every node's mask into its
// neighbours'.  The reference has no counterpart (its fork pool, sparse_sampler.py:90-114, merges
// dictionaries on the host); this belongs to the multi-GPU row of the hot path.

#include "grf_common.cuh"

namespace grf {

__global__ void __launch_bounds__(256) reach_init_kernel(const int64_t *__restrict__ bounds, int32_t world,
                                                         int64_t n_nodes, unsigned long long *__restrict__ mask_a,
                                                         unsigned long long *__restrict__ mask_b) {
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < n_nodes;
         u += (int64_t)gridDim.x * blockDim.x) {
        int g = 0;
        while (g + 1 < world && u >= bounds[g + 1]) ++g;  // world <= 64: a short scan
        const unsigned long long m = 1ull << g;
        mask_a[u] = m;
        mask_b[u] = m;
    }
}

// one warp per node: OR the node's mask into its neighbours' (mask_out starts as a copy of mask_in)
__global__ void __launch_bounds__(256) reach_round_kernel(const int32_t *__restrict__ row_ptr,
                                                          const int32_t *__restrict__ col_idx, int64_t n_nodes,
                                                          const unsigned long long *__restrict__ mask_in,
                                                          unsigned long long *__restrict__ mask_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t u = warp0; u < n_nodes; u += nwarps) {
        const unsigned long long m = mask_in[u];
        const int32_t b = row_ptr[u], e = row_ptr[u + 1];
        for (int32_t i = b + lane; i < e; i += 32) {
            const int32_t v = col_idx[i];
            if ((mask_in[v] & m) != m) atomicOr(mask_out + v, m);
        }
    }
}

}  // namespace grf

extern "C" int grf_shard_reach(const GrfGraph *graph, const int64_t *bounds, int32_t world, int32_t hops,
                               unsigned long long *mask, unsigned long long *scratch, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, mask);
    using namespace grf;
    GRF_REQUIRE(graph && bounds, "grf_shard_reach: null graph/bounds");
    GRF_REQUIRE(world >= 1 && world <= 64, "grf_shard_reach: world must be in [1, 64]");
    GRF_REQUIRE(hops >= 0, "grf_shard_reach: hops must be >= 0");
    const int64_t n = graph->n_nodes;
    if (n == 0) return GRF_OK;
    GRF_REQUIRE(mask && scratch && graph->row_ptr, "grf_shard_reach: null buffer");
    GRF_REQUIRE(graph->nnz == 0 || graph->col_idx, "grf_shard_reach: null col_idx");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)kSmCount * 16) g = (int64_t)kSmCount * 16;
    // after an even number of rounds the result sits in `mask`: start in the buffer that makes it so
    unsigned long long *cur = (hops % 2 == 0) ? mask : scratch;
    unsigned long long *nxt = (hops % 2 == 0) ? scratch : mask;
    reach_init_kernel<<<(int)g, 256, 0, st>>>(bounds, world, n, cur, nxt);
    GRF_CUDA_OK(cudaGetLastError());
    int64_t gw = (n + 7) / 8;
    if (gw > (int64_t)kSmCount * 32) gw = (int64_t)kSmCount * 32;
    for (int32_t h = 0; h < hops; ++h) {
        if (h > 0)  // nxt must start as a copy of cur (round 0: both hold the initial masks)
            GRF_CUDA_OK(cudaMemcpyAsync(nxt, cur, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        reach_round_kernel<<<(int)gw, 256, 0, st>>>(graph->row_ptr, graph->col_idx, n, cur, nxt);
        GRF_CUDA_OK(cudaGetLastError());
        unsigned long long *t = cur;
        cur = nxt;
        nxt = t;
    }
    return GRF_OK;
}
