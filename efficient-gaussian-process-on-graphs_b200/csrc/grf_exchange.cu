// The one exchange of the GRF path: U = sum over the GPUs of the partials U_g = Phi_g[x2]^T V_g
// (SURVEY.md 8e), as ONE kernel per rank over NVLink peer memory.
//
// Reference: none -- its fork pool merges dictionaries on the host (sparse_sampler.py:90-114) and its
// matvec runs on one device.  Round 1 did  index_select -> ncclAllReduce -> index_copy  (three launches
// plus the collective's latency).  Here every rank's U lives in a buffer that all ranks have mapped
// (torch symmetric memory: cudaMalloc + IPC handles; plumbing), and a rank
//   1. tells its peers that its partial is complete and waits until theirs are (one flag per CTA and
//      peer, system-scope release / acquire),
//   2. owns 1/G of the rows: loads that slice from all G buffers (G - 1 of them over NVLink), adds the
//      G partials in rank order -- every rank computes every element in the same order, so all copies
//      of U end up bit-identical, which NCCL does not promise -- and stores the sum back into all G
//      buffers (reduce-scatter and all-gather of a two-shot all-reduce, fused per 16-byte vector),
//   3. signals and waits once more, so that the second half of the product (the next kernel on the
//      stream) reads a complete U.
// Bytes over NVLink per rank: (G-1)/G * |U| in and the same out; at config 4 on 8 GPUs 235 MB each way,
// 0.3 ms at the measured 770 GB/s per direction, against a ring/tree all-reduce's 2 (G-1)/G |U| / busbw.
// The flags only ever grow (epoch), so nothing is reset between products.

#include "grf_common.cuh"

namespace grf {

constexpr int kExMaxWorld = 16;
constexpr int kExThreads = 512;
constexpr int kExCtas = kSmCount;  // one CTA per SM; flags are sized for this many

struct ExchangeArgs {
    float *u[kExMaxWorld];          // every rank's U (peer-mapped), same layout
    uint32_t *flags[kExMaxWorld];   // every rank's flag block: [2 phases][kExCtas][world]
    int32_t world, rank;
    int64_t n_vec;                  // float4s in U
    uint32_t epoch;
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}
__device__ __forceinline__ void st_peer(float4 *p, const float4 &v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

// all CTAs with the same blockIdx on all ranks meet here (phase 0 or 1 of this epoch)
__device__ __forceinline__ void cross_rank_barrier(const ExchangeArgs &a, int phase) {
    __syncthreads();  // this CTA's stores precede the fence below
    const int g = threadIdx.x;
    if (g < a.world) {
        const int64_t slot = ((int64_t)phase * kExCtas + blockIdx.x) * a.world;
        __threadfence_system();
        st_release_sys(a.flags[g] + slot + a.rank, a.epoch);
        const uint32_t *mine = a.flags[a.rank] + slot + g;
        while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0) {
        }
    }
    __syncthreads();
}

template <int WORLD>
__global__ void __launch_bounds__(kExThreads, 1) exchange_sum_kernel(const ExchangeArgs a) {
    cross_rank_barrier(a, 0);  // every rank's partial U_g is complete and visible
    const int64_t lo = a.n_vec * a.rank / a.world, hi = a.n_vec * (a.rank + 1) / a.world;
    const int64_t stride = (int64_t)gridDim.x * kExThreads;
    for (int64_t i = lo + (int64_t)blockIdx.x * kExThreads + threadIdx.x; i < hi; i += stride) {
        float4 part[WORLD];
#pragma unroll
        for (int g = 0; g < WORLD; ++g) part[g] = ld_peer(reinterpret_cast<const float4 *>(a.u[g]) + i);
        float4 s = part[0];
#pragma unroll
        for (int g = 1; g < WORLD; ++g) {  // rank order, on every rank: bit-identical sums everywhere
            s.x += part[g].x;
            s.y += part[g].y;
            s.z += part[g].z;
            s.w += part[g].w;
        }
#pragma unroll
        for (int g = 0; g < WORLD; ++g) st_peer(reinterpret_cast<float4 *>(a.u[g]) + i, s);
    }
    cross_rank_barrier(a, 1);  // every rank's slice has landed in this rank's U
}

// The same exchange through the NVSwitch (NVLS): `mc` is the MULTICAST address of U (every rank's copy behind one
// pointer).  multimem.ld_reduce has the switch read all copies of a vector and return their sum -- 1/G of U
// enters this GPU instead of (G-1)/G -- and multimem.st writes the sum to every copy with one store.  Each slice
// is reduced by one rank and multicast, so all copies still come out bit-identical.
__device__ __forceinline__ float4 mc_ld_reduce(const float4 *mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(mc)
                 : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float4 *mc, const float4 &v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(kExThreads, 1) exchange_nvls_kernel(const ExchangeArgs a, float *mc_u) {
    cross_rank_barrier(a, 0);  // every rank's partial U_g is complete and visible
    const int64_t lo = a.n_vec * a.rank / a.world, hi = a.n_vec * (a.rank + 1) / a.world;
    const int64_t stride = (int64_t)gridDim.x * kExThreads;
    float4 *mc = reinterpret_cast<float4 *>(mc_u);
    for (int64_t i0 = lo + (int64_t)blockIdx.x * kExThreads + threadIdx.x; i0 < hi; i0 += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k * stride < hi) v[k] = mc_ld_reduce(mc + i0 + k * stride);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (i0 + k * stride < hi) mc_st(mc + i0 + k * stride, v[k]);
    }
    cross_rank_barrier(a, 1);  // every rank's slice has landed in this rank's U
}

}  // namespace grf

using namespace grf;

extern "C" int grf_exchange_sum_nvls(float *multicast_u, uint32_t *const *peer_flags, int32_t world, int32_t rank,
                                     int64_t n_floats, uint32_t epoch, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, (peer_flags && rank >= 0 && rank < world ? peer_flags[rank] : nullptr));
    GRF_REQUIRE(multicast_u && peer_flags, "grf_exchange_sum_nvls: null pointer");
    GRF_REQUIRE(world >= 1 && world <= kExMaxWorld && rank >= 0 && rank < world,
                "grf_exchange_sum_nvls: world must be 1..%d and rank inside it", kExMaxWorld);
    GRF_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "grf_exchange_sum_nvls: U must hold a multiple of 4 floats");
    GRF_REQUIRE(((uintptr_t)multicast_u & 15u) == 0, "grf_exchange_sum_nvls: U must be 16-byte aligned");
    GRF_REQUIRE(epoch != 0, "grf_exchange_sum_nvls: epoch 0 is the flags' initial value; start at 1");
    if (world == 1 || n_floats == 0) return GRF_OK;
    ExchangeArgs a;
    for (int g = 0; g < world; ++g) {
        GRF_REQUIRE(peer_flags[g], "grf_exchange_sum_nvls: null peer flag block");
        a.u[g] = nullptr;
        a.flags[g] = peer_flags[g];
    }
    a.world = world;
    a.rank = rank;
    a.n_vec = n_floats / 4;
    a.epoch = epoch;
    exchange_nvls_kernel<<<kExCtas, kExThreads, 0, (cudaStream_t)stream>>>(a, multicast_u);
    return check_cuda(cudaGetLastError(), "exchange_nvls_kernel launch");
}

extern "C" int64_t grf_exchange_flag_bytes(int32_t world) {
    return world < 1 ? 0 : (int64_t)2 * kExCtas * world * (int64_t)sizeof(uint32_t);
}

extern "C" int grf_exchange_sum(float *const *peer_u, uint32_t *const *peer_flags, int32_t world, int32_t rank,
                                int64_t n_floats, uint32_t epoch, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, (peer_u && rank >= 0 && rank < world ? peer_u[rank] : nullptr));
    GRF_REQUIRE(peer_u && peer_flags, "grf_exchange_sum: null pointer tables");
    GRF_REQUIRE(world >= 1 && world <= kExMaxWorld && rank >= 0 && rank < world,
                "grf_exchange_sum: world must be 1..%d and rank inside it", kExMaxWorld);
    GRF_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, "grf_exchange_sum: U must hold a multiple of 4 floats");
    GRF_REQUIRE(epoch != 0, "grf_exchange_sum: epoch 0 is the flags' initial value; start at 1");
    if (world == 1 || n_floats == 0) return GRF_OK;
    ExchangeArgs a;
    for (int g = 0; g < world; ++g) {
        GRF_REQUIRE(peer_u[g] && peer_flags[g], "grf_exchange_sum: null peer buffer");
        GRF_REQUIRE(((uintptr_t)peer_u[g] & 15u) == 0, "grf_exchange_sum: U must be 16-byte aligned");
        a.u[g] = peer_u[g];
        a.flags[g] = peer_flags[g];
    }
    a.world = world;
    a.rank = rank;
    a.n_vec = n_floats / 4;
    a.epoch = epoch;
    cudaStream_t st = (cudaStream_t)stream;
    switch (world) {
        case 2: exchange_sum_kernel<2><<<kExCtas, kExThreads, 0, st>>>(a); break;
        case 3: exchange_sum_kernel<3><<<kExCtas, kExThreads, 0, st>>>(a); break;
        case 4: exchange_sum_kernel<4><<<kExCtas, kExThreads, 0, st>>>(a); break;
        case 5: exchange_sum_kernel<5><<<kExCtas, kExThreads, 0, st>>>(a); break;
        case 6: exchange_sum_kernel<6><<<kExCtas, kExThreads, 0, st>>>(a); break;
        case 7: exchange_sum_kernel<7><<<kExCtas, kExThreads, 0, st>>>(a); break;
        case 8: exchange_sum_kernel<8><<<kExCtas, kExThreads, 0, st>>>(a); break;
        default:
            return fail(GRF_ERR_UNSUPPORTED, "grf_exchange_sum: built for 2..8 GPUs of one NVSwitch box (got %d)", world);
    }
    return check_cuda(cudaGetLastError(), "exchange_sum_kernel launch");
}
