// Fused vector kernels of batched conjugate gradients on (K + sigma^2 I) X = B.
//
// The reference calls upstream linear_cg (models/sparse_grf_model.py:43); every iteration
// there is one kernel matvec plus ~12 small elementwise / reduction launches on n x t
// temporaries.  Here an iteration is the two spmm launches plus three fused launches:
//   grf_cg_dot        Ad = Kd + sigma^2 d (in place),  per-block partials of <d, Ad>
//   grf_cg_update     alpha = rs / <d, Ad>;  x += alpha d;  r -= alpha Ad;  partials of <r, r>
//   grf_cg_direction  rs_new = <r, r>;  d = r + (rs_new / rs) d;  rs = rs_new
// Dot products are reduced in a fixed order (per-block partials, summed block by block),
// so a solve is deterministic.  All HBM-bound streaming passes over n x t floats.

#include "grf_common.cuh"

namespace grf {

// Layout: the n x t operands are contiguous (ld == t), so they are streamed as flat arrays.
// A thread owns VEC consecutive elements per step and strides by a multiple of t / VEC
// "column groups", so it always sees the same VEC columns and can keep their partial dot
// products in registers: fully coalesced 16-byte accesses, 4 steps unrolled (ncu on the first
// version -- one scalar load in flight per thread, 296 serial partial reads -- ran at ~0.3 TB/s).
constexpr int kCgThreads = 256;
constexpr int kCgBlocks = kSmCount * 4;
constexpr int kCgUnroll = 4;

template <int VEC>
struct Pack;
template <>
struct Pack<4> {
    float v[4];
    __device__ __forceinline__ void load(const float *p) {
        const float4 q = *reinterpret_cast<const float4 *>(p);
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    }
    __device__ __forceinline__ void store(float *p) const {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Pack<1> {
    float v[1];
    __device__ __forceinline__ void load(const float *p) { v[0] = *p; }
    __device__ __forceinline__ void store(float *p) const { *p = v[0]; }
};

// threads of a block: tid = j * period + cg, j < J; sums v[] over j for every column group cg
template <int VEC>
__device__ __forceinline__ void block_sum_by_group(float (&v)[VEC], int period, int J, float *sh) {
    const int tid = threadIdx.x;
    if (tid < period * J) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) sh[tid * VEC + q] = v[q];
    }
    __syncthreads();
    int span = 1;
    while (span < J) span <<= 1;
    const int j = tid / period;
    for (int s = span >> 1; s > 0; s >>= 1) {
        if (tid < period * J && j < s && j + s < J) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) sh[tid * VEC + q] += sh[(tid + s * period) * VEC + q];
        }
        __syncthreads();
    }
    const int cg = tid % period;
#pragma unroll
    for (int q = 0; q < VEC; ++q) v[q] = sh[cg * VEC + q];
    __syncthreads();
}

// total over the producer's per-block partials for this thread's columns (cooperative, fixed order)
template <int VEC>
__device__ __forceinline__ void total_of_partials(const float *__restrict__ partial, int n_partial, int t, int period,
                                                  int J, float *sh, float (&out)[VEC]) {
    const int tid = threadIdx.x;
    const int cg = tid % period, j = tid / period;
#pragma unroll
    for (int q = 0; q < VEC; ++q) out[q] = 0.f;
    if (tid < period * J) {
        for (int b = j; b < n_partial; b += J) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) out[q] += partial[(int64_t)b * t + cg * VEC + q];
        }
    }
    block_sum_by_group<VEC>(out, period, J, sh);
}

template <int VEC>
__global__ void __launch_bounds__(kCgThreads) cg_dot_kernel(float *__restrict__ ad, const float *__restrict__ d,
                                                            float sigma2, int64_t n_elem, int32_t t,
                                                            float *__restrict__ partial) {
    __shared__ float sh[kCgThreads * VEC];
    const int period = t / VEC, J = kCgThreads / period;
    const int tid = threadIdx.x;
    float acc[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) acc[q] = 0.f;
    if (tid < period * J) {
        const int64_t stride = (int64_t)gridDim.x * period * J * VEC;
        for (int64_t e0 = ((int64_t)blockIdx.x * period * J + tid) * VEC; e0 < n_elem; e0 += stride * kCgUnroll) {
            Pack<VEC> dv[kCgUnroll], av[kCgUnroll];
#pragma unroll
            for (int u = 0; u < kCgUnroll; ++u) {
                const int64_t e = e0 + u * stride;
                if (e < n_elem) {
                    dv[u].load(d + e);
                    av[u].load(ad + e);
                }
            }
#pragma unroll
            for (int u = 0; u < kCgUnroll; ++u) {
                const int64_t e = e0 + u * stride;
                if (e < n_elem) {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        av[u].v[q] = fmaf(sigma2, dv[u].v[q], av[u].v[q]);
                        acc[q] = fmaf(dv[u].v[q], av[u].v[q], acc[q]);
                    }
                    av[u].store(ad + e);
                }
            }
        }
    }
    block_sum_by_group<VEC>(acc, period, J, sh);
    if (tid < period) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) partial[(int64_t)blockIdx.x * t + tid * VEC + q] = acc[q];
    }
}

template <int VEC>
__global__ void __launch_bounds__(kCgThreads) cg_update_kernel(float *__restrict__ x, float *__restrict__ r,
                                                               const float *__restrict__ d,
                                                               const float *__restrict__ ad,
                                                               const float *__restrict__ rs,
                                                               const float *__restrict__ dad_partial,
                                                               int32_t n_partial, int64_t n_elem, int32_t t,
                                                               float eps, float *__restrict__ rr_partial) {
    __shared__ float sh[kCgThreads * VEC];
    const int period = t / VEC, J = kCgThreads / period;
    const int tid = threadIdx.x;
    const int cg = tid % period;
    float alpha[VEC], acc[VEC];
    total_of_partials<VEC>(dad_partial, n_partial, t, period, J, sh, alpha);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        // upstream linear_cg: a vanishing <d, Ad> means the column has converged -> alpha = 0
        alpha[q] = alpha[q] > eps ? rs[cg * VEC + q] / alpha[q] : 0.f;
        acc[q] = 0.f;
    }
    if (tid < period * J) {
        const int64_t stride = (int64_t)gridDim.x * period * J * VEC;
        for (int64_t e0 = ((int64_t)blockIdx.x * period * J + tid) * VEC; e0 < n_elem; e0 += stride * kCgUnroll) {
            Pack<VEC> dv[kCgUnroll], av[kCgUnroll], xv[kCgUnroll], rv[kCgUnroll];
#pragma unroll
            for (int u = 0; u < kCgUnroll; ++u) {
                const int64_t e = e0 + u * stride;
                if (e < n_elem) {
                    dv[u].load(d + e);
                    av[u].load(ad + e);
                    xv[u].load(x + e);
                    rv[u].load(r + e);
                }
            }
#pragma unroll
            for (int u = 0; u < kCgUnroll; ++u) {
                const int64_t e = e0 + u * stride;
                if (e < n_elem) {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        xv[u].v[q] = fmaf(alpha[q], dv[u].v[q], xv[u].v[q]);
                        rv[u].v[q] = fmaf(-alpha[q], av[u].v[q], rv[u].v[q]);
                        acc[q] = fmaf(rv[u].v[q], rv[u].v[q], acc[q]);
                    }
                    xv[u].store(x + e);
                    rv[u].store(r + e);
                }
            }
        }
    }
    block_sum_by_group<VEC>(acc, period, J, sh);
    if (tid < period) {
#pragma unroll
        for (int q = 0; q < VEC; ++q) rr_partial[(int64_t)blockIdx.x * t + tid * VEC + q] = acc[q];
    }
}

template <int VEC>
__global__ void __launch_bounds__(kCgThreads) cg_direction_kernel(float *__restrict__ d, const float *__restrict__ r,
                                                                  const float *__restrict__ rs,
                                                                  const float *__restrict__ rr_partial,
                                                                  int32_t n_partial, int64_t n_elem, int32_t t,
                                                                  float eps, float *__restrict__ rs_out) {
    __shared__ float sh[kCgThreads * VEC];
    const int period = t / VEC, J = kCgThreads / period;
    const int tid = threadIdx.x;
    const int cg = tid % period;
    float beta[VEC];
    total_of_partials<VEC>(rr_partial, n_partial, t, period, J, sh, beta);
#pragma unroll
    for (int q = 0; q < VEC; ++q) {
        const float rs_new = beta[q];
        const float rs_old = rs[cg * VEC + q];
        if (blockIdx.x == 0 && tid < period) rs_out[cg * VEC + q] = rs_new;
        beta[q] = rs_old > eps ? rs_new / rs_old : 0.f;
    }
    if (tid < period * J) {
        const int64_t stride = (int64_t)gridDim.x * period * J * VEC;
        for (int64_t e0 = ((int64_t)blockIdx.x * period * J + tid) * VEC; e0 < n_elem; e0 += stride * kCgUnroll) {
            Pack<VEC> dv[kCgUnroll], rv[kCgUnroll];
#pragma unroll
            for (int u = 0; u < kCgUnroll; ++u) {
                const int64_t e = e0 + u * stride;
                if (e < n_elem) {
                    dv[u].load(d + e);
                    rv[u].load(r + e);
                }
            }
#pragma unroll
            for (int u = 0; u < kCgUnroll; ++u) {
                const int64_t e = e0 + u * stride;
                if (e < n_elem) {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) dv[u].v[q] = fmaf(beta[q], dv[u].v[q], rv[u].v[q]);
                    dv[u].store(d + e);
                }
            }
        }
    }
}

static int cg_blocks(int64_t n, int32_t t) {
    const int64_t per_block = (int64_t)kCgThreads * 4 * kCgUnroll;
    int64_t g = (n * t + per_block - 1) / per_block;
    if (g > kCgBlocks) g = kCgBlocks;
    if (g < 1) g = 1;
    return (int)g;
}

static inline bool al16(const void *p) { return ((uintptr_t)p & 15u) == 0; }

}  // namespace grf

using namespace grf;

extern "C" int32_t grf_cg_num_partials(int64_t n, int32_t t) { return cg_blocks(n, t); }

#define GRF_CG_CHECK(name)                                                                              \
    GRF_REQUIRE(n >= 0 && t >= 1 && t <= kCgThreads, name ": bad shape (1 <= t <= 256)");                 \
    const int64_t n_elem = n * t;                                                                       \
    const int g = cg_blocks(n, t);                                                                      \
    cudaStream_t st = (cudaStream_t)stream

extern "C" int grf_cg_dot(float *ad, int64_t ldad, const float *d, int64_t ldd, float sigma2, int64_t n, int32_t t,
                          float *partial, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, ad);
    GRF_CG_CHECK("grf_cg_dot");
    GRF_REQUIRE(ad && d && partial && ldad == t && ldd == t, "grf_cg_dot: operands must be contiguous n x t");
    if (t % 4 == 0 && al16(ad) && al16(d))
        cg_dot_kernel<4><<<g, kCgThreads, 0, st>>>(ad, d, sigma2, n_elem, t, partial);
    else
        cg_dot_kernel<1><<<g, kCgThreads, 0, st>>>(ad, d, sigma2, n_elem, t, partial);
    return check_cuda(cudaGetLastError(), "cg_dot_kernel launch");
}

extern "C" int grf_cg_update(float *x, int64_t ldx, float *r, int64_t ldr, const float *d, int64_t ldd,
                             const float *ad, int64_t ldad, const float *rs, const float *dad_partial, int64_t n,
                             int32_t t, float eps, float *rr_partial, void *stream) {
    GRF_ON_STREAM_DEVICE(stream, x);
    GRF_CG_CHECK("grf_cg_update");
    GRF_REQUIRE(x && r && d && ad && rs && dad_partial && rr_partial, "grf_cg_update: null buffer");
    GRF_REQUIRE(ldx == t && ldr == t && ldd == t && ldad == t, "grf_cg_update: operands must be contiguous n x t");
    if (t % 4 == 0 && al16(x) && al16(r) && al16(d) && al16(ad))
        cg_update_kernel<4><<<g, kCgThreads, 0, st>>>(x, r, d, ad, rs, dad_partial, g, n_elem, t, eps, rr_partial);
    else
        cg_update_kernel<1><<<g, kCgThreads, 0, st>>>(x, r, d, ad, rs, dad_partial, g, n_elem, t, eps, rr_partial);
    return check_cuda(cudaGetLastError(), "cg_update_kernel launch");
}

extern "C" int grf_cg_direction(float *d, int64_t ldd, const float *r, int64_t ldr, const float *rs,
                                const float *rr_partial, int64_t n, int32_t t, float eps, float *rs_out,
                                void *stream) {
    GRF_ON_STREAM_DEVICE(stream, d);
    GRF_CG_CHECK("grf_cg_direction");
    GRF_REQUIRE(d && r && rs && rr_partial && rs_out && rs != rs_out, "grf_cg_direction: bad buffers");
    GRF_REQUIRE(ldd == t && ldr == t, "grf_cg_direction: operands must be contiguous n x t");
    if (t % 4 == 0 && al16(d) && al16(r))
        cg_direction_kernel<4><<<g, kCgThreads, 0, st>>>(d, r, rs, rr_partial, g, n_elem, t, eps, rs_out);
    else
        cg_direction_kernel<1><<<g, kCgThreads, 0, st>>>(d, r, rs, rr_partial, g, n_elem, t, eps, rs_out);
    return check_cuda(cudaGetLastError(), "cg_direction_kernel launch");
}
