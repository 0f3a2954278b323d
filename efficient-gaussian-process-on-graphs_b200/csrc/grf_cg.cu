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

constexpr int kCgThreads = 256;
constexpr int kCgMaxBlocks = kSmCount * 2;

struct CgShape {
    int cb;  // columns handled per block (power of two <= 32)
    int rb;  // rows per block iteration
};

__host__ __device__ inline CgShape cg_shape(int t) {
    CgShape s;
    s.cb = 1;
    while (s.cb < t && s.cb < 32) s.cb <<= 1;
    s.rb = kCgThreads / s.cb;
    return s;
}

// block-level reduction over the row lanes (ty) for every column lane (tx); result valid for ty == 0
__device__ __forceinline__ float reduce_rows(float v, int tx, int ty, int cb, int rb, float *sh) {
    sh[ty * cb + tx] = v;
    __syncthreads();
    for (int s = rb >> 1; s > 0; s >>= 1) {
        if (ty < s) sh[ty * cb + tx] += sh[(ty + s) * cb + tx];
        __syncthreads();
    }
    const float out = sh[tx];
    __syncthreads();
    return out;
}

// sum of the per-block partials of one column, in block order
__device__ __forceinline__ float sum_partials(const float *__restrict__ partial, int n_blocks, int ldp, int c) {
    float acc = 0.f;
    for (int b = 0; b < n_blocks; ++b) acc += partial[(int64_t)b * ldp + c];
    return acc;
}

__global__ void __launch_bounds__(kCgThreads) cg_dot_kernel(float *__restrict__ ad, int64_t ldad,
                                                            const float *__restrict__ d, int64_t ldd, float sigma2,
                                                            int64_t n, int32_t t, float *__restrict__ partial,
                                                            int32_t ldp) {
    __shared__ float sh[kCgThreads];
    const CgShape s = cg_shape(t);
    const int tx = threadIdx.x % s.cb, ty = threadIdx.x / s.cb;
    for (int c0 = 0; c0 < t; c0 += s.cb) {
        const int c = c0 + tx;
        float acc = 0.f;
        if (c < t) {
            for (int64_t i = (int64_t)blockIdx.x * s.rb + ty; i < n; i += (int64_t)gridDim.x * s.rb) {
                const float dv = d[i * ldd + c];
                const float a = fmaf(sigma2, dv, ad[i * ldad + c]);
                ad[i * ldad + c] = a;
                acc = fmaf(dv, a, acc);
            }
        }
        const float tot = reduce_rows(acc, tx, ty, s.cb, s.rb, sh);
        if (ty == 0 && c < t) partial[(int64_t)blockIdx.x * ldp + c] = tot;
    }
}

__global__ void __launch_bounds__(kCgThreads) cg_update_kernel(float *__restrict__ x, int64_t ldx,
                                                               float *__restrict__ r, int64_t ldr,
                                                               const float *__restrict__ d, int64_t ldd,
                                                               const float *__restrict__ ad, int64_t ldad,
                                                               const float *__restrict__ rs,
                                                               const float *__restrict__ dad_partial,
                                                               int32_t n_partial, int64_t n, int32_t t, float eps,
                                                               float *__restrict__ rr_partial, int32_t ldp) {
    __shared__ float sh[kCgThreads];
    const CgShape s = cg_shape(t);
    const int tx = threadIdx.x % s.cb, ty = threadIdx.x / s.cb;
    for (int c0 = 0; c0 < t; c0 += s.cb) {
        const int c = c0 + tx;
        float acc = 0.f;
        if (c < t) {
            const float dad = sum_partials(dad_partial, n_partial, ldp, c);
            // upstream linear_cg: a vanishing <d, Ad> means the column has converged -> alpha = 0
            const float alpha = dad > eps ? rs[c] / dad : 0.f;
            for (int64_t i = (int64_t)blockIdx.x * s.rb + ty; i < n; i += (int64_t)gridDim.x * s.rb) {
                x[i * ldx + c] = fmaf(alpha, d[i * ldd + c], x[i * ldx + c]);
                const float rv = fmaf(-alpha, ad[i * ldad + c], r[i * ldr + c]);
                r[i * ldr + c] = rv;
                acc = fmaf(rv, rv, acc);
            }
        }
        const float tot = reduce_rows(acc, tx, ty, s.cb, s.rb, sh);
        if (ty == 0 && c < t) rr_partial[(int64_t)blockIdx.x * ldp + c] = tot;
    }
}

__global__ void __launch_bounds__(kCgThreads) cg_direction_kernel(float *__restrict__ d, int64_t ldd,
                                                                  const float *__restrict__ r, int64_t ldr,
                                                                  float *__restrict__ rs,
                                                                  const float *__restrict__ rr_partial,
                                                                  int32_t n_partial, int64_t n, int32_t t, float eps,
                                                                  int32_t ldp, float *__restrict__ rs_out) {
    const CgShape s = cg_shape(t);
    const int tx = threadIdx.x % s.cb, ty = threadIdx.x / s.cb;
    for (int c0 = 0; c0 < t; c0 += s.cb) {
        const int c = c0 + tx;
        if (c >= t) continue;
        const float rs_new = sum_partials(rr_partial, n_partial, ldp, c);
        const float rs_old = rs[c];
        const float beta = rs_old > eps ? rs_new / rs_old : 0.f;
        for (int64_t i = (int64_t)blockIdx.x * s.rb + ty; i < n; i += (int64_t)gridDim.x * s.rb)
            d[i * ldd + c] = fmaf(beta, d[i * ldd + c], r[i * ldr + c]);
        if (blockIdx.x == 0 && ty == 0) rs_out[c] = rs_new;  // rs is read by every block: write the copy
    }
}

static int cg_blocks(int64_t n, int32_t t) {
    const CgShape s = cg_shape(t);
    int64_t g = (n + s.rb - 1) / s.rb;
    if (g > kCgMaxBlocks) g = kCgMaxBlocks;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace grf

using namespace grf;

extern "C" int32_t grf_cg_num_partials(int64_t n, int32_t t) { return cg_blocks(n, t); }

extern "C" int grf_cg_dot(float *ad, int64_t ldad, const float *d, int64_t ldd, float sigma2, int64_t n, int32_t t,
                          float *partial, void *stream) {
    GRF_REQUIRE(n >= 0 && t >= 1, "grf_cg_dot: bad shape");
    GRF_REQUIRE(ad && d && partial && ldad >= t && ldd >= t, "grf_cg_dot: bad buffers");
    cg_dot_kernel<<<cg_blocks(n, t), kCgThreads, 0, (cudaStream_t)stream>>>(ad, ldad, d, ldd, sigma2, n, t, partial,
                                                                             t);
    return check_cuda(cudaGetLastError(), "cg_dot_kernel launch");
}

extern "C" int grf_cg_update(float *x, int64_t ldx, float *r, int64_t ldr, const float *d, int64_t ldd,
                             const float *ad, int64_t ldad, const float *rs, const float *dad_partial, int64_t n,
                             int32_t t, float eps, float *rr_partial, void *stream) {
    GRF_REQUIRE(n >= 0 && t >= 1, "grf_cg_update: bad shape");
    GRF_REQUIRE(x && r && d && ad && rs && dad_partial && rr_partial, "grf_cg_update: null buffer");
    const int g = cg_blocks(n, t);
    cg_update_kernel<<<g, kCgThreads, 0, (cudaStream_t)stream>>>(x, ldx, r, ldr, d, ldd, ad, ldad, rs, dad_partial, g,
                                                                 n, t, eps, rr_partial, t);
    return check_cuda(cudaGetLastError(), "cg_update_kernel launch");
}

extern "C" int grf_cg_direction(float *d, int64_t ldd, const float *r, int64_t ldr, const float *rs,
                                const float *rr_partial, int64_t n, int32_t t, float eps, float *rs_out,
                                void *stream) {
    GRF_REQUIRE(n >= 0 && t >= 1, "grf_cg_direction: bad shape");
    GRF_REQUIRE(d && r && rs && rr_partial && rs_out && rs != rs_out, "grf_cg_direction: bad buffers");
    const int g = cg_blocks(n, t);
    cg_direction_kernel<<<g, kCgThreads, 0, (cudaStream_t)stream>>>(d, ldd, r, ldr, const_cast<float *>(rs),
                                                                    rr_partial, g, n, t, eps, t, rs_out);
    return check_cuda(cudaGetLastError(), "cg_direction_kernel launch");
}
