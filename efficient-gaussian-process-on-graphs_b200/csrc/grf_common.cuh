// Shared helpers of the grf_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "grf_b200.h"

namespace grf {

// thread-local error text returned by grf_last_error()
char *error_buffer();
int fail(int code, const char *fmt, ...);
int check_cuda(cudaError_t err, const char *what);

#define GRF_CUDA_OK(expr)                                         \
    do {                                                          \
        int _rc = ::grf::check_cuda((expr), #expr);               \
        if (_rc != GRF_OK) return _rc;                            \
    } while (0)

#define GRF_REQUIRE(cond, ...)                                    \
    do {                                                          \
        if (!(cond)) return ::grf::fail(GRF_ERR_INVALID, __VA_ARGS__); \
    } while (0)

// Every entry point runs on the device that owns `stream`, whatever device is current in the calling
// thread (a caller that holds cuda:1 tensors while cuda:0 is current would otherwise launch on the
// wrong device: "invalid resource handle").  The previous device is restored on return.
struct StreamDeviceGuard {
    int prev = -1;
    bool switched = false;
    int rc = GRF_OK;
    // `anchor`: any device buffer of the call.  torch's current stream is usually the null stream, which names no
    // device: then the device is the one that owns the anchor (cudaPointerGetAttributes, ~1 us).
    StreamDeviceGuard(void *stream, const void *anchor) {
        int dev = -1;
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();  // no usable device: argument checks still run, the first CUDA call reports it
            return;
        }
        const bool null_stream =
            stream == nullptr || stream == (void *)cudaStreamLegacy || stream == (void *)cudaStreamPerThread;
        if (!null_stream) {
            // a capturing stream is left alone: the capture was begun with its device current, and stream
            // queries other than this one may invalidate the capture
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing((cudaStream_t)stream, &cap) != cudaSuccess) {
                cudaGetLastError();
                return;
            }
            if (cap != cudaStreamCaptureStatusNone) return;
            if (cudaStreamGetDevice((cudaStream_t)stream, &dev) != cudaSuccess) {
                cudaGetLastError();
                return;
            }
        } else if (anchor != nullptr) {
            cudaPointerAttributes attr;
            if (cudaPointerGetAttributes(&attr, anchor) != cudaSuccess) {
                cudaGetLastError();
                return;
            }
            if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) return;
            dev = attr.device;
        } else {
            return;
        }
        if (dev >= 0 && dev != prev) {
            rc = check_cuda(cudaSetDevice(dev), "cudaSetDevice(device of the call's buffers)");
            switched = rc == GRF_OK;
        }
    }
    ~StreamDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

#define GRF_ON_STREAM_DEVICE(stream, anchor)                 \
    ::grf::StreamDeviceGuard _grf_guard(stream, anchor);     \
    if (_grf_guard.rc != GRF_OK) return _grf_guard.rc

// Phi / Phi^T entries carry their walk length in the top bits of `col`
// (GRF_ENTRY_STEP_SHIFT in grf_b200.h): the matvec then needs one pointer pair
// per row instead of one per (row, length).
constexpr int kStepShift = GRF_ENTRY_STEP_SHIFT;
constexpr uint32_t kColMask = (1u << GRF_ENTRY_STEP_SHIFT) - 1u;
constexpr int kMaxSteps = 1 << (32 - GRF_ENTRY_STEP_SHIFT);

__host__ __device__ __forceinline__ int32_t pack_col(int32_t col, int step) {
    return (int32_t)(((uint32_t)step << kStepShift) | (uint32_t)col);
}
__host__ __device__ __forceinline__ int32_t entry_col(int32_t packed) { return (int32_t)((uint32_t)packed & kColMask); }
__host__ __device__ __forceinline__ int entry_step(int32_t packed) { return (int)((uint32_t)packed >> kStepShift); }
constexpr int kSmCount = 148;  // B200

__host__ __device__ inline uint32_t next_pow2(uint32_t x) {
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

__host__ __device__ inline int bit_width(uint64_t x) {
    int b = 0;
    while (x) {
        ++b;
        x >>= 1;
    }
    return b;
}

// Philox4x32-10 (Salmon et al., SC'11).  Same counter/key convention as
// oracle/grf_oracle.c: ctr = (walk_lo, walk_hi, step, 0), key = (seed_lo, seed_hi).
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0);
        const uint32_t lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2);
        const uint32_t lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0;
        const uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

// Bitonic sort of 32*KPL keys held KPL per lane (element e = lane*KPL + r): exchanges at
// distance < KPL stay in registers, larger ones are one shuffle per key.  ~4x fewer issue
// slots than the shared-memory network it replaces (ncu: the sort was 45 % of all
// instructions of the kernel).
// All-ascending form: the first stage of every merge pairs e with its mirror e ^ (k - 1), the
// remaining ones e with e ^ j, and every comparator puts the smaller key at the lower index, so
// the in-register exchanges are a plain min/max pair (no per-lane direction select).
template <typename KeyT, int KPL>
__device__ __forceinline__ void warp_bitonic_sort(KeyT (&key)[KPL], const int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * KPL; k <<= 1) {
        if (k <= KPL) {
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                const int pr = r ^ (k - 1);
                if (r < pr) {
                    const KeyT a = key[r], b = key[pr];
                    key[r] = a < b ? a : b;
                    key[pr] = a < b ? b : a;
                }
            }
        } else {
            const bool lower = (lane & (k / 2 / KPL)) == 0;
            KeyT other[KPL];
#pragma unroll
            for (int r = 0; r < KPL; ++r) other[r] = __shfl_xor_sync(0xffffffffu, key[KPL - 1 - r], k / KPL - 1);
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                const KeyT mine = key[r];
                const KeyT mn = mine < other[r] ? mine : other[r];
                const KeyT mx = mine < other[r] ? other[r] : mine;
                key[r] = lower ? mn : mx;
            }
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j >= KPL) {
                const int lj = j / KPL;
                const bool lower = (lane & lj) == 0;
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    const KeyT mine = key[r];
                    const KeyT other = __shfl_xor_sync(0xffffffffu, mine, lj);
                    const KeyT mn = mine < other ? mine : other;
                    const KeyT mx = mine < other ? other : mine;
                    key[r] = lower ? mn : mx;
                }
            } else {
#pragma unroll
                for (int r = 0; r < KPL; ++r) {
                    if ((r & j) == 0) {
                        const KeyT a = key[r], b = key[r | j];
                        key[r] = a < b ? a : b;
                        key[r | j] = a < b ? b : a;
                    }
                }
            }
        }
    }
}


// All-ascending bitonic network on C keys in registers (C a power of two, indices compile-time).
template <int C>
__device__ __forceinline__ void sort_network_u32(uint32_t (&k)[C]) {
#pragma unroll
    for (int kk = 2; kk <= C; kk <<= 1) {
        const int half = kk >> 1;
#pragma unroll
        for (int c = 0; c < C / 2; ++c) {
            const int blk = c / half, off = c % half;
            const int lo = blk * kk + off, hi = blk * kk + kk - 1 - off;
            const uint32_t a = k[lo], b = k[hi];
            k[lo] = min(a, b);
            k[hi] = max(a, b);
        }
#pragma unroll
        for (int j = half >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int c = 0; c < C / 2; ++c) {
                const int lo = ((c & ~(j - 1)) << 1) | (c & (j - 1));
                const int hi = lo + j;
                const uint32_t a = k[lo], b = k[hi];
                k[lo] = min(a, b);
                k[hi] = max(a, b);
            }
        }
    }
}

}  // namespace grf
