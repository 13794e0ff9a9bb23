// Shared device/host helpers for the fissure_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <float.h>

#include "../../include/fissure_b200.h"

#ifndef FS_NUM_SMS
#define FS_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
#endif

#define FS_FULL_MASK 0xffffffffu
#define FS_IDX_PAD 0x7fffffff

// Every entry point returns 0, a negative fs_status, or a positive cudaError_t.
#define FS_RETURN_IF_LAUNCH_FAILED()                     \
    do {                                                 \
        cudaError_t fs_e_ = cudaGetLastError();          \
        if (fs_e_ != cudaSuccess) return (int)fs_e_;     \
    } while (0)

#define FS_CUDA_TRY(expr)                                \
    do {                                                 \
        cudaError_t fs_e_ = (expr);                      \
        if (fs_e_ != cudaSuccess) return (int)fs_e_;     \
    } while (0)

// RAII device guard: the library never assumes the caller's current device.
struct FsDeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t err = cudaSuccess;
    explicit FsDeviceGuard(int device) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) {
            err = cudaSetDevice(device);
            switched = (err == cudaSuccess);
        }
    }
    ~FsDeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

#define FS_ENTER(device)                                 \
    FsDeviceGuard fs_guard_(device);                     \
    if (fs_guard_.err != cudaSuccess) return (int)fs_guard_.err

static inline int fs_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

__device__ __forceinline__ int fs_lane() { return threadIdx.x & 31; }

__device__ __forceinline__ float fs_leaky(float z) { return z > 0.f ? z : 0.2f * z; }

// 128-bit read-only global load.
__device__ __forceinline__ float4 fs_ldg4(const float* p) {
    return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ void fs_bf16x8_to_float(const uint4& v, float* f) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// Load VEC consecutive channels of a table row as fp32 (VEC=4 for fp32 rows, 8 for bf16 rows;
// both are one 128-bit transaction).
template <typename T> struct FsRow;
template <> struct FsRow<float> {
    static constexpr int VEC = 4;
    __device__ __forceinline__ static void load(const float* p, float* f) {
        float4 v = __ldg(reinterpret_cast<const float4*>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    __device__ __forceinline__ static void store(float* p, const float* f) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};
template <> struct FsRow<__nv_bfloat16> {
    static constexpr int VEC = 8;
    __device__ __forceinline__ static void load(const __nv_bfloat16* p, float* f) {
        uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
        fs_bf16x8_to_float(v, f);
    }
    __device__ __forceinline__ static void store(__nv_bfloat16* p, const float* f) {
        uint4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
};

__device__ __forceinline__ double fs_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FS_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float fs_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FS_FULL_MASK, v, o);
    return v;
}
