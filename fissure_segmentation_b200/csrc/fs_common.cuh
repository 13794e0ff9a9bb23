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

// ------------------------------------------------------------------------------------------------
// Per-channel statistics buffers (fp64). Layout, C channels, S = fs_stat_slots(C):
//   [0, 2C)  final sums (sum1 | sum2)     [2C, 3C)  pivot     [3C, 3C + S*2C)  slot partials     [+0]  ticket
// Blocks add their partials into slot (blockIdx.x % S) (S-fold less contention than one set of addresses);
// the last block to finish sums the slots into the final area. The caller zero-fills the buffer.
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline int fs_stat_slots(int C) {
    int s = 2048 / C;
    return s < 1 ? 1 : (s > 32 ? 32 : s);
}
__host__ __device__ inline long long fs_stats_doubles(int C) { return 3ll * C + (long long)fs_stat_slots(C) * 2 * C + 2; }

// Threads whose (threadIdx.x % period) agree own the same NCH channels chan[0..NCH).
// buf: shared double[2 * NCH * blockDim.x]. Every thread of the block must call this exactly once per kernel.
// BatchNorm finalisation folded into the kernel that applies it (saves one tiny launch per BatchNorm layer and step):
// every block derives the coefficients of the channels it needs from the finished statistics; ONE designated thread per
// channel also publishes them (coef = [mean | invstd | scale | beta], the layout of fs_bn_finalize) and updates the
// running statistics / num_batches_tracked exactly like bn_finalize_kernel.
struct FsBnFin {
    const double* stats;        // nullptr: no inline finalisation (coefficients are read from `coef`)
    double count;
    const float* gamma;
    const float* beta;
    float eps, momentum;
    float* running_mean;
    float* running_var;
    long long* nbt;
    float* coef_out;
};
__device__ __forceinline__ void fs_bn_fin_channel(const FsBnFin& f, int c, int C, bool publish, float& mean_f, float& invstd,
                                                  float& scale, float& beta) {
    // fp64 only where it matters (E[y^2] - mean^2 cancels); the reciprocal square root is taken in fp32 (correctly rounded
    // sqrtf and division: every block derives bit-identical coefficients) - the fp64 divide / sqrt routines cost hundreds
    // of instructions per channel
    const double ic = 1.0 / f.count;
    const double m1 = f.stats[c] * ic;
    double var = f.stats[C + c] * ic - m1 * m1;
    if (var < 0.0) var = 0.0;
    const double mean = m1 + f.stats[2 * C + c];
    invstd = 1.0f / sqrtf((float)var + f.eps);
    mean_f = (float)mean;
    scale = __ldg(f.gamma + c) * invstd;
    beta = __ldg(f.beta + c);
    if (publish) {
        f.coef_out[c] = mean_f;
        f.coef_out[C + c] = invstd;
        f.coef_out[2 * C + c] = scale;
        f.coef_out[3 * C + c] = beta;
        if (f.running_mean) f.running_mean[c] = (1.f - f.momentum) * f.running_mean[c] + f.momentum * mean_f;
        if (f.running_var) {
            const double unbiased = f.count > 1.0 ? var * f.count / (f.count - 1.0) : var;
            f.running_var[c] = (1.f - f.momentum) * f.running_var[c] + f.momentum * (float)unbiased;
        }
        if (c == 0 && f.nbt) *f.nbt += 1;
    }
}

template <int NCH>
__device__ __forceinline__ void fs_stats_commit_impl(double* buf, const double* s1, const double* s2, const int* chan,
                                                     int period, int C, double* gstats, unsigned block_linear,
                                                     unsigned num_blocks) {
    const int T = blockDim.x, t = threadIdx.x;
    __shared__ bool fs_last_block;
#pragma unroll
    for (int e = 0; e < NCH; ++e) {
        buf[e * T + t] = s1[e];
        buf[(NCH + e) * T + t] = s2[e];
    }
    __syncthreads();
    const int slots = fs_stat_slots(C);
    double* slot = gstats + 3 * C + (block_linear % slots) * 2 * C;
    if (t < period) {
#pragma unroll
        for (int e = 0; e < NCH; ++e) {
            double a = 0.0, b = 0.0;
            for (int j = t; j < T; j += period) { a += buf[e * T + j]; b += buf[(NCH + e) * T + j]; }
            atomicAdd(slot + chan[e], a);
            atomicAdd(slot + C + chan[e], b);
        }
    }
    __threadfence();
    __syncthreads();
    unsigned* ticket = reinterpret_cast<unsigned*>(gstats + 3 * C + slots * 2 * C);
    if (t == 0) fs_last_block = atomicAdd(ticket, 1u) == num_blocks - 1;
    __syncthreads();
    if (fs_last_block) {
        __threadfence();
        for (int c = t; c < 2 * C; c += T) {
            double a = 0.0;
            for (int sidx = 0; sidx < slots; ++sidx) a += __ldcg(gstats + 3 * C + sidx * 2 * C + c);
            gstats[c] = a;
        }
    }
}

template <int NCH>
__device__ __forceinline__ void fs_stats_commit(double* buf, const double* s1, const double* s2, const int* chan,
                                                int period, int C, double* gstats) {
    fs_stats_commit_impl<NCH>(buf, s1, s2, chan, period, C, gstats, blockIdx.x, gridDim.x);
}
template <int NCH>
__device__ __forceinline__ void fs_stats_commit_2d(double* buf, const double* s1, const double* s2, const int* chan,
                                                   int period, int C, double* gstats) {
    fs_stats_commit_impl<NCH>(buf, s1, s2, chan, period, C, gstats, blockIdx.y * gridDim.x + blockIdx.x,
                              gridDim.x * gridDim.y);
}
