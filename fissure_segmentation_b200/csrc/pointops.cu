// pointops_cuda operator family behind models/pointtransformer/pointops.py (upstream
// POSTECH-CVLab/point-transformer lib/pointops, not vendored by the reference). knnquery lives in
// knn.cu; this file holds farthest point sampling and the gather-style operators, plus the fused
// Adam step and the library's version / error-string entry points.
#include "fs_common.cuh"

namespace {

constexpr int FPS_THREADS = 1024;

// One CTA per segment. Each round: every thread relaxes tmp[j] = min(tmp[j], |p_j - p_last|^2) for its
// points and proposes its farthest point; a two-level (shuffle, shared) arg-max picks the winner,
// ties -> lower index.
__global__ void __launch_bounds__(FPS_THREADS)
fps_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offset, const int32_t* __restrict__ new_offset,
           float* __restrict__ tmp, int32_t* __restrict__ idx) {
    __shared__ float s_d[32];
    __shared__ int s_i[32];
    __shared__ int s_last;
    const int s = blockIdx.x;
    const int start = s == 0 ? 0 : offset[s - 1];
    const int end = offset[s];
    const int ostart = s == 0 ? 0 : new_offset[s - 1];
    const int m = new_offset[s] - ostart;
    if (m <= 0 || end <= start) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int last = start;
    if (threadIdx.x == 0) idx[ostart] = start;
    for (int r = 1; r < m; ++r) {
        const float lx = __ldg(xyz + 3ll * last), ly = __ldg(xyz + 3ll * last + 1), lz = __ldg(xyz + 3ll * last + 2);
        float bd = -1.f;
        int bi = FS_IDX_PAD;
        for (int j = start + threadIdx.x; j < end; j += FPS_THREADS) {
            const float dx = __ldg(xyz + 3ll * j) - lx, dy = __ldg(xyz + 3ll * j + 1) - ly, dz = __ldg(xyz + 3ll * j + 2) - lz;
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const float t = fminf(tmp[j], d);
            tmp[j] = t;
            if (t > bd) { bd = t; bi = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(FS_FULL_MASK, bd, o);
            const int oi = __shfl_xor_sync(FS_FULL_MASK, bi, o);
            if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
        }
        if (lane == 0) { s_d[warp] = bd; s_i[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            bd = s_d[lane]; bi = s_i[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(FS_FULL_MASK, bd, o);
                const int oi = __shfl_xor_sync(FS_FULL_MASK, bi, o);
                if (od > bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (lane == 0) { s_last = bi; idx[ostart + r] = bi; }
        }
        __syncthreads();
        last = s_last;
        __syncthreads();
    }
}

// Register-resident variant for segments of up to FPS_THREADS * PPT points: the segment's coordinates are staged once
// in shared memory (the winner's coordinates are read from there), every thread keeps its PPT points and their running
// minimum distances in registers, and one round costs ONE block barrier: each warp publishes its best (distance, index)
// as a 64-bit key into a buffer selected by the round's parity and every warp reduces the 32 keys redundantly.
// key = (float bits of the distance << 32) | (0xffffffff - index): the maximum is the farthest point, ties -> lower index.
template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS)
fps_regs_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ offset, const int32_t* __restrict__ new_offset,
                float* __restrict__ tmp, int32_t* __restrict__ idx) {
    extern __shared__ float4 s_pts[];
    __shared__ unsigned long long s_key[2][32];
    const int s = blockIdx.x;
    const int start = s == 0 ? 0 : offset[s - 1];
    const int end = offset[s];
    const int ostart = s == 0 ? 0 : new_offset[s - 1];
    const int m = new_offset[s] - ostart;
    const int n = end - start;
    if (m <= 0 || n <= 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float px[PPT], py[PPT], pz[PPT], pt[PPT];
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
        const int j = threadIdx.x + i * FPS_THREADS;
        px[i] = py[i] = pz[i] = 0.f;
        pt[i] = -1.f;                                   // slots past the segment never win
        if (j < n) {
            px[i] = __ldg(xyz + 3ll * (start + j)); py[i] = __ldg(xyz + 3ll * (start + j) + 1); pz[i] = __ldg(xyz + 3ll * (start + j) + 2);
            pt[i] = tmp[start + j];
            s_pts[j] = make_float4(px[i], py[i], pz[i], 0.f);
        }
    }
    if (threadIdx.x == 0) idx[ostart] = start;
    __syncthreads();
    int last = 0;
    for (int r = 1; r < m; ++r) {
        const float4 lp = s_pts[last];
        unsigned long long best = 0ull;
#pragma unroll
        for (int i = 0; i < PPT; ++i) {
            const int j = threadIdx.x + i * FPS_THREADS;
            const float dx = px[i] - lp.x, dy = py[i] - lp.y, dz = pz[i] - lp.z;
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (pt[i] >= 0.f) {
                pt[i] = fminf(pt[i], d);
                const unsigned long long key = ((unsigned long long)__float_as_uint(pt[i]) << 32) | (unsigned)(0xffffffffu - (unsigned)j);
                best = key > best ? key : best;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(FS_FULL_MASK, best, o);
            best = ob > best ? ob : best;
        }
        if (lane == 0) s_key[r & 1][warp] = best;
        __syncthreads();
        best = s_key[r & 1][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ob = __shfl_xor_sync(FS_FULL_MASK, best, o);
            best = ob > best ? ob : best;
        }
        last = (int)(0xffffffffu - (unsigned)(best & 0xffffffffull));
        if (threadIdx.x == 0) idx[ostart + r] = start + last;
    }
#pragma unroll
    for (int i = 0; i < PPT; ++i) {                     // the caller's workspace ends up as the global kernel leaves it
        const int j = threadIdx.x + i * FPS_THREADS;
        if (j < n) tmp[start + j] = pt[i];
    }
}

__global__ void grouping_fwd_kernel(long long total, int nsample, int c, const float* __restrict__ in,
                                    const int32_t* __restrict__ idx, float* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ms = e / c;
        const int ch = (int)(e - ms * c);
        out[e] = __ldg(in + (long long)__ldg(idx + ms) * c + ch);
    }
}
__global__ void grouping_bwd_kernel(long long total, int nsample, int c, const float* __restrict__ go,
                                    const int32_t* __restrict__ idx, float* __restrict__ gi) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ms = e / c;
        const int ch = (int)(e - ms * c);
        atomicAdd(gi + (long long)__ldg(idx + ms) * c + ch, __ldg(go + e));
    }
}
__global__ void interpolation_fwd_kernel(long long total, int c, int k, const float* __restrict__ in,
                                         const int32_t* __restrict__ idx, const float* __restrict__ w,
                                         float* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / c;
        const int ch = (int)(e - n * c);
        float acc = 0.f;
        for (int i = 0; i < k; ++i) acc = fmaf(__ldg(w + n * k + i), __ldg(in + (long long)__ldg(idx + n * k + i) * c + ch), acc);
        out[e] = acc;
    }
}
__global__ void interpolation_bwd_kernel(long long total, int c, int k, const float* __restrict__ go,
                                         const int32_t* __restrict__ idx, const float* __restrict__ w,
                                         float* __restrict__ gi) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / c;
        const int ch = (int)(e - n * c);
        const float g = __ldg(go + e);
        for (int i = 0; i < k; ++i) atomicAdd(gi + (long long)__ldg(idx + n * k + i) * c + ch, g * __ldg(w + n * k + i));
    }
}
__global__ void subtraction_fwd_kernel(long long total, int nsample, int c, const float* __restrict__ in1,
                                       const float* __restrict__ in2, const int32_t* __restrict__ idx,
                                       float* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ns = e / c;
        const int ch = (int)(e - ns * c);
        const long long n = ns / nsample;
        out[e] = __ldg(in1 + n * c + ch) - __ldg(in2 + (long long)__ldg(idx + ns) * c + ch);
    }
}
__global__ void subtraction_bwd_kernel(long long total, int nsample, int c, const int32_t* __restrict__ idx,
                                       const float* __restrict__ go, float* __restrict__ g1, float* __restrict__ g2) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long ns = e / c;
        const int ch = (int)(e - ns * c);
        const long long n = ns / nsample;
        const float g = __ldg(go + e);
        atomicAdd(g1 + n * c + ch, g);
        atomicAdd(g2 + (long long)__ldg(idx + ns) * c + ch, -g);
    }
}
__global__ void aggregation_fwd_kernel(long long total, int nsample, int c, int w_c, const float* __restrict__ in,
                                       const float* __restrict__ pos, const float* __restrict__ w,
                                       const int32_t* __restrict__ idx, float* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / c;
        const int ch = (int)(e - n * c);
        const int wc = ch % w_c;
        float acc = 0.f;
        for (int s = 0; s < nsample; ++s) {
            const long long ns = n * nsample + s;
            acc = fmaf(__ldg(in + (long long)__ldg(idx + ns) * c + ch) + __ldg(pos + ns * c + ch), __ldg(w + ns * w_c + wc), acc);
        }
        out[e] = acc;
    }
}
__global__ void aggregation_bwd_kernel(long long total, int nsample, int c, int w_c, const float* __restrict__ in,
                                       const float* __restrict__ pos, const float* __restrict__ w,
                                       const int32_t* __restrict__ idx, const float* __restrict__ go,
                                       float* __restrict__ gi, float* __restrict__ gp, float* __restrict__ gw) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long n = e / c;
        const int ch = (int)(e - n * c);
        const int wc = ch % w_c;
        const float g = __ldg(go + e);
        for (int s = 0; s < nsample; ++s) {
            const long long ns = n * nsample + s;
            const long long src = (long long)__ldg(idx + ns) * c + ch;
            const float wv = __ldg(w + ns * w_c + wc);
            atomicAdd(gi + src, g * wv);
            gp[ns * c + ch] = g * wv;
            atomicAdd(gw + ns * w_c + wc, g * (__ldg(in + src) + __ldg(pos + ns * c + ch)));
        }
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2_sqrt, float gscale, const float* __restrict__ dyn) {
    if (dyn) {  // device-resident [step, lr]: lets a captured CUDA graph replay with a moving step / schedule
        const float st = __ldg(dyn);
        lr = __ldg(dyn + 1);
        bc1 = 1.f - powf(b1, st);
        bc2_sqrt = sqrtf(1.f - powf(b2, st));
    }
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float pv = p[e];
        const float gv = fmaf(wd, pv, g[e] * gscale);
        const float mv = fmaf(1.f - b1, gv - m[e], m[e]);          // lerp(m, g, 1-b1)
        const float vv = fmaf(b2, v[e], (1.f - b2) * gv * gv);
        m[e] = mv;
        v[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[e] = pv - (lr / bc1) * (mv / denom);
    }
}

// Tail bucket of the data-parallel step: all-reduce and Adam in ONE kernel over peer memory. Every rank reads the same
// slice of all ranks' gradient buffers (symmetric allocations mapped over NVLink / NVSwitch; peers[r] = base pointer of
// rank r) in the same order r = 0..world-1, so all ranks form the bit-identical sum and apply the bit-identical update;
// nothing is written to a peer. The caller brackets the launch with device-side barriers over the same symmetric
// allocation (the peers' gradients are complete before, and no peer starts overwriting its gradients before every rank
// has read them).
__global__ void adam_peers_kernel(float* __restrict__ p, const unsigned long long* __restrict__ peers, int world,
                                  long long offset, float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                  float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                                  const float* __restrict__ dyn, float* __restrict__ g_out) {
    if (dyn) {
        const float st = __ldg(dyn);
        lr = __ldg(dyn + 1);
        bc1 = 1.f - powf(b1, st);
        bc2_sqrt = sqrtf(1.f - powf(b2, st));
    }
    auto update = [&](long long e, float gs) {
        if (g_out) g_out[e] = gs;
        const float pv = p[e];
        const float gv = fmaf(wd, pv, gs * gscale);
        const float mv = fmaf(1.f - b1, gv - m[e], m[e]);
        const float vv = fmaf(b2, v[e], (1.f - b2) * gv * gv);
        m[e] = mv;
        v[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[e] = pv - (lr / bc1) * (mv / denom);
    };
    // 16-byte peer loads (all ranks' loads of a quad are in flight before the first add) when the slice is aligned
    const bool vec = (offset & 3) == 0;                    // the symmetric buffers themselves are at least 16-byte aligned
    const long long nq = vec ? n >> 2 : 0;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += (long long)gridDim.x * blockDim.x) {
        float4 gs = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int r = 0; r < world; ++r) {
            const float4 t = __ldcv(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(peers[r]) + offset) + q);
            gs.x += t.x; gs.y += t.y; gs.z += t.z; gs.w += t.w;   // rank order: identical on every rank
        }
        update(4 * q, gs.x); update(4 * q + 1, gs.y); update(4 * q + 2, gs.z); update(4 * q + 3, gs.w);
    }
    for (long long e = 4 * nq + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float gs = 0.f;
        for (int r = 0; r < world; ++r) gs += __ldcv(reinterpret_cast<const float*>(peers[r]) + offset + e);
        update(e, gs);
    }
}

// 30-bit Morton (Z-order) code of the first three channels, 10 bits per axis over the fixed box [-2, 2)^3
// (grid coordinates live in [-1, 1] before augmentation, augmentations.py:52-75).
__device__ __forceinline__ uint32_t spread10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void morton_kernel(const float* __restrict__ coords, long long bs, long long cs, long long ps, int N,
                              long long total, int32_t* __restrict__ codes) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / N;
        const long long n = e - b * N;
        const float* p = coords + b * bs + n * ps;
        uint32_t q[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = (__ldg(p + c * cs) + 2.0f) * 256.0f;          // [-2,2) -> [0,1024)
            v = v != v ? 0.f : fminf(fmaxf(v, 0.f), 1023.f);        // NaN -> 0, clamp
            q[c] = (uint32_t)v;
        }
        codes[e] = (int32_t)(spread10(q[0]) | (spread10(q[1]) << 1) | (spread10(q[2]) << 2));
    }
}

int flat_grid(long long total) {
    long long g = (total + 255) / 256;
    const long long cap = (long long)FS_NUM_SMS * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int fs_version(void) { return 101; }

extern "C" size_t fs_stats_buffer_doubles(int C) { return C > 0 ? (size_t)fs_stats_doubles(C) : 0; }

extern "C" const char* fs_error_string(int code) {
    switch (code) {
        case FS_OK: return "success";
        case FS_ERR_BAD_ARG: return "fissure_b200: bad argument";
        case FS_ERR_UNSUPPORTED: return "fissure_b200: unsupported shape or dtype";
        case FS_ERR_ALIGNMENT: return "fissure_b200: pointer or leading dimension not 16-byte aligned";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "fissure_b200: unknown error";
}

extern "C" int fs_furthestsampling(int device, fs_stream_t stream_, int b, int n_max, const float* xyz, const int32_t* offset,
                                   const int32_t* new_offset, float* tmp, int32_t* idx) {
    if (!xyz || !offset || !new_offset || !tmp || !idx || b < 0) return FS_ERR_BAD_ARG;
    if (b == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    // n_max (the largest segment, pointops.py:25-27 computes it for the upstream kernel) selects the register-resident
    // variant; 0 = unknown -> the global-memory kernel
    if (n_max > 0 && n_max <= FPS_THREADS * 8) {
        const size_t smem = (size_t)n_max * sizeof(float4);
#define FPS_GO(PPT)                                                                                                   \
    do {                                                                                                              \
        if (smem > 48 * 1024) FS_CUDA_TRY(cudaFuncSetAttribute(fps_regs_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        fps_regs_kernel<PPT><<<b, FPS_THREADS, smem, stream>>>(xyz, offset, new_offset, tmp, idx);                     \
    } while (0)
        if (n_max <= FPS_THREADS) FPS_GO(1);
        else if (n_max <= 2 * FPS_THREADS) FPS_GO(2);
        else if (n_max <= 4 * FPS_THREADS) FPS_GO(4);
        else FPS_GO(8);
#undef FPS_GO
    } else {
        fps_kernel<<<b, FPS_THREADS, 0, stream>>>(xyz, offset, new_offset, tmp, idx);
    }
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

#define FLAT_LAUNCH(kernel, total, ...)                                                      \
    do {                                                                                     \
        if ((total) <= 0) return FS_OK;                                                      \
        FS_ENTER(device);                                                                    \
        kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream_>>>(total, __VA_ARGS__);     \
        FS_RETURN_IF_LAUNCH_FAILED();                                                        \
        return FS_OK;                                                                        \
    } while (0)

extern "C" int fs_grouping_fwd(int device, fs_stream_t stream_, int m, int nsample, int c, const float* in,
                               const int32_t* idx, float* out) {
    if (!in || !idx || !out || m < 0 || nsample < 0 || c < 0) return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(grouping_fwd_kernel, (long long)m * nsample * c, nsample, c, in, idx, out);
}
extern "C" int fs_grouping_bwd(int device, fs_stream_t stream_, int m, int nsample, int c, const float* grad_out,
                               const int32_t* idx, float* grad_in) {
    if (!grad_out || !idx || !grad_in || m < 0 || nsample < 0 || c < 0) return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(grouping_bwd_kernel, (long long)m * nsample * c, nsample, c, grad_out, idx, grad_in);
}
extern "C" int fs_interpolation_fwd(int device, fs_stream_t stream_, int n, int c, int k, const float* in,
                                    const int32_t* idx, const float* weight, float* out) {
    if (!in || !idx || !weight || !out || n < 0 || c < 0 || k < 0) return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(interpolation_fwd_kernel, (long long)n * c, c, k, in, idx, weight, out);
}
extern "C" int fs_interpolation_bwd(int device, fs_stream_t stream_, int n, int c, int k, const float* grad_out,
                                    const int32_t* idx, const float* weight, float* grad_in) {
    if (!grad_out || !idx || !weight || !grad_in || n < 0 || c < 0 || k < 0) return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(interpolation_bwd_kernel, (long long)n * c, c, k, grad_out, idx, weight, grad_in);
}
extern "C" int fs_subtraction_fwd(int device, fs_stream_t stream_, int n, int nsample, int c, const float* in1,
                                  const float* in2, const int32_t* idx, float* out) {
    if (!in1 || !in2 || !idx || !out || n < 0 || nsample < 0 || c < 0) return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(subtraction_fwd_kernel, (long long)n * nsample * c, nsample, c, in1, in2, idx, out);
}
extern "C" int fs_subtraction_bwd(int device, fs_stream_t stream_, int n, int nsample, int c, const int32_t* idx,
                                  const float* grad_out, float* grad_in1, float* grad_in2) {
    if (!idx || !grad_out || !grad_in1 || !grad_in2 || n < 0 || nsample < 0 || c < 0) return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(subtraction_bwd_kernel, (long long)n * nsample * c, nsample, c, idx, grad_out, grad_in1, grad_in2);
}
extern "C" int fs_aggregation_fwd(int device, fs_stream_t stream_, int n, int nsample, int c, int w_c, const float* in,
                                  const float* pos, const float* weight, const int32_t* idx, float* out) {
    if (!in || !pos || !weight || !idx || !out || n < 0 || nsample < 0 || c < 0 || w_c <= 0) return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(aggregation_fwd_kernel, (long long)n * c, nsample, c, w_c, in, pos, weight, idx, out);
}
extern "C" int fs_aggregation_bwd(int device, fs_stream_t stream_, int n, int nsample, int c, int w_c, const float* in,
                                  const float* pos, const float* weight, const int32_t* idx, const float* grad_out,
                                  float* grad_in, float* grad_pos, float* grad_weight) {
    if (!in || !pos || !weight || !idx || !grad_out || !grad_in || !grad_pos || !grad_weight || n < 0 || nsample < 0 ||
        c < 0 || w_c <= 0)
        return FS_ERR_BAD_ARG;
    FLAT_LAUNCH(aggregation_bwd_kernel, (long long)n * c, nsample, c, w_c, in, pos, weight, idx, grad_out, grad_in,
                grad_pos, grad_weight);
}

extern "C" int fs_morton_codes(int device, fs_stream_t stream_, const float* coords, long long batch_stride,
                               long long chan_stride, long long point_stride, int B, int N, int32_t* codes) {
    if (B < 0 || N < 0) return FS_ERR_BAD_ARG;
    if (B == 0 || N == 0) return FS_OK;
    if (!coords || !codes) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    const long long total = (long long)B * N;
    morton_kernel<<<flat_grid(total), 256, 0, (cudaStream_t)stream_>>>(coords, batch_stride, chan_stride, point_stride, N,
                                                                     total, codes);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_adam_step(int device, fs_stream_t stream_, float* param, const float* grad, float* exp_avg,
                            float* exp_avg_sq, long long n, float lr, float beta1, float beta2, float eps,
                            float weight_decay, int step, float grad_scale, const float* dyn_step_lr) {
    if (!param || !grad || !exp_avg || !exp_avg_sq || n < 0 || (step <= 0 && !dyn_step_lr)) return FS_ERR_BAD_ARG;
    if (n == 0) return FS_OK;
    FS_ENTER(device);
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    adam_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream_>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                               weight_decay, bc1, bc2_sqrt, grad_scale, dyn_step_lr);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_adam_step_peers(int device, fs_stream_t stream_, float* param, const unsigned long long* peer_grad_ptrs,
                                  int world, long long elem_offset, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                                  const float* dyn_step_lr, float* grad_sum_out) {
    if (!param || !peer_grad_ptrs || !exp_avg || !exp_avg_sq || n < 0 || world < 1 || world > 64 || elem_offset < 0)
        return FS_ERR_BAD_ARG;
    if (n == 0) return FS_OK;
    FS_ENTER(device);
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    adam_peers_kernel<<<flat_grid(n), 256, 0, (cudaStream_t)stream_>>>(param, peer_grad_ptrs, world, elem_offset, exp_avg,
                                                                     exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                     bc1, bc2_sqrt, grad_scale, dyn_step_lr, grad_sum_out);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
