// BatchNorm + LeakyReLU on point-major tables (the dense layers around the EdgeConvs: SharedFullyConnected
// with dim=1, models/dgcnn.py:282-323, used at :123-137) and the fused global max-pool
// (AdaptiveMaxPool1d after Conv1d+BN+LeakyReLU, models/dgcnn.py:123-126, 156).
//
// The 1x1 convolutions themselves are cuBLAS GEMMs; these kernels replace the BatchNorm / activation /
// pooling passes around them: column statistics in one read, normalise+activate in one read+write, and for
// the global feature only per-cloud max/min of the GEMM output is kept (LeakyReLU(BN(.)) is monotone), so the
// B*N x 1024 activation is never written.
//
// An optional per-cloud row bias (B x C, fp32) is added to x on load: segmentation[0] acts on
// [local | broadcast global] (models/dgcnn.py:159), i.e. local GEMM + one bias row per cloud.
#include "fs_common.cuh"

namespace {

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    __device__ __forceinline__ static void load(const float* p, float* f) {
        float4 v = __ldg(reinterpret_cast<const float4*>(p));
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    __device__ __forceinline__ static void store(float* p, const float* f) {
        *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
    }
};
template <> struct Vec<__nv_bfloat16> {
    static constexpr int N = 8;
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

constexpr int DN_THREADS = 256;

// thread -> (row within block pass, channel vector). tpr = threads per row = C / VEC (power of two, <= 256)
struct RowMap {
    int tpr, rows_per_pass, r, c0;
    __device__ RowMap(int C, int vec) {
        tpr = C / vec;
        if (tpr > DN_THREADS) tpr = DN_THREADS;
        rows_per_pass = DN_THREADS / tpr;
        r = threadIdx.x / tpr;
        c0 = (threadIdx.x - r * tpr) * vec;
    }
};

template <int V>
__device__ __forceinline__ void add_bias(float* f, const float* __restrict__ bias, long long row, int N, int C, int c) {
    if (bias) {
        const float* bp = bias + (row / N) * C + c;
#pragma unroll
        for (int i = 0; i < V; ++i) f[i] += __ldg(bp + i);
    }
}

// Per-column sums of (x - pivot), (x - pivot)^2 in fp64; pivot = row 0 (layout: fs_stats_commit in fs_common.cuh).
// Requires C <= tpr * V, i.e. one channel vector per thread (C <= 1024 fp32, <= 2048 bf16).
template <typename T>
__global__ void __launch_bounds__(DN_THREADS)
colstats_kernel(const T* __restrict__ x, int ld, long long rows, int C, const float* __restrict__ bias, int N,
                double* __restrict__ stats) {
    constexpr int V = Vec<T>::N;
    extern __shared__ double red[];     // [2 * V * DN_THREADS]
    RowMap m(C, V);
    const int cc = m.c0;
    float piv[V];
    Vec<T>::load(x + cc, piv);
    add_bias<V>(piv, bias, 0, N, C, cc);
    float s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { s1[i] = 0.f; s2[i] = 0.f; }
    const long long step = (long long)gridDim.x * m.rows_per_pass;
    long long row = (long long)blockIdx.x * m.rows_per_pass + m.r;
    for (; row + 3 * step < rows; row += 4 * step) {       // 4 independent 128-bit loads in flight per thread
        float f[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) Vec<T>::load(x + (row + u * step) * ld + cc, f[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            add_bias<V>(f[u], bias, row + u * step, N, C, cc);
#pragma unroll
            for (int i = 0; i < V; ++i) { const float d = f[u][i] - piv[i]; s1[i] += d; s2[i] = fmaf(d, d, s2[i]); }
        }
    }
    for (; row < rows; row += step) {
        float f[V];
        Vec<T>::load(x + row * ld + cc, f);
        add_bias<V>(f, bias, row, N, C, cc);
#pragma unroll
        for (int i = 0; i < V; ++i) { const float d = f[i] - piv[i]; s1[i] += d; s2[i] = fmaf(d, d, s2[i]); }
    }
    double d1[V], d2[V];
    int chans[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { d1[i] = (double)s1[i]; d2[i] = (double)s2[i]; chans[i] = cc + i; }
    fs_stats_commit<V>(red, d1, d2, chans, m.tpr, C, stats);
    if (blockIdx.x == 0 && m.r == 0) {
#pragma unroll
        for (int i = 0; i < V; ++i) stats[2 * C + cc + i] = (double)piv[i];
    }
}

// y = LeakyReLU(scale * (x - mu) + beta)
template <typename T, typename OT>
__global__ void __launch_bounds__(DN_THREADS)
bn_act_apply_kernel(const T* __restrict__ x, int ld, long long rows, int C, const float* __restrict__ bias, int N,
                    const float* __restrict__ coef, const FsBnFin fin, float slope, OT* __restrict__ out, int ld_out) {
    constexpr int V = Vec<T>::N;
    RowMap m(C, V);
    extern __shared__ float ba_coef[];          // [3][C] when the finalisation is folded in: one channel per thread, once
    if (fin.stats) {                            // per block (the fp64 divide / sqrt per channel is not free)
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            float mu1, inv, sc1, be1;
            fs_bn_fin_channel(fin, c, C, blockIdx.x == 0, mu1, inv, sc1, be1);
            ba_coef[c] = mu1; ba_coef[C + c] = sc1; ba_coef[2 * C + c] = be1;
        }
        __syncthreads();
    }
    for (int cc = m.c0; cc < C; cc += m.tpr * V) {
        float mu[V], sc[V], be[V];
        if (fin.stats) {
#pragma unroll
            for (int i = 0; i < V; ++i) { mu[i] = ba_coef[cc + i]; sc[i] = ba_coef[C + cc + i]; be[i] = ba_coef[2 * C + cc + i]; }
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) { mu[i] = __ldg(coef + cc + i); sc[i] = __ldg(coef + 2 * C + cc + i); be[i] = __ldg(coef + 3 * C + cc + i); }
        }
        for (long long row = (long long)blockIdx.x * m.rows_per_pass + m.r; row < rows; row += (long long)gridDim.x * m.rows_per_pass) {
            float f[V];
            Vec<T>::load(x + row * ld + cc, f);
            add_bias<V>(f, bias, row, N, C, cc);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float z = fmaf(sc[i], f[i] - mu[i], be[i]);
                f[i] = z > 0.f ? z : slope * z;
            }
            if (V == 8 && sizeof(OT) == 4) {
                Vec<float>::store(reinterpret_cast<float*>(out) + row * ld_out + cc, f);
                Vec<float>::store(reinterpret_cast<float*>(out) + row * ld_out + cc + 4, f + 4);
            } else if (V == 4 && sizeof(OT) == 2) {
                uint2 v;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
                h[0] = __floats2bfloat162_rn(f[0], f[1]);
                h[1] = __floats2bfloat162_rn(f[2], f[3]);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + row * ld_out + cc) = v;
            } else {
                Vec<OT>::store(out + row * ld_out + cc, f);
            }
        }
    }
}

// d = g * LeakyReLU'(z); dgb = [sum d | sum d * xhat] in fp64 (layout: fs_stats_commit)
__device__ __forceinline__ void load4_any(const float* p, float* f) { Vec<float>::load(p, f); }
__device__ __forceinline__ void load4_any(const __nv_bfloat16* p, float* f) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    float2 a = __bfloat1622float2(h[0]), b2 = __bfloat1622float2(h[1]);
    f[0] = a.x; f[1] = a.y; f[2] = b2.x; f[3] = b2.y;
}

template <typename GT, typename T>
__global__ void __launch_bounds__(DN_THREADS)
bn_act_bwd_reduce_kernel(const GT* __restrict__ g, int ldg, const T* __restrict__ x, int ld, long long rows, int C,
                         const float* __restrict__ bias, int N, const float* __restrict__ coef, float slope,
                         double* __restrict__ dgb) {
    constexpr int V = 4;
    extern __shared__ double red[];     // [2 * V * DN_THREADS]
    RowMap m(C, V);
    const int cc = m.c0;
    float mu[V], inv[V], sc[V], be[V], s1[V], s2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        mu[i] = __ldg(coef + cc + i); inv[i] = __ldg(coef + C + cc + i); sc[i] = __ldg(coef + 2 * C + cc + i);
        be[i] = __ldg(coef + 3 * C + cc + i); s1[i] = 0.f; s2[i] = 0.f;
    }
    const long long step = (long long)gridDim.x * m.rows_per_pass;
    long long row = (long long)blockIdx.x * m.rows_per_pass + m.r;
    for (; row < rows; row += 2 * step) {
        float gv[2][V], f[2][V];
        const bool two = row + step < rows;
        load4_any(g + row * ldg + cc, gv[0]);
        load4_any(x + row * ld + cc, f[0]);
        if (two) { load4_any(g + (row + step) * ldg + cc, gv[1]); load4_any(x + (row + step) * ld + cc, f[1]); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (u == 1 && !two) break;
            add_bias<V>(f[u], bias, row + u * step, N, C, cc);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float xc = f[u][i] - mu[i];
                const float z = fmaf(sc[i], xc, be[i]);
                const float d = z > 0.f ? gv[u][i] : slope * gv[u][i];
                s1[i] += d;
                s2[i] = fmaf(d, xc * inv[i], s2[i]);
            }
        }
    }
    double d1[V], d2[V];
    int chans[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { d1[i] = (double)s1[i]; d2[i] = (double)s2[i]; chans[i] = cc + i; }
    fs_stats_commit<V>(red, d1, d2, chans, m.tpr, C, dgb);
}

// dx = scale * (d - dbeta/M - xhat * dgamma/M)      (train_stats = 0: dx = scale * d)
template <typename GT, typename T, typename OT>
__global__ void __launch_bounds__(DN_THREADS)
bn_act_bwd_apply_kernel(const GT* __restrict__ g, int ldg, const T* __restrict__ x, int ld, long long rows, int C,
                        const float* __restrict__ bias, int N, const float* __restrict__ coef, float slope,
                        const double* __restrict__ dgb, double count, int train_stats, OT* __restrict__ dx, int ld_dx) {
    constexpr int V = 4;
    RowMap m(C, V);
    for (int cc = m.c0; cc < C; cc += m.tpr * V) {
        // mean(d) and mean(d xhat) as hi + lo float pairs: the per-cloud sums of dx (the gradient of a per-cloud bias
        // row, i.e. of the global feature) cancel to rounding level, so an fp32-rounded mean would leave a systematic
        // residue of rows * ulp(mean) in them
        float mu[V], inv[V], sc[V], be[V], mb[V], mg[V], mbl[V], mgl[V];
#pragma unroll
        for (int i = 0; i < V; ++i) {
            mu[i] = __ldg(coef + cc + i); inv[i] = __ldg(coef + C + cc + i); sc[i] = __ldg(coef + 2 * C + cc + i);
            be[i] = __ldg(coef + 3 * C + cc + i);
            const double mbd = train_stats ? dgb[cc + i] / count : 0.0;
            const double mgd = train_stats ? dgb[C + cc + i] / count * (double)inv[i] : 0.0;
            mb[i] = (float)mbd; mbl[i] = (float)(mbd - (double)mb[i]);
            mg[i] = (float)mgd; mgl[i] = (float)(mgd - (double)mg[i]);
        }
        for (long long row = (long long)blockIdx.x * m.rows_per_pass + m.r; row < rows; row += (long long)gridDim.x * m.rows_per_pass) {
            float gv[V], f[V], o[V];
            if (sizeof(GT) == 4) Vec<float>::load(reinterpret_cast<const float*>(g) + row * ldg + cc, gv);
            else {
                uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(g) + row * ldg + cc));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
                float2 a = __bfloat1622float2(h[0]), b2 = __bfloat1622float2(h[1]);
                gv[0] = a.x; gv[1] = a.y; gv[2] = b2.x; gv[3] = b2.y;
            }
            if (sizeof(T) == 4) Vec<float>::load(reinterpret_cast<const float*>(x) + row * ld + cc, f);
            else {
                uint2 v = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + row * ld + cc));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
                float2 a = __bfloat1622float2(h[0]), b2 = __bfloat1622float2(h[1]);
                f[0] = a.x; f[1] = a.y; f[2] = b2.x; f[3] = b2.y;
            }
            add_bias<V>(f, bias, row, N, C, cc);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                const float xc = f[i] - mu[i];
                const float z = fmaf(sc[i], xc, be[i]);
                const float d = z > 0.f ? gv[i] : slope * gv[i];
                o[i] = sc[i] * (((d - mb[i]) - mbl[i]) - fmaf(mgl[i], xc, mg[i] * xc));
            }
            if (sizeof(OT) == 4) Vec<float>::store(reinterpret_cast<float*>(dx) + row * ld_dx + cc, o);
            else {
                uint2 v;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
                h[0] = __floats2bfloat162_rn(o[0], o[1]);
                h[1] = __floats2bfloat162_rn(o[2], o[3]);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(dx) + row * ld_dx + cc) = v;
            }
        }
    }
}

// Global max-pool fused with the statistics. grid = (chunks, B): each block scans a chunk of the N rows of one cloud
// for all C channels (128-bit loads, four rows in flight per thread), keeps max of the sign-flipped key (so that
// gamma < 0 channels take the minimum) with its row, and publishes (key, row) per channel with one 64-bit atomicMax;
// ties go to the lower row like a sequential arg-max. Column statistics go through fs_stats_commit.
__device__ __forceinline__ unsigned int ordered_u32(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_u32(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

template <typename T>
__global__ void __launch_bounds__(DN_THREADS, 3)
pool_reduce_kernel(const T* __restrict__ x, int ld, int N, int C, const float* __restrict__ gamma,
                   unsigned long long* __restrict__ packed, double* __restrict__ stats) {
    constexpr int V = Vec<T>::N;
    extern __shared__ double red[];     // [2 * V * DN_THREADS]
    RowMap m(C, V);
    const int cc = m.c0;
    const int b = blockIdx.y;
    const int rows_per_chunk = (N + gridDim.x - 1) / gridDim.x;
    const int r_begin = blockIdx.x * rows_per_chunk;
    const int r_end = min(N, r_begin + rows_per_chunk);
    const T* xb = x + (long long)b * N * ld + cc;
    unsigned int flip[V];
    float piv[V], f1[V], f2[V], best[V];
    int barg[V];
    {
        float p0[V];
        Vec<T>::load(x + cc, p0);
#pragma unroll
        for (int i = 0; i < V; ++i) {
            flip[i] = __ldg(gamma + cc + i) >= 0.f ? 0u : 0x80000000u;
            piv[i] = stats ? p0[i] : 0.f; f1[i] = 0.f; f2[i] = 0.f; best[i] = -INFINITY; barg[i] = 0x7fffffff;
        }
    }
    const int step = m.rows_per_pass;
    int r = r_begin + m.r;
    for (; r < r_end; r += 4 * step) {
        float f[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) Vec<T>::load(xb + (long long)min(r + u * step, r_end - 1) * ld, f[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (r + u * step < r_end) {
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float key = __uint_as_float(__float_as_uint(f[u][i]) ^ flip[i]);
                    const bool better = key > best[i];
                    best[i] = better ? key : best[i];
                    barg[i] = better ? r + u * step : barg[i];
                    const float d = f[u][i] - piv[i];
                    f1[i] += d;
                    f2[i] = fmaf(d, d, f2[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
        if (barg[i] != 0x7fffffff) {
            const unsigned long long pk = ((unsigned long long)ordered_u32(best[i]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)barg[i]);
            atomicMax(packed + (long long)b * C + cc + i, pk);
        }
    }
    if (stats) {
        double d1[V], d2[V];
        int chans[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { d1[i] = (double)f1[i]; d2[i] = (double)f2[i]; chans[i] = cc + i; }
        // the grid is 2-D: fold it so that the ticket logic of fs_stats_commit sees one linear block index
        fs_stats_commit_2d<V>(red, d1, d2, chans, m.tpr, C, stats);
        if (blockIdx.x == 0 && blockIdx.y == 0 && m.r == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i) stats[2 * C + cc + i] = (double)piv[i];
        }
    }
}

__global__ void pool_decode_kernel(const unsigned long long* __restrict__ packed, const float* __restrict__ gamma, int C,
                                   long long total, float* __restrict__ sel, int32_t* __restrict__ arg) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const unsigned long long pk = packed[e];
    const unsigned int flip = __ldg(gamma + (int)(e % C)) >= 0.f ? 0u : 0x80000000u;
    sel[e] = __uint_as_float(__float_as_uint(from_ordered_u32((unsigned int)(pk >> 32))) ^ flip);
    arg[e] = (int32_t)(0xffffffffu - (unsigned int)(pk & 0xffffffffu));
}

// dX[r, c] = scale * ( [r == arg[b,c]] * d[b,c] - dbeta/M - xhat[r,c] * dgamma/M ),  d = g * LeakyReLU'(z_sel)
// grid (row chunks, clouds): the cloud of a CTA is fixed, so the arg-max row and the routed gradient of the
// thread's channels are loaded ONCE into registers (the first revision re-read arg per element and divided by N per
// row); four independent row loads are in flight per thread.
template <typename T, typename OT>
__global__ void __launch_bounds__(DN_THREADS)
pool_bwd_kernel(const T* __restrict__ x, int ld, int N, long long rows, int C, const float* __restrict__ g,
                const float* __restrict__ sel, const int32_t* __restrict__ arg, const float* __restrict__ coef, float slope,
                const double* __restrict__ dgb, double count, int train_stats, OT* __restrict__ dx, int ld_dx) {
    constexpr int V = Vec<T>::N;
    constexpr int U = 4;
    RowMap m(C, V);
    const int cc = m.c0;
    const long long b = blockIdx.y;
    float mu[V], sc[V], mb[V], mg[V], routed[V];
    int ar[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        mu[i] = __ldg(coef + cc + i); sc[i] = __ldg(coef + 2 * C + cc + i);
        mb[i] = train_stats ? (float)(dgb[cc + i] / count) : 0.f;
        mg[i] = train_stats ? (float)(dgb[C + cc + i] / count) * __ldg(coef + C + cc + i) : 0.f;
        ar[i] = __ldg(arg + b * C + cc + i);
        const float zs = fmaf(sc[i], __ldg(sel + b * C + cc + i) - mu[i], __ldg(coef + 3 * C + cc + i));
        const float gv = __ldg(g + b * C + cc + i);
        routed[i] = zs > 0.f ? gv : slope * gv;
    }
    const int step = gridDim.x * m.rows_per_pass;
    const T* xb = x + b * N * (long long)ld + cc;
    OT* db = dx + b * N * (long long)ld_dx + cc;
    for (int r0 = blockIdx.x * m.rows_per_pass + m.r; r0 < N; r0 += U * step) {
        float f[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + u * step;
            Vec<T>::load(xb + (long long)(r < N ? r : r0) * ld, f[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = r0 + u * step;
            if (r < N) {
                float o[V];
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float rt = ar[i] == r ? routed[i] : 0.f;
                    o[i] = train_stats ? sc[i] * (rt - mb[i] - mg[i] * (f[u][i] - mu[i])) : sc[i] * rt;
                }
                OT* op = db + (long long)r * ld_dx;
                if (V == 8 && sizeof(OT) == 4) {
                    Vec<float>::store(reinterpret_cast<float*>(op), o);
                    Vec<float>::store(reinterpret_cast<float*>(op) + 4, o + 4);
                } else if (V == 4 && sizeof(OT) == 2) {
                    uint2 v;
                    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
                    h[0] = __floats2bfloat162_rn(o[0], o[1]);
                    h[1] = __floats2bfloat162_rn(o[2], o[3]);
                    *reinterpret_cast<uint2*>(op) = v;
                } else {
                    Vec<OT>::store(op, o);
                }
            }
        }
    }
}

int dn_grid(long long rows, int rows_per_pass) {
    long long need = (rows + rows_per_pass - 1) / rows_per_pass;
    const long long cap = (long long)FS_NUM_SMS * 4;     // measured in the training step: 4 per SM beats 2, 3, 6 and 8
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}
bool pow2_width(int C, int vec) { return C >= 64 && C <= 1024 && (C & (C - 1)) == 0 && C % vec == 0; }
int rows_per_pass(int C, int vec) { int tpr = C / vec; if (tpr > DN_THREADS) tpr = DN_THREADS; return DN_THREADS / tpr; }

}  // namespace

extern "C" int fs_colstats(int device, fs_stream_t stream_, const void* x, int dtype, int ld, long long rows, int C,
                           const float* rowbias, int N, double* stats) {
    if (!x || !stats || rows <= 0 || ld < C || (rowbias && N <= 0)) return FS_ERR_BAD_ARG;
    const int vec = dtype == FS_BF16 ? 8 : 4;
    if (!pow2_width(C, vec)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    int grid = dn_grid(rows, rows_per_pass(C, vec));
    // two blocks are resident per SM (96 registers): one wave, half the fp64 commits of the generic 4-per-SM sizing
    // (measured in the training step: 2.092 ms against 2.108 ms)
    if (grid > FS_NUM_SMS * 2) grid = FS_NUM_SMS * 2;
    const size_t smem = (size_t)2 * vec * DN_THREADS * sizeof(double);
    if (dtype == FS_BF16) colstats_kernel<<<grid, DN_THREADS, smem, stream>>>((const __nv_bfloat16*)x, ld, rows, C, rowbias, N, stats);
    else colstats_kernel<<<grid, DN_THREADS, smem, stream>>>((const float*)x, ld, rows, C, rowbias, N, stats);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

static int bn_act_apply_launch(int device, fs_stream_t stream_, const void* x, int dtype, int ld, long long rows, int C,
                               const float* rowbias, int N, const float* coef, const FsBnFin& fin, float slope, void* out,
                               int out_dtype, int ld_out) {
    if (!x || !out || rows <= 0 || ld < C || ld_out < C || (rowbias && N <= 0)) return FS_ERR_BAD_ARG;
    const int vec = dtype == FS_BF16 ? 8 : 4;
    if (!pow2_width(C, vec)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = dn_grid(rows, rows_per_pass(C, vec));
    const size_t smem = fin.stats ? (size_t)3 * C * sizeof(float) : 0;
#define GO(T, OT) bn_act_apply_kernel<<<grid, DN_THREADS, smem, stream>>>((const T*)x, ld, rows, C, rowbias, N, coef, fin, slope, (OT*)out, ld_out)
    if (dtype == FS_BF16 && out_dtype == FS_BF16) GO(__nv_bfloat16, __nv_bfloat16);
    else if (dtype == FS_BF16) GO(__nv_bfloat16, float);
    else if (out_dtype == FS_BF16) GO(float, __nv_bfloat16);
    else GO(float, float);
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_bn_act_apply(int device, fs_stream_t stream_, const void* x, int dtype, int ld, long long rows, int C,
                               const float* rowbias, int N, const float* coef, float slope, void* out, int out_dtype,
                               int ld_out) {
    if (!coef) return FS_ERR_BAD_ARG;
    FsBnFin fin{};
    return bn_act_apply_launch(device, stream_, x, dtype, ld, rows, C, rowbias, N, coef, fin, slope, out, out_dtype, ld_out);
}

// The same with the BatchNorm finalisation (fs_bn_finalize) folded in: coefficients from `stats`, published to coef_out,
// running statistics updated.
extern "C" int fs_bn_act_apply_fin(int device, fs_stream_t stream_, const void* x, int dtype, int ld, long long rows, int C,
                                   const float* rowbias, int N, const double* stats, double count, const float* gamma,
                                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                   long long* num_batches_tracked, float* coef_out, float slope, void* out, int out_dtype,
                                   int ld_out) {
    if (!stats || !gamma || !beta || !coef_out || count <= 0) return FS_ERR_BAD_ARG;
    FsBnFin fin{stats, count, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, coef_out};
    return bn_act_apply_launch(device, stream_, x, dtype, ld, rows, C, rowbias, N, nullptr, fin, slope, out, out_dtype, ld_out);
}

extern "C" int fs_bn_act_bwd(int device, fs_stream_t stream_, const void* g, int g_dtype, int ldg, const void* x, int dtype,
                             int ld, long long rows, int C, const float* rowbias, int N, const float* coef, float slope,
                             double* dgb, double count, int train_stats, void* dx, int dx_dtype, int ld_dx) {
    if (!g || !x || !coef || !dgb || rows <= 0 || ld < C || ldg < C || (dx && ld_dx < C) || count <= 0) return FS_ERR_BAD_ARG;
    if (!pow2_width(C, 4)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = dn_grid(rows, rows_per_pass(C, 4));
    const size_t smem = (size_t)2 * 4 * DN_THREADS * sizeof(double);
#define RED(GT, T) bn_act_bwd_reduce_kernel<<<grid, DN_THREADS, smem, stream>>>((const GT*)g, ldg, (const T*)x, ld, rows, C, rowbias, N, coef, slope, dgb)
    if (g_dtype == FS_BF16 && dtype == FS_BF16) RED(__nv_bfloat16, __nv_bfloat16);
    else if (g_dtype == FS_BF16) RED(__nv_bfloat16, float);
    else if (dtype == FS_BF16) RED(float, __nv_bfloat16);
    else RED(float, float);
#undef RED
    FS_RETURN_IF_LAUNCH_FAILED();
    if (!dx) return FS_OK;      // reduction only (pooled layers: the dense gradient is written by fs_pool_bwd)
#define APP(GT, T, OT) bn_act_bwd_apply_kernel<<<grid, DN_THREADS, 0, stream>>>((const GT*)g, ldg, (const T*)x, ld, rows, C, rowbias, N, coef, slope, dgb, count, train_stats, (OT*)dx, ld_dx)
    const bool gb = g_dtype == FS_BF16, xb = dtype == FS_BF16, ob = dx_dtype == FS_BF16;
    if (gb && xb && ob) APP(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16);
    else if (gb && xb) APP(__nv_bfloat16, __nv_bfloat16, float);
    else if (gb && ob) APP(__nv_bfloat16, float, __nv_bfloat16);
    else if (gb) APP(__nv_bfloat16, float, float);
    else if (xb && ob) APP(float, __nv_bfloat16, __nv_bfloat16);
    else if (xb) APP(float, __nv_bfloat16, float);
    else if (ob) APP(float, float, __nv_bfloat16);
    else APP(float, float, float);
#undef APP
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_pool_reduce(int device, fs_stream_t stream_, const void* x, int dtype, int ld, int B, int N, int C,
                              const float* gamma, float* sel, int32_t* arg, double* stats, unsigned long long* packed_ws) {
    if (!x || !gamma || !sel || !arg || !packed_ws || B <= 0 || N <= 0 || ld < C) return FS_ERR_BAD_ARG;
    const int vec = dtype == FS_BF16 ? 8 : 4;
    if (!pow2_width(C, vec)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    // three blocks are resident per SM (__launch_bounds__): the grid is at most ONE full wave - a grid of 2.05 waves
    // (the first sizing) spent half of its time in a nearly empty third wave
    int chunks = (3 * FS_NUM_SMS) / B;
    const int rpp = rows_per_pass(C, vec);
    if (chunks > (N + 4 * rpp - 1) / (4 * rpp)) chunks = (N + 4 * rpp - 1) / (4 * rpp);
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, B);
    const size_t smem = (size_t)2 * vec * DN_THREADS * sizeof(double);
    if (dtype == FS_BF16) pool_reduce_kernel<<<grid, DN_THREADS, smem, stream>>>((const __nv_bfloat16*)x, ld, N, C, gamma, packed_ws, stats);
    else pool_reduce_kernel<<<grid, DN_THREADS, smem, stream>>>((const float*)x, ld, N, C, gamma, packed_ws, stats);
    FS_RETURN_IF_LAUNCH_FAILED();
    const long long total = (long long)B * C;
    pool_decode_kernel<<<fs_div_up(total, 256), 256, 0, stream>>>(packed_ws, gamma, C, total, sel, arg);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_pool_bwd(int device, fs_stream_t stream_, const void* x, int dtype, int ld, int B, int N, int C,
                           const float* g, const float* sel, const int32_t* arg, const float* coef, float slope,
                           const double* dgb, double count, int train_stats, void* dx, int dx_dtype, int ld_dx) {
    if (!x || !g || !sel || !arg || !coef || !dx || B <= 0 || N <= 0 || ld < C || ld_dx < C) return FS_ERR_BAD_ARG;
    if (train_stats && (!dgb || count <= 0)) return FS_ERR_BAD_ARG;
    const int vec = dtype == FS_BF16 ? 8 : 4;
    if (!pow2_width(C, vec)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long rows = (long long)B * N;
    const int rpp = rows_per_pass(C, vec);
    int chunks = (8 * FS_NUM_SMS + B - 1) / B;           // about eight blocks per SM in total
    if (chunks > (N + 4 * rpp - 1) / (4 * rpp)) chunks = (N + 4 * rpp - 1) / (4 * rpp);
    if (chunks < 1) chunks = 1;
    if (B > 65535) return FS_ERR_UNSUPPORTED;
    const dim3 grid(chunks, B);
#define GO(T, OT) pool_bwd_kernel<<<grid, DN_THREADS, 0, stream>>>((const T*)x, ld, N, rows, C, g, sel, arg, coef, slope, dgb, count, train_stats, (OT*)dx, ld_dx)
    if (dtype == FS_BF16 && dx_dtype == FS_BF16) GO(__nv_bfloat16, __nv_bfloat16);
    else if (dtype == FS_BF16) GO(__nv_bfloat16, float);
    else if (dx_dtype == FS_BF16) GO(float, __nv_bfloat16);
    else GO(float, float);
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

// sel / arg from the packed (ordered key << 32 | ~row) table that fs_pool_reduce and fs_pool_gemm fill
extern "C" int fs_pool_decode(int device, fs_stream_t stream_, const unsigned long long* packed, const float* gamma, int B, int C,
                              float* sel, int32_t* arg) {
    if (!packed || !gamma || !sel || !arg || B <= 0 || C <= 0) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long total = (long long)B * C;
    pool_decode_kernel<<<fs_div_up(total, 256), 256, 0, stream>>>(packed, gamma, C, total, sel, arg);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
