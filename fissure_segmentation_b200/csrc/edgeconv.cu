// Fused EdgeConv (models/dgcnn.py:212-243) on point-major tables.
//
// Identity used throughout: with W = [W1 | W2] acting on [x_j - x_i, x_i],
//   W [x_j - x_i ; x_i] = W1 x_j + (W2 - W1) x_i = a_j + b_i,
// so one per-point GEMM produces the table T = [a | b] and the N x k x C edge tensor never exists.
// BatchNorm(affine) followed by LeakyReLU is monotone per channel, so max over k commutes with it:
// only max_j a_j (gamma >= 0) or min_j a_j (gamma < 0) is needed, plus the batch sums of y = a_j+b_i.
#include "fs_common.cuh"
#include "edgeconv_smem.cuh"
#include <stdlib.h>
#include <string.h>

namespace {

constexpr int EC_THREADS = 256;
constexpr int EC_WARPS = EC_THREADS / 32;

template <typename TT, int CP>
struct EcMap {
    static constexpr int VEC = FsRow<TT>::VEC;                    // channels per 128-bit load
    static constexpr int LPP = (CP / VEC) < 32 ? (CP / VEC) : 32; // lanes per point
    static constexpr int NV = CP / (VEC * LPP);                   // vectors per lane
    static constexpr int PPW = 32 / LPP;                          // points per warp
    static constexpr int NCH = NV * VEC;                          // channels per thread
    static_assert(CP % VEC == 0 && NV >= 1 && NV * VEC * LPP == CP, "unsupported channel count");
    __device__ __forceinline__ static int chan(int l, int v) { return (v * LPP + l) * VEC; }
};

// generic 4-channel loads for the elementwise kernels
__device__ __forceinline__ void load4(const float* p, float* f) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float* f) {
    uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    float2 t0 = __bfloat1622float2(h[0]), t1 = __bfloat1622float2(h[1]);
    f[0] = t0.x; f[1] = t0.y; f[2] = t1.x; f[3] = t1.y;
}
__device__ __forceinline__ void store4(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float* f) {
    uint2 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
    h[0] = __floats2bfloat162_rn(f[0], f[1]);
    h[1] = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = v;
}

// ---------------------------------------------------------------------------------------------
// Pass 1 (train) / single pass (eval): gather the k neighbour rows of a, select max/min.
// MODE 0: train gather  -> sel, arg, sy, stats
// MODE 1: eval fused    -> out (= LeakyReLU(scale*(sel+b-mu)+beta)), optional arg
//
// Per edge and channel the loop does four instructions: S1 += a_j, compare, select value, select slot.
// The batch statistics of y = a_j + b_i over all edges are NOT accumulated per edge; with the in-degree of
// every point (reverse graph) they follow from per-point quantities, all shifted by pivots pa, pb (row 0):
//     sum_e y'   = sum_i ( S1'_i + k b'_i )                         y' = y - (pa + pb), a' = a - pa, b' = b - pb
//     sum_e y'^2 = sum_i ( indeg_i a'_i^2 + k b'_i^2 + 2 b'_i S1'_i )   S1'_i = sum_{j in N(i)} a'_j
// evaluated in fp64 per point and channel (a_j ~ -b_i cancels, so the products must not round in fp32).
// Each CTA owns a contiguous range of points: with spatially sorted clouds the neighbour rows of
// consecutive points overlap and are served by L1.
// ---------------------------------------------------------------------------------------------
template <typename TT, int CP, int MODE, typename OT>
__global__ void __launch_bounds__(EC_THREADS)
edgeconv_gather_kernel(const TT* __restrict__ table, int ld, const int32_t* __restrict__ idx, long long P,
                       int N, int k, const float* __restrict__ gamma_or_coef, const int32_t* __restrict__ rev_ptr,
                       float* __restrict__ sel_out, uint8_t* __restrict__ arg_out, float* __restrict__ sy_out,
                       double* __restrict__ stats, OT* __restrict__ out, int ld_out) {
    using M = EcMap<TT, CP>;
    constexpr int VEC = M::VEC, NV = M::NV, NCH = M::NCH;
    __shared__ double red[MODE == 0 ? 2 * NCH * EC_THREADS : 1];

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane / M::LPP;
    const int l = lane - sub * M::LPP;

    uint32_t flip[NCH];      // 0x80000000 where the minimum is wanted (gamma < 0): -a is maximised instead
    float pa[NCH], pb[NCH];
    float mu[NCH], scale[NCH], beta[NCH];
    int chans[NCH];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int c0 = M::chan(l, v);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
            const int c = c0 + e;
            chans[v * VEC + e] = c;
            if (MODE == 0) {
                flip[v * VEC + e] = __ldg(gamma_or_coef + c) >= 0.f ? 0u : 0x80000000u;
            } else {
                mu[v * VEC + e] = __ldg(gamma_or_coef + c);
                scale[v * VEC + e] = __ldg(gamma_or_coef + 2 * CP + c);
                beta[v * VEC + e] = __ldg(gamma_or_coef + 3 * CP + c);
                flip[v * VEC + e] = scale[v * VEC + e] >= 0.f ? 0u : 0x80000000u;
            }
        }
        if (MODE == 0 && stats) {
            FsRow<TT>::load(table + c0, pa + v * VEC);
            FsRow<TT>::load(table + CP + c0, pb + v * VEC);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) { pa[v * VEC + e] = 0.f; pb[v * VEC + e] = 0.f; }
        }
    }
    (void)mu; (void)scale; (void)beta;

    double s1[NCH], s2[NCH];
#pragma unroll
    for (int e = 0; e < NCH; ++e) { s1[e] = 0.0; s2[e] = 0.0; }

    // contiguous point range of this CTA, interleaved over its warps / point slots
    const long long per_cta = (P + gridDim.x - 1) / gridDim.x;
    const long long p_begin = (long long)blockIdx.x * per_cta;
    const long long p_end = p_begin + per_cta < P ? p_begin + per_cta : P;
    const float kf = (float)k;
    for (long long pt = p_begin + warp * M::PPW + sub; pt < p_end; pt += EC_WARPS * M::PPW) {
        const long long cloud0 = (pt / N) * N;
        const int32_t* irow = idx + pt * k;
        // keys: a with the sign bit flipped on channels that take the minimum (gamma < 0), so the loop is a
        // pure arg-max; four neighbours per step are reduced as a tree (one dependent step on `best` per chunk)
        float best[NCH], S1[NCH];
        int barg[NCH];
#pragma unroll
        for (int e = 0; e < NCH; ++e) { best[e] = -INFINITY; barg[e] = 0; S1[e] = 0.f; }
        const int klast = __ldg(irow + k - 1);
        for (int t0 = 0; t0 < k; t0 += 4) {
            int jn[4];
            float w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool in = t0 + u < k;
                jn[u] = in ? __ldg(irow + t0 + u) : klast;      // tail: repeat the last neighbour (max unchanged),
                w[u] = in ? 1.f : 0.f;                          //       weight 0 in the sum
            }
            float a[4][NCH];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    FsRow<TT>::load(table + (cloud0 + jn[u]) * ld + M::chan(l, v), a[u] + v * VEC);
            }
#pragma unroll
            for (int e = 0; e < NCH; ++e) {
                if (MODE == 0) S1[e] += fmaf(w[3], a[3][e], fmaf(w[2], a[2][e], fmaf(w[1], a[1][e], w[0] * a[0][e])));
                const float k0 = __uint_as_float(__float_as_uint(a[0][e]) ^ flip[e]);
                const float k1 = __uint_as_float(__float_as_uint(a[1][e]) ^ flip[e]);
                const float k2 = __uint_as_float(__float_as_uint(a[2][e]) ^ flip[e]);
                const float k3 = __uint_as_float(__float_as_uint(a[3][e]) ^ flip[e]);
                const bool p01 = k1 > k0, p23 = k3 > k2;
                const float m01 = fmaxf(k0, k1), m23 = fmaxf(k2, k3);
                const bool pm = m23 > m01;
                const float m = fmaxf(m01, m23);
                const int am = pm ? (p23 ? 3 : 2) : (p01 ? 1 : 0);
                const bool better = m > best[e];
                best[e] = better ? m : best[e];
                barg[e] = better ? t0 + am : barg[e];
            }
        }
#pragma unroll
        for (int e = 0; e < NCH; ++e) best[e] = __uint_as_float(__float_as_uint(best[e]) ^ flip[e]);
        float bi[NCH];
#pragma unroll
        for (int v = 0; v < NV; ++v) FsRow<TT>::load(table + pt * ld + CP + M::chan(l, v), bi + v * VEC);
        if (MODE == 0) {
            float sy[NCH];
#pragma unroll
            for (int e = 0; e < NCH; ++e) sy[e] = fmaf(kf, bi[e], S1[e]);
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c0 = M::chan(l, v);
#pragma unroll
                for (int h = 0; h < VEC; h += 4) {
                    FsRow<float>::store(sel_out + pt * CP + c0 + h, best + v * VEC + h);
                    if (sy_out) FsRow<float>::store(sy_out + pt * CP + c0 + h, sy + v * VEC + h);
                }
            }
            if (stats) {
                float ai[NCH];
#pragma unroll
                for (int v = 0; v < NV; ++v) FsRow<TT>::load(table + pt * ld + M::chan(l, v), ai + v * VEC);
                const double deg = (double)(__ldg(rev_ptr + pt + 1) - __ldg(rev_ptr + pt));
                const double kd = (double)k;
#pragma unroll
                for (int e = 0; e < NCH; ++e) {
                    const double ap = (double)ai[e] - (double)pa[e];
                    const double bp = (double)bi[e] - (double)pb[e];
                    const double S1p = (double)S1[e] - kd * (double)pa[e];
                    s1[e] += S1p + kd * bp;
                    s2[e] += deg * ap * ap + bp * (kd * bp + 2.0 * S1p);
                }
            }
        } else {
            float o[NCH];
#pragma unroll
            for (int e = 0; e < NCH; ++e) o[e] = fs_leaky(fmaf(scale[e], (best[e] + bi[e]) - mu[e], beta[e]));
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c0 = M::chan(l, v);
#pragma unroll
                for (int h = 0; h < VEC; h += 4) store4(out + pt * ld_out + c0 + h, o + v * VEC + h);
            }
        }
        if (arg_out) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int c0 = M::chan(l, v);
#pragma unroll
                for (int h = 0; h < VEC; h += 4) {
                    const int* ba = barg + v * VEC + h;
                    uint32_t pk = (uint32_t)ba[0] | ((uint32_t)ba[1] << 8) | ((uint32_t)ba[2] << 16) | ((uint32_t)ba[3] << 24);
                    *reinterpret_cast<uint32_t*>(arg_out + pt * CP + c0 + h) = pk;
                }
            }
        }
    }
    if (MODE == 0 && stats) {
        fs_stats_commit<NCH>(red, s1, s2, chans, M::LPP, CP, stats);
        if (blockIdx.x == 0 && warp == 0 && sub == 0) {
#pragma unroll
            for (int e = 0; e < NCH; ++e) stats[2 * CP + chans[e]] = (double)pa[e] + (double)pb[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm coefficient kernels. coef = [mu | invstd | scale | beta], each Cp floats.
// ---------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, int Cp,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, float* __restrict__ coef, float* running_mean,
                                   float* running_var, long long* nbt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= Cp) return;
    const double m1 = stats[c] / count;
    double var = stats[Cp + c] / count - m1 * m1;
    if (var < 0.0) var = 0.0;
    const double mean = m1 + stats[2 * Cp + c];
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    coef[c] = (float)mean;
    coef[Cp + c] = invstd;
    coef[2 * Cp + c] = gamma[c] * invstd;
    coef[3 * Cp + c] = beta[c];
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__global__ void bn_coef_eval_kernel(int Cp, const float* __restrict__ gamma, const float* __restrict__ beta,
                                    const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                    float* __restrict__ coef) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    const float invstd = 1.0f / sqrtf(rv[c] + eps);
    coef[c] = rm[c];
    coef[Cp + c] = invstd;
    coef[2 * Cp + c] = gamma[c] * invstd;
    coef[3 * Cp + c] = beta[c];
}

// ---------------------------------------------------------------------------------------------
// Pass 2: out = LeakyReLU(scale*((sel + b) - mu) + beta). One thread per 4 channels.
// ---------------------------------------------------------------------------------------------
template <typename TT, typename OT>
__global__ void __launch_bounds__(256)
edgeconv_apply_kernel(const float* __restrict__ sel, const TT* __restrict__ table, int ld, long long P, int Cp,
                      const float* __restrict__ coef, const FsBnFin fin, OT* __restrict__ out, int ld_out) {
    extern __shared__ float ea_coef[];          // [3][Cp]: mean | scale | beta (from `coef`, or from the statistics)
    for (int c = threadIdx.x; c < Cp; c += blockDim.x) {
        float mu, inv, sc, be;
        if (fin.stats) {
            fs_bn_fin_channel(fin, c, Cp, blockIdx.x == 0, mu, inv, sc, be);
        } else {
            mu = __ldg(coef + c); sc = __ldg(coef + 2 * Cp + c); be = __ldg(coef + 3 * Cp + c);
        }
        ea_coef[c] = mu; ea_coef[Cp + c] = sc; ea_coef[2 * Cp + c] = be;
    }
    __syncthreads();
    const int q = Cp >> 2;
    const long long total = P * q;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long pt = e / q;
        const int c0 = (int)(e - pt * q) << 2;
        float s[4], b[4], o[4];
        load4(sel + pt * Cp + c0, s);
        if (table) load4(table + pt * ld + Cp + c0, b);
        else { b[0] = b[1] = b[2] = b[3] = 0.f; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float mu = ea_coef[c0 + i], sc = ea_coef[Cp + c0 + i], be = ea_coef[2 * Cp + c0 + i];
            o[i] = fs_leaky(fmaf(sc, (s[i] + b[i]) - mu, be));
        }
        store4(out + pt * ld_out + c0, o);
    }
}

// ---------------------------------------------------------------------------------------------
// Reverse graph: counting sort of the edges by target, three grid-wide passes over a zeroed rev_ptr:
//   (1) histogram: rev_ptr[J + 1] += 1 for every edge with global target J;
//   (2) per-cloud exclusive scan, in place: rev_ptr[J + 1] = first slot of J (the fill cursor of J);
//   (3) fill: slot = atomicAdd(&rev_ptr[J + 1], 1) - when every edge is placed the cursor of J has advanced to the first
//       slot of J + 1, which is exactly the value rev_ptr[J + 1] must hold. rev_ptr[0] stays 0.
// (The first version sorted each cloud inside ONE CTA with shared-memory atomics: 27 us at B = 32, N = 2048, k = 20 but
// 242 us for a single cloud of 8192 points with k = 40.) The order of the sources inside a target's list depends on the
// atomics' arrival order.
// ---------------------------------------------------------------------------------------------
constexpr int RG_THREADS = 1024;

__global__ void __launch_bounds__(256)
reverse_hist_kernel(const int32_t* __restrict__ idx, int N, int k, long long edges, int32_t* __restrict__ rev_ptr) {
    const long long ek = (long long)N * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < edges; e += (long long)gridDim.x * blockDim.x) {
        const int t = __ldg(idx + e);
        if ((unsigned)t < (unsigned)N) atomicAdd(rev_ptr + (e / ek) * N + t + 1, 1);
    }
}

__global__ void __launch_bounds__(RG_THREADS)
reverse_scan_kernel(int N, int k, int32_t* __restrict__ rev_ptr) {
    __shared__ int wsum[32];
    const int b = blockIdx.x;
    int32_t* cnt = rev_ptr + (long long)b * N + 1;          // counts of this cloud's targets, then their cursors
    const long long e0 = (long long)b * N * k;
    // exclusive scan of cnt[0..N): each thread owns a contiguous run
    const int per = (N + RG_THREADS - 1) / RG_THREADS;
    const int j0 = threadIdx.x * per;
    int local = 0;
    for (int j = j0; j < min(j0 + per, N); ++j) local += cnt[j];
    int incl = local;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(FS_FULL_MASK, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int v = __shfl_up_sync(FS_FULL_MASK, wi, o);
            if (lane >= o) wi += v;
        }
        wsum[lane] = wi - w;
    }
    __syncthreads();
    int run = wsum[warp] + incl - local;
    for (int j = j0; j < min(j0 + per, N); ++j) {
        const int c = cnt[j];
        cnt[j] = (int32_t)(e0 + run);
        run += c;
    }
}

__global__ void __launch_bounds__(256)
reverse_fill_kernel(const int32_t* __restrict__ idx, int N, int k, long long edges, int32_t* __restrict__ rev_ptr,
                    int32_t* __restrict__ rev_src) {
    const long long ek = (long long)N * k;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < edges; e += (long long)gridDim.x * blockDim.x) {
        const int t = __ldg(idx + e);
        if ((unsigned)t < (unsigned)N) {
            const int pos = atomicAdd(rev_ptr + (e / ek) * N + t + 1, 1);
            rev_src[pos] = (int32_t)(e / k);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Backward step 1: d = g * LeakyReLU'(z); dbeta, dgamma (fp64 accumulation).
// ---------------------------------------------------------------------------------------------
template <typename GT, typename TT, int CP>
__global__ void __launch_bounds__(256)
edgeconv_bwd_reduce_kernel(const GT* __restrict__ g, int ldg, const float* __restrict__ sel,
                           const TT* __restrict__ table, int ld, long long P, const float* __restrict__ coef,
                           float* __restrict__ d, double* __restrict__ dgb) {
    constexpr int Q = CP / 4;              // threads per point row
    constexpr int ROWS = 256 / Q;          // rows per block pass
    static_assert(256 % Q == 0, "CP must divide 1024");
    __shared__ double red[2 * 4 * 256];
    const int l = threadIdx.x % Q;
    const int r = threadIdx.x / Q;
    const int c0 = l * 4;
    float mu[4], inv[4], sc[4], be[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mu[i] = __ldg(coef + c0 + i); inv[i] = __ldg(coef + CP + c0 + i);
        sc[i] = __ldg(coef + 2 * CP + c0 + i); be[i] = __ldg(coef + 3 * CP + c0 + i);
    }
    double s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    for (long long pt = (long long)blockIdx.x * ROWS + r; pt < P; pt += (long long)gridDim.x * ROWS) {
        float gv[4], sv[4], bv[4], dv[4];
        load4(g + pt * ldg + c0, gv);
        load4(sel + pt * CP + c0, sv);
        if (table) load4(table + pt * ld + CP + c0, bv);
        else { bv[0] = bv[1] = bv[2] = bv[3] = 0.f; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float yc = (sv[i] + bv[i]) - mu[i];
            const float z = fmaf(sc[i], yc, be[i]);
            dv[i] = z > 0.f ? gv[i] : 0.2f * gv[i];
            s1[i] += (double)dv[i];
            s2[i] += (double)(dv[i] * (yc * inv[i]));
        }
        store4(d + pt * CP + c0, dv);
    }
    int chans[4] = {c0, c0 + 1, c0 + 2, c0 + 3};
    fs_stats_commit<4>(red, s1, s2, chans, Q, CP, dgb);
}

// ---------------------------------------------------------------------------------------------
// Backward step 2: BatchNorm-coupled per-point gradient (writes every element of dT).
// ---------------------------------------------------------------------------------------------
template <typename TT, int CP>
__global__ void __launch_bounds__(256)
edgeconv_bwd_point_kernel(const float* __restrict__ d, const float* __restrict__ sy, const TT* __restrict__ table,
                          int ld, const int32_t* __restrict__ rev_ptr, const int32_t* __restrict__ rev_src,
                          long long P, int k, const float* __restrict__ coef, const double* __restrict__ dgb,
                          double count, int train_stats, float* __restrict__ dT, float* __restrict__ dgamma_dbeta) {
    constexpr int Q = CP / 4;
    constexpr int ROWS = 256 / Q;
    const int l = threadIdx.x % Q;
    const int r = threadIdx.x / Q;
    const int c0 = l * 4;
    float mu[4], inv[4], sc[4], mb[4], mg[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        mu[i] = __ldg(coef + c0 + i); inv[i] = __ldg(coef + CP + c0 + i); sc[i] = __ldg(coef + 2 * CP + c0 + i);
        mb[i] = train_stats ? (float)(dgb[c0 + i] / count) : 0.f;                 // dbeta / M
        mg[i] = train_stats ? (float)(dgb[CP + c0 + i] / count) * inv[i] : 0.f;   // dgamma / M * invstd
    }
    if (dgamma_dbeta && blockIdx.x == 0 && r == 0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            dgamma_dbeta[c0 + i] = (float)dgb[CP + c0 + i];
            dgamma_dbeta[CP + c0 + i] = (float)dgb[c0 + i];
        }
    }
    for (long long pt = (long long)blockIdx.x * ROWS + r; pt < P; pt += (long long)gridDim.x * ROWS) {
        float dv[4], da[4], db[4];
        load4(d + pt * CP + c0, dv);
        if (train_stats) {
            float syv[4], av[4], R[4] = {0.f, 0.f, 0.f, 0.f};
            load4(sy + pt * CP + c0, syv);
            load4(table + pt * ld + c0, av);
            const int e0 = __ldg(rev_ptr + pt), e1 = __ldg(rev_ptr + pt + 1);
            for (int e = e0; e < e1; ++e) {
                const long long src = __ldg(rev_src + e);
                float bv[4];
                load4(table + src * ld + CP + c0, bv);
#pragma unroll
                for (int i = 0; i < 4; ++i) R[i] += bv[i];
            }
            const float indeg = (float)(e1 - e0);
            const float kf = (float)k;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                db[i] = sc[i] * (dv[i] - kf * mb[i] - mg[i] * (syv[i] - kf * mu[i]));
                da[i] = -sc[i] * (indeg * mb[i] + mg[i] * (indeg * (av[i] - mu[i]) + R[i]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { db[i] = sc[i] * dv[i]; da[i] = 0.f; }
        }
        store4(dT + pt * (2 * CP) + c0, da);
        store4(dT + pt * (2 * CP) + CP + c0, db);
    }
}

// ---------------------------------------------------------------------------------------------
// Backward step 3: argmax-routed scatter into the a-half of dT.
// ---------------------------------------------------------------------------------------------
template <int CP>
__global__ void __launch_bounds__(256)
edgeconv_bwd_route_kernel(const float* __restrict__ d, const uint8_t* __restrict__ arg, const int32_t* __restrict__ idx,
                          long long P, int N, int k, const float* __restrict__ coef, float* __restrict__ dT) {
    constexpr int Q = CP / 4;
    constexpr int ROWS = 256 / Q;
    const int l = threadIdx.x % Q;
    const int r = threadIdx.x / Q;
    const int c0 = l * 4;
    float sc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sc[i] = __ldg(coef + 2 * CP + c0 + i);
    for (long long pt = (long long)blockIdx.x * ROWS + r; pt < P; pt += (long long)gridDim.x * ROWS) {
        const long long cloud0 = (pt / N) * N;
        float dv[4];
        load4(d + pt * CP + c0, dv);
        const uint32_t pk = __ldg(reinterpret_cast<const uint32_t*>(arg + pt * CP + c0));
        int j[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) j[i] = __ldg(idx + pt * k + ((pk >> (8 * i)) & 0xff));
        if (j[0] == j[1] && j[1] == j[2] && j[2] == j[3]) {
            // all four channels route to the same neighbour row: one 128-bit vector reduction
            float* p = dT + (cloud0 + j[0]) * (2 * CP) + c0;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(sc[0] * dv[0]),
                         "f"(sc[1] * dv[1]), "f"(sc[2] * dv[2]), "f"(sc[3] * dv[3])
                         : "memory");
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) atomicAdd(dT + (cloud0 + j[i]) * (2 * CP) + c0 + i, sc[i] * dv[i]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Materialised edge tensors (two-layer EdgeConv, first revision).
// ---------------------------------------------------------------------------------------------
template <typename TT, typename YT, int CP>
__global__ void __launch_bounds__(256)
edge_build_kernel(const TT* __restrict__ table, int ld, const int32_t* __restrict__ idx, long long P, int N, int k,
                  YT* __restrict__ y) {
    constexpr int Q = CP / 4;
    constexpr int ROWS = 256 / Q;
    const int l = threadIdx.x % Q;
    const int r = threadIdx.x / Q;
    const int c0 = l * 4;
    for (long long pt = (long long)blockIdx.x * ROWS + r; pt < P; pt += (long long)gridDim.x * ROWS) {
        const long long cloud0 = (pt / N) * N;
        float bv[4];
        load4(table + pt * ld + CP + c0, bv);
        for (int t = 0; t < k; ++t) {
            const int j = __ldg(idx + pt * k + t);
            float av[4], o[4];
            load4(table + (cloud0 + j) * ld + c0, av);
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = av[i] + bv[i];
            store4(y + (pt * k + t) * CP + c0, o);
        }
    }
}

template <typename YT, int CP>
__global__ void __launch_bounds__(256)
edge_build_bwd_kernel(const YT* __restrict__ dy, const int32_t* __restrict__ idx, long long P, int N, int k,
                      float* __restrict__ dT) {
    constexpr int Q = CP / 4;
    constexpr int ROWS = 256 / Q;
    const int l = threadIdx.x % Q;
    const int r = threadIdx.x / Q;
    const int c0 = l * 4;
    for (long long pt = (long long)blockIdx.x * ROWS + r; pt < P; pt += (long long)gridDim.x * ROWS) {
        const long long cloud0 = (pt / N) * N;
        float db[4] = {0.f, 0.f, 0.f, 0.f};
        for (int t = 0; t < k; ++t) {
            const int j = __ldg(idx + pt * k + t);
            float g[4];
            load4(dy + (pt * k + t) * CP + c0, g);
#pragma unroll
            for (int i = 0; i < 4; ++i) db[i] += g[i];
            float* p = dT + (cloud0 + j) * (2 * CP) + c0;
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(g[0]), "f"(g[1]), "f"(g[2]),
                         "f"(g[3])
                         : "memory");
        }
        float* pb = dT + pt * (2 * CP) + CP + c0;
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(pb), "f"(db[0]), "f"(db[1]), "f"(db[2]),
                     "f"(db[3])
                     : "memory");
    }
}

// 128-bit row segments: VEC = 4 channels (fp32) or 8 channels (bf16) per thread, four rows in flight per thread.
template <typename ZT, int CP>
__global__ void __launch_bounds__(256)
edge_reduce_kernel(const ZT* __restrict__ z, long long P, int k, const float* __restrict__ gamma,
                   float* __restrict__ sel, uint8_t* __restrict__ arg, float* __restrict__ sy, double* __restrict__ stats) {
    constexpr int V = FsRow<ZT>::VEC;
    constexpr int Q = CP / V;              // threads per point
    constexpr int ROWS = 256 / Q;
    __shared__ double red[2 * V * 256];
    const int l = threadIdx.x % Q;
    const int r = threadIdx.x / Q;
    const int c0 = l * V;
    uint32_t flip[V];
    float pivot[V];
    {
        float z0[V];
        FsRow<ZT>::load(z + c0, z0);
#pragma unroll
        for (int i = 0; i < V; ++i) { flip[i] = __ldg(gamma + c0 + i) >= 0.f ? 0u : 0x80000000u; pivot[i] = stats ? z0[i] : 0.f; }
    }
    float t1[V], t2[V];
#pragma unroll
    for (int i = 0; i < V; ++i) { t1[i] = 0.f; t2[i] = 0.f; }
    for (long long pt = (long long)blockIdx.x * ROWS + r; pt < P; pt += (long long)gridDim.x * ROWS) {
        float best[V], ysum[V], f1[V], f2[V];
        int barg[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { best[i] = -INFINITY; barg[i] = 0; ysum[i] = 0.f; f1[i] = 0.f; f2[i] = 0.f; }
        const ZT* zp = z + pt * k * CP + c0;
        for (int t0 = 0; t0 < k; t0 += 4) {
            float zv[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) FsRow<ZT>::load(zp + (long long)(t0 + u < k ? t0 + u : k - 1) * CP, zv[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (t0 + u < k) {
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float key = __uint_as_float(__float_as_uint(zv[u][i]) ^ flip[i]);
                        const bool better = key > best[i];
                        best[i] = better ? key : best[i];
                        barg[i] = better ? t0 + u : barg[i];
                        ysum[i] += zv[u][i];
                        const float ys = zv[u][i] - pivot[i];
                        f1[i] += ys;
                        f2[i] = fmaf(ys, ys, f2[i]);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) { best[i] = __uint_as_float(__float_as_uint(best[i]) ^ flip[i]); t1[i] += f1[i]; t2[i] += f2[i]; }
#pragma unroll
        for (int h = 0; h < V; h += 4) {
            store4(sel + pt * CP + c0 + h, best + h);
            if (sy) store4(sy + pt * CP + c0 + h, ysum + h);
            *reinterpret_cast<uint32_t*>(arg + pt * CP + c0 + h) =
                (uint32_t)barg[h] | ((uint32_t)barg[h + 1] << 8) | ((uint32_t)barg[h + 2] << 16) | ((uint32_t)barg[h + 3] << 24);
        }
    }
    if (stats) {
        double d1[V], d2[V];
        int chans[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { d1[i] = (double)t1[i]; d2[i] = (double)t2[i]; chans[i] = c0 + i; }
        fs_stats_commit<V>(red, d1, d2, chans, Q, CP, stats);
        if (blockIdx.x == 0 && r == 0) {
#pragma unroll
            for (int i = 0; i < V; ++i) stats[2 * CP + c0 + i] = (double)pivot[i];
        }
    }
}

template <typename ZT, typename DT, int CP>
__global__ void __launch_bounds__(256)
edge_reduce_bwd_kernel(const ZT* __restrict__ z, const float* __restrict__ d, const uint8_t* __restrict__ arg,
                       long long P, int k, const float* __restrict__ coef, const double* __restrict__ dgb, double count,
                       int train_stats, DT* __restrict__ dz) {
    constexpr int V = FsRow<ZT>::VEC;
    constexpr int Q = CP / V;
    constexpr int ROWS = 256 / Q;
    const int l = threadIdx.x % Q;
    const int r = threadIdx.x / Q;
    const int c0 = l * V;
    float mu[V], sc[V], mb[V], mg[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        mu[i] = __ldg(coef + c0 + i); sc[i] = __ldg(coef + 2 * CP + c0 + i);
        mb[i] = train_stats ? (float)(dgb[c0 + i] / count) : 0.f;
        mg[i] = train_stats ? (float)(dgb[CP + c0 + i] / count) * __ldg(coef + CP + c0 + i) : 0.f;
    }
    for (long long pt = (long long)blockIdx.x * ROWS + r; pt < P; pt += (long long)gridDim.x * ROWS) {
        float dv[V];
        int ar[V];
#pragma unroll
        for (int h = 0; h < V; h += 4) {
            load4(d + pt * CP + c0 + h, dv + h);
            const uint32_t pk = __ldg(reinterpret_cast<const uint32_t*>(arg + pt * CP + c0 + h));
#pragma unroll
            for (int i = 0; i < 4; ++i) ar[h + i] = (int)((pk >> (8 * i)) & 0xff);
        }
        for (int t0 = 0; t0 < k; t0 += 2) {
            float zv[2][V];
            if (train_stats) {
#pragma unroll
                for (int u = 0; u < 2; ++u) FsRow<ZT>::load(z + (pt * k + (t0 + u < k ? t0 + u : k - 1)) * CP + c0, zv[u]);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (t0 + u < k) {
                    float o[V];
#pragma unroll
                    for (int i = 0; i < V; ++i) {
                        const float routed = ar[i] == t0 + u ? dv[i] : 0.f;
                        o[i] = train_stats ? sc[i] * (routed - mb[i] - mg[i] * (zv[u][i] - mu[i])) : sc[i] * routed;
                    }
                    DT* op = dz + (pt * k + t0 + u) * CP + c0;
                    if (sizeof(DT) == 2 && V == 8) FsRow<__nv_bfloat16>::store(reinterpret_cast<__nv_bfloat16*>(op), o);
                    else {
#pragma unroll
                        for (int h = 0; h < V; h += 4) store4(op + h, o + h);
                    }
                }
            }
        }
    }
}

// contiguous point ranges: a multiple of the SM count, at least one pass of points per CTA
int gather_grid(long long P, int points_per_pass) {
    long long need = (P + points_per_pass - 1) / points_per_pass;
    const long long cap = (long long)FS_NUM_SMS * 4;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

int ec_grid(long long rows, int rows_per_block) {
    long long need = (rows + rows_per_block - 1) / rows_per_block;
    long long cap = (long long)FS_NUM_SMS * 8;     // measured in the training step: 8 blocks per SM beats 2, 4, 16 and 32
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
#define EC_CHECK_COMMON(table, ld, Cp)                                                    \
    if (!(table)) return FS_ERR_BAD_ARG;                                                  \
    if ((Cp) != 64 && (Cp) != 128 && (Cp) != 256) return FS_ERR_UNSUPPORTED;               \
    if ((ld) < 2 * (Cp)) return FS_ERR_BAD_ARG

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// FS_GATHER=global forces the global-memory gather kernel (A/B measurements); default: shared-memory-resident
// slices whenever the shape fits (edgeconv_smem.cu).
static bool use_smem_gather() {
    const char* e = getenv("FS_GATHER");     // read per call: tests and benchmarks toggle it at run time
    return !(e && strcmp(e, "global") == 0);
}

#define EC_DISPATCH_CP(Cp, MACRO) \
    switch (Cp) {                 \
        case 64: MACRO(64); break; \
        case 128: MACRO(128); break; \
        case 256: MACRO(256); break; \
        default: return FS_ERR_UNSUPPORTED; \
    }

extern "C" int fs_edgeconv_gather(int device, fs_stream_t stream_, const void* table, int dtype, int ld,
                                  const int32_t* idx, int B, int N, int k, int Cp, const float* gamma,
                                  const int32_t* rev_ptr, float* sel, uint8_t* arg, float* sy, double* stats) {
    EC_CHECK_COMMON(table, ld, Cp);
    if (!idx || !gamma || !sel || !arg || B < 0 || N <= 0 || k <= 0 || k > 255) return FS_ERR_BAD_ARG;
    if (stats && !rev_ptr) return FS_ERR_BAD_ARG;
    const int esz = dtype == FS_BF16 ? 2 : 4;
    if (!aligned16(table) || (ld * esz) % 16 != 0 || !aligned16(sel)) return FS_ERR_ALIGNMENT;
    if (B == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    if (dtype == FS_F32 && use_smem_gather() && aligned16(sel) && (!sy || aligned16(sy))) {
        const int rc = fs_gather_smem_train(stream, (const float*)table, ld, idx, B, N, k, Cp, gamma, rev_ptr, sel, arg,
                                            sy, stats);
        if (rc != FS_SMEM_GATHER_UNSUPPORTED) return rc;
    }
#define GO(CP)                                                                                                  \
    if (dtype == FS_BF16) {                                                                                     \
        using M = EcMap<__nv_bfloat16, CP>;                                                                     \
        edgeconv_gather_kernel<__nv_bfloat16, CP, 0, float><<<gather_grid(P, EC_WARPS * M::PPW), EC_THREADS, 0, stream>>>( \
            (const __nv_bfloat16*)table, ld, idx, P, N, k, gamma, rev_ptr, sel, arg, sy, stats, nullptr, 0);             \
    } else {                                                                                                    \
        using M = EcMap<float, CP>;                                                                             \
        edgeconv_gather_kernel<float, CP, 0, float><<<gather_grid(P, EC_WARPS * M::PPW), EC_THREADS, 0, stream>>>(  \
            (const float*)table, ld, idx, P, N, k, gamma, rev_ptr, sel, arg, sy, stats, nullptr, 0);                     \
    }
    EC_DISPATCH_CP(Cp, GO)
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edgeconv_fused_eval(int device, fs_stream_t stream_, const void* table, int dtype, int ld,
                                      const int32_t* idx, int B, int N, int k, int Cp, const float* coef, void* out,
                                      int out_dtype, int ld_out, uint8_t* arg) {
    EC_CHECK_COMMON(table, ld, Cp);
    if (!idx || !coef || !out || B < 0 || N <= 0 || k <= 0 || k > 255 || ld_out < Cp) return FS_ERR_BAD_ARG;
    const int esz = dtype == FS_BF16 ? 2 : 4, osz = out_dtype == FS_BF16 ? 2 : 4;
    if (!aligned16(table) || (ld * esz) % 16 != 0 || (reinterpret_cast<uintptr_t>(out) % (4 * osz)) != 0 ||
        (ld_out * osz) % (4 * osz) != 0)
        return FS_ERR_ALIGNMENT;
    if (B == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    if (dtype == FS_F32 && use_smem_gather()) {
        const int rc = fs_gather_smem_eval(stream, (const float*)table, ld, idx, B, N, k, Cp, coef, out,
                                           out_dtype == FS_BF16, ld_out, arg);
        if (rc != FS_SMEM_GATHER_UNSUPPORTED) return rc;
    }
#define GO2(CP, TT, OT)                                                                                         \
    edgeconv_gather_kernel<TT, CP, 1, OT><<<gather_grid(P, EC_WARPS * EcMap<TT, CP>::PPW), EC_THREADS, 0, stream>>>( \
        (const TT*)table, ld, idx, P, N, k, coef, nullptr, nullptr, arg, nullptr, nullptr, (OT*)out, ld_out)
#define GO(CP)                                                                      \
    if (dtype == FS_BF16 && out_dtype == FS_BF16) GO2(CP, __nv_bfloat16, __nv_bfloat16); \
    else if (dtype == FS_BF16) GO2(CP, __nv_bfloat16, float);                       \
    else if (out_dtype == FS_BF16) GO2(CP, float, __nv_bfloat16);                   \
    else GO2(CP, float, float);
    EC_DISPATCH_CP(Cp, GO)
#undef GO
#undef GO2
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_bn_finalize(int device, fs_stream_t stream_, const double* stats, double count, int Cp,
                              const float* gamma, const float* beta, float eps, float momentum, float* coef,
                              float* running_mean, float* running_var, long long* num_batches_tracked) {
    if (!stats || !gamma || !beta || !coef || Cp <= 0 || count <= 0) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    bn_finalize_kernel<<<fs_div_up(Cp, 128), 128, 0, (cudaStream_t)stream_>>>(stats, count, Cp, gamma, beta, eps, momentum,
                                                                            coef, running_mean, running_var,
                                                                            num_batches_tracked);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_bn_coef_eval(int device, fs_stream_t stream_, int Cp, const float* gamma, const float* beta,
                               const float* running_mean, const float* running_var, float eps, float* coef) {
    if (!gamma || !beta || !running_mean || !running_var || !coef || Cp <= 0) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    bn_coef_eval_kernel<<<fs_div_up(Cp, 128), 128, 0, (cudaStream_t)stream_>>>(Cp, gamma, beta, running_mean, running_var,
                                                                             eps, coef);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

static int edgeconv_apply_launch(int device, fs_stream_t stream_, const float* sel, const void* table, int dtype, int ld,
                                 long long P, int Cp, const float* coef, const FsBnFin& fin, void* out, int out_dtype,
                                 int ld_out) {
    if (!sel || !out || P < 0 || Cp <= 0 || Cp % 4 || Cp > 2048 || (table && ld < 2 * Cp) || ld_out < Cp) return FS_ERR_BAD_ARG;
    if (P == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = ec_grid(P * (Cp / 4), 256);
    const size_t smem = (size_t)3 * Cp * sizeof(float);
    if (dtype == FS_BF16 && out_dtype == FS_BF16)
        edgeconv_apply_kernel<<<grid, 256, smem, stream>>>(sel, (const __nv_bfloat16*)table, ld, P, Cp, coef, fin, (__nv_bfloat16*)out, ld_out);
    else if (dtype == FS_BF16)
        edgeconv_apply_kernel<<<grid, 256, smem, stream>>>(sel, (const __nv_bfloat16*)table, ld, P, Cp, coef, fin, (float*)out, ld_out);
    else if (out_dtype == FS_BF16)
        edgeconv_apply_kernel<<<grid, 256, smem, stream>>>(sel, (const float*)table, ld, P, Cp, coef, fin, (__nv_bfloat16*)out, ld_out);
    else
        edgeconv_apply_kernel<<<grid, 256, smem, stream>>>(sel, (const float*)table, ld, P, Cp, coef, fin, (float*)out, ld_out);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edgeconv_apply(int device, fs_stream_t stream_, const float* sel, const void* table, int dtype,
                                 int ld, long long P, int Cp, const float* coef, void* out, int out_dtype, int ld_out) {
    if (!coef) return FS_ERR_BAD_ARG;
    FsBnFin fin{};
    return edgeconv_apply_launch(device, stream_, sel, table, dtype, ld, P, Cp, coef, fin, out, out_dtype, ld_out);
}

// The same with fs_bn_finalize folded in (coefficients from `stats`, published to coef_out, running statistics updated).
extern "C" int fs_edgeconv_apply_fin(int device, fs_stream_t stream_, const float* sel, const void* table, int dtype, int ld,
                                     long long P, int Cp, const double* stats, double count, const float* gamma,
                                     const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, float* coef_out, void* out, int out_dtype, int ld_out) {
    if (!stats || !gamma || !beta || !coef_out || count <= 0) return FS_ERR_BAD_ARG;
    FsBnFin fin{stats, count, gamma, beta, eps, momentum, running_mean, running_var, num_batches_tracked, coef_out};
    return edgeconv_apply_launch(device, stream_, sel, table, dtype, ld, P, Cp, nullptr, fin, out, out_dtype, ld_out);
}

extern "C" int fs_reverse_graph(int device, fs_stream_t stream_, const int32_t* idx, int B, int N, int k,
                                int32_t* rev_ptr, int32_t* rev_src) {
    if (!idx || !rev_ptr || !rev_src || B < 0 || N <= 0 || k <= 0) return FS_ERR_BAD_ARG;
    if ((long long)B * N * k >= 0x7fffffffLL) return FS_ERR_UNSUPPORTED;
    if (B == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N, edges = P * k;
    FS_CUDA_TRY(cudaMemsetAsync(rev_ptr, 0, (size_t)(P + 1) * sizeof(int32_t), stream));
    const long long want = fs_div_up(edges, 256 * 4);
    const int grid = (int)(want < 1 ? 1 : (want > (long long)FS_NUM_SMS * 16 ? (long long)FS_NUM_SMS * 16 : want));
    reverse_hist_kernel<<<grid, 256, 0, stream>>>(idx, N, k, edges, rev_ptr);
    FS_RETURN_IF_LAUNCH_FAILED();
    reverse_scan_kernel<<<B, RG_THREADS, 0, stream>>>(N, k, rev_ptr);
    FS_RETURN_IF_LAUNCH_FAILED();
    reverse_fill_kernel<<<grid, 256, 0, stream>>>(idx, N, k, edges, rev_ptr, rev_src);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edgeconv_bwd_reduce(int device, fs_stream_t stream_, const void* g, int g_dtype, int ldg,
                                      const float* sel, const void* table, int dtype, int ld, long long P, int Cp,
                                      const float* coef, float* d, double* dgb) {
    if (Cp != 64 && Cp != 128 && Cp != 256) return FS_ERR_UNSUPPORTED;
    if (!g || !sel || !coef || !d || !dgb || P < 0 || ldg < Cp || (table && ld < 2 * Cp)) return FS_ERR_BAD_ARG;
    if (P == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
#define GO(CP)                                                                                             \
    {                                                                                                      \
        const int grid = ec_grid(P, 256 / (CP / 4));                                                       \
        if (g_dtype == FS_BF16 && dtype == FS_BF16)                                                        \
            edgeconv_bwd_reduce_kernel<__nv_bfloat16, __nv_bfloat16, CP><<<grid, 256, 0, stream>>>(        \
                (const __nv_bfloat16*)g, ldg, sel, (const __nv_bfloat16*)table, ld, P, coef, d, dgb);      \
        else if (g_dtype == FS_BF16)                                                                       \
            edgeconv_bwd_reduce_kernel<__nv_bfloat16, float, CP><<<grid, 256, 0, stream>>>(                \
                (const __nv_bfloat16*)g, ldg, sel, (const float*)table, ld, P, coef, d, dgb);              \
        else if (dtype == FS_BF16)                                                                         \
            edgeconv_bwd_reduce_kernel<float, __nv_bfloat16, CP><<<grid, 256, 0, stream>>>(                \
                (const float*)g, ldg, sel, (const __nv_bfloat16*)table, ld, P, coef, d, dgb);              \
        else                                                                                               \
            edgeconv_bwd_reduce_kernel<float, float, CP><<<grid, 256, 0, stream>>>(                        \
                (const float*)g, ldg, sel, (const float*)table, ld, P, coef, d, dgb);                      \
    }
    EC_DISPATCH_CP(Cp, GO)
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edgeconv_bwd_point(int device, fs_stream_t stream_, const float* d, const float* sy,
                                     const void* table, int dtype, int ld, const int32_t* rev_ptr,
                                     const int32_t* rev_src, long long P, int k, int Cp, const float* coef,
                                     const double* dgb, double count, int train_stats, float* dT,
                                     float* dgamma_dbeta) {
    EC_CHECK_COMMON(table, ld, Cp);
    if (!d || !coef || !dT || P < 0) return FS_ERR_BAD_ARG;
    if (train_stats && (!sy || !rev_ptr || !rev_src || !dgb || count <= 0)) return FS_ERR_BAD_ARG;
    if (P == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
#define GO(CP)                                                                                              \
    {                                                                                                       \
        const int grid = ec_grid(P, 256 / (CP / 4));                                                        \
        if (dtype == FS_BF16)                                                                               \
            edgeconv_bwd_point_kernel<__nv_bfloat16, CP><<<grid, 256, 0, stream>>>(                         \
                d, sy, (const __nv_bfloat16*)table, ld, rev_ptr, rev_src, P, k, coef, dgb, count, train_stats, dT, \
                dgamma_dbeta);                                                                              \
        else                                                                                                \
            edgeconv_bwd_point_kernel<float, CP><<<grid, 256, 0, stream>>>(                                 \
                d, sy, (const float*)table, ld, rev_ptr, rev_src, P, k, coef, dgb, count, train_stats, dT,  \
                dgamma_dbeta);                                                                              \
    }
    EC_DISPATCH_CP(Cp, GO)
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edgeconv_bwd_route(int device, fs_stream_t stream_, const float* d, const uint8_t* arg,
                                     const int32_t* idx, int B, int N, int k, int Cp, const float* coef, float* dT) {
    if (!d || !arg || !idx || !coef || !dT || B < 0 || N <= 0 || k <= 0) return FS_ERR_BAD_ARG;
    if (Cp != 64 && Cp != 128 && Cp != 256) return FS_ERR_UNSUPPORTED;
    if (B == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
#define GO(CP) edgeconv_bwd_route_kernel<CP><<<ec_grid(P, 256 / (CP / 4)), 256, 0, stream>>>(d, arg, idx, P, N, k, coef, dT);
    EC_DISPATCH_CP(Cp, GO)
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edge_build(int device, fs_stream_t stream_, const void* table, int dtype, int ld,
                             const int32_t* idx, int B, int N, int k, int Cp, void* y, int y_dtype) {
    EC_CHECK_COMMON(table, ld, Cp);
    if (!idx || !y || B < 0 || N <= 0 || k <= 0) return FS_ERR_BAD_ARG;
    if (B == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
#define GO2(CP, TT, YT) \
    edge_build_kernel<TT, YT, CP><<<ec_grid(P, 256 / (CP / 4)), 256, 0, stream>>>((const TT*)table, ld, idx, P, N, k, (YT*)y)
#define GO(CP)                                                                     \
    if (dtype == FS_BF16 && y_dtype == FS_BF16) GO2(CP, __nv_bfloat16, __nv_bfloat16); \
    else if (dtype == FS_BF16) GO2(CP, __nv_bfloat16, float);                      \
    else if (y_dtype == FS_BF16) GO2(CP, float, __nv_bfloat16);                    \
    else GO2(CP, float, float);
    EC_DISPATCH_CP(Cp, GO)
#undef GO
#undef GO2
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edge_build_bwd(int device, fs_stream_t stream_, const void* dy, int dy_dtype, const int32_t* idx,
                                 int B, int N, int k, int Cp, float* dT) {
    if (!dy || !idx || !dT || B < 0 || N <= 0 || k <= 0) return FS_ERR_BAD_ARG;
    if (B == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
#define GO(CP)                                                                                                     \
    if (dy_dtype == FS_BF16)                                                                                       \
        edge_build_bwd_kernel<__nv_bfloat16, CP><<<ec_grid(P, 256 / (CP / 4)), 256, 0, stream>>>((const __nv_bfloat16*)dy, idx, P, N, k, dT); \
    else                                                                                                           \
        edge_build_bwd_kernel<float, CP><<<ec_grid(P, 256 / (CP / 4)), 256, 0, stream>>>((const float*)dy, idx, P, N, k, dT);
    EC_DISPATCH_CP(Cp, GO)
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edge_reduce(int device, fs_stream_t stream_, const void* z, int z_dtype, long long P, int k,
                              int Cp, const float* gamma, float* sel, uint8_t* arg, float* sy, double* stats) {
    if (!z || !gamma || !sel || !arg || P < 0 || k <= 0 || k > 255) return FS_ERR_BAD_ARG;
    if (P == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
#define GO(CP)                                                                                                   \
    if (z_dtype == FS_BF16)                                                                                      \
        edge_reduce_kernel<__nv_bfloat16, CP><<<ec_grid(P, 256 / (CP / 8)), 256, 0, stream>>>((const __nv_bfloat16*)z, P, k, gamma, sel, arg, sy, stats); \
    else                                                                                                         \
        edge_reduce_kernel<float, CP><<<ec_grid(P, 256 / (CP / 4)), 256, 0, stream>>>((const float*)z, P, k, gamma, sel, arg, sy, stats);
    EC_DISPATCH_CP(Cp, GO)
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edge_reduce_bwd(int device, fs_stream_t stream_, const void* z, int z_dtype, const float* d,
                                  const uint8_t* arg, long long P, int k, int Cp, const float* coef, const double* dgb,
                                  double count, int train_stats, void* dz, int dz_dtype) {
    if (!z || !d || !arg || !coef || !dz || P < 0 || k <= 0) return FS_ERR_BAD_ARG;
    if (train_stats && (!dgb || count <= 0)) return FS_ERR_BAD_ARG;
    if (P == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
#define GO2(CP, ZT, DT)                                                                                   \
    edge_reduce_bwd_kernel<ZT, DT, CP><<<ec_grid(P, 256 / (CP / FsRow<ZT>::VEC)), 256, 0, stream>>>(                   \
        (const ZT*)z, d, arg, P, k, coef, dgb, count, train_stats, (DT*)dz)
#define GO(CP)                                                                       \
    if (z_dtype == FS_BF16 && dz_dtype == FS_BF16) GO2(CP, __nv_bfloat16, __nv_bfloat16); \
    else if (z_dtype == FS_BF16) GO2(CP, __nv_bfloat16, float);                      \
    else if (dz_dtype == FS_BF16) GO2(CP, float, __nv_bfloat16);                     \
    else GO2(CP, float, float);
    EC_DISPATCH_CP(Cp, GO)
#undef GO
#undef GO2
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
