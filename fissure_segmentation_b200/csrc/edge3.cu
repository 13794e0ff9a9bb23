// First layer of an EdgeConv on raw 3-D inputs (ec1 of DGCNNSeg and the spatial transformer's EdgeConv:
// models/dgcnn.py:119, 251 with in_features = 3): Conv2d(6 -> Cp, 1x1) + BatchNorm2d + LeakyReLU(0.2) on the edge
// features e = [x_j - x_i, x_i] (models/dgcnn.py:15-36), producing the hidden edge tensor H [P*k, Cp] that feeds
// the second shared layer.
//
// The pre-activation y_c = w_c . e is linear in the 6-vector e, so the BatchNorm batch statistics of all Cp
// channels follow from the first and second moments of e over the edges (6 + 21 numbers):
//     mean_c = w_c . E[e]          var_c = w_c^T (E[e e^T] - E[e] E[e]^T) w_c
// One light pass over the edges (12 bytes gathered per edge) replaces a statistics pass over the P*k x Cp
// tensor, and y is recomputed from the coordinates wherever it is needed (6 FMA per channel) instead of being
// stored: forward writes only H, backward reads only dH. With a 3-channel input that needs no gradient the
// weight gradient is accumulated directly as sum_e dy_e (x) e — no scatter over the graph.
#include "fs_common.cuh"

namespace {

constexpr int E3_THREADS = 256;
constexpr int E3_NMOM = 27;          // 6 first moments + 21 upper-triangular second moments
constexpr int E3_SLOTS = 32;

__device__ __forceinline__ void load_e(const float* __restrict__ x, int ldx, long long i, long long j, float* e) {
    const float xi0 = __ldg(x + i * ldx), xi1 = __ldg(x + i * ldx + 1), xi2 = __ldg(x + i * ldx + 2);
    e[0] = __ldg(x + j * ldx) - xi0; e[1] = __ldg(x + j * ldx + 1) - xi1; e[2] = __ldg(x + j * ldx + 2) - xi2;
    e[3] = xi0; e[4] = xi1; e[5] = xi2;
}

// mom layout (doubles): [0,27) final | [27, 27 + 32*27) slot partials | ticket
__global__ void __launch_bounds__(E3_THREADS)
edge3_moments_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx, long long P, int N, int k,
                     double* __restrict__ mom) {
    float m[E3_NMOM];
#pragma unroll
    for (int a = 0; a < E3_NMOM; ++a) m[a] = 0.f;
    for (long long pt = (long long)blockIdx.x * E3_THREADS + threadIdx.x; pt < P; pt += (long long)gridDim.x * E3_THREADS) {
        const long long cloud0 = (pt / N) * N;
        const float xi0 = __ldg(x + pt * ldx), xi1 = __ldg(x + pt * ldx + 1), xi2 = __ldg(x + pt * ldx + 2);
        float s[3] = {0.f, 0.f, 0.f}, q[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // sums of d and of d d^T (d = x_j - x_i)
        for (int t = 0; t < k; ++t) {
            const long long j = cloud0 + __ldg(idx + pt * k + t);
            const float d0 = __ldg(x + j * ldx) - xi0, d1 = __ldg(x + j * ldx + 1) - xi1, d2 = __ldg(x + j * ldx + 2) - xi2;
            s[0] += d0; s[1] += d1; s[2] += d2;
            q[0] = fmaf(d0, d0, q[0]); q[1] = fmaf(d0, d1, q[1]); q[2] = fmaf(d0, d2, q[2]);
            q[3] = fmaf(d1, d1, q[3]); q[4] = fmaf(d1, d2, q[4]); q[5] = fmaf(d2, d2, q[5]);
        }
        const float kf = (float)k;
        // first moments: e = [d, x_i]
        m[0] += s[0]; m[1] += s[1]; m[2] += s[2]; m[3] += kf * xi0; m[4] += kf * xi1; m[5] += kf * xi2;
        // second moments, upper triangle row-major over (0..5): (0,0)(0,1)...(0,5)(1,1)...(5,5)
        m[6] += q[0]; m[7] += q[1]; m[8] += q[2]; m[9] += s[0] * xi0; m[10] += s[0] * xi1; m[11] += s[0] * xi2;
        m[12] += q[3]; m[13] += q[4]; m[14] += s[1] * xi0; m[15] += s[1] * xi1; m[16] += s[1] * xi2;
        m[17] += q[5]; m[18] += s[2] * xi0; m[19] += s[2] * xi1; m[20] += s[2] * xi2;
        m[21] += kf * xi0 * xi0; m[22] += kf * xi0 * xi1; m[23] += kf * xi0 * xi2;
        m[24] += kf * xi1 * xi1; m[25] += kf * xi1 * xi2;
        m[26] += kf * xi2 * xi2;
    }
    __shared__ double red[E3_THREADS / 32][E3_NMOM];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < E3_NMOM; ++a) {
        const double v = fs_warp_sum((double)m[a]);
        if (lane == 0) red[warp][a] = v;
    }
    __syncthreads();
    double* slot = mom + E3_NMOM + (blockIdx.x % E3_SLOTS) * E3_NMOM;
    if (threadIdx.x < E3_NMOM) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < E3_THREADS / 32; ++w) v += red[w][threadIdx.x];
        atomicAdd(slot + threadIdx.x, v);
    }
    __threadfence();
    __syncthreads();
    unsigned* ticket = reinterpret_cast<unsigned*>(mom + E3_NMOM + E3_SLOTS * E3_NMOM);
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < E3_NMOM) {
        __threadfence();
        double v = 0.0;
        for (int sidx = 0; sidx < E3_SLOTS; ++sidx) v += __ldcg(mom + E3_NMOM + sidx * E3_NMOM + threadIdx.x);
        mom[threadIdx.x] = v;
    }
}

// coef [mu | invstd | scale | beta] of all Cp channels from the moments and the weight w [Cp, 6].
__global__ void edge3_coef_kernel(const double* __restrict__ mom, double count, const float* __restrict__ w, int Cp,
                                  const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                  float* __restrict__ coef, float* running_mean, float* running_var, long long* nbt) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c == 0 && nbt) *nbt += 1;
    if (c >= Cp) return;
    double wc[6], m1[6];
#pragma unroll
    for (int a = 0; a < 6; ++a) { wc[a] = (double)w[c * 6 + a]; m1[a] = mom[a] / count; }
    double mean = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) mean += wc[a] * m1[a];
    double ey2 = 0.0;
    int p = 6;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int b2 = a; b2 < 6; ++b2) { ey2 += (a == b2 ? 1.0 : 2.0) * wc[a] * wc[b2] * (mom[p] / count); ++p; }
    double var = ey2 - mean * mean;
    if (var < 0.0) var = 0.0;
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

// H[(i,t), c] = LeakyReLU(scale_c (w_c . e - mu_c) + beta_c).
// One warp per point, lane = CP/32 consecutive channels: x_i is loaded once, every neighbour costs one index load,
// three coordinate loads (warp-uniform addresses) and one coalesced 32*NC-channel row store.
template <typename HT, int CP>
__global__ void __launch_bounds__(E3_THREADS)
edge3_hidden_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx, long long P, int N, int k,
                    const float* __restrict__ w, const float* __restrict__ coef, HT* __restrict__ h) {
    constexpr int NC = CP / 32;                     // channels per lane (2 or 4)
    const int lane = threadIdx.x & 31;
    const int c0 = lane * NC;
    float wr[NC][6], mu[NC], sc[NC], be[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
#pragma unroll
        for (int a = 0; a < 6; ++a) wr[i][a] = __ldg(w + (c0 + i) * 6 + a);
        mu[i] = __ldg(coef + c0 + i); sc[i] = __ldg(coef + 2 * CP + c0 + i); be[i] = __ldg(coef + 3 * CP + c0 + i);
    }
    const long long nwarps = (long long)gridDim.x * (E3_THREADS / 32);
    for (long long pt = (long long)blockIdx.x * (E3_THREADS / 32) + (threadIdx.x >> 5); pt < P; pt += nwarps) {
        const long long cloud0 = (pt / N) * N;
        const float xi0 = __ldg(x + pt * ldx), xi1 = __ldg(x + pt * ldx + 1), xi2 = __ldg(x + pt * ldx + 2);
        float base[NC];                               // contribution of x_i: w[3:6] . x_i
#pragma unroll
        for (int i = 0; i < NC; ++i) base[i] = fmaf(wr[i][5], xi2, fmaf(wr[i][4], xi1, wr[i][3] * xi0));
        for (int t0 = 0; t0 < k; t0 += 32) {
            // one neighbour per lane: all index / coordinate gathers of the point are in flight together
            const int tl = t0 + lane < k ? t0 + lane : k - 1;
            const long long jl = cloud0 + __ldg(idx + pt * k + tl);
            const float n0 = __ldg(x + jl * ldx), n1 = __ldg(x + jl * ldx + 1), n2 = __ldg(x + jl * ldx + 2);
            const int tn = k - t0 < 32 ? k - t0 : 32;
            for (int tt = 0; tt < tn; ++tt) {
                const float d0 = __shfl_sync(FS_FULL_MASK, n0, tt) - xi0, d1 = __shfl_sync(FS_FULL_MASK, n1, tt) - xi1,
                            d2 = __shfl_sync(FS_FULL_MASK, n2, tt) - xi2;
                float o[NC];
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    const float y = fmaf(wr[i][2], d2, fmaf(wr[i][1], d1, fmaf(wr[i][0], d0, base[i])));
                    o[i] = fs_leaky(fmaf(sc[i], y - mu[i], be[i]));
                }
                const long long eo = (pt * k + t0 + tt) * CP + c0;
                if (sizeof(HT) == 2) {
                    __nv_bfloat162 v[NC / 2];
#pragma unroll
                    for (int i = 0; i < NC / 2; ++i) v[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
                    if (NC == 2) *reinterpret_cast<__nv_bfloat162*>(reinterpret_cast<__nv_bfloat16*>(h) + eo) = v[0];
                    else *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(h) + eo) = *reinterpret_cast<uint2*>(v);
                } else {
                    if (NC == 2) *reinterpret_cast<float2*>(reinterpret_cast<float*>(h) + eo) = make_float2(o[0], o[1]);
                    else *reinterpret_cast<float4*>(reinterpret_cast<float*>(h) + eo) = make_float4(o[0], o[1], o[NC - 2], o[NC - 1]);
                }
            }
        }
    }
}

template <typename T, int NC> __device__ __forceinline__ void load_nc(const T* p, float* f);
template <> __device__ __forceinline__ void load_nc<float, 2>(const float* p, float* f) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(p)); f[0] = v.x; f[1] = v.y;
}
template <> __device__ __forceinline__ void load_nc<float, 4>(const float* p, float* f) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p)); f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
template <> __device__ __forceinline__ void load_nc<__nv_bfloat16, 2>(const __nv_bfloat16* p, float* f) {
    const unsigned v = __ldg(reinterpret_cast<const unsigned*>(p));
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v)); f[0] = t.x; f[1] = t.y;
}
template <> __device__ __forceinline__ void load_nc<__nv_bfloat16, 4>(const __nv_bfloat16* p, float* f) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&v);
    const float2 a = __bfloat1622float2(hh[0]), b = __bfloat1622float2(hh[1]); f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

// Backward, ONE pass over dH: per channel  s1 = sum d,  s2 = sum d*yhat,  A[a] = sum d * e_a   (d = dH * LeakyReLU'(z)).
// The BatchNorm-coupled weight gradient then follows in closed form from the edge moments (edge3_dw_kernel).
// acc layout (float): [Cp][6] A sums (atomics; zeroed by the caller).
//
// One warp per point. A dH row (Cp channels) is read by LPRW = Cp/4 lanes with one 8-byte (bf16) or 16-byte (fp32)
// load each, so a warp-wide load covers EPW = 32/LPRW edges and four such loads are in flight per lane
// (1 KB per warp); the first revision used 4-byte loads of one edge at a time and was latency bound at 0.8 TB/s.
template <typename GT, int CP>
__global__ void __launch_bounds__(E3_THREADS)
edge3_bwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx, long long P, int N, int k,
                 const float* __restrict__ w, const float* __restrict__ coef, const GT* __restrict__ dh, double* __restrict__ dgb,
                 float* __restrict__ acc_out) {
    constexpr int NC = 4;
    constexpr int LPRW = CP / NC;          // lanes per dH row (16 or 32)
    constexpr int EPW = 32 / LPRW;         // edges per warp-wide load (2 or 1)
    constexpr int U = 4;                   // loads in flight per lane
    static_assert(LPRW <= 32 && EPW * LPRW == 32, "Cp must be 64 or 128");
    __shared__ double red[2 * NC * E3_THREADS];
    __shared__ float ared[E3_THREADS / 32][CP * 6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / LPRW;
    const int c0 = (lane % LPRW) * NC;
    float wr[NC][6], mu[NC], sc[NC], be[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
#pragma unroll
        for (int a = 0; a < 6; ++a) wr[i][a] = __ldg(w + (c0 + i) * 6 + a);
        mu[i] = __ldg(coef + c0 + i); sc[i] = __ldg(coef + 2 * CP + c0 + i); be[i] = __ldg(coef + 3 * CP + c0 + i);
    }
    float s1[NC], s2[NC], A[NC][6];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        s1[i] = 0.f; s2[i] = 0.f;
#pragma unroll
        for (int a = 0; a < 6; ++a) A[i][a] = 0.f;
    }
    const long long nwarps = (long long)gridDim.x * (E3_THREADS / 32);
    for (long long pt = (long long)blockIdx.x * (E3_THREADS / 32) + warp; pt < P; pt += nwarps) {
        const long long cloud0 = (pt / N) * N;
        const float xi0 = __ldg(x + pt * ldx), xi1 = __ldg(x + pt * ldx + 1), xi2 = __ldg(x + pt * ldx + 2);
        float base[NC], dsum[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) { base[i] = fmaf(wr[i][5], xi2, fmaf(wr[i][4], xi1, wr[i][3] * xi0)); dsum[i] = 0.f; }
        for (int t0 = 0; t0 < k; t0 += 32) {
            const int tl = t0 + lane < k ? t0 + lane : k - 1;
            const long long jl = cloud0 + __ldg(idx + pt * k + tl);
            const float n0 = __ldg(x + jl * ldx), n1 = __ldg(x + jl * ldx + 1), n2 = __ldg(x + jl * ldx + 2);
            const int tn = k - t0 < 32 ? k - t0 : 32;
            for (int tb = 0; tb < tn; tb += EPW * U) {
                float g[U][NC];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int tt = tb + u * EPW + sub;
                    load_nc<GT, NC>(dh + (pt * k + t0 + (tt < tn ? tt : tn - 1)) * CP + c0, g[u]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (tb + u * EPW < tn) {                     // warp-uniform
                        const int tt = tb + u * EPW + sub;
                        const int ts = tt < tn ? tt : tn - 1;
                        const float live = tt < tn ? 1.f : 0.f;      // the clamped duplicate contributes nothing
                        const float d0 = __shfl_sync(FS_FULL_MASK, n0, ts) - xi0, d1 = __shfl_sync(FS_FULL_MASK, n1, ts) - xi1,
                                    d2 = __shfl_sync(FS_FULL_MASK, n2, ts) - xi2;
#pragma unroll
                        for (int i = 0; i < NC; ++i) {
                            const float y = fmaf(wr[i][2], d2, fmaf(wr[i][1], d1, fmaf(wr[i][0], d0, base[i])));
                            const float yc = y - mu[i];
                            const float z = fmaf(sc[i], yc, be[i]);
                            const float gl = g[u][i] * live;
                            const float d = z > 0.f ? gl : 0.2f * gl;
                            s2[i] = fmaf(d, yc, s2[i]);
                            dsum[i] += d;
                            A[i][0] = fmaf(d, d0, A[i][0]); A[i][1] = fmaf(d, d1, A[i][1]); A[i][2] = fmaf(d, d2, A[i][2]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NC; ++i) {      // the x_i components of e are constant over the point's edges
            s1[i] += dsum[i];
            A[i][3] = fmaf(dsum[i], xi0, A[i][3]); A[i][4] = fmaf(dsum[i], xi1, A[i][4]); A[i][5] = fmaf(dsum[i], xi2, A[i][5]);
        }
    }
    // the EPW edge sub-groups of a warp own the same channels: fold them into sub-group 0
#pragma unroll
    for (int o = LPRW; o < 32; o <<= 1) {
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            s1[i] += __shfl_xor_sync(FS_FULL_MASK, s1[i], o);
            s2[i] += __shfl_xor_sync(FS_FULL_MASK, s2[i], o);
#pragma unroll
            for (int a = 0; a < 6; ++a) A[i][a] += __shfl_xor_sync(FS_FULL_MASK, A[i][a], o);
        }
    }
    if (sub == 0) {
#pragma unroll
        for (int i = 0; i < NC; ++i)
#pragma unroll
            for (int a = 0; a < 6; ++a) ared[warp][(c0 + i) * 6 + a] = A[i][a];
    }
    double d1v[NC], d2v[NC];
    int chans[NC];
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        d1v[i] = sub == 0 ? (double)s1[i] : 0.0;
        d2v[i] = sub == 0 ? (double)s2[i] * (double)__ldg(coef + CP + c0 + i) : 0.0;     // yhat = yc * invstd
        chans[i] = c0 + i;
    }
    fs_stats_commit<NC>(red, d1v, d2v, chans, LPRW, CP, dgb);     // contains the __syncthreads that publishes ared
    for (int o = threadIdx.x; o < CP * 6; o += E3_THREADS) {
        float v = 0.f;
#pragma unroll
        for (int ww = 0; ww < E3_THREADS / 32; ++ww) v += ared[ww][o];
        atomicAdd(acc_out + o, v);
    }
}

// dW[c, a] = scale_c ( A[c,a] - dbeta_c m1[a] - dgamma_c invstd_c ( (M2 w_c)[a] - mu_c m1[a] ) )   (train)
//          = scale_c A[c,a]                                                                         (eval)
__global__ void edge3_dw_kernel(const float* __restrict__ acc, const double* __restrict__ dgb, const double* __restrict__ mom,
                                double count, const float* __restrict__ w, const float* __restrict__ coef, int Cp,
                                int train_stats, float* __restrict__ dw) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cp) return;
    const double sc = (double)coef[2 * Cp + c], inv = (double)coef[Cp + c], mu = (double)coef[c];
    double M2[6][6], m1[6], wc[6];
    if (train_stats) {
        int p = 6;
        for (int a = 0; a < 6; ++a) { m1[a] = mom[a] / count; wc[a] = (double)w[c * 6 + a]; }
        for (int a = 0; a < 6; ++a)
            for (int b2 = a; b2 < 6; ++b2) { M2[a][b2] = M2[b2][a] = mom[p] / count; ++p; }
    }
    for (int a = 0; a < 6; ++a) {
        double v = (double)acc[c * 6 + a];
        if (train_stats) {
            double m2w = 0.0;
            for (int b2 = 0; b2 < 6; ++b2) m2w += M2[a][b2] * wc[b2];
            v -= dgb[c] * m1[a] + dgb[Cp + c] * inv * (m2w - mu * m1[a]);
        }
        dw[c * 6 + a] = (float)(sc * v);
    }
}

int e3_grid(long long items, int per_block) {
    long long need = (items + per_block - 1) / per_block;
    const long long cap = (long long)FS_NUM_SMS * 4;
    return (int)(need < 1 ? 1 : (need > cap ? cap : need));
}

}  // namespace

extern "C" size_t fs_edge3_moment_doubles(void) { return (size_t)E3_NMOM * (1 + E3_SLOTS) + 2; }

extern "C" int fs_edge3_bn_coef(int device, fs_stream_t stream_, const float* x, int ldx, const int32_t* idx, int B, int N,
                                int k, const float* w, int Cp, const float* gamma, const float* beta, float eps,
                                float momentum, double* moments, float* coef, float* running_mean, float* running_var,
                                long long* num_batches_tracked) {
    if (!x || !idx || !w || !gamma || !beta || !moments || !coef || B <= 0 || N <= 0 || k <= 0 || Cp <= 0 || ldx < 3)
        return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    edge3_moments_kernel<<<e3_grid(P, E3_THREADS), E3_THREADS, 0, stream>>>(x, ldx, idx, P, N, k, moments);
    FS_RETURN_IF_LAUNCH_FAILED();
    edge3_coef_kernel<<<fs_div_up(Cp, 64), 64, 0, stream>>>(moments, (double)P * k, w, Cp, gamma, beta, eps, momentum, coef,
                                                           running_mean, running_var, num_batches_tracked);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edge3_hidden(int device, fs_stream_t stream_, const float* x, int ldx, const int32_t* idx, int B, int N,
                               int k, const float* w, int Cp, const float* coef, void* h, int h_dtype) {
    if (!x || !idx || !w || !coef || !h || B <= 0 || N <= 0 || k <= 0 || ldx < 3) return FS_ERR_BAD_ARG;
    if (Cp != 64 && Cp != 128) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    const int grid = e3_grid(P, E3_THREADS / 32);
#define GO(CP)                                                                                                        \
    if (h_dtype == FS_BF16)                                                                                           \
        edge3_hidden_kernel<__nv_bfloat16, CP><<<grid, E3_THREADS, 0, stream>>>(x, ldx, idx, P, N, k, w, coef, (__nv_bfloat16*)h); \
    else                                                                                                              \
        edge3_hidden_kernel<float, CP><<<grid, E3_THREADS, 0, stream>>>(x, ldx, idx, P, N, k, w, coef, (float*)h);
    if (Cp == 64) { GO(64) } else { GO(128) }
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edge3_bwd(int device, fs_stream_t stream_, const float* x, int ldx, const int32_t* idx, int B, int N, int k,
                            const float* w, int Cp, const float* coef, const void* dh, int dh_dtype, int train_stats,
                            const double* moments, double* dgb, float* acc_ws, float* dw) {
    if (!x || !idx || !w || !coef || !dh || !dgb || !acc_ws || !dw || B <= 0 || N <= 0 || k <= 0 || ldx < 3) return FS_ERR_BAD_ARG;
    if (train_stats && !moments) return FS_ERR_BAD_ARG;
    if (Cp != 64 && Cp != 128) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    const int grid = e3_grid(P, E3_THREADS / 32);
#define GO(CP)                                                                                                        \
    if (dh_dtype == FS_BF16)                                                                                          \
        edge3_bwd_kernel<__nv_bfloat16, CP><<<grid, E3_THREADS, 0, stream>>>(x, ldx, idx, P, N, k, w, coef,           \
                                                                             (const __nv_bfloat16*)dh, dgb, acc_ws);  \
    else                                                                                                              \
        edge3_bwd_kernel<float, CP><<<grid, E3_THREADS, 0, stream>>>(x, ldx, idx, P, N, k, w, coef, (const float*)dh, dgb, acc_ws);
    if (Cp == 64) { GO(64) } else { GO(128) }
#undef GO
    FS_RETURN_IF_LAUNCH_FAILED();
    edge3_dw_kernel<<<fs_div_up(Cp, 64), 64, 0, stream>>>(acc_ws, dgb, moments, (double)P * k, w, coef, Cp, train_stats, dw);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

// The closed-form weight gradient alone (second half of fs_edge3_bwd): for callers that accumulate the per-channel sums
// and the 64 x 6 moment sums themselves (fs_edge2_bwd).
extern "C" int fs_edge3_dw(int device, fs_stream_t stream_, const float* acc, const double* dgb, const double* moments,
                           double count, const float* w, const float* coef, int Cp, int train_stats, float* dw) {
    if (!acc || !dgb || !w || !coef || !dw || Cp <= 0 || count <= 0) return FS_ERR_BAD_ARG;
    if (train_stats && !moments) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    edge3_dw_kernel<<<fs_div_up(Cp, 64), 64, 0, (cudaStream_t)stream_>>>(acc, dgb, moments, count, w, coef, Cp, train_stats, dw);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
