// EdgeConv gather/max pass with the neighbour table resident in shared memory.
//
// The gather of models/dgcnn.py:31-36 (take_along_dim over the k neighbours) reads k = 20 rows per point:
// B*N*k*Cp*4 bytes = 335 MB per launch at B = 32, N = 2048, Cp = 64 -- 4.4x the compulsory HBM bytes, all of it
// served by L2 (about 6300 B/clk for the whole chip) when rows are fetched straight from global memory.
// Here one CTA owns one (cloud, channel-slice) pair: the slice a[cloud, :, cs*CS : (cs+1)*CS] (N x CS fp32,
// 128 KB at N = 2048, CS = 16) is copied ONCE into shared memory (cp.async, 16 B per request) and every
// neighbour row is then a 16-byte LDS per lane. Shared memory delivers 128 B/clk per SM (148 SMs: 3x the L2
// rate), HBM traffic drops to the compulsory bytes (table once, idx once per slice via L2, outputs once).
//
// Thread mapping: LPR = CS/4 lanes per point (each lane owns 4 channels = one float4 per neighbour row),
// 32/LPR points per warp. The indices, b_i row and in-degree of the NEXT point are prefetched into registers
// while the current one is reduced. Batch statistics: same shifted fp64 per-point formula as
// edgeconv_gather_kernel (edgeconv.cu), committed through fs_stats_commit.
#include "fs_common.cuh"
#include "edgeconv_smem.cuh"

namespace {

constexpr int GS_THREADS = 512;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void st4(float* p, const float* f) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float* f) {
    uint2 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
    h[0] = __floats2bfloat162_rn(f[0], f[1]);
    h[1] = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = v;
}

// CS: channels per slice (16, 8 or 4); KQ: index registers per lane (KQ * LPR >= k, KQ * LPR % 4 == 0)
// MODE 0: train gather -> sel, arg, sy, stats        MODE 1: eval -> out (+ optional arg)
template <int CS, int KQ, int MODE, typename OT>
__global__ void __launch_bounds__(GS_THREADS, 1)
edgeconv_gather_smem_kernel(const float* __restrict__ table, int ld, const int32_t* __restrict__ idx, int N, int k,
                            int CP, const float* __restrict__ gamma_or_coef, const int32_t* __restrict__ rev_ptr,
                            float* __restrict__ sel_out, uint8_t* __restrict__ arg_out, float* __restrict__ sy_out,
                            double* __restrict__ stats, OT* __restrict__ out, int ld_out) {
    constexpr int LPR = CS / 4;            // lanes per point
    constexpr int RPC = 4 / LPR;           // index registers consumed per chunk of 4 neighbours
    constexpr int NCHUNK = KQ / RPC;       // chunks of 4 neighbours
    constexpr int G = GS_THREADS / LPR;    // points in flight per CTA
    static_assert(KQ % RPC == 0, "KQ must cover whole chunks");
    extern __shared__ __align__(16) unsigned char gs_smem[];
    float* srow = reinterpret_cast<float*>(gs_smem);

    const int cs = blockIdx.x;                                     // channel slice
    const long long cloud0 = (long long)blockIdx.y * N;            // first row of this cloud
    const int lane = threadIdx.x & 31;
    const int q = threadIdx.x % LPR;                               // float4 within the slice row
    const int grp = threadIdx.x / LPR;
    const int gbase = lane - q;                                    // first lane of this point's group
    const int c_glob = cs * CS + q * 4;                            // first of this lane's 4 channels

    // ---- stage the slice: rows of CS floats, 16-byte async copies -------------------------------------------
    {
        const float* src = table + cloud0 * ld + cs * CS;
        for (int e = threadIdx.x; e < N * LPR; e += GS_THREADS) {
            const int r = e / LPR, qq = e - r * LPR;
            cp_async16(srow + r * CS + qq * 4, src + (long long)r * ld + qq * 4);
        }
    }

    uint32_t flip[4];
    float pa[4], pb[4], mu[4], scale[4], beta[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int c = c_glob + e;
        if (MODE == 0) {
            flip[e] = __ldg(gamma_or_coef + c) >= 0.f ? 0u : 0x80000000u;
            pa[e] = stats ? __ldg(table + c) : 0.f;
            pb[e] = stats ? __ldg(table + CP + c) : 0.f;
        } else {
            mu[e] = __ldg(gamma_or_coef + c);
            scale[e] = __ldg(gamma_or_coef + 2 * CP + c);
            beta[e] = __ldg(gamma_or_coef + 3 * CP + c);
            flip[e] = scale[e] >= 0.f ? 0u : 0x80000000u;
        }
    }
    (void)pa; (void)pb; (void)mu; (void)scale; (void)beta;

    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};

    // point range of this CTA (gridDim.z CTAs share one slice when the batch is small)
    const int per = (N + gridDim.z - 1) / gridDim.z;
    const int p_begin = blockIdx.z * per;
    const int p_end = min(p_begin + per, N);

    // prefetch registers for one point: KQ indices (lane q holds neighbours q, q+LPR, ...), b_i, in-degree
    int jcur[KQ];
    float4 bcur = make_float4(0.f, 0.f, 0.f, 0.f);
    int dcur = 0;
    auto fetch = [&](int pt, int (&j)[KQ], float4& b, int& deg) {
        const int32_t* irow = idx + (cloud0 + pt) * k;
#pragma unroll
        for (int m = 0; m < KQ; ++m) {
            const int t = m * LPR + q;
            j[m] = __ldg(irow + (t < k ? t : k - 1));             // tail: repeat the last neighbour
        }
        b = __ldg(reinterpret_cast<const float4*>(table + (cloud0 + pt) * ld + CP + c_glob));
        if (MODE == 0 && stats) deg = __ldg(rev_ptr + cloud0 + pt + 1) - __ldg(rev_ptr + cloud0 + pt);
    };
    // The trip count is uniform over the CTA (full-mask shuffles inside): groups past the end of the range redo
    // the last point and skip the stores.
    int pt_raw = p_begin + grp;
    fetch(min(pt_raw, p_end - 1), jcur, bcur, dcur);

    cp_async_wait_all();
    __syncthreads();

    const float kf = (float)k;
    for (int base = p_begin; base < p_end; base += G, pt_raw += G) {
        const bool live = pt_raw < p_end;
        const int pt = live ? pt_raw : p_end - 1;
        int jnext[KQ];
        float4 bnext = make_float4(0.f, 0.f, 0.f, 0.f);
        int dnext = 0;
        if (base + G < p_end) fetch(min(pt_raw + G, p_end - 1), jnext, bnext, dnext);

        float best[4], S1[4];
        int barg[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { best[e] = -INFINITY; barg[e] = 0; S1[e] = 0.f; }
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int t0 = ch * 4;
            if (t0 < k) {                                          // warp-uniform
                float4 a[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = __shfl_sync(FS_FULL_MASK, jcur[ch * RPC + u / LPR], gbase + (u % LPR));
                    a[u] = *reinterpret_cast<const float4*>(srow + j * CS + q * 4);
                }
                float w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) w[u] = (t0 + u < k) ? 1.f : 0.f;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float a0 = reinterpret_cast<const float*>(&a[0])[e];
                    const float a1 = reinterpret_cast<const float*>(&a[1])[e];
                    const float a2 = reinterpret_cast<const float*>(&a[2])[e];
                    const float a3 = reinterpret_cast<const float*>(&a[3])[e];
                    if (MODE == 0) S1[e] += fmaf(w[3], a3, fmaf(w[2], a2, fmaf(w[1], a1, w[0] * a0)));
                    const float k0 = __uint_as_float(__float_as_uint(a0) ^ flip[e]);
                    const float k1 = __uint_as_float(__float_as_uint(a1) ^ flip[e]);
                    const float k2 = __uint_as_float(__float_as_uint(a2) ^ flip[e]);
                    const float k3 = __uint_as_float(__float_as_uint(a3) ^ flip[e]);
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
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) best[e] = __uint_as_float(__float_as_uint(best[e]) ^ flip[e]);
        const float bi[4] = {bcur.x, bcur.y, bcur.z, bcur.w};
        const long long row = cloud0 + pt;
        if (MODE == 0) {
            float sy[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) sy[e] = fmaf(kf, bi[e], S1[e]);
            if (live) {
                st4(sel_out + row * CP + c_glob, best);
                if (sy_out) st4(sy_out + row * CP + c_glob, sy);
            }
            if (stats && live) {
                const float4 av = *reinterpret_cast<const float4*>(srow + pt * CS + q * 4);
                const float ai[4] = {av.x, av.y, av.z, av.w};
                const double deg = (double)dcur, kd = (double)k;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double ap = (double)ai[e] - (double)pa[e];
                    const double bp = (double)bi[e] - (double)pb[e];
                    const double S1p = (double)S1[e] - kd * (double)pa[e];
                    s1[e] += S1p + kd * bp;
                    s2[e] += deg * ap * ap + bp * (kd * bp + 2.0 * S1p);
                }
            }
        } else {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) o[e] = fs_leaky(fmaf(scale[e], (best[e] + bi[e]) - mu[e], beta[e]));
            if (live) st4(out + row * ld_out + c_glob, o);
        }
        if (arg_out && live) {
            const uint32_t pk = (uint32_t)barg[0] | ((uint32_t)barg[1] << 8) | ((uint32_t)barg[2] << 16) |
                                ((uint32_t)barg[3] << 24);
            *reinterpret_cast<uint32_t*>(arg_out + row * CP + c_glob) = pk;
        }
#pragma unroll
        for (int m = 0; m < KQ; ++m) jcur[m] = jnext[m];
        bcur = bnext;
        dcur = dnext;
    }

    if (MODE == 0 && stats) {
        __syncthreads();                                           // every gather is done: reuse the slice as scratch
        int chans[4] = {c_glob, c_glob + 1, c_glob + 2, c_glob + 3};
        const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
        const unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        fs_stats_commit_impl<4>(reinterpret_cast<double*>(gs_smem), s1, s2, chans, LPR, CP, stats, lin, nblocks);
        if (blockIdx.y == 0 && blockIdx.z == 0 && grp == 0) {     // every slice publishes the pivots of its channels
#pragma unroll
            for (int e = 0; e < 4; ++e) stats[2 * CP + c_glob + e] = (double)pa[e] + (double)pb[e];
        }
    }
}

constexpr size_t GS_SMEM_CAP = 200 * 1024;
constexpr size_t GS_SCRATCH = 2 * 4 * GS_THREADS * sizeof(double);   // fs_stats_commit buffer

int pick_cs(int N) {
    for (int cs = 16; cs >= 4; cs >>= 1)
        if ((size_t)N * cs * 4 <= GS_SMEM_CAP) return cs;
    return 0;
}

template <int CS, int KQ, int MODE, typename OT>
int launch(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k, int CP,
           const float* gc, const int32_t* rev_ptr, float* sel, uint8_t* arg, float* sy, double* stats, OT* out,
           int ld_out) {
    size_t smem = (size_t)N * CS * 4;
    if (smem < GS_SCRATCH) smem = GS_SCRATCH;
    auto kern = edgeconv_gather_smem_kernel<CS, KQ, MODE, OT>;
    static bool configured = false;      // idempotent attribute; a race only repeats the call
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_SMEM_CAP);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    const int slices = CP / CS;
    int split = 1;                       // small batches: several CTAs share a slice (each stages it again from L2)
    while (split < 8 && (long long)B * slices * split * 2 <= FS_NUM_SMS && N / (split * 2) >= GS_THREADS / (CS / 4)) split *= 2;
    dim3 grid(slices, B, split);
    kern<<<grid, GS_THREADS, smem, stream>>>(table, ld, idx, N, k, CP, gc, rev_ptr, sel, arg, sy, stats, out, ld_out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <int MODE, typename OT>
int dispatch(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k, int CP,
             const float* gc, const int32_t* rev_ptr, float* sel, uint8_t* arg, float* sy, double* stats, OT* out,
             int ld_out) {
    const int cs = pick_cs(N);
    if (cs == 0 || B > 65535) return FS_SMEM_GATHER_UNSUPPORTED;
#define GS_GO(CS, KQ) \
    return launch<CS, KQ, MODE, OT>(stream, table, ld, idx, B, N, k, CP, gc, rev_ptr, sel, arg, sy, stats, out, ld_out)
    if (cs == 16) {              // 4 lanes per point: KQ = ceil(k / 4)
        if (k <= 20) GS_GO(16, 5);
        if (k <= 40) GS_GO(16, 10);
        if (k <= 64) GS_GO(16, 16);
    } else if (cs == 8) {        // 2 lanes per point: KQ = 2 * ceil(k / 4)
        if (k <= 20) GS_GO(8, 10);
        if (k <= 40) GS_GO(8, 20);
    } else {                     // 1 lane per point: KQ = 4 * ceil(k / 4)
        if (k <= 20) GS_GO(4, 20);
        if (k <= 40) GS_GO(4, 40);
    }
#undef GS_GO
    return FS_SMEM_GATHER_UNSUPPORTED;
}

}  // namespace

int fs_gather_smem_train(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k,
                         int CP, const float* gamma, const int32_t* rev_ptr, float* sel, uint8_t* arg, float* sy,
                         double* stats) {
    return dispatch<0, float>(stream, table, ld, idx, B, N, k, CP, gamma, rev_ptr, sel, arg, sy, stats, nullptr, 0);
}

int fs_gather_smem_eval(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k,
                        int CP, const float* coef, void* out, int out_bf16, int ld_out, uint8_t* arg) {
    if (out_bf16)
        return dispatch<1, __nv_bfloat16>(stream, table, ld, idx, B, N, k, CP, coef, nullptr, nullptr, arg, nullptr,
                                          nullptr, (__nv_bfloat16*)out, ld_out);
    return dispatch<1, float>(stream, table, ld, idx, B, N, k, CP, coef, nullptr, nullptr, arg, nullptr, nullptr,
                              (float*)out, ld_out);
}
