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
// 32/LPR points per warp. The slice is stored as KEYS (sign bit flipped on channels that take the minimum), so the
// loop is a pure arg-max. The indices, b_i rows and in-degrees of a warp's NEXT 32/LPR points are staged by
// cp.async into a per-warp double buffer while the current ones are reduced. Batch statistics: same shifted fp64
// per-point formula as edgeconv_gather_kernel (edgeconv.cu), committed through fs_stats_commit.
#include "fs_common.cuh"
#include "edgeconv_smem.cuh"

namespace {

constexpr int GS_THREADS = 512;    // 1024 threads (64 registers) measured slower: the loop is LSU/ALU-throughput bound

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

// Per-warp staging of the NEXT iteration's indices, b rows and in-degrees: cp.async into shared memory, so the
// hot loop holds no register scoreboard on a global load (register prefetch made every SHFL/LDS of the loop wait
// for the outstanding LDGs that shared its scoreboard slot: 16 % of all stall samples).
__device__ __forceinline__ void cp_async_zfill16(void* smem, const void* gmem, int src_bytes) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_PENDING) : "memory"); }

__host__ __device__ inline int gs_align16(int bytes) { return (bytes + 15) & ~15; }
// staging bytes of one warp and one buffer: idx (PPW*k + 4 slack ints) | b (PPW*CS floats) | rev_ptr (PPW+1 ints)
__host__ __device__ inline int gs_stage_bytes(int ppw, int k, int cs) {
    return gs_align16((ppw * k + 4) * 4) + ppw * cs * 4 + gs_align16((ppw + 1) * 4);
}

// CS: channels per slice (16 or 8); NCHUNK: chunks of 4 neighbours covered (4 * NCHUNK >= k)
// MODE 0: train gather -> sel, arg, sy, stats        MODE 1: eval -> out (+ optional arg)
template <int CS, int NCHUNK, int MODE, typename OT>
__global__ void __launch_bounds__(GS_THREADS, 1)
edgeconv_gather_smem_kernel(const float* __restrict__ table, int ld, const int32_t* __restrict__ idx, int N, int k,
                            int CP, const float* __restrict__ gamma_or_coef, const int32_t* __restrict__ rev_ptr,
                            float* __restrict__ sel_out, uint8_t* __restrict__ arg_out, float* __restrict__ sy_out,
                            double* __restrict__ stats, OT* __restrict__ out, int ld_out, int table_bytes) {
    constexpr int LPR = CS / 4;            // lanes per point
    constexpr int PPW = 32 / LPR;          // points per warp and iteration
    constexpr int G = GS_THREADS / LPR;    // points per CTA and iteration
    extern __shared__ __align__(16) unsigned char gs_smem[];
    float* srow = reinterpret_cast<float*>(gs_smem);

    const int cs = blockIdx.x;                                     // channel slice
    const long long cloud0 = (long long)blockIdx.y * N;            // first row of this cloud
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int q = threadIdx.x % LPR;                               // float4 within the slice row
    const int sub = lane / LPR;                                    // point of this lane within the warp's block
    const int c_glob = cs * CS + q * 4;                            // first of this lane's 4 channels

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

    // point range of this CTA (gridDim.z CTAs share one slice when the batch is small)
    const int per = (N + gridDim.z - 1) / gridDim.z;
    const int p_begin = blockIdx.z * per;
    const int p_end = min(p_begin + per, N);

    // ---- per-warp staging buffers (two per warp) behind the table slice --------------------------------------
    const int idx_bytes = gs_align16((PPW * k + 4) * 4);
    const int stage_bytes = gs_stage_bytes(PPW, k, CS);
    unsigned char* stage0 = gs_smem + table_bytes + (size_t)warp * 2 * stage_bytes;
    const bool idx_vec = ((reinterpret_cast<uintptr_t>(idx + (cloud0 + p_begin + warp * PPW) * k) & 15) == 0) &&
                         ((G * k) % 4 == 0);
    auto stage = [&](int it, int buf) {
        unsigned char* st = stage0 + buf * stage_bytes;
        const int pt0 = p_begin + it * G + warp * PPW;
        const int32_t* gi = idx + (cloud0 + pt0) * k;
        const int nvalid = (min(p_end, pt0 + PPW) - pt0) * k;          // ints of this block inside the range (may be <= 0)
        if (idx_vec) {
            for (int i = lane; i < PPW * k / 4; i += 32) {
                const int bytes = max(0, min(16, (nvalid - 4 * i) * 4));
                cp_async_zfill16(st + 16 * i, bytes > 0 ? (const void*)(gi + 4 * i) : (const void*)idx, bytes);
            }
        } else {
            for (int i = lane; i < PPW * k; i += 32)
                cp_async4(st + 4 * i, i < nvalid ? gi + i : idx);
        }
        const int ptc = min(pt0 + sub, p_end - 1);
        cp_async_zfill16(st + idx_bytes + (sub * CS + q * 4) * 4, table + (cloud0 + ptc) * ld + CP + c_glob, 16);
        if (MODE == 0 && stats && lane <= PPW)
            cp_async4(st + idx_bytes + PPW * CS * 4 + lane * 4, rev_ptr + cloud0 + min(pt0 + lane, p_end));
        cp_async_commit();
    };
    if (lane < 4) {                                                    // slack behind the index block of both buffers
        reinterpret_cast<int*>(stage0)[PPW * k + lane] = 0;
        reinterpret_cast<int*>(stage0 + stage_bytes)[PPW * k + lane] = 0;
    }
    stage(0, 0);

    // ---- stage the slice as KEYS: the sign bit is flipped on channels that take the minimum (gamma < 0), so the
    //      hot loop is a pure arg-max with no per-edge sign handling; sums of keys are negated back exactly.
    //      Chunk q of every row belongs to the same channels as this lane's q (LPR divides the thread count).
    {
        const float* src = table + cloud0 * ld + cs * CS + q * 4;
        constexpr int RPP = GS_THREADS / LPR;          // rows per pass
        int r = threadIdx.x / LPR;
        for (; r + 3 * RPP < N; r += 4 * RPP) {        // four independent 16-byte loads in flight per thread
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint4*>(src + (long long)(r + u * RPP) * ld));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                v[u].x ^= flip[0]; v[u].y ^= flip[1]; v[u].z ^= flip[2]; v[u].w ^= flip[3];
                *reinterpret_cast<uint4*>(srow + (r + u * RPP) * CS + q * 4) = v[u];
            }
        }
        for (; r < N; r += RPP) {
            uint4 v = __ldg(reinterpret_cast<const uint4*>(src + (long long)r * ld));
            v.x ^= flip[0]; v.y ^= flip[1]; v.z ^= flip[2]; v.w ^= flip[3];
            *reinterpret_cast<uint4*>(srow + r * CS + q * 4) = v;
        }
    }
    __syncthreads();

    double s1[4] = {0.0, 0.0, 0.0, 0.0}, s2[4] = {0.0, 0.0, 0.0, 0.0};
    const float kf = (float)k;
    const int n_iter = (p_end - p_begin + G - 1) / G;              // uniform over the CTA
    for (int it = 0; it < n_iter; ++it) {
        if (it + 1 < n_iter) { stage(it + 1, (it + 1) & 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncwarp();
        const unsigned char* st = stage0 + (it & 1) * stage_bytes;
        const int* si = reinterpret_cast<const int*>(st) + sub * k;
        const int pt_raw = p_begin + it * G + warp * PPW + sub;
        const bool live = pt_raw < p_end;
        const int pt = live ? pt_raw : p_end - 1;

        float best[4], S1[4];
        int barg[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { best[e] = -INFINITY; barg[e] = 0; S1[e] = 0.f; }
        const int jlast = si[k - 1];
        const bool k4 = (k & 3) == 0;                              // uniform
#pragma unroll
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int t0 = ch * 4;
            if (t0 < k) {                                          // uniform
                const bool full = t0 + 4 <= k;                     // uniform
                float4 a[4];
                int jn[4];
                if (k4) {                                          // 16-byte aligned index rows: one LDS.128 per chunk
                    const int4 jv = *reinterpret_cast<const int4*>(si + t0);
                    jn[0] = jv.x; jn[1] = jv.y; jn[2] = jv.z; jn[3] = jv.w;
                } else {
#pragma unroll
                    for (int u = 0; u < 4; ++u) jn[u] = si[t0 + u];   // tail: slack / next row, replaced below
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!full && t0 + u >= k) jn[u] = jlast;
                    a[u] = *reinterpret_cast<const float4*>(srow + jn[u] * CS + q * 4);
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float k0 = reinterpret_cast<const float*>(&a[0])[e];
                    const float k1 = reinterpret_cast<const float*>(&a[1])[e];
                    const float k2 = reinterpret_cast<const float*>(&a[2])[e];
                    const float k3 = reinterpret_cast<const float*>(&a[3])[e];
                    if (MODE == 0) {
                        if (full) S1[e] += (k0 + k1) + (k2 + k3);
                        else      S1[e] += k0 + (t0 + 1 < k ? k1 : 0.f) + (t0 + 2 < k ? k2 : 0.f) + (t0 + 3 < k ? k3 : 0.f);
                    }
                    const bool p01 = k1 > k0, p23 = k3 > k2;
                    const float m01 = fmaxf(k0, k1), m23 = fmaxf(k2, k3);
                    const bool pm = m23 > m01;
                    const float m = fmaxf(m01, m23);
                    const int am = pm ? (p23 ? 3 : 2) : (p01 ? 1 : 0);
                    const bool better = m > best[e];
                    best[e] = fmaxf(best[e], m);
                    barg[e] = better ? t0 + am : barg[e];
                }
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {                              // keys -> values (negation is exact)
            best[e] = __uint_as_float(__float_as_uint(best[e]) ^ flip[e]);
            S1[e] = __uint_as_float(__float_as_uint(S1[e]) ^ flip[e]);
        }
        const float4 bv = *reinterpret_cast<const float4*>(st + idx_bytes + (sub * CS + q * 4) * 4);
        const float bi[4] = {bv.x, bv.y, bv.z, bv.w};
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
                const uint4 av = *reinterpret_cast<const uint4*>(srow + pt * CS + q * 4);
                const float ai[4] = {__uint_as_float(av.x ^ flip[0]), __uint_as_float(av.y ^ flip[1]),
                                     __uint_as_float(av.z ^ flip[2]), __uint_as_float(av.w ^ flip[3])};
                const int* sd = reinterpret_cast<const int*>(st + idx_bytes + PPW * CS * 4);
                const double deg = (double)(sd[sub + 1] - sd[sub]), kd = (double)k;
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
        __syncwarp();                                              // buffer (it & 1) is refilled by the next stage()
    }

    if (MODE == 0 && stats) {
        __syncthreads();                                           // every gather is done: reuse the slice as scratch
        int chans[4] = {c_glob, c_glob + 1, c_glob + 2, c_glob + 3};
        const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
        const unsigned lin = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        fs_stats_commit_impl<4>(reinterpret_cast<double*>(gs_smem), s1, s2, chans, LPR, CP, stats, lin, nblocks);
        if (blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x < LPR) {   // every slice publishes the pivots of its channels
#pragma unroll
            for (int e = 0; e < 4; ++e) stats[2 * CP + c_glob + e] = (double)pa[e] + (double)pb[e];
        }
    }
}

constexpr size_t GS_SMEM_CAP = 220 * 1024;
constexpr size_t GS_SCRATCH = 2 * 4 * GS_THREADS * sizeof(double);   // fs_stats_commit buffer

size_t gs_smem_bytes(int N, int k, int cs, int* table_bytes) {
    size_t tb = (size_t)N * cs * 4;
    if (tb < GS_SCRATCH) tb = GS_SCRATCH;
    *table_bytes = (int)tb;
    return tb + (size_t)(GS_THREADS / 32) * 2 * gs_stage_bytes(32 / (cs / 4), k, cs);
}

int pick_cs(int N, int k) {
    for (int cs = 16; cs >= 8; cs >>= 1) {
        int tb;
        if (gs_smem_bytes(N, k, cs, &tb) <= GS_SMEM_CAP) return cs;
    }
    return 0;
}

template <int CS, int NCHUNK, int MODE, typename OT>
int launch(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k, int CP,
           const float* gc, const int32_t* rev_ptr, float* sel, uint8_t* arg, float* sy, double* stats, OT* out,
           int ld_out) {
    int table_bytes;
    const size_t smem = gs_smem_bytes(N, k, CS, &table_bytes);
    auto kern = edgeconv_gather_smem_kernel<CS, NCHUNK, MODE, OT>;
    {   // per launch: the attribute is per device and the caller may use several devices from one process
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GS_SMEM_CAP);
        if (e != cudaSuccess) return (int)e;
    }
    const int slices = CP / CS;
    int split = 1;                       // small batches: several CTAs share a slice (each stages it again from L2)
    while (split < 8 && (long long)B * slices * split * 2 <= FS_NUM_SMS && N / (split * 2) >= GS_THREADS / (CS / 4)) split *= 2;
    dim3 grid(slices, B, split);
    kern<<<grid, GS_THREADS, smem, stream>>>(table, ld, idx, N, k, CP, gc, rev_ptr, sel, arg, sy, stats, out, ld_out,
                                             table_bytes);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <int MODE, typename OT>
int dispatch(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k, int CP,
             const float* gc, const int32_t* rev_ptr, float* sel, uint8_t* arg, float* sy, double* stats, OT* out,
             int ld_out) {
    const int cs = pick_cs(N, k);
    if (cs == 0 || B > 65535 || k > 64) return FS_SMEM_GATHER_UNSUPPORTED;
#define GS_GO(CS, NCH) \
    return launch<CS, NCH, MODE, OT>(stream, table, ld, idx, B, N, k, CP, gc, rev_ptr, sel, arg, sy, stats, out, ld_out)
    if (cs == 16) {
        if (k <= 20) GS_GO(16, 5);
        if (k <= 40) GS_GO(16, 10);
        GS_GO(16, 16);
    } else {
        if (k <= 20) GS_GO(8, 5);
        if (k <= 40) GS_GO(8, 10);
        GS_GO(8, 16);
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
