// Fused two-layer EdgeConv on raw coordinates (ec1 of DGCNNSeg: models/dgcnn.py:119 `EdgeConv(in, [64, 64])` with the
// 3-channel coordinate input, loop at :237-241): the P*k x 64 hidden tensor H, the P*k x 64 pre-activation Z and their
// gradients are never written to memory.
//
//   layer 1 (6 -> 64, BatchNorm + LeakyReLU) is RECOMPUTED from the coordinates wherever it is needed: its batch
//            statistics follow from the 27 moments of the edge vectors (edge3.cu), y = w.e is 6 FMAs per value;
//   layer 2 (64 -> 64) runs on the 5th-generation tensor cores: producer warps write bf16 tiles of H (edges x channels,
//            SWIZZLE_128B canonical layout) into shared memory, one elected thread issues tcgen05.mma with the
//            TRANSPOSED product  Z^T (channels x edges) = W2' (channels x 64) . H^T, so that in the TMEM accumulator a
//            lane is an output channel and a column is an edge: the max over the k edges of a point (and sum z, sum z^2 for
//            the second BatchNorm) is a per-thread reduction over registers read with tcgen05.ld - no shuffles.
//            Two groups of points share one M = 128 MMA through a block-diagonal weight operand (K = 128).
//            W2' = sign(gamma2) . W2: LeakyReLU(BN(.)) is monotone per channel with the sign of gamma, so only the max of
//            the sign-flipped pre-activation is needed.
//   training additionally accumulates the Gram matrix S = sum_e h_e h_e^T on the tensor cores (the SAME shared-memory tiles
//            read as MN-major operands, accumulator resident in TMEM for the whole kernel) and sum_e h_e: the BatchNorm-2
//            coupling of the backward pass (every edge receives -(gamma/sigma)/M (dbeta + zhat dgamma)) needs exactly these.
#include "fs_common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcx;

constexpr int E2_C = 64;                 // hidden width = output width
constexpr int E2_G = 2;                  // point groups per unit (block-diagonal M = 128 operand)
constexpr int E2_PROD_WARPS = 8;
constexpr int E2_EPI_WARPS = 8;
constexpr int E2_PROD_THREADS = E2_PROD_WARPS * 32;
constexpr int E2_WARP_MMA = E2_PROD_WARPS + E2_EPI_WARPS;
constexpr int E2_THREADS = (E2_PROD_WARPS + E2_EPI_WARPS + 1) * 32;      // 544
constexpr int E2_ATOM_ROWS_A = 128;      // rows of the weight operand
constexpr int E2_GRAM_COL = 384;         // TMEM columns [384, 512): Gram accumulator; [0, 2 * NE): forward accumulators

__host__ __device__ constexpr int e2_gcd(int a, int b) { return b == 0 ? a : e2_gcd(b, a % b); }
// points per group: the largest even PP with PP * k <= 160 edges and PP * k a multiple of 16 (UMMA N)
__host__ __device__ constexpr int e2_pp(int k) {
    int step = 4 / e2_gcd(k / 4, 4);
    if (step < 2) step = 2;
    return (160 / k) / step * step;
}

template <int K>
struct E2Cfg {
    static constexpr int PP = e2_pp(K);
    static constexpr int NE = PP * K;                              // edges per group = UMMA N
    static constexpr int UNIT_PTS = E2_G * PP;
    static constexpr int UNIT_EDGES = E2_G * NE;
    static constexpr int A_BYTES = 2 * E2_ATOM_ROWS_A * 128;       // two K atoms of the block-diagonal weight
    static constexpr int H_ATOM = NE * 128;                        // one group's H tile
    static constexpr int H_STAGE = E2_G * H_ATOM;
    static constexpr int D_STAGE = UNIT_EDGES * 16;                // float4 per edge
    static constexpr int B_STAGE = UNIT_PTS * E2_C * 4;            // w[3:6].x_i per point and channel
    static constexpr int OFF_H = A_BYTES;
    static constexpr int OFF_D = OFF_H + 2 * H_STAGE;
    static constexpr int OFF_B = OFF_D + 2 * D_STAGE;
    static constexpr int OFF_PAR = OFF_B + 2 * B_STAGE;            // w1 [64][6] | mu | sc | be
    static constexpr int OFF_RED = OFF_PAR + (E2_C * 6 + 3 * E2_C) * 4;
    static constexpr int OFF_BAR = OFF_RED + 2 * E2_THREADS * 8;
    static constexpr int SMEM = OFF_BAR + 128 + 1024;              // + alignment slack
    static_assert(PP >= 2 && PP % 2 == 0 && NE % 16 == 0 && NE <= 160 && NE >= 16, "unsupported k");
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}

// Coordinates and centre terms of one unit: registers of the producer threads (prefetched one unit ahead).
template <int K>
struct E2Stage {
    float4 d[2];          // edges E = tid, tid + 256: (x_j - x_i, valid)
    float base[3][4];     // points tid / 16 (+16, +32), channels (tid % 16) * 4 .. +4: w[3:6] . x_i
};

template <int K>
__device__ __forceinline__ void e2_prefetch(E2Stage<K>& st, const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx,
                                            long long P, int N, long long p0, const float* __restrict__ par, int tid) {
    using Cfg = E2Cfg<K>;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int E = tid + i * E2_PROD_THREADS;
        st.d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (E < Cfg::UNIT_EDGES) {
            const int g = E / Cfg::NE, e = E - g * Cfg::NE;
            const int pl = e / K, t = e - pl * K;
            const long long pt = p0 + g * Cfg::PP + pl;
            if (pt < P) {
                const long long cloud0 = (pt / N) * N;
                const long long j = cloud0 + __ldg(idx + pt * K + t);
                const float xi0 = __ldg(x + pt * ldx), xi1 = __ldg(x + pt * ldx + 1), xi2 = __ldg(x + pt * ldx + 2);
                st.d[i] = make_float4(__ldg(x + j * ldx) - xi0, __ldg(x + j * ldx + 1) - xi1, __ldg(x + j * ldx + 2) - xi2, 1.f);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const int ptl = (tid >> 4) + 16 * i;
        const int c0 = (tid & 15) * 4;
        if (ptl < Cfg::UNIT_PTS) {
            const long long pt = p0 + ptl;
            float x0 = 0.f, x1 = 0.f, x2 = 0.f;
            if (pt < P) { x0 = __ldg(x + pt * ldx); x1 = __ldg(x + pt * ldx + 1); x2 = __ldg(x + pt * ldx + 2); }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float* w = par + (c0 + c) * 6;
                st.base[i][c] = fmaf(w[5], x2, fmaf(w[4], x1, w[3] * x0));      // edge3_hidden_kernel's arithmetic
            }
        }
    }
}

// TRAIN: batch statistics of layer 2, Gram matrix and column sums of H.
template <int K, bool TRAIN>
__global__ void __launch_bounds__(E2_THREADS, 1)
edge2_fwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx, long long P, int N,
                 const float* __restrict__ w1, const float* __restrict__ coef1, const float* __restrict__ w2,
                 const float* __restrict__ gamma2, float* __restrict__ sel, uint8_t* __restrict__ arg,
                 double* __restrict__ stats, float* __restrict__ gram, float* __restrict__ hsum) {
    using Cfg = E2Cfg<K>;
    constexpr int NE = Cfg::NE, PP = Cfg::PP;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* par = reinterpret_cast<float*>(smem + Cfg::OFF_PAR);            // w1 [64][6] | mu [64] | sc [64] | be [64]
    double* red = reinterpret_cast<double*>(smem + Cfg::OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* h_full = bars;            // [2] producers -> MMA
    uint64_t* h_empty = bars + 2;       // [2] MMA -> producers
    uint64_t* acc_full = bars + 4;      // [2] MMA -> epilogue
    uint64_t* acc_empty = bars + 6;     // [2] epilogue -> MMA
    uint64_t* gram_full = bars + 8;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_units = (P + Cfg::UNIT_PTS - 1) / Cfg::UNIT_PTS;

    // ---- one-time setup: parameters, block-diagonal bf16 weight operand (sign of gamma2 folded in), barriers, TMEM
    for (int i = tid; i < E2_C * 6; i += E2_THREADS) par[i] = __ldg(w1 + i);
    for (int i = tid; i < E2_C; i += E2_THREADS) {
        par[E2_C * 6 + i] = __ldg(coef1 + i);                   // mu
        par[E2_C * 7 + i] = __ldg(coef1 + 2 * E2_C + i);        // scale = gamma / sigma
        par[E2_C * 8 + i] = __ldg(coef1 + 3 * E2_C + i);        // beta
    }
    // A operand: rows 0..63 = [W2' | 0], rows 64..127 = [0 | W2'] (K = 128 = two 64-wide atoms), K-major SWIZZLE_128B
    for (int i = tid; i < 128 * 16; i += E2_THREADS) {
        const int row = i >> 4, chunk16 = i & 15;               // 16 chunks of 8 bf16 per 256-byte row
        const int atom = chunk16 >> 3, chunk = chunk16 & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((row >> 6) == atom) {
            const int c2 = row & 63;
            const float sg = __ldg(gamma2 + c2) >= 0.f ? 1.f : -1.f;
            const float* wr = w2 + c2 * E2_C + chunk * 8;
            v.x = pack_bf16(sg * __ldg(wr), sg * __ldg(wr + 1));
            v.y = pack_bf16(sg * __ldg(wr + 2), sg * __ldg(wr + 3));
            v.z = pack_bf16(sg * __ldg(wr + 4), sg * __ldg(wr + 5));
            v.w = pack_bf16(sg * __ldg(wr + 6), sg * __ldg(wr + 7));
        }
        *reinterpret_cast<uint4*>(smem + atom * (E2_ATOM_ROWS_A * 128) + sw128_offset(row, chunk)) = v;
    }
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(h_full + s), E2_PROD_WARPS);
            mbar_init(smem_u32(h_empty + s), 1);
            mbar_init(smem_u32(acc_full + s), 1);
            mbar_init(smem_u32(acc_empty + s), E2_EPI_WARPS);
        }
        mbar_init(smem_u32(gram_full), 1);
        mbar_init_fence();
    }
    if (warp == E2_WARP_MMA) tmem_alloc512(smem_u32(tmem_slot));
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    double st_s1 = 0.0, st_s2 = 0.0;      // statistics of this thread's channel (epilogue threads)
    int st_chan = tid & (E2_C - 1);

    if (warp < E2_PROD_WARPS) {
        // ===================== producers: H tiles of the units of this CTA =====================
        const int chunk = tid & 7;                    // channels 8 * chunk .. + 8
        const int el = tid >> 3;                      // edge lane 0..31
        float w0[8], w1r[8], w2r[8], mu[8], sc[8], be[8], hs[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int ch = chunk * 8 + c;
            w0[c] = par[ch * 6]; w1r[c] = par[ch * 6 + 1]; w2r[c] = par[ch * 6 + 2];
            mu[c] = par[E2_C * 6 + ch]; sc[c] = par[E2_C * 7 + ch]; be[c] = par[E2_C * 8 + ch];
            hs[c] = 0.f;
        }
        E2Stage<K> pre;
        long long u = blockIdx.x;
        if (u < n_units) e2_prefetch<K>(pre, x, ldx, idx, P, N, u * Cfg::UNIT_PTS, par, tid);
        int s = 0;
        uint32_t ph = 0;
        for (; u < n_units; u += gridDim.x) {
            float4* d_s = reinterpret_cast<float4*>(smem + Cfg::OFF_D + s * Cfg::D_STAGE);
            float* b_s = reinterpret_cast<float*>(smem + Cfg::OFF_B + s * Cfg::B_STAGE);
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int E = tid + i * E2_PROD_THREADS;
                if (E < Cfg::UNIT_EDGES) d_s[E] = pre.d[i];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const int ptl = (tid >> 4) + 16 * i;
                if (ptl < Cfg::UNIT_PTS)
                    *reinterpret_cast<float4*>(b_s + ptl * E2_C + (tid & 15) * 4) =
                        make_float4(pre.base[i][0], pre.base[i][1], pre.base[i][2], pre.base[i][3]);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(E2_PROD_THREADS) : "memory");
            if (u + gridDim.x < n_units) e2_prefetch<K>(pre, x, ldx, idx, P, N, (u + gridDim.x) * Cfg::UNIT_PTS, par, tid);
            mbar_wait(smem_u32(h_empty + s), ph ^ 1);                  // the MMAs that read this stage have retired
            uint8_t* h_s = smem + Cfg::OFF_H + s * Cfg::H_STAGE;
            for (int E = el; E < Cfg::UNIT_EDGES; E += 32) {
                const int g = E / NE, e = E - g * NE;
                const int pl = e / K;
                const float4 dd = d_s[E];
                const float4 ba = *reinterpret_cast<const float4*>(b_s + (g * PP + pl) * E2_C + chunk * 8);
                const float4 bb = *reinterpret_cast<const float4*>(b_s + (g * PP + pl) * E2_C + chunk * 8 + 4);
                const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                float o[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float y = fmaf(w2r[c], dd.z, fmaf(w1r[c], dd.y, fmaf(w0[c], dd.x, bs[c])));
                    const float z = fmaf(sc[c], y - mu[c], be[c]);
                    o[c] = dd.w != 0.f ? fs_leaky(z) : 0.f;
                    if (TRAIN) hs[c] += o[c];
                }
                uint4 v;
                v.x = pack_bf16(o[0], o[1]); v.y = pack_bf16(o[2], o[3]); v.z = pack_bf16(o[4], o[5]); v.w = pack_bf16(o[6], o[7]);
                *reinterpret_cast<uint4*>(h_s + g * Cfg::H_ATOM + sw128_offset(e, chunk)) = v;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(h_full + s));
            if (++s == 2) { s = 0; ph ^= 1; }
        }
        if (TRAIN) {
            // column sums of H: per-thread partials -> shared (reusing the parameter block is not safe: other producers may
            // still read it) -> the statistics scratch is free until the commit at the end
            float* hred = reinterpret_cast<float*>(red);
            asm volatile("bar.sync 1, %0;" ::"n"(E2_PROD_THREADS) : "memory");
            if (tid < E2_C) hred[tid] = 0.f;
            asm volatile("bar.sync 1, %0;" ::"n"(E2_PROD_THREADS) : "memory");
#pragma unroll
            for (int c = 0; c < 8; ++c) atomicAdd(hred + chunk * 8 + c, hs[c]);
            asm volatile("bar.sync 1, %0;" ::"n"(E2_PROD_THREADS) : "memory");
            if (tid < E2_C) atomicAdd(hsum + tid, hred[tid]);
        }
    } else if (warp == E2_WARP_MMA) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC_FWD = instr_desc_f16(128, NE, 1, 0, 0);
        constexpr uint32_t IDESC_GRAM = instr_desc_f16(128, 128, 1, 1, 1);
        const uint32_t a_base = smem_u32(smem);
        int s = 0;
        uint32_t ph = 0;
        bool first = true;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            mbar_wait(smem_u32(h_full + s), ph);
            mbar_wait(smem_u32(acc_empty + s), ph ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t h_s = smem_u32(smem + Cfg::OFF_H + s * Cfg::H_STAGE);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t da = smem_desc_sw128(a_base + (ks >> 2) * (E2_ATOM_ROWS_A * 128) + (ks & 3) * 32, 16, 1024);
                    const uint64_t db = smem_desc_sw128(h_s + (ks >> 2) * Cfg::H_ATOM + (ks & 3) * 32, 16, 1024);
                    umma_ss(tmem_base + s * NE, da, db, IDESC_FWD, ks ? 1u : 0u);
                }
                umma_commit(smem_u32(acc_full + s));
                if (TRAIN) {
                    // Gram of the two groups' channels over the edges of this unit: the H tile as MN-major operand on both
                    // sides (M = N = 128 channels = two 64-wide atoms H_ATOM bytes apart, K = 16 edges = two 8-row groups)
#pragma unroll 2
                    for (int ks = 0; ks < NE / 16; ++ks) {
                        const uint64_t dh = smem_desc_sw128(h_s + ks * 2048, Cfg::H_ATOM, 1024);
                        umma_ss(tmem_base + E2_GRAM_COL, dh, dh, IDESC_GRAM, (first && ks == 0) ? 0u : 1u);
                    }
                }
                umma_commit(smem_u32(h_empty + s));
            }
            __syncwarp();
            first = false;
            if (++s == 2) { s = 0; ph ^= 1; }
        }
        if (TRAIN && elect_one_sync()) umma_commit(smem_u32(gram_full));
        __syncwarp();
    } else {
        // ===================== epilogue: TMEM lane = (group, output channel), column = edge =====================
        const int ew = warp - E2_PROD_WARPS;
        const int quarter = ew & 3;                   // TMEM lane quarter this warp may access
        const int half = ew >> 2;                     // which half of the group's points
        const int r = quarter * 32 + lane;            // accumulator row
        const int g = r >> 6, c = r & 63;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const float sgn = __ldg(gamma2 + c) >= 0.f ? 1.f : -1.f;
        st_chan = c;
        float piv = 0.f;
        bool have_piv = false;
        double n_seen = 0.0;
        int s = 0;
        uint32_t ph = 0;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            mbar_wait(smem_u32(acc_full + s), ph);
            tc_fence_after();
            const long long pt0 = u * Cfg::UNIT_PTS + g * PP + half * (PP / 2);
#pragma unroll 1
            for (int pl = 0; pl < PP / 2; ++pl) {
                float v[K];
                const uint32_t col = tmem_base + lane_base + (uint32_t)(s * NE + (half * (PP / 2) + pl) * K);
#pragma unroll
                for (int q4 = 0; q4 < K / 4; ++q4) tmem_ld4_nowait(col + q4 * 4, v + q4 * 4);
                tmem_ld_wait();
                const long long pt = pt0 + pl;
                if (pt < P) {                          // warp-uniform
                    float best = v[0];
                    int barg = 0;
#pragma unroll
                    for (int t = 1; t < K; ++t) {
                        const bool better = v[t] > best;
                        best = better ? v[t] : best;
                        barg = better ? t : barg;
                    }
                    sel[pt * E2_C + c] = sgn * best;
                    arg[pt * E2_C + c] = (uint8_t)barg;
                    if (TRAIN) {
                        if (!have_piv) { piv = v[0]; have_piv = true; }
                        float f1 = 0.f, f2 = 0.f;
#pragma unroll
                        for (int t = 0; t < K; ++t) {
                            const float ys = v[t] - piv;
                            f1 += ys;
                            f2 = fmaf(ys, ys, f2);
                        }
                        st_s1 += (double)f1;
                        st_s2 += (double)f2;
                        n_seen += (double)K;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + s));
            if (++s == 2) { s = 0; ph ^= 1; }
        }
        if (TRAIN) {
            // shifted sums -> plain sums of the UNFLIPPED pre-activation z = sgn * z' (pivot 0 in the statistics buffer)
            const double p = (double)piv;
            const double sum1 = st_s1 + n_seen * p;
            const double sum2 = st_s2 + 2.0 * p * st_s1 + n_seen * p * p;
            st_s1 = (double)sgn * sum1;
            st_s2 = sum2;
            // Gram accumulator: rows = (group, channel), columns = (group', channel'); the two diagonal blocks are the
            // groups' contributions to S = sum_e h_e h_e^T
            if (half == 0) {
                mbar_wait(smem_u32(gram_full), 0);
                tc_fence_after();
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    float gv[16];
                    tmem_ld16_nowait(tmem_base + lane_base + (uint32_t)(E2_GRAM_COL + g * 64 + q * 16), gv);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) atomicAdd(gram + c * E2_C + q * 16 + i, gv[i]);
                }
            }
        } else {
            st_s1 = 0.0; st_s2 = 0.0;
        }
    }
    if (TRAIN) {
        if (warp < E2_PROD_WARPS || warp == E2_WARP_MMA) { st_s1 = 0.0; st_s2 = 0.0; }
        __syncthreads();                                  // the statistics scratch held the column sums of H
        fs_stats_commit<1>(red, &st_s1, &st_s2, &st_chan, E2_C, E2_C, stats);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == E2_WARP_MMA) tmem_dealloc512(tmem_base);
}

template <int K>
int e2_launch_fwd(cudaStream_t stream, const float* x, int ldx, const int32_t* idx, long long P, int N, const float* w1,
                  const float* coef1, const float* w2, const float* gamma2, float* sel, uint8_t* arg, double* stats,
                  float* gram, float* hsum) {
    using Cfg = E2Cfg<K>;
    const long long n_units = (P + Cfg::UNIT_PTS - 1) / Cfg::UNIT_PTS;
    const int grid = (int)(n_units < FS_NUM_SMS ? n_units : FS_NUM_SMS);
    if (stats) {
        FS_CUDA_TRY(cudaFuncSetAttribute(edge2_fwd_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        edge2_fwd_kernel<K, true><<<grid, E2_THREADS, Cfg::SMEM, stream>>>(x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg,
                                                                          stats, gram, hsum);
    } else {
        FS_CUDA_TRY(cudaFuncSetAttribute(edge2_fwd_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        edge2_fwd_kernel<K, false><<<grid, E2_THREADS, Cfg::SMEM, stream>>>(x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg,
                                                                           nullptr, nullptr, nullptr);
    }
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

}  // namespace

extern "C" int fs_edge2_supported(int k, int C1, int C2) {
    return (C1 == 64 && C2 == 64 && (k == 8 || k == 12 || k == 16 || k == 20 || k == 40)) ? 1 : 0;
}

extern "C" int fs_edge2_fwd(int device, fs_stream_t stream_, const float* x, int ldx, const int32_t* idx, int B, int N, int k,
                            const float* w1, const float* coef1, const float* w2, int C2, const float* gamma2, float* sel,
                            uint8_t* arg, double* stats, float* gram, float* hsum) {
    if (!x || !idx || !w1 || !coef1 || !w2 || !gamma2 || !sel || !arg || B <= 0 || N <= 0 || ldx < 3) return FS_ERR_BAD_ARG;
    if (stats && (!gram || !hsum)) return FS_ERR_BAD_ARG;
    if (!fs_edge2_supported(k, 64, C2)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    switch (k) {
        case 8: return e2_launch_fwd<8>(stream, x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg, stats, gram, hsum);
        case 12: return e2_launch_fwd<12>(stream, x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg, stats, gram, hsum);
        case 16: return e2_launch_fwd<16>(stream, x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg, stats, gram, hsum);
        case 20: return e2_launch_fwd<20>(stream, x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg, stats, gram, hsum);
        case 40: return e2_launch_fwd<40>(stream, x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg, stats, gram, hsum);
    }
    return FS_ERR_UNSUPPORTED;
}
