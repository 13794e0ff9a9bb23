// Fused two-layer EdgeConv on raw coordinates (ec1 of DGCNNSeg: models/dgcnn.py:119 `EdgeConv(in, [64, 64])` with the
// 3-channel coordinate input, loop at :237-241): the P*k x 64 hidden tensor H, the P*k x 64 pre-activation Z and their
// gradients are never written to memory.
//
//   layer 1 (6 -> 64, BatchNorm + LeakyReLU) is RECOMPUTED from the coordinates wherever it is needed: its batch
//            statistics follow from the 27 moments of the edge vectors (edge3.cu); with BatchNorm folded into the weights a
//            hidden value is three FMAs, issued two channels at a time (fma.rn.f32x2 -> FFMA2);
//   layer 2 (64 -> 64) runs on the 5th-generation tensor cores: producer warps write bf16 tiles of H (edges x channels,
//            SWIZZLE_128B canonical layout) into shared memory, one elected thread issues tcgen05.mma with the
//            TRANSPOSED product  Z^T (channels x edges) = W2' (channels x 64) . H^T, so that in the TMEM accumulator a
//            lane is an output channel and a column is an edge: the max over the k edges of a point is a per-thread
//            reduction over registers read with tcgen05.ld - no shuffles. Two groups of points share one M = 128 MMA
//            through a block-diagonal weight operand (K = 128).
//            W2' = sign(gamma2) . W2: LeakyReLU(BN(.)) is monotone per channel with the sign of gamma, so only the max of
//            the sign-flipped pre-activation is needed.
//   training accumulates the Gram matrix S = sum_e h_e h_e^T on the tensor cores as well (the SAME shared-memory tiles
//            read as MN-major operands, accumulator resident in TMEM for the whole kernel) and sum_e h_e. They give the
//            batch statistics of layer 2 without touching the edges again (sum z = W2 sum h, sum z_c^2 = w_c^T S w_c), and
//            the BatchNorm-2 coupling of the backward pass (every edge receives -(gamma/sigma)/M (dbeta + zhat dgamma))
//            needs exactly these two.
//   backward: dz_e = R_e + alpha + beta' (.) z_e (arg-max routed gradient + coupling) and z_e = W2 h_e give
//                 dh_e  = [W2^T | Gm] . [R_e ; h_e] + a0,        Gm = W2^T diag(beta') W2,  a0 = W2^T alpha
//                 dW2   = sum_e R_e h_e^T + alpha hsum^T + diag(beta') W2 S
//            one tensor-core product per tile gives dH^T (channels x edges) from the bf16 tiles [R | H] the producers
//            rebuild, and the same tiles read as MN-major operands accumulate sum_e R_e h_e^T in TMEM. The epilogue
//            (TMEM lane = hidden channel, column = edge) applies LeakyReLU' of layer 1 and accumulates what edge3_bwd_kernel
//            accumulates from a materialised dH: sum t, sum t (y - mu) and the 64 x 6 moments sum t (x) e.
#include "fs_common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcx;

constexpr int E2_C = 64;                 // hidden width = output width
constexpr int E2_G = 2;                  // point groups per unit (block-diagonal M = 128 operand)
constexpr int E2_PROD_WARPS = 12;
constexpr int E2_EPI_WARPS = 8;
constexpr int E2_PROD_THREADS = E2_PROD_WARPS * 32;                      // 384
constexpr int E2_EDGE_LANES = E2_PROD_THREADS / 8;                       // 48 edges in flight, 8 channel chunks each
constexpr int E2_WARP_MMA = E2_PROD_WARPS + E2_EPI_WARPS;
constexpr int E2_THREADS = (E2_PROD_WARPS + E2_EPI_WARPS + 1) * 32;      // 672
constexpr int E2_ATOM_ROWS_A = 128;      // rows of the weight operand
constexpr int E2_GRAM_COL = 384;         // TMEM columns [384, 512): Gram / R^T H accumulator; [0, 2 * NE): tile accumulators

__host__ __device__ constexpr int e2_gcd(int a, int b) { return b == 0 ? a : e2_gcd(b, a % b); }
// points per group: the largest even PP with PP * k <= 160 edges and PP * k a multiple of 16 (UMMA N)
__host__ __device__ constexpr int e2_pp(int k) {
    int step = 4 / e2_gcd(k / 4, 4);
    if (step < 2) step = 2;
    return (160 / k) / step * step;
}

// parameter block in shared memory (floats)
constexpr int PAR_WF = 0;                       // [64][4] folded layer-1 weights sc * w[0:3] (+ pad)
constexpr int PAR_W = PAR_WF + E2_C * 4;        // [64][6] raw weights
constexpr int PAR_MU = PAR_W + E2_C * 6;
constexpr int PAR_SC = PAR_MU + E2_C;
constexpr int PAR_BE = PAR_SC + E2_C;
constexpr int PAR_A0 = PAR_BE + E2_C;           // backward only
constexpr int PAR_FLOATS = PAR_A0 + E2_C;

template <int K>
struct E2Cfg {
    static constexpr int PP = e2_pp(K);
    static constexpr int NE = PP * K;                              // edges per group = UMMA N
    static constexpr int UNIT_PTS = E2_G * PP;
    static constexpr int UNIT_EDGES = E2_G * NE;
    static constexpr int H_ATOM = NE * 128;                        // one group's bf16 tile (128-byte rows)
    static_assert(PP >= 2 && PP % 2 == 0 && NE % 16 == 0 && NE <= 160 && NE >= 16, "unsupported k");
    static_assert(UNIT_EDGES <= E2_PROD_THREADS && UNIT_PTS * 16 <= 2 * E2_PROD_THREADS, "staging item budget");
};

template <int K, int NS>
struct E2FwdCfg {
    using C = E2Cfg<K>;
    static constexpr int A_BYTES = 2 * E2_ATOM_ROWS_A * 128;       // two K atoms of the block-diagonal weight
    static constexpr int H_STAGE = E2_G * C::H_ATOM;
    static constexpr int D_STAGE = C::UNIT_EDGES * 16;             // float4 per edge
    static constexpr int B_STAGE = C::UNIT_PTS * E2_C * 4;         // folded centre term per point and channel
    static constexpr int OFF_H = A_BYTES;
    static constexpr int OFF_D = OFF_H + NS * H_STAGE;
    static constexpr int OFF_B = OFF_D + 2 * D_STAGE;
    static constexpr int OFF_PAR = OFF_B + 2 * B_STAGE;
    static constexpr int OFF_HS = OFF_PAR + PAR_FLOATS * 4;        // [64] column sums of H of this CTA
    static constexpr int OFF_BAR = OFF_HS + E2_C * 4;
    static constexpr int SMEM = OFF_BAR + 128 + 1024;              // + alignment slack
    static_assert(SMEM <= 232448, "shared memory budget");
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// packed fp32 pairs (FFMA2 / FMUL2 / FADD2 on sm_100)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// Coordinates and centre terms of one unit, in registers of the producer threads (prefetched one unit ahead).
struct E2Stage {
    float4 d;             // edge tid: (x_j - x_i, valid)
    float4 base[2];       // (point, channel quad) items tid, tid + 384: folded centre term sc (w[3:6].x_i - mu) + be
};

template <int K>
__device__ __forceinline__ void e2_prefetch(E2Stage& st, const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx,
                                            long long P, int N, long long p0, const float* __restrict__ par, int tid) {
    using Cfg = E2Cfg<K>;
    st.d = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < Cfg::UNIT_EDGES) {
        const int g = tid / Cfg::NE, e = tid - g * Cfg::NE;
        const int pl = e / K, t = e - pl * K;
        const long long pt = p0 + g * Cfg::PP + pl;
        if (pt < P) {
            const long long cloud0 = (pt / N) * N;
            const long long j = cloud0 + __ldg(idx + pt * K + t);
            const float xi0 = __ldg(x + pt * ldx), xi1 = __ldg(x + pt * ldx + 1), xi2 = __ldg(x + pt * ldx + 2);
            st.d = make_float4(__ldg(x + j * ldx) - xi0, __ldg(x + j * ldx + 1) - xi1, __ldg(x + j * ldx + 2) - xi2, 1.f);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int item = tid + i * E2_PROD_THREADS;
        st.base[i] = make_float4(0.f, 0.f, 0.f, 0.f);              // points past the end: h = LeakyReLU(0) = 0
        if (item < Cfg::UNIT_PTS * 16) {
            const int ptl = item >> 4, c0 = (item & 15) * 4;
            const long long pt = p0 + ptl;
            if (pt < P) {
                const float x0 = __ldg(x + pt * ldx), x1 = __ldg(x + pt * ldx + 1), x2 = __ldg(x + pt * ldx + 2);
                float b[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* w = par + PAR_W + (c0 + c) * 6;
                    const float yb = fmaf(w[5], x2, fmaf(w[4], x1, w[3] * x0));
                    b[c] = fmaf(par[PAR_SC + c0 + c], yb - par[PAR_MU + c0 + c], par[PAR_BE + c0 + c]);
                }
                st.base[i] = make_float4(b[0], b[1], b[2], b[3]);
            }
        }
    }
}

__device__ __forceinline__ void e2_load_params(float* par, const float* __restrict__ w1, const float* __restrict__ coef1,
                                               const float* __restrict__ a0, int tid) {
    for (int i = tid; i < E2_C * 6; i += E2_THREADS) par[PAR_W + i] = __ldg(w1 + i);
    for (int i = tid; i < E2_C; i += E2_THREADS) {
        const float sc = __ldg(coef1 + 2 * E2_C + i);
        par[PAR_MU + i] = __ldg(coef1 + i);
        par[PAR_SC + i] = sc;                                     // gamma / sigma
        par[PAR_BE + i] = __ldg(coef1 + 3 * E2_C + i);
        par[PAR_A0 + i] = a0 ? __ldg(a0 + i) : 0.f;
        par[PAR_WF + i * 4] = sc * __ldg(w1 + i * 6);
        par[PAR_WF + i * 4 + 1] = sc * __ldg(w1 + i * 6 + 1);
        par[PAR_WF + i * 4 + 2] = sc * __ldg(w1 + i * 6 + 2);
        par[PAR_WF + i * 4 + 3] = 0.f;
    }
}

// Folded layer-1 weights of a producer thread: channels 8 * chunk .. + 8 as four pairs.
struct E2ProdW {
    f32x2 w0[4], w1[4], w2[4];
    __device__ __forceinline__ void load(const float* par, int chunk) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const float* a = par + PAR_WF + (chunk * 8 + 2 * p) * 4;
            w0[p] = pk(a[0], a[4]); w1[p] = pk(a[1], a[5]); w2[p] = pk(a[2], a[6]);
        }
    }
};

// One 16-byte chunk (8 channels) of one row of H: h = LeakyReLU(w' . d + base'). hsum (nullable) accumulates the values.
__device__ __forceinline__ uint4 e2_hidden_chunk(const E2ProdW& W, const float4 dd, const float4 ba, const float4 bb, f32x2* hsum) {
    const f32x2 d0 = pk(dd.x, dd.x), d1 = pk(dd.y, dd.y), d2 = pk(dd.z, dd.z), slope = pk(0.2f, 0.2f);
    const f32x2 bs[4] = {pk(ba.x, ba.y), pk(ba.z, ba.w), pk(bb.x, bb.y), pk(bb.z, bb.w)};
    uint32_t out[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const f32x2 z = fma2(W.w2[p], d2, fma2(W.w1[p], d1, fma2(W.w0[p], d0, bs[p])));
        const f32x2 lz = mul2(z, slope);
        float z0, z1, l0, l1;
        upk(z, z0, z1);
        upk(lz, l0, l1);
        const float h0 = fmaxf(z0, l0), h1 = fmaxf(z1, l1);       // LeakyReLU(0.2)
        if (hsum) hsum[p] = add2(hsum[p], pk(h0, h1));
        out[p] = pack_bf16(h0, h1);
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

// TRAIN: Gram matrix and column sums of H (batch statistics of layer 2 and the backward coupling).
template <int K, bool TRAIN>
__global__ void __launch_bounds__(E2_THREADS, 1)
edge2_fwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx, long long P, int N,
                 const float* __restrict__ w1, const float* __restrict__ coef1, const float* __restrict__ w2,
                 const float* __restrict__ gamma2, float* __restrict__ sel, uint8_t* __restrict__ arg,
                 float* __restrict__ gram, float* __restrict__ hsum) {
    using Cfg = E2Cfg<K>;
    constexpr int NS = TRAIN ? 3 : 2;              // H stages: the Gram MMAs keep a stage busy longer
    using FC = E2FwdCfg<K, NS>;
    constexpr int NE = Cfg::NE, PP = Cfg::PP;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* par = reinterpret_cast<float*>(smem + FC::OFF_PAR);
    float* hs_s = reinterpret_cast<float*>(smem + FC::OFF_HS);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FC::OFF_BAR);
    uint64_t* h_full = bars;            // [3] producers -> MMA
    uint64_t* h_empty = bars + 3;       // [3] MMA -> producers
    uint64_t* acc_full = bars + 6;      // [2] MMA -> epilogue
    uint64_t* acc_empty = bars + 8;     // [2] epilogue -> MMA
    uint64_t* gram_full = bars + 10;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_units = (P + Cfg::UNIT_PTS - 1) / Cfg::UNIT_PTS;

    // ---- one-time setup: parameters, block-diagonal bf16 weight operand (sign of gamma2 folded in), barriers, TMEM
    e2_load_params(par, w1, coef1, nullptr, tid);
    if (tid < E2_C) hs_s[tid] = 0.f;
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
        for (int s = 0; s < 3; ++s) {
            mbar_init(smem_u32(h_full + s), E2_PROD_WARPS);
            mbar_init(smem_u32(h_empty + s), 1);
        }
        for (int s = 0; s < 2; ++s) {
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

    if (warp < E2_PROD_WARPS) {
        // ===================== producers: H tiles of the units of this CTA =====================
        const int chunk = tid & 7;                    // channels 8 * chunk .. + 8
        const int el = tid >> 3;                      // edge lane 0..47
        E2ProdW W;
        W.load(par, chunk);
        f32x2 hs[4] = {pk(0.f, 0.f), pk(0.f, 0.f), pk(0.f, 0.f), pk(0.f, 0.f)};
        E2Stage pre;
        long long u = blockIdx.x;
        if (u < n_units) e2_prefetch<K>(pre, x, ldx, idx, P, N, u * Cfg::UNIT_PTS, par, tid);
        int s = 0, sd = 0;
        uint32_t ph = 0;
        for (; u < n_units; u += gridDim.x) {
            float4* d_s = reinterpret_cast<float4*>(smem + FC::OFF_D + sd * FC::D_STAGE);
            float* b_s = reinterpret_cast<float*>(smem + FC::OFF_B + sd * FC::B_STAGE);
            if (tid < Cfg::UNIT_EDGES) d_s[tid] = pre.d;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int item = tid + i * E2_PROD_THREADS;
                // item = (point, channel quad q): chunk = q / 2, half = q % 2 -> [point][half][chunk][4], so that the two
                // 128-bit reads of a chunk are bank-conflict-free across the eight chunk lanes of a warp
                if (item < Cfg::UNIT_PTS * 16)
                    *reinterpret_cast<float4*>(b_s + (item >> 4) * E2_C + (item & 1) * 32 + ((item & 15) >> 1) * 4) = pre.base[i];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(E2_PROD_THREADS) : "memory");
            if (u + gridDim.x < n_units) e2_prefetch<K>(pre, x, ldx, idx, P, N, (u + gridDim.x) * Cfg::UNIT_PTS, par, tid);
            mbar_wait(smem_u32(h_empty + s), ph ^ 1);                  // the MMAs that read this stage have retired
            uint8_t* h_s = smem + FC::OFF_H + s * FC::H_STAGE;
            for (int E = el; E < Cfg::UNIT_EDGES; E += E2_EDGE_LANES) {
                const int g = E / NE, e = E - g * NE;
                const int pl = e / K;
                const float* bp = b_s + (g * PP + pl) * E2_C + chunk * 4;
                const uint4 v = e2_hidden_chunk(W, d_s[E], *reinterpret_cast<const float4*>(bp),
                                                *reinterpret_cast<const float4*>(bp + 32), TRAIN ? hs : nullptr);
                *reinterpret_cast<uint4*>(h_s + g * Cfg::H_ATOM + sw128_offset(e, chunk)) = v;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(h_full + s));
            if (++s == NS) { s = 0; ph ^= 1; }
            sd ^= 1;
        }
        if (TRAIN) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float a, b;
                upk(hs[p], a, b);
                atomicAdd(hs_s + chunk * 8 + 2 * p, a);
                atomicAdd(hs_s + chunk * 8 + 2 * p + 1, b);
            }
        }
    } else if (warp == E2_WARP_MMA) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC_FWD = instr_desc_f16(128, NE, 1, 0, 0);
        constexpr uint32_t IDESC_GRAM = instr_desc_f16(128, 128, 1, 1, 1);
        const uint32_t a_base = smem_u32(smem);
        int s = 0, a = 0;
        uint32_t ph = 0, pha = 0;
        bool first = true;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            mbar_wait(smem_u32(h_full + s), ph);
            mbar_wait(smem_u32(acc_empty + a), pha ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t h_s = smem_u32(smem + FC::OFF_H + s * FC::H_STAGE);
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    const uint64_t da = smem_desc_sw128(a_base + (ks >> 2) * (E2_ATOM_ROWS_A * 128) + (ks & 3) * 32, 16, 1024);
                    const uint64_t db = smem_desc_sw128(h_s + (ks >> 2) * Cfg::H_ATOM + (ks & 3) * 32, 16, 1024);
                    umma_ss(tmem_base + a * NE, da, db, IDESC_FWD, ks ? 1u : 0u);
                }
                umma_commit(smem_u32(acc_full + a));
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
            if (++s == NS) { s = 0; ph ^= 1; }
            if (++a == 2) { a = 0; pha ^= 1; }
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
        int a = 0;
        uint32_t pha = 0;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            mbar_wait(smem_u32(acc_full + a), pha);
            tc_fence_after();
            const long long pt0 = u * Cfg::UNIT_PTS + g * PP + half * (PP / 2);
#pragma unroll 1
            for (int pl = 0; pl < PP / 2; ++pl) {
                float v[K];
                const uint32_t col = tmem_base + lane_base + (uint32_t)(a * NE + (half * (PP / 2) + pl) * K);
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
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + a));
            if (++a == 2) { a = 0; pha ^= 1; }
        }
        if (TRAIN && half == 0) {
            // Gram accumulator: rows = (group, channel), columns = (group', channel'); the two diagonal blocks are the
            // groups' contributions to S = sum_e h_e h_e^T
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
    }
    tc_fence_before();
    __syncthreads();
    if (TRAIN && tid < E2_C) atomicAdd(hsum + tid, hs_s[tid]);
    if (warp == E2_WARP_MMA) tmem_dealloc512(tmem_base);
}

// Batch statistics of layer 2 from the Gram matrix and the column sums of H (bf16-rounded W2, like the tensor-core product):
//   sum_e z_c = w_c . hsum,   sum_e z_c^2 = w_c^T S w_c       -> stats[c], stats[C + c] (pivot 0), fp64
// grid = 64 channels, block = 64 threads (thread i: row i of S).
__global__ void __launch_bounds__(E2_C)
edge2_stats_kernel(const float* __restrict__ w2, const float* __restrict__ gram, const float* __restrict__ hsum,
                   double* __restrict__ stats) {
    __shared__ float s_w[E2_C];
    __shared__ double s_r[2][2];
    const int c = blockIdx.x, i = threadIdx.x;
    s_w[i] = __bfloat162float(__float2bfloat16_rn(__ldg(w2 + c * E2_C + i)));
    __syncthreads();
    double row = 0.0;
#pragma unroll 8
    for (int j = 0; j < E2_C; ++j) row += (double)__ldg(gram + i * E2_C + j) * (double)s_w[j];
    double p2 = (double)s_w[i] * row, p1 = (double)s_w[i] * (double)__ldg(hsum + i);
    p1 = fs_warp_sum(p1);
    p2 = fs_warp_sum(p2);
    if ((i & 31) == 0) { s_r[i >> 5][0] = p1; s_r[i >> 5][1] = p2; }
    __syncthreads();
    if (i == 0) {
        const double s2 = s_r[0][1] + s_r[1][1];
        stats[c] = s_r[0][0] + s_r[1][0];
        stats[E2_C + c] = s2 < 0.0 ? 0.0 : s2;
    }
}

// =====================================================================================================================
// Backward
// =====================================================================================================================
template <int K>
struct E2BwdCfg {
    using C = E2Cfg<K>;
    static constexpr int A_BYTES = 4 * E2_ATOM_ROWS_A * 128;       // four K atoms: [W2^T | Gm | 0 | 0] / [0 | 0 | W2^T | Gm]
    static constexpr int T_ATOM = C::NE * 128;
    static constexpr int T_BYTES = 4 * T_ATOM;                     // [R_a | H_a | R_b | H_b], single stage
    static constexpr int D_STAGE = C::UNIT_EDGES * 16;             // float4 (d, valid) per edge, for the producers
    static constexpr int S_STAGE = C::UNIT_EDGES * 4;              // one component per edge (SoA copy for the epilogue pairs)
    static constexpr int B_STAGE = C::UNIT_PTS * E2_C * 4;
    static constexpr int X_STAGE = C::UNIT_PTS * 16;
    static constexpr int STAGE = D_STAGE + 4 * S_STAGE + 2 * B_STAGE + X_STAGE;   // d | d0 d1 d2 valid | base' | base_y | x_i
    static constexpr int OFF_T = A_BYTES;
    static constexpr int OFF_ST = OFF_T + T_BYTES;
    static constexpr int OFF_PAR = OFF_ST + 2 * STAGE;
    static constexpr int OFF_ACC = OFF_PAR + PAR_FLOATS * 4;       // [64][6] moment sums of the CTA
    static constexpr int OFF_RED = OFF_ACC + E2_C * 6 * 4;
    static constexpr int OFF_BAR = OFF_RED + 2 * E2_THREADS * 8;
    static constexpr int SMEM = OFF_BAR + 128 + 1024;
    static_assert(SMEM <= 232448, "shared memory budget");
};

struct E2BwdStage {
    float4 basey[2];      // (point, channel quad) items: w[3:6] . x_i (unfolded, the epilogue needs y - mu)
    float4 xi;            // point tid (< UNIT_PTS)
    float4 rv[2];         // (point, channel quad) items tid + i * 384: s2 * d2
    uint32_t ra[2];       // their arg-max slots (4 x uint8)
};

template <int K>
__device__ __forceinline__ void e2_bwd_prefetch(E2BwdStage& st, const float* __restrict__ x, int ldx, long long P, long long p0,
                                                const float* __restrict__ d2, const uint8_t* __restrict__ arg,
                                                const float* __restrict__ scale2, const float* __restrict__ par, int tid) {
    using Cfg = E2Cfg<K>;
    st.xi = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < Cfg::UNIT_PTS && p0 + tid < P) {
        const long long pt = p0 + tid;
        st.xi = make_float4(__ldg(x + pt * ldx), __ldg(x + pt * ldx + 1), __ldg(x + pt * ldx + 2), 1.f);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int item = tid + i * E2_PROD_THREADS;              // (point, 4 channels): UNIT_PTS * 16 items
        st.rv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        st.basey[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        st.ra[i] = 0;
        if (item < Cfg::UNIT_PTS * 16) {
            const int ptl = item >> 4, c0 = (item & 15) * 4;
            const long long pt = p0 + ptl;
            if (pt < P) {
                const float4 dv = __ldg(reinterpret_cast<const float4*>(d2 + pt * E2_C + c0));
                const float4 sv = __ldg(reinterpret_cast<const float4*>(scale2 + c0));
                st.rv[i] = make_float4(dv.x * sv.x, dv.y * sv.y, dv.z * sv.z, dv.w * sv.w);
                st.ra[i] = __ldg(reinterpret_cast<const uint32_t*>(arg + pt * E2_C + c0));
                const float x0 = __ldg(x + pt * ldx), x1 = __ldg(x + pt * ldx + 1), x2 = __ldg(x + pt * ldx + 2);
                float b[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float* w = par + PAR_W + (c0 + c) * 6;
                    b[c] = fmaf(w[5], x2, fmaf(w[4], x1, w[3] * x0));
                }
                st.basey[i] = make_float4(b[0], b[1], b[2], b[3]);
            }
        }
    }
}

template <int K>
__global__ void __launch_bounds__(E2_THREADS, 1)
edge2_bwd_kernel(const float* __restrict__ x, int ldx, const int32_t* __restrict__ idx, long long P, int N,
                 const float* __restrict__ w1, const float* __restrict__ coef1, const float* __restrict__ w2,
                 const float* __restrict__ gm, const float* __restrict__ a0, const float* __restrict__ scale2,
                 const float* __restrict__ d2, const uint8_t* __restrict__ arg, double* __restrict__ dgb1,
                 float* __restrict__ acc1, float* __restrict__ rh) {
    using Cfg = E2Cfg<K>;
    using BC = E2BwdCfg<K>;
    constexpr int NE = Cfg::NE, PP = Cfg::PP;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* par = reinterpret_cast<float*>(smem + BC::OFF_PAR);
    float* acc_s = reinterpret_cast<float*>(smem + BC::OFF_ACC);
    double* red = reinterpret_cast<double*>(smem + BC::OFF_RED);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BC::OFF_BAR);
    uint64_t* t_full = bars;            // producers -> MMA (single tile stage)
    uint64_t* t_empty = bars + 1;       // MMA -> producers
    uint64_t* acc_full = bars + 2;      // [2]
    uint64_t* acc_empty = bars + 4;     // [2]
    uint64_t* rh_full = bars + 6;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_units = (P + Cfg::UNIT_PTS - 1) / Cfg::UNIT_PTS;

    e2_load_params(par, w1, coef1, a0, tid);
    for (int i = tid; i < E2_C * 6; i += E2_THREADS) acc_s[i] = 0.f;
    // A operand (K-major, SWIZZLE_128B), 128 rows x 256 bf16 = four 64-wide atoms:
    //   rows 0..63   (group a, hidden channel c'): atom 0 = W2^T[c'][c2] = W2[c2][c'], atom 1 = Gm[c'][:], atoms 2, 3 = 0
    //   rows 64..127 (group b)                   : atoms 0, 1 = 0, atom 2 = W2^T, atom 3 = Gm
    for (int i = tid; i < 128 * 32; i += E2_THREADS) {
        const int row = i >> 5, chunk32 = i & 31;
        const int atom = chunk32 >> 3, chunk = chunk32 & 7;
        const int grp = row >> 6, cp = row & 63;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = 0.f;
        if ((atom >> 1) == grp) {
            if ((atom & 1) == 0) {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __ldg(w2 + (chunk * 8 + e) * E2_C + cp);
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] = __ldg(gm + cp * E2_C + chunk * 8 + e);
            }
        }
        uint4 v;
        v.x = pack_bf16(f[0], f[1]); v.y = pack_bf16(f[2], f[3]); v.z = pack_bf16(f[4], f[5]); v.w = pack_bf16(f[6], f[7]);
        *reinterpret_cast<uint4*>(smem + atom * (E2_ATOM_ROWS_A * 128) + sw128_offset(row, chunk)) = v;
    }
    if (tid == 0) {
        mbar_init(smem_u32(t_full), E2_PROD_WARPS);
        mbar_init(smem_u32(t_empty), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(acc_full + s), 1);
            mbar_init(smem_u32(acc_empty + s), E2_EPI_WARPS);
        }
        mbar_init(smem_u32(rh_full), 1);
        mbar_init_fence();
    }
    if (warp == E2_WARP_MMA) tmem_alloc512(smem_u32(tmem_slot));
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    double st_s1 = 0.0, st_s2 = 0.0;
    int st_chan = tid & (E2_C - 1);

    if (warp < E2_PROD_WARPS) {
        // ===================== producers: [R | H] tiles =====================
        const int chunk = tid & 7;
        const int el = tid >> 3;
        E2ProdW W;
        W.load(par, chunk);
        E2Stage pre;
        E2BwdStage preb;
        long long u = blockIdx.x;
        if (u < n_units) {
            e2_prefetch<K>(pre, x, ldx, idx, P, N, u * Cfg::UNIT_PTS, par, tid);
            e2_bwd_prefetch<K>(preb, x, ldx, P, u * Cfg::UNIT_PTS, d2, arg, scale2, par, tid);
        }
        int s = 0;
        uint32_t ph = 0, pht = 0;
        uint8_t* tiles = smem + BC::OFF_T;
        for (; u < n_units; u += gridDim.x) {
            // the staging block of this parity was last read by the epilogue of unit u - 2
            mbar_wait(smem_u32(acc_empty + s), ph ^ 1);
            uint8_t* stg = smem + BC::OFF_ST + s * BC::STAGE;
            float4* d_s = reinterpret_cast<float4*>(stg);
            float* so = reinterpret_cast<float*>(stg + BC::D_STAGE);                       // d0 | d1 | d2 | valid
            float* b_s = reinterpret_cast<float*>(stg + BC::D_STAGE + 4 * BC::S_STAGE);    // folded centre term
            float* y_s = b_s + Cfg::UNIT_PTS * E2_C;                                       // unfolded centre term
            float4* x_s = reinterpret_cast<float4*>(y_s + Cfg::UNIT_PTS * E2_C);
            if (tid < Cfg::UNIT_EDGES) {
                d_s[tid] = pre.d;
                so[tid] = pre.d.x; so[Cfg::UNIT_EDGES + tid] = pre.d.y; so[2 * Cfg::UNIT_EDGES + tid] = pre.d.z;
                so[3 * Cfg::UNIT_EDGES + tid] = pre.d.w;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int item = tid + i * E2_PROD_THREADS;
                if (item < Cfg::UNIT_PTS * 16) {
                    *reinterpret_cast<float4*>(b_s + (item >> 4) * E2_C + (item & 1) * 32 + ((item & 15) >> 1) * 4) = pre.base[i];   // [pt][half][chunk][4]
                    *reinterpret_cast<float4*>(y_s + (item >> 4) * E2_C + (item & 15) * 4) = preb.basey[i];
                }
            }
            if (tid < Cfg::UNIT_PTS) x_s[tid] = preb.xi;
            // the tile stage is single-buffered: the MMAs of the previous unit must have retired
            mbar_wait(smem_u32(t_empty), pht ^ 1);
            // zero the two R atoms (atoms 0 and 2)
            for (int i = tid; i < 2 * (BC::T_ATOM / 16); i += E2_PROD_THREADS) {
                const int atom = i / (BC::T_ATOM / 16);
                *reinterpret_cast<uint4*>(tiles + (2 * atom) * BC::T_ATOM + (i - atom * (BC::T_ATOM / 16)) * 16) = make_uint4(0, 0, 0, 0);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(E2_PROD_THREADS) : "memory");
            // scatter the routed gradients: R[g][pl * K + arg][c2] = bf16(s2 * d2)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int item = tid + i * E2_PROD_THREADS;
                if (item < Cfg::UNIT_PTS * 16) {
                    const int ptl = item >> 4, c0 = (item & 15) * 4;
                    const int g = ptl / PP, pl = ptl - g * PP;
                    const float rvv[4] = {preb.rv[i].x, preb.rv[i].y, preb.rv[i].z, preb.rv[i].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int slot = (preb.ra[i] >> (8 * c)) & 0xff;
                        const int e = pl * K + (slot < K ? slot : 0);
                        const int c2 = c0 + c;
                        *reinterpret_cast<__nv_bfloat16*>(tiles + (2 * g) * BC::T_ATOM + sw128_offset(e, c2 >> 3) + (c2 & 7) * 2) =
                            __float2bfloat16_rn(rvv[c]);
                    }
                }
            }
            if (u + gridDim.x < n_units) {
                e2_prefetch<K>(pre, x, ldx, idx, P, N, (u + gridDim.x) * Cfg::UNIT_PTS, par, tid);
                e2_bwd_prefetch<K>(preb, x, ldx, P, (u + gridDim.x) * Cfg::UNIT_PTS, d2, arg, scale2, par, tid);
            }
            for (int E = el; E < Cfg::UNIT_EDGES; E += E2_EDGE_LANES) {
                const int g = E / NE, e = E - g * NE;
                const int pl = e / K;
                const float* bp = b_s + (g * PP + pl) * E2_C + chunk * 4;
                const uint4 v = e2_hidden_chunk(W, d_s[E], *reinterpret_cast<const float4*>(bp),
                                                *reinterpret_cast<const float4*>(bp + 32), nullptr);
                *reinterpret_cast<uint4*>(tiles + (2 * g + 1) * BC::T_ATOM + sw128_offset(e, chunk)) = v;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(t_full));
            pht ^= 1;
            if (++s == 2) { s = 0; ph ^= 1; }
        }
    } else if (warp == E2_WARP_MMA) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC_DH = instr_desc_f16(128, NE, 1, 0, 0);
        constexpr uint32_t IDESC_RH = instr_desc_f16(128, 128, 1, 1, 1);
        const uint32_t a_base = smem_u32(smem);
        const uint32_t t_base = smem_u32(smem + BC::OFF_T);
        int s = 0;
        uint32_t ph = 0, pht = 0;
        bool first = true;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            mbar_wait(smem_u32(t_full), pht);
            mbar_wait(smem_u32(acc_empty + s), ph ^ 1);
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < 16; ++ks) {
                    const uint64_t da = smem_desc_sw128(a_base + (ks >> 2) * (E2_ATOM_ROWS_A * 128) + (ks & 3) * 32, 16, 1024);
                    const uint64_t db = smem_desc_sw128(t_base + (ks >> 2) * BC::T_ATOM + (ks & 3) * 32, 16, 1024);
                    umma_ss(tmem_base + s * NE, da, db, IDESC_DH, ks ? 1u : 0u);
                }
                umma_commit(smem_u32(acc_full + s));
                // sum_e R_e h_e^T: M = (group, c2) from the R atoms 0 and 2, N = (group, c') from the H atoms 1 and 3,
                // K = edges (MN-major operands: 64-wide blocks 2 * T_ATOM apart, 8-edge groups 1024 bytes apart)
#pragma unroll 2
                for (int ks = 0; ks < NE / 16; ++ks) {
                    const uint64_t dr = smem_desc_sw128(t_base + ks * 2048, 2 * BC::T_ATOM, 1024);
                    const uint64_t dh = smem_desc_sw128(t_base + BC::T_ATOM + ks * 2048, 2 * BC::T_ATOM, 1024);
                    umma_ss(tmem_base + E2_GRAM_COL, dr, dh, IDESC_RH, (first && ks == 0) ? 0u : 1u);
                }
                umma_commit(smem_u32(t_empty));
            }
            __syncwarp();
            first = false;
            pht ^= 1;
            if (++s == 2) { s = 0; ph ^= 1; }
        }
        if (elect_one_sync()) umma_commit(smem_u32(rh_full));
        __syncwarp();
    } else {
        // ===================== epilogue: TMEM lane = (group, hidden channel), column = edge; two edges per step ==========
        const int ew = warp - E2_PROD_WARPS;
        const int quarter = ew & 3;
        const int half = ew >> 2;
        const int r = quarter * 32 + lane;
        const int g = r >> 6, c = r & 63;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        st_chan = c;
        const float* wc = par + PAR_W + c * 6;
        const f32x2 w0 = pk(wc[0], wc[0]), w1c = pk(wc[1], wc[1]), w2c = pk(wc[2], wc[2]);
        const float mu = par[PAR_MU + c], sc = par[PAR_SC + c], be = par[PAR_BE + c], a0c = par[PAR_A0 + c];
        const f32x2 nmu2 = pk(-mu, -mu), sc2 = pk(sc, sc), be2 = pk(be, be), a02 = pk(a0c, a0c);
        f32x2 s1p = pk(0.f, 0.f), s2p = pk(0.f, 0.f), A0p = pk(0.f, 0.f), A1p = pk(0.f, 0.f), A2p = pk(0.f, 0.f);
        float A3 = 0.f, A4 = 0.f, A5 = 0.f;
        int s = 0;
        uint32_t ph = 0;
        for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
            mbar_wait(smem_u32(acc_full + s), ph);
            tc_fence_after();
            const uint8_t* stg = smem + BC::OFF_ST + s * BC::STAGE;
            const float* so = reinterpret_cast<const float*>(stg + BC::D_STAGE);
            const float* y_s = reinterpret_cast<const float*>(stg + BC::D_STAGE + 4 * BC::S_STAGE) + Cfg::UNIT_PTS * E2_C;
            const float4* x_s = reinterpret_cast<const float4*>(y_s + Cfg::UNIT_PTS * E2_C);
#pragma unroll 1
            for (int pl = 0; pl < PP / 2; ++pl) {
                const int plg = half * (PP / 2) + pl;
                float v[K];
                const uint32_t col = tmem_base + lane_base + (uint32_t)(s * NE + plg * K);
#pragma unroll
                for (int q4 = 0; q4 < K / 4; ++q4) tmem_ld4_nowait(col + q4 * 4, v + q4 * 4);
                tmem_ld_wait();
                const float by = y_s[(g * PP + plg) * E2_C + c];
                const f32x2 by2 = pk(by, by);
                const float4 xi = x_s[g * PP + plg];
                const float* e0 = so + g * NE + plg * K;                 // K is even and the offset is even: 8-byte aligned pairs
                f32x2 dsum = pk(0.f, 0.f);
#pragma unroll
                for (int t = 0; t < K; t += 2) {
                    const f32x2 d0 = *reinterpret_cast<const f32x2*>(e0 + t);                       // broadcast reads
                    const f32x2 d1 = *reinterpret_cast<const f32x2*>(e0 + Cfg::UNIT_EDGES + t);
                    const f32x2 d2 = *reinterpret_cast<const f32x2*>(e0 + 2 * Cfg::UNIT_EDGES + t);
                    const f32x2 vl = *reinterpret_cast<const f32x2*>(e0 + 3 * Cfg::UNIT_EDGES + t);
                    const f32x2 y = fma2(w2c, d2, fma2(w1c, d1, fma2(w0, d0, by2)));
                    const f32x2 yc = add2(y, nmu2);
                    const f32x2 z = fma2(sc2, yc, be2);
                    float z0, z1;
                    upk(z, z0, z1);
                    const f32x2 slope = pk(z0 > 0.f ? 1.f : 0.2f, z1 > 0.f ? 1.f : 0.2f);
                    const f32x2 gl = mul2(add2(pk(v[t], v[t + 1]), a02), vl);       // valid = 0 for points past the end
                    const f32x2 tt = mul2(gl, slope);
                    s2p = fma2(tt, yc, s2p);
                    dsum = add2(dsum, tt);
                    A0p = fma2(tt, d0, A0p); A1p = fma2(tt, d1, A1p); A2p = fma2(tt, d2, A2p);
                }
                s1p = add2(s1p, dsum);
                float ds0, ds1;
                upk(dsum, ds0, ds1);
                const float ds = ds0 + ds1;
                A3 = fmaf(ds, xi.x, A3); A4 = fmaf(ds, xi.y, A4); A5 = fmaf(ds, xi.z, A5);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + s));
            if (++s == 2) { s = 0; ph ^= 1; }
        }
        float lo, hi;
        upk(s1p, lo, hi); st_s1 = (double)lo + (double)hi;
        upk(s2p, lo, hi); st_s2 = ((double)lo + (double)hi) * (double)__ldg(coef1 + E2_C + c);   // yhat = (y - mu) * invstd
        upk(A0p, lo, hi); atomicAdd(acc_s + c * 6 + 0, lo + hi);
        upk(A1p, lo, hi); atomicAdd(acc_s + c * 6 + 1, lo + hi);
        upk(A2p, lo, hi); atomicAdd(acc_s + c * 6 + 2, lo + hi);
        atomicAdd(acc_s + c * 6 + 3, A3); atomicAdd(acc_s + c * 6 + 4, A4); atomicAdd(acc_s + c * 6 + 5, A5);
        if (half == 0) {
            mbar_wait(smem_u32(rh_full), 0);
            tc_fence_after();
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                float gv[16];
                tmem_ld16_nowait(tmem_base + lane_base + (uint32_t)(E2_GRAM_COL + g * 64 + q * 16), gv);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) atomicAdd(rh + c * E2_C + q * 16 + i, gv[i]);
            }
        }
    }
    if (warp < E2_PROD_WARPS || warp == E2_WARP_MMA) { st_s1 = 0.0; st_s2 = 0.0; }
    fs_stats_commit<1>(red, &st_s1, &st_s2, &st_chan, E2_C, E2_C, dgb1);       // contains the __syncthreads that publishes acc_s
    for (int i = tid; i < E2_C * 6; i += E2_THREADS) atomicAdd(acc1 + i, acc_s[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == E2_WARP_MMA) tmem_dealloc512(tmem_base);
}

// Coupling coefficients of the backward pass (one CTA): from the BatchNorm-2 coefficients coef2 = (mu | 1/sigma |
// gamma/sigma | beta), the sums dgb2 = (sum d | sum d zhat) of fs_edgeconv_bwd_reduce and W2 (bf16-rounded like the
// tensor-core operand):  alpha, beta', a0 = W2^T alpha, Gm = W2^T diag(beta') W2.  ws = alpha | beta' | a0 | gm.
__global__ void __launch_bounds__(256)
edge2_bwd_prep_kernel(const float* __restrict__ w2, const float* __restrict__ coef2, const double* __restrict__ dgb2,
                      double count, int train_stats, float* __restrict__ ws) {
    // grid = 16 blocks: every block derives alpha / beta' (64 values), block b writes Gm rows 4 b .. 4 b + 4; block 0 also
    // writes alpha, beta' and a0
    __shared__ float s_w[E2_C * E2_C];
    __shared__ float s_al[E2_C], s_bp[E2_C];
    for (int i = threadIdx.x; i < E2_C * E2_C; i += blockDim.x) s_w[i] = __bfloat162float(__float2bfloat16_rn(__ldg(w2 + i)));
    if (threadIdx.x < E2_C) {
        const int c = threadIdx.x;
        float al = 0.f, bp = 0.f;
        if (train_stats) {
            const double mu = coef2[c], inv = coef2[E2_C + c], s2 = coef2[2 * E2_C + c];
            const double mb = dgb2[c] / count, mg = dgb2[E2_C + c] / count;
            al = (float)(-s2 * (mb - mu * inv * mg));
            bp = (float)(-s2 * inv * mg);
        }
        s_al[c] = al; s_bp[c] = bp;
        if (blockIdx.x == 0) { ws[c] = al; ws[E2_C + c] = bp; }
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < E2_C) {
        float a = 0.f;
        for (int c2 = 0; c2 < E2_C; ++c2) a = fmaf(s_w[c2 * E2_C + threadIdx.x], s_al[c2], a);
        ws[2 * E2_C + threadIdx.x] = a;
    }
    const int o = blockIdx.x * 256 + threadIdx.x;           // Gm[i][j] = sum_c2 W2[c2][i] beta'[c2] W2[c2][j]
    const int i = o >> 6, j = o & 63;
    float a = 0.f;
#pragma unroll 8
    for (int c2 = 0; c2 < E2_C; ++c2) a = fmaf(s_w[c2 * E2_C + i] * s_bp[c2], s_w[c2 * E2_C + j], a);
    ws[3 * E2_C + o] = a;
}

// dW2 = rh + alpha hsum^T + diag(beta') W2 S
__global__ void edge2_dw2_kernel(const float* __restrict__ w2, const float* __restrict__ ws, const float* __restrict__ rh,
                                 const float* __restrict__ gram, const float* __restrict__ hsum, int train_stats,
                                 float* __restrict__ dw2) {
    __shared__ float s_w[E2_C * E2_C];
    __shared__ float s_g[E2_C * E2_C];
    for (int i = threadIdx.x; i < E2_C * E2_C; i += blockDim.x) {
        s_w[i] = __bfloat162float(__float2bfloat16_rn(__ldg(w2 + i)));
        s_g[i] = train_stats ? gram[i] : 0.f;
    }
    __syncthreads();
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < E2_C * E2_C; o += gridDim.x * blockDim.x) {
        const int c2 = o >> 6, cp = o & 63;
        float v = rh[o];
        if (train_stats) {
            float a = 0.f;
            for (int i = 0; i < E2_C; ++i) a = fmaf(s_w[c2 * E2_C + i], s_g[i * E2_C + cp], a);
            v += ws[c2] * hsum[cp] + ws[E2_C + c2] * a;
        }
        dw2[o] = v;
    }
}

template <int K>
int e2_launch_fwd(cudaStream_t stream, const float* x, int ldx, const int32_t* idx, long long P, int N, const float* w1,
                  const float* coef1, const float* w2, const float* gamma2, float* sel, uint8_t* arg, double* stats,
                  float* gram, float* hsum) {
    using Cfg = E2Cfg<K>;
    const long long n_units = (P + Cfg::UNIT_PTS - 1) / Cfg::UNIT_PTS;
    const int grid = (int)(n_units < FS_NUM_SMS ? n_units : FS_NUM_SMS);
    if (stats) {
        constexpr int SM = E2FwdCfg<K, 3>::SMEM;
        FS_CUDA_TRY(cudaFuncSetAttribute(edge2_fwd_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
        edge2_fwd_kernel<K, true><<<grid, E2_THREADS, SM, stream>>>(x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg, gram, hsum);
        FS_RETURN_IF_LAUNCH_FAILED();
        edge2_stats_kernel<<<E2_C, E2_C, 0, stream>>>(w2, gram, hsum, stats);
    } else {
        constexpr int SM = E2FwdCfg<K, 2>::SMEM;
        FS_CUDA_TRY(cudaFuncSetAttribute(edge2_fwd_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
        edge2_fwd_kernel<K, false><<<grid, E2_THREADS, SM, stream>>>(x, ldx, idx, P, N, w1, coef1, w2, gamma2, sel, arg, nullptr, nullptr);
    }
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

template <int K>
int e2_launch_bwd(cudaStream_t stream, const float* x, int ldx, const int32_t* idx, long long P, int N, const float* w1,
                  const float* coef1, const float* w2, const float* ws, const float* scale2, const float* d2,
                  const uint8_t* arg, double* dgb1, float* acc1, float* rh) {
    using Cfg = E2Cfg<K>;
    using BC = E2BwdCfg<K>;
    const long long n_units = (P + Cfg::UNIT_PTS - 1) / Cfg::UNIT_PTS;
    const int grid = (int)(n_units < FS_NUM_SMS ? n_units : FS_NUM_SMS);
    FS_CUDA_TRY(cudaFuncSetAttribute(edge2_bwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, BC::SMEM));
    edge2_bwd_kernel<K><<<grid, E2_THREADS, BC::SMEM, stream>>>(x, ldx, idx, P, N, w1, coef1, w2, ws + 3 * E2_C, ws + 2 * E2_C, scale2,
                                                                d2, arg, dgb1, acc1, rh);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

}  // namespace

extern "C" int fs_edge2_supported(int k, int C1, int C2) {
    return (C1 == 64 && C2 == 64 && (k == 8 || k == 12 || k == 16 || k == 20 || k == 40)) ? 1 : 0;
}

// floats of the scratch block of fs_edge2_bwd: alpha | beta' | a0 | gm [64*64] | rh [64*64] (zero-filled by the caller)
extern "C" size_t fs_edge2_bwd_scratch_floats(void) { return 3 * E2_C + 2 * E2_C * E2_C; }

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

extern "C" int fs_edge2_bwd(int device, fs_stream_t stream_, const float* x, int ldx, const int32_t* idx, int B, int N, int k,
                            const float* w1, const float* coef1, const float* w2, int C2, const float* coef2,
                            const double* dgb2, int train_stats, const float* gram, const float* hsum, const float* d2,
                            const uint8_t* arg, float* scratch, double* dgb1, float* acc1, float* dw2) {
    if (!x || !idx || !w1 || !coef1 || !w2 || !coef2 || !dgb2 || !d2 || !arg || !scratch || !dgb1 || !acc1 || !dw2 || B <= 0 ||
        N <= 0 || ldx < 3)
        return FS_ERR_BAD_ARG;
    if (train_stats && (!gram || !hsum)) return FS_ERR_BAD_ARG;
    if (!fs_edge2_supported(k, 64, C2)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    float* rh = scratch + 3 * E2_C + E2_C * E2_C;
    edge2_bwd_prep_kernel<<<16, 256, 0, stream>>>(w2, coef2, dgb2, (double)P * k, train_stats, scratch);
    FS_RETURN_IF_LAUNCH_FAILED();
    int e = FS_ERR_UNSUPPORTED;
    const float* scale2 = coef2 + 2 * E2_C;
    switch (k) {
        case 8: e = e2_launch_bwd<8>(stream, x, ldx, idx, P, N, w1, coef1, w2, scratch, scale2, d2, arg, dgb1, acc1, rh); break;
        case 12: e = e2_launch_bwd<12>(stream, x, ldx, idx, P, N, w1, coef1, w2, scratch, scale2, d2, arg, dgb1, acc1, rh); break;
        case 16: e = e2_launch_bwd<16>(stream, x, ldx, idx, P, N, w1, coef1, w2, scratch, scale2, d2, arg, dgb1, acc1, rh); break;
        case 20: e = e2_launch_bwd<20>(stream, x, ldx, idx, P, N, w1, coef1, w2, scratch, scale2, d2, arg, dgb1, acc1, rh); break;
        case 40: e = e2_launch_bwd<40>(stream, x, ldx, idx, P, N, w1, coef1, w2, scratch, scale2, d2, arg, dgb1, acc1, rh); break;
    }
    if (e) return e;
    edge2_dw2_kernel<<<16, 256, 0, stream>>>(w2, scratch, rh, gram, hsum, train_stats, dw2);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
