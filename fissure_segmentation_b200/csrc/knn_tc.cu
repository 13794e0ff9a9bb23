// Feature-space kNN on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), C = 64 channels.
//
// The reference builds the graph of the feature-space EdgeConvs from a dense N x N distance matrix
// (utils/general_utils.py:43-53 + :315-327, called from models/dgcnn.py:26 on 64-channel features).
// Here the -2 X X^T contraction runs on tcgen05.mma and the N x N scores never leave the SM:
//
//   1. prep      : per cloud, centre the features, split every value into bf16 hi + lo and write two
//                  operand tables A', B' (P x 256 bf16) such that one K = 208 tensor-core product gives
//                      A'_i . B'_j = |x_i|^2 + |x_j|^2 - 2 x_i.x_j       (error ~ 2^-16 |x_i||x_j|)
//                  (hi*hi + lo*hi + hi*lo in K = 192, the squared norms ride in 16 extra K columns).
//   2. candidates: one CTA per (cloud, 128 queries). TMA stages 128-row operand tiles (128B swizzle),
//                  one elected thread issues tcgen05.mma (M = 128 queries, N = 128 candidates) into a
//                  double-buffered TMEM accumulator, and four epilogue warps read the scores with
//                  tcgen05.ld — one query row per thread — keeping the KP smallest (distance,index)
//                  keys of their row in registers (branch-free compare-exchange chain).
//   3. re-rank   : exact FP32 distances in the reference's own arithmetic for the KP candidates, warp
//                  bitonic top-k, and a per-row certificate that no non-candidate can belong to the
//                  top-k; rows without certificate are redone by the exact SIMT kernel (knn.cu).
#include <cuda.h>

#include "warp_select.cuh"

namespace {

constexpr int TC_M = 128;          // queries per CTA  (UMMA M, TMEM lanes)
constexpr int TC_N = 128;          // candidates per tile (UMMA N, TMEM columns per accumulator)
constexpr int TC_C = 64;           // feature channels handled by this path
constexpr int TC_KROW = 256;       // bf16 elements per operand row: 3 x 64 data + 64 (16 used) extras
constexpr int TC_BOXES = 4;        // TMA boxes (64 bf16 = 128 B wide) per operand row
constexpr int TC_BOX_BYTES = TC_M * 128;            // one 128-row x 128-B box in shared memory
constexpr int TC_TILE_BYTES = TC_BOXES * TC_BOX_BYTES;
constexpr int TC_STAGES = 2;
constexpr int TC_THREADS = 192;    // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer
constexpr int TC_KP = 32;          // candidates kept per query
constexpr int TC_SMEM_BYTES = TC_TILE_BYTES * (1 + TC_STAGES) + 1024 /*align*/ + 256 /*barriers*/;

// ----------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    // K-major, SWIZZLE_128B canonical layout: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1),
    // descriptor version 1 (Blackwell), layout type 2 (cute/arch/mma_sm100_desc.hpp SmemDescriptor).
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128.
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Candidate tiles are visited nearest-first around the query tile (qt, qt+1, qt-1, qt+2, ...): with
// spatially sorted points the selection threshold is tight after the first tile.
__device__ __forceinline__ int tile_at(int t, int qt, int T) {
    const int L = qt, R = T - 1 - qt;
    const int m = L < R ? L : R;
    if (t <= 2 * m) return qt + ((t & 1) ? ((t + 1) >> 1) : -(t >> 1));
    return R > L ? qt + t - m : qt - (t - m);
}

// ----------------------------------------------------------------------------------------------- prep
__global__ void tc_colsum_kernel(const float* __restrict__ x, int ldx, int N, float* __restrict__ sums) {
    // grid (chunks, B), block 256 = 4 row-groups x 64 channels
    const int b = blockIdx.y;
    const int c = threadIdx.x & 63;
    const int rg = threadIdx.x >> 6;
    float acc = 0.f;
    for (int r = blockIdx.x * 4 + rg; r < N; r += gridDim.x * 4) acc += __ldg(x + ((long long)b * N + r) * ldx + c);
    __shared__ float red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (rg == 0) atomicAdd(sums + b * TC_C + c, red[c] + red[64 + c] + red[128 + c] + red[192 + c]);
}

__device__ __forceinline__ void split3(float v, __nv_bfloat16& h, __nv_bfloat16& l, __nv_bfloat16& l2) {
    h = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(h);
    l = __float2bfloat16_rn(r1);
    l2 = __float2bfloat16_rn(r1 - __bfloat162float(l));
}

// One warp per point: lane handles channels 2*lane, 2*lane+1.
__global__ void __launch_bounds__(256)
tc_split_kernel(const float* __restrict__ x, int ldx, long long P, int N, const float* __restrict__ sums,
                __nv_bfloat16* __restrict__ A, __nv_bfloat16* __restrict__ Bm, float* __restrict__ cnorm,
                int* __restrict__ cnorm_max_bits) {
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= P) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(row / N);
    const float inv_n = 1.0f / (float)N;
    const float2 xv = *reinterpret_cast<const float2*>(x + row * ldx + 2 * lane);
    const float v0 = xv.x - __ldg(sums + b * TC_C + 2 * lane) * inv_n;
    const float v1 = xv.y - __ldg(sums + b * TC_C + 2 * lane + 1) * inv_n;
    __nv_bfloat16 h0, l0, t0, h1, l1, t1;
    split3(v0, h0, l0, t0);
    split3(v1, h1, l1, t1);
    // the norm that rides in the GEMM is the norm of the values the GEMM actually multiplies (hi + lo)
    const float e0 = __bfloat162float(h0) + __bfloat162float(l0), e1 = __bfloat162float(h1) + __bfloat162float(l1);
    float nrm = fs_warp_sum(e0 * e0 + e1 * e1);
    __nv_bfloat162* Ar = reinterpret_cast<__nv_bfloat162*>(A + row * TC_KROW);
    __nv_bfloat162* Br = reinterpret_cast<__nv_bfloat162*>(Bm + row * TC_KROW);
    const __nv_bfloat162 hh = __halves2bfloat162(h0, h1), ll = __halves2bfloat162(l0, l1);
    const __nv_bfloat162 hh2 = __hmul2(hh, __float2bfloat162_rn(-2.0f)), ll2 = __hmul2(ll, __float2bfloat162_rn(-2.0f));
    Ar[lane] = hh;        Br[lane] = hh2;        // hi * hi
    Ar[32 + lane] = ll;   Br[32 + lane] = hh2;   // lo * hi
    Ar[64 + lane] = hh;   Br[64 + lane] = ll2;   // hi * lo
    // extras: A = [1 1 1 n_h n_l n_l2 0...], B = [n_h n_l n_l2 1 1 1 0...]
    __nv_bfloat16 nh, nl, nl2;
    split3(nrm, nh, nl, nl2);
    const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f);
    __nv_bfloat16 ea0 = zero, ea1 = zero, eb0 = zero, eb1 = zero;
    if (lane == 0) { ea0 = one; ea1 = one; eb0 = nh; eb1 = nl; }
    if (lane == 1) { ea0 = one; ea1 = nh; eb0 = nl2; eb1 = one; }
    if (lane == 2) { ea0 = nl; ea1 = nl2; eb0 = one; eb1 = one; }
    Ar[96 + lane] = __halves2bfloat162(ea0, ea1);
    Br[96 + lane] = __halves2bfloat162(eb0, eb1);
    if (lane == 0) {
        cnorm[row] = nrm;
        atomicMax(cnorm_max_bits + b, __float_as_int(nrm));
    }
}

// ----------------------------------------------------------------------------------------------- candidates
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_candidates_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         int N, int idx_bits, int32_t* __restrict__ cand /* [P, TC_KP] packed keys */) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + TC_TILE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_TILE_BYTES * (1 + TC_STAGES));
    // bars: 0 a_full | 1,2 b_full | 3,4 b_empty | 5,6 acc_full | 7,8 acc_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = (N + TC_N - 1) / TC_N;          // candidate tiles per cloud
    const int qt = blockIdx.x;                     // query tile of this CTA
    const int b = blockIdx.y;
    const long long cloud0 = (long long)b * N;
    const int q_row0 = qt * TC_M;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bars + 0), 1);
        for (int s = 0; s < TC_STAGES; ++s) {
            mbar_init(smem_u32(bars + 1 + s), 1);
            mbar_init(smem_u32(bars + 3 + s), 1);
            mbar_init(smem_u32(bars + 5 + s), 1);
            mbar_init(smem_u32(bars + 7 + s), 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
            mbar_expect_tx(smem_u32(bars + 0), TC_TILE_BYTES);
            for (int bx = 0; bx < TC_BOXES; ++bx)
                tma_load_2d(smem_u32(smem_a + bx * TC_BOX_BYTES), &map_a, smem_u32(bars + 0), bx * 64,
                            (int)(cloud0 + q_row0));
            for (int t = 0; t < T; ++t) {
                const int s = t % TC_STAGES;
                const uint32_t ph = (t / TC_STAGES) & 1;
                mbar_wait(smem_u32(bars + 3 + s), ph ^ 1);
                mbar_expect_tx(smem_u32(bars + 1 + s), TC_TILE_BYTES);
                const int row = (int)(cloud0 + tile_at(t, qt, T) * TC_N);
                for (int bx = 0; bx < TC_BOXES; ++bx)
                    tma_load_2d(smem_u32(smem_b + s * TC_TILE_BYTES + bx * TC_BOX_BYTES), &map_b, smem_u32(bars + 1 + s),
                                bx * 64, row);
            }
        }
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        mbar_wait(smem_u32(bars + 0), 0);
        for (int t = 0; t < T; ++t) {
            const int s = t % TC_STAGES;
            const uint32_t ph = (t / TC_STAGES) & 1;
            mbar_wait(smem_u32(bars + 1 + s), ph);          // operands landed
            mbar_wait(smem_u32(bars + 7 + s), ph ^ 1);      // accumulator buffer drained by the epilogue
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t acc = tmem_base + s * TC_N;
                uint32_t accumulate = 0;
                for (int bx = 0; bx < TC_BOXES; ++bx) {
                    const int nk = bx < 3 ? 4 : 1;          // extras box: only its first 16 K columns are non-zero
                    for (int kk = 0; kk < nk; ++kk) {
                        const uint64_t da = umma_desc_sw128(smem_u32(smem_a + bx * TC_BOX_BYTES) + kk * 32);
                        const uint64_t db = umma_desc_sw128(smem_u32(smem_b + s * TC_TILE_BYTES + bx * TC_BOX_BYTES) + kk * 32);
                        umma_bf16(acc, da, db, accumulate);
                        accumulate = 1;
                    }
                }
                umma_commit(smem_u32(bars + 3 + s));        // smem stage free once these MMAs retire
                umma_commit(smem_u32(bars + 5 + s));        // accumulator ready
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue: one query row per thread =====================
        int list[TC_KP];
#pragma unroll
        for (int i = 0; i < TC_KP; ++i) list[i] = 0x7fffffff;
        const int idx_mask = (1 << idx_bits) - 1;
        for (int t = 0; t < T; ++t) {
            const int s = t % TC_STAGES;
            const uint32_t ph = (t / TC_STAGES) & 1;
            const int col0 = tile_at(t, qt, T) * TC_N;
            mbar_wait(smem_u32(bars + 5 + s), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int cb = 0; cb < TC_N / 32; ++cb) {
                int v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * TC_N + cb * 32), v);
                const int jbase = col0 + cb * 32;
                const int nvalid = N - jbase;           // columns >= nvalid belong to the next cloud / padding
                const int thr_hi = list[TC_KP - 1] | idx_mask;
                unsigned hits = 0;
#pragma unroll
                for (int e = 0; e < 32; ++e) hits |= (v[e] <= thr_hi) ? (1u << e) : 0u;
                if (nvalid < 32) hits &= nvalid <= 0 ? 0u : ((1u << nvalid) - 1u);
                unsigned any = __reduce_or_sync(FS_FULL_MASK, hits);
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    if (any & (1u << e)) {      // warp-uniform: some row of this warp accepts column e
                        int key = (hits & (1u << e)) ? ((v[e] & ~idx_mask) | (jbase + e)) : 0x7fffffff;
#pragma unroll
                        for (int i = 0; i < TC_KP; ++i) {
                            const int lo = min(list[i], key);
                            key = max(list[i], key);
                            list[i] = lo;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bars + 7 + s));
        }
        const int q = q_row0 + warp * 32 + lane;
        if (q < N) {
            int4* out = reinterpret_cast<int4*>(cand + (cloud0 + q) * TC_KP);
#pragma unroll
            for (int i = 0; i < TC_KP / 4; ++i) out[i] = make_int4(list[4 * i], list[4 * i + 1], list[4 * i + 2], list[4 * i + 3]);
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
    }
}

// ----------------------------------------------------------------------------------------------- re-rank
// One warp per query: lane t evaluates candidate t exactly (same FP32 arithmetic as knn_feat_kernel),
// the warp sorts, writes the k nearest and certifies the row.
__global__ void __launch_bounds__(256)
knn_tc_rerank_kernel(const float* __restrict__ x, int ldx, int N, long long P, int k, int self_loop, int diag_zero,
                     int idx_bits, const int32_t* __restrict__ cand, const float* __restrict__ sqnorm,
                     const float* __restrict__ cnorm, const int* __restrict__ cnorm_max_bits,
                     int32_t* __restrict__ idx, float* __restrict__ dist2, uint8_t* __restrict__ redo) {
    __shared__ float qd_all[8 * 64];
    __shared__ int qi_all[8 * 64];
    __shared__ float xq[8][TC_C];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + warp;
    if (row >= P) return;
    const int b = (int)(row / N);
    const long long cloud0 = (long long)b * N;
    const int q = (int)(row - cloud0);
    const int kk = k + (self_loop ? 0 : 1);
    const int idx_mask = (1 << idx_bits) - 1;

    xq[warp][lane] = __ldg(x + row * ldx + lane);
    xq[warp][lane + 32] = __ldg(x + row * ldx + lane + 32);
    __syncwarp();
    const int key = __ldg(cand + row * TC_KP + lane);
    const bool valid = key != 0x7fffffff && (key & idx_mask) < N;
    const int j = valid ? (key & idx_mask) : 0;
    const float* xr = x + (cloud0 + j) * ldx;
    float acc = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < TC_C / 4; ++c4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + c4);
        acc = fmaf(xq[warp][4 * c4], v.x, acc);
        acc = fmaf(xq[warp][4 * c4 + 1], v.y, acc);
        acc = fmaf(xq[warp][4 * c4 + 2], v.z, acc);
        acc = fmaf(xq[warp][4 * c4 + 3], v.w, acc);
    }
    const float qq = __ldg(sqnorm + row), nj = __ldg(sqnorm + cloud0 + j);
    float d = diag_zero ? __fadd_rn(__fsub_rn(qq, 2.0f * acc), nj) : __fadd_rn(__fsub_rn(nj, 2.0f * acc), qq);
    if (diag_zero && j == q) d = 0.f;

    FsWarpSelect<1> sel;
    sel.init(qd_all + warp * 64, qi_all + warp * 64, kk);
    sel.offer(d, j, valid);
    sel.finish();
    sel.store(self_loop ? 0 : 1, idx + row * k, dist2 ? dist2 + row * k : nullptr, 0, 0, INFINITY);

    // certificate: every non-candidate has approximate distance >= floor(key_KP) and the approximation is
    // within err of the exact FP32 form, so it cannot beat the exact kk-th candidate if that is below bound.
    float dk; int ik;
    sel.get(kk - 1, dk, ik);
    const int last_key = __shfl_sync(FS_FULL_MASK, key, TC_KP - 1);
    const float approx_floor = __int_as_float(last_key & ~idx_mask);
    const float cmax = __int_as_float(__ldg(cnorm_max_bits + b));
    const float err = 6.2e-5f * (__ldg(cnorm + row) + cmax) + 4e-6f * (qq + __ldg(sqnorm + cloud0 + ik)) + 1e-30f;
    const bool certified = (last_key == 0x7fffffff) /* fewer than KP points: everything was a candidate */
                           || (ik != FS_IDX_PAD && dk < approx_floor - err);
    if (lane == 0) redo[row] = certified ? 0 : 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_operand_map(EncodeTiledFn fn, CUtensorMap* map, void* base, long long rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)TC_KROW, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)TC_KROW * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)TC_M};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// Declared in knn.cu: exact SIMT kernel restricted to the rows flagged in `redo`.
int fs_knn_feat_masked(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                       int diag_zero, int32_t* idx, float* dist2, const float* sqnorm, const uint8_t* redo);
void fs_row_sqnorm(cudaStream_t stream, const float* x, int ldx, long long P, int C, float* out);

extern "C" size_t fs_knn_feat_tc_workspace_bytes(int B, int N, int C, int k) {
    (void)C; (void)k;
    const size_t P = (size_t)B * N;
    size_t bytes = 0;
    bytes += align_up(P * TC_KROW * 2, 256) * 2;        // A', B'
    bytes += align_up(P * TC_KP * 4, 256);              // candidate keys
    bytes += align_up(P * 4, 256) * 2;                  // exact norms, centred norms
    bytes += align_up((size_t)B * TC_C * 4, 256);       // channel sums
    bytes += align_up((size_t)B * 4, 256);              // max centred norm per cloud
    bytes += align_up(P, 256);                          // redo flags
    return bytes;
}

extern "C" int fs_knn_feat_tc_supported(int B, int N, int C, int k, int self_loop) {
    const int kk = k + (self_loop ? 0 : 1);
    return (C == TC_C && kk + 4 <= TC_KP && N >= TC_KP && N <= 8192 && (long long)B * N <= 0x7fffffff / TC_KROW) ? 1 : 0;
}

extern "C" int fs_knn_feat_tc(int device, fs_stream_t stream_, const float* x, int ldx, int B, int N, int C, int k,
                              int self_loop, int diag_zero, int32_t* idx, float* dist2, void* workspace,
                              size_t workspace_bytes) {
    if (B < 0 || N <= 0 || k <= 0 || ldx < C) return FS_ERR_BAD_ARG;
    if (B == 0) return FS_OK;
    if (!x || !idx || !workspace) return FS_ERR_BAD_ARG;
    if (!fs_knn_feat_tc_supported(B, N, C, k, self_loop)) return FS_ERR_UNSUPPORTED;
    if (workspace_bytes < fs_knn_feat_tc_workspace_bytes(B, N, C, k)) return FS_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (ldx & 3) || (reinterpret_cast<uintptr_t>(workspace) & 255)) return FS_ERR_ALIGNMENT;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;

    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(ws); ws += align_up((size_t)P * TC_KROW * 2, 256);
    __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(ws); ws += align_up((size_t)P * TC_KROW * 2, 256);
    int32_t* cand = reinterpret_cast<int32_t*>(ws); ws += align_up((size_t)P * TC_KP * 4, 256);
    float* sqnorm = reinterpret_cast<float*>(ws); ws += align_up((size_t)P * 4, 256);
    float* cnorm = reinterpret_cast<float*>(ws); ws += align_up((size_t)P * 4, 256);
    float* sums = reinterpret_cast<float*>(ws); ws += align_up((size_t)B * TC_C * 4, 256);
    int* cmax = reinterpret_cast<int*>(ws); ws += align_up((size_t)B * 4, 256);
    uint8_t* redo = ws;

    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FS_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    CUtensorMap map_a, map_b;
    int e = make_operand_map(encode, &map_a, A, P);
    if (e) return e;
    e = make_operand_map(encode, &map_b, Bm, P);
    if (e) return e;

    // 1. prep
    FS_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)B * TC_C * 4 + 0, stream));
    FS_CUDA_TRY(cudaMemsetAsync(cmax, 0, (size_t)B * 4, stream));
    tc_colsum_kernel<<<dim3(16, B), 256, 0, stream>>>(x, ldx, N, sums);
    tc_split_kernel<<<fs_div_up(P, 8), 256, 0, stream>>>(x, ldx, P, N, sums, A, Bm, cnorm, cmax);
    fs_row_sqnorm(stream, x, ldx, P, C, sqnorm);
    FS_RETURN_IF_LAUNCH_FAILED();

    // 2. tensor-core candidate search
    int idx_bits = 1;
    while ((1 << idx_bits) < N) ++idx_bits;
    FS_CUDA_TRY(cudaFuncSetAttribute(knn_tc_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    dim3 grid(fs_div_up(N, TC_M), B);
    knn_tc_candidates_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(map_a, map_b, N, idx_bits, cand);
    FS_RETURN_IF_LAUNCH_FAILED();

    // 3. exact re-rank + certificate, then the exact kernel on uncertified rows
    knn_tc_rerank_kernel<<<fs_div_up(P, 8), 256, 0, stream>>>(x, ldx, N, P, k, self_loop, diag_zero, idx_bits, cand, sqnorm,
                                                              cnorm, cmax, idx, dist2, redo);
    FS_RETURN_IF_LAUNCH_FAILED();
    return fs_knn_feat_masked(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, dist2, sqnorm, redo);
}
