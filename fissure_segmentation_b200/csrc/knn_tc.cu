// Feature-space kNN on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), C = 64 channels.
//
// The reference builds the graph of the feature-space EdgeConvs from a dense N x N distance matrix
// (utils/general_utils.py:43-53 + :315-327, called from models/dgcnn.py:26 on 64-channel features).
// Here the -2 X X^T contraction runs on tcgen05.mma and the N x N scores never leave the SM:
//
//   1. prep     : per cloud, centre the features, split every value into bf16 hi + lo and write two
//                 operand tables A', B' (P x 256 bf16) such that one K = 208 tensor-core product gives
//                     A'_i . B'_j = |x_i|^2 + |x_j|^2 - 2 x_i.x_j        (error ~ 2^-15 (|x_i|^2+|x_j|^2))
//                 (hi*hi + lo*hi + hi*lo in K = 192; the squared norms ride in 16 extra K columns).
//   2. select   : one CTA per (cloud, 256 queries). TMA stages 128B-swizzled candidate tiles (5 stages), one
//                 thread chosen with elect.sync issues tcgen05.mma (M = 128 queries x N = 64 candidates, the
//                 query operand of both query tiles lives in TMEM) into double-buffered TMEM accumulators;
//                 16 epilogue warps read the scores with tcgen05.ld, one query row per TMEM lane and one warp per
//                 32-column half of the tile. The distance matrix is swept twice: sweep 1 keeps, per thread, the
//                 minimum of each of 32 interleaved column classes (one FMNMX per score) - 64 disjoint classes
//                 per row; the kk-th smallest class minimum is an upper bound tau of the kk-th smallest
//                 distance. Sweep 2 lists every (index, distance) with distance <= tau + 2 err (about 1.2 kk
//                 entries per row) in two per-half survivor lists.
//   3. finalize : one warp per query sorts its survivors (packed 64-bit keys, bitonic network) by approximate
//                 distance. Entries further than 2 err from the kk-th are decided by the approximation alone;
//                 the few within 2 err are re-evaluated exactly in the reference's FP32 arithmetic (same code
//                 as knn.cu), so the neighbour SET equals the exact kernel's. Rows whose survivor lists
//                 overflowed are recomputed by the exact SIMT kernel.
// What bounds the select kernel and what was tried is recorded in DESIGN.md section 4.
#include <cuda.h>

#include "warp_select.cuh"

namespace {

constexpr int TC_M = 128;          // queries per query tile (UMMA M, TMEM lanes)
constexpr int TC_QT = 2;           // query tiles per CTA (share each candidate tile)
constexpr int TC_NB = 64;          // candidates per tile (UMMA N)
constexpr int TC_C = 64;           // feature channels handled by this path
constexpr int TC_KROW = 256;       // bf16 elements per operand row: 3 x 64 data + 64 (16 used) extras
constexpr int TC_BOXES = 4;        // TMA boxes (64 bf16 = 128 B wide) per operand row
constexpr int TC_ABOX_BYTES = TC_M * 128;
constexpr int TC_ATILE_BYTES = TC_BOXES * TC_ABOX_BYTES;     // 64 KB
constexpr int TC_BBOX_BYTES = TC_NB * 128;
constexpr int TC_BTILE_BYTES = TC_BOXES * TC_BBOX_BYTES;     // 32 KB
constexpr int TC_STAGES = 5;         // smem stages of the candidate (B) tiles (32 KB each)
constexpr int TC_ACC = 2;            // TMEM accumulator buffers (each holds TC_QT tiles of TC_NB columns)
constexpr int TC_KSTEPS = 13;        // K = 16 steps: 12 data (hi*hi, lo*hi, hi*lo) + 1 extras
constexpr int TC_A_COLS = TC_KSTEPS * 8;   // TMEM columns of one query tile's A operand (2 bf16 per 32-bit column)
constexpr int TC_HALVES = TC_NB / 32;     // 32-column halves of a candidate tile, one epilogue warp each
constexpr int TC_EPI_WARPS = 4 * TC_QT * TC_HALVES;   // warp w: TMEM lane quarter w & 3, query tile (w >> 2) & 1, column half w >> 3
constexpr int TC_WARP_TMA = TC_EPI_WARPS;
constexpr int TC_WARP_MMA = TC_EPI_WARPS + 1;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_STAGE_BYTES = 32 * 32 * 4;  // per-warp staging of one 32 x 32 score block (sweep 2)
constexpr int TC_CAP = 128;        // survivor slots per query (observed: mean 26-33, max 68), TC_CAP / 2 per column half
constexpr int TC_NCLS = 32;        // interleaved column classes of sweep 1, per 32-column half (64 per row)
constexpr int TC_MAX_KK = 32;      // the merged list holds the 32 smallest of the 64 class minima
constexpr int TC_TMEM_ACC0 = 256;     // accumulators at columns [256, 512), A operands at [0, 2*104)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_BTILE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + TC_EPI_WARPS * TC_STAGE_BYTES;
static_assert(TC_QT * TC_A_COLS <= TC_TMEM_ACC0 && TC_TMEM_ACC0 + TC_ACC * TC_QT * TC_NB <= TC_TMEM_COLS, "TMEM budget");
static_assert(TC_SMEM_BYTES <= 232448, "shared memory budget");

constexpr float TC_ERR_CENTRED = 6.2e-5f;   // 2^-14: bound of the bf16 hi/lo product error, per (|xi|^2+|xj|^2)
constexpr float TC_ERR_RAW = 1.0e-6f;       // rounding of the exact FP32 expansion form, per raw squared norm

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
// Spin on the phase parity. FS_TC_BOUNDED_WAIT (debug builds) traps after ~2 s instead of hanging on a protocol bug;
// the release loop carries no clock reads (they cost six extra issue slots per spin in the hot epilogue warps).
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#ifdef FS_TC_BOUNDED_WAIT
    const long long t0 = clock64();
#endif
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
#ifdef FS_TC_BOUNDED_WAIT
        if (clock64() - t0 > 4000000000LL) __trap();
#endif
    }
}
// One elected lane of a converged warp (ptxas then knows a single thread is active and issues the tcgen05 /
// TMA instructions straight from uniform registers instead of a per-lane waterfall loop).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
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
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, N = 64, M = 128.
constexpr uint32_t TC_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_NB >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
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

// One warp per 4 consecutive points (four independent row loads in flight per lane): lane handles channels
// 2*lane, 2*lane+1 of the operand rows; the raw squared norm is summed exactly like row_sqnorm_kernel (knn.cu)
// so the exact re-evaluation matches the exact kernel bit for bit.
constexpr int TC_SPLIT_ROWS = 4;
__global__ void __launch_bounds__(256)
tc_split_kernel(const float* __restrict__ x, int ldx, long long P, int N, const float* __restrict__ sums,
                __nv_bfloat16* __restrict__ A, __nv_bfloat16* __restrict__ Bm, float* __restrict__ cnorm,
                float* __restrict__ sqnorm, int* __restrict__ norm_max_bits /* [B][2]: centred, raw */) {
    // per-cloud maxima of the norms: one atomic per block and cloud instead of one per row (the per-row version
    // serialised 2048 atomics on each address and dominated the kernel)
    __shared__ int s_max[2];
    const long long blk_row0 = (long long)blockIdx.x * 8 * TC_SPLIT_ROWS;
    const long long blk_row1 = blk_row0 + 8 * TC_SPLIT_ROWS - 1 < P - 1 ? blk_row0 + 8 * TC_SPLIT_ROWS - 1 : P - 1;
    const bool one_cloud = blk_row0 / N == blk_row1 / N;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    __syncthreads();
    const long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * TC_SPLIT_ROWS;
    const int lane = threadIdx.x & 31;
    const float inv_n = 1.0f / (float)N;
    if (row0 < P) {
    float2 xv[TC_SPLIT_ROWS];
    float ra[TC_SPLIT_ROWS], rb[TC_SPLIT_ROWS];
#pragma unroll
    for (int r = 0; r < TC_SPLIT_ROWS; ++r) {
        const long long row = row0 + r < P ? row0 + r : P - 1;
        xv[r] = __ldg(reinterpret_cast<const float2*>(x + row * ldx + 2 * lane));
        ra[r] = __ldg(x + row * ldx + lane);
        rb[r] = __ldg(x + row * ldx + lane + 32);
    }
    const __nv_bfloat16 one = __float2bfloat16_rn(1.0f), zero = __float2bfloat16_rn(0.0f);
    const __nv_bfloat162 m2 = __float2bfloat162_rn(-2.0f);
#pragma unroll
    for (int r = 0; r < TC_SPLIT_ROWS; ++r) {
        const long long row = row0 + r;
        if (row >= P) break;
        const int b = (int)(row / N);
        const float v0 = xv[r].x - __ldg(sums + b * TC_C + 2 * lane) * inv_n;
        const float v1 = xv[r].y - __ldg(sums + b * TC_C + 2 * lane + 1) * inv_n;
        __nv_bfloat16 h0, l0, t0, h1, l1, t1;
        split3(v0, h0, l0, t0);
        split3(v1, h1, l1, t1);
        // the norm that rides in the GEMM is the norm of the values the GEMM actually multiplies (hi + lo)
        const float e0 = __bfloat162float(h0) + __bfloat162float(l0), e1 = __bfloat162float(h1) + __bfloat162float(l1);
        const float nrm = fs_warp_sum(e0 * e0 + e1 * e1);
        const float raw = fs_warp_sum(__fadd_rn(__fadd_rn(0.f, __fmul_rn(ra[r], ra[r])), __fmul_rn(rb[r], rb[r])));
        __nv_bfloat162* Ar = reinterpret_cast<__nv_bfloat162*>(A + row * TC_KROW);
        __nv_bfloat162* Br = reinterpret_cast<__nv_bfloat162*>(Bm + row * TC_KROW);
        const __nv_bfloat162 hh = __halves2bfloat162(h0, h1), ll = __halves2bfloat162(l0, l1);
        const __nv_bfloat162 hh2 = __hmul2(hh, m2), ll2 = __hmul2(ll, m2);
        Ar[lane] = hh;        Br[lane] = hh2;        // hi * hi
        Ar[32 + lane] = ll;   Br[32 + lane] = hh2;   // lo * hi
        Ar[64 + lane] = hh;   Br[64 + lane] = ll2;   // hi * lo
        // extras: A = [1 1 1 n_h n_l n_l2 0...], B = [n_h n_l n_l2 1 1 1 0...]
        __nv_bfloat16 nh, nl, nl2;
        split3(nrm, nh, nl, nl2);
        __nv_bfloat16 ea0 = zero, ea1 = zero, eb0 = zero, eb1 = zero;
        if (lane == 0) { ea0 = one; ea1 = one; eb0 = nh; eb1 = nl; }
        if (lane == 1) { ea0 = one; ea1 = nh; eb0 = nl2; eb1 = one; }
        if (lane == 2) { ea0 = nl; ea1 = nl2; eb0 = one; eb1 = one; }
        Ar[96 + lane] = __halves2bfloat162(ea0, ea1);
        Br[96 + lane] = __halves2bfloat162(eb0, eb1);
        if (lane == 0) {
            cnorm[row] = nrm;
            sqnorm[row] = raw;
            if (one_cloud) {
                atomicMax(&s_max[0], __float_as_int(nrm));
                atomicMax(&s_max[1], __float_as_int(raw));
            } else {
                atomicMax(norm_max_bits + 2 * b, __float_as_int(nrm));
                atomicMax(norm_max_bits + 2 * b + 1, __float_as_int(raw));
            }
        }
    }
    }
    __syncthreads();
    if (one_cloud && threadIdx.x < 2) atomicMax(norm_max_bits + 2 * (int)(blk_row0 / N) + threadIdx.x, s_max[threadIdx.x]);
}

__device__ __forceinline__ float tc_row_err(float cn, float cmax, float rn, float rmax) {
    return TC_ERR_CENTRED * (cn + cmax) + TC_ERR_RAW * (rn + rmax) + 1e-30f;
}

// Ascending bitonic sorting network on 32 registers; every index is a compile-time constant.
template <int SIZE, int STRIDE>
__device__ __forceinline__ void bitonic_layer32(float (&g)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        constexpr int dummy = 0; (void)dummy;
        const int p = i ^ STRIDE;
        if (p > i) {
            const bool up = (i & SIZE) == 0;
            const float lo = fminf(g[i], g[p]), hi = fmaxf(g[i], g[p]);
            g[i] = up ? lo : hi;
            g[p] = up ? hi : lo;
        }
    }
}
__device__ __forceinline__ void bitonic_sort32(float (&g)[32]) {
    bitonic_layer32<2, 1>(g);
    bitonic_layer32<4, 2>(g); bitonic_layer32<4, 1>(g);
    bitonic_layer32<8, 4>(g); bitonic_layer32<8, 2>(g); bitonic_layer32<8, 1>(g);
    bitonic_layer32<16, 8>(g); bitonic_layer32<16, 4>(g); bitonic_layer32<16, 2>(g); bitonic_layer32<16, 1>(g);
    bitonic_layer32<32, 16>(g); bitonic_layer32<32, 8>(g); bitonic_layer32<32, 4>(g); bitonic_layer32<32, 2>(g);
    bitonic_layer32<32, 1>(g);
}

// ----------------------------------------------------------------------------------------------- select
// tcgen05.mma with the A operand (queries) in TMEM: D[tmem] (+)= A[tmem] * B[smem].
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& lo, const uint4& hi) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}

// Shared-memory bandwidth is the scarce resource of this kernel (an SS-mode MMA re-reads its 128-row A slab for
// every 64 candidates), so the query operands live in TMEM (TS mode): each epilogue thread copies its own query
// row (TMEM lane = row) from global memory with tcgen05.st once, and shared memory only carries the streamed
// candidate tiles (6 TMA stages) plus the sweep-2 staging blocks.
__global__ void __launch_bounds__(TC_THREADS, 1)
knn_tc_select_kernel(const __nv_bfloat16* __restrict__ a_rows, long long P, const __grid_constant__ CUtensorMap map_b,
                     int N, int kk, int diag_zero, const float* __restrict__ cnorm, const float* __restrict__ sqnorm,
                     const int* __restrict__ norm_max_bits, int32_t* __restrict__ cand_j, float* __restrict__ cand_d,
                     int32_t* __restrict__ cand_n) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + TC_STAGES * TC_BTILE_BYTES);
    // bars: b_full[STAGES] | b_empty[STAGES] | acc_full[ACC] | acc_empty[ACC] | a_full
    uint64_t* b_full = bars;
    uint64_t* b_empty = bars + TC_STAGES;
    uint64_t* acc_full = bars + 2 * TC_STAGES;
    uint64_t* acc_empty = acc_full + TC_ACC;
    uint64_t* a_full = acc_empty + TC_ACC;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 1);
    float* stage_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = (N + TC_NB - 1) / TC_NB;        // candidate tiles per sweep
    const int b = blockIdx.y;
    const long long cloud0 = (long long)b * N;
    const int q_row0 = blockIdx.x * (TC_QT * TC_M);

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(smem_u32(b_full + s), 1); mbar_init(smem_u32(b_empty + s), 1); }
        for (int a = 0; a < TC_ACC; ++a) { mbar_init(smem_u32(acc_full + a), 1); mbar_init(smem_u32(acc_empty + a), TC_EPI_WARPS); }
        mbar_init(smem_u32(a_full), TC_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_WARP_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TC_WARP_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
            for (int it = 0; it < 2 * T; ++it) {
                const int s = it % TC_STAGES;
                const uint32_t ph = (it / TC_STAGES) & 1;
                mbar_wait(smem_u32(b_empty + s), ph ^ 1);
                mbar_expect_tx(smem_u32(b_full + s), TC_BTILE_BYTES);
                const int row = (int)(cloud0 + (it % T) * TC_NB);
                for (int bx = 0; bx < TC_BOXES; ++bx)
                    tma_load_2d(smem_u32(smem_b + s * TC_BTILE_BYTES + bx * TC_BBOX_BYTES), &map_b, smem_u32(b_full + s),
                                bx * 64, row);
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer: one thread, straight-line issue (descriptors = base + constant) ========
        mbar_wait(smem_u32(a_full), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t db0 = umma_desc_sw128(smem_u32(smem_b));
        for (int it = 0; it < 2 * T; ++it) {
            const int s = it % TC_STAGES;
            const int a = it % TC_ACC;
            mbar_wait(smem_u32(b_full + s), (it / TC_STAGES) & 1);            // operands landed
            mbar_wait(smem_u32(acc_empty + a), ((it / TC_ACC) & 1) ^ 1);      // accumulator buffer drained by the epilogue
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one_sync()) {
                const uint64_t db_s = db0 + (uint64_t)((s * TC_BTILE_BYTES) >> 4);
                const uint32_t acc_a = tmem_base + TC_TMEM_ACC0 + a * TC_QT * TC_NB;
#pragma unroll
                for (int st = 0; st < TC_KSTEPS; ++st) {
                    const int bx = st < 12 ? st / 4 : 3, kq = st < 12 ? st % 4 : 0;   // extras box: first 16 K columns only
                    const uint64_t db = db_s + (uint64_t)((bx * TC_BBOX_BYTES + kq * 32) >> 4);
#pragma unroll
                    for (int u = 0; u < TC_QT; ++u)        // the two query tiles alternate: independent back-to-back MMAs
                        umma_bf16_ts(acc_a + u * TC_NB, tmem_base + u * TC_A_COLS + st * 8, db, st ? 1u : 0u);
                }
                umma_commit(smem_u32(b_empty + s));         // smem stage free once these MMAs retire
                umma_commit(smem_u32(acc_full + a));        // accumulators ready
            }
            __syncwarp();
        }
    } else {
        // ===================== epilogue: TMEM lane = query row; two warps share a row, one per 32-column half ======
        const int u = (warp >> 2) & (TC_QT - 1);       // query tile
        const int w4 = warp & 3;                       // TMEM lane quarter this warp may access
        const int cb = warp >> 3;                      // column half of every candidate tile
        const int qrow = q_row0 + u * TC_M + w4 * 32 + lane;
        const bool row_ok = qrow < N;
        const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
        {
            // A operand: this thread's query row -> TMEM (columns u*104 .. +104), the K steps split between the two
            // warps of the row; rows past the table are zero
            const long long grow = cloud0 + qrow;
            const uint4* src = reinterpret_cast<const uint4*>(a_rows + (grow < P ? grow : 0) * TC_KROW);
#pragma unroll
            for (int st = 0; st < TC_KSTEPS; ++st) {
                if ((st < (TC_KSTEPS + 1) / 2) == (cb == 0)) {          // warp-uniform
                    uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
                    if (grow < P) { lo = __ldg(src + 2 * st); hi = __ldg(src + 2 * st + 1); }
                    tmem_st8(tmem_base + lane_base + (uint32_t)(u * TC_A_COLS + st * 8), lo, hi);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(a_full));
        }
        float* stage = stage_all + warp * (TC_STAGE_BYTES / 4) + lane * 32;   // this thread's 32 staged scores
        float gm[TC_NCLS];
#pragma unroll
        for (int e = 0; e < TC_NCLS; ++e) gm[e] = INFINITY;
        float tau = INFINITY;
        int cnt = 0;
        const long long out_base = (cloud0 + (row_ok ? qrow : 0)) * TC_CAP + cb * (TC_CAP / TC_HALVES);
        for (int it = 0; it < 2 * T; ++it) {
            const int a = it % TC_ACC;
            const uint32_t ph = (it / TC_ACC) & 1;
            const int jb = (it % T) * TC_NB + cb * 32;
            if (it == T) {
                // between the sweeps: the two column halves of a row hold 32 class minima each over DISJOINT candidate
                // sets = 64 classes. Each thread sorts its own 32, the pair exchanges them through the staging blocks
                // (named barrier per warp pair); min(own[i], other[31-i]) is the (bitonic) lower half of the union,
                // one bitonic merge sorts it. tau = kk-th smallest of the 64 class minima (+ 2 err): on average
                // 24 survivors per row for kk = 20 instead of 30.6 with 32 classes.
                const int pair_bar = 1 + (warp & 7);
                bitonic_sort32(gm);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(stage + ((c ^ (lane & 7)) << 2)) = make_float4(gm[4 * c], gm[4 * c + 1], gm[4 * c + 2], gm[4 * c + 3]);
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                const float* other = stage_all + (warp ^ 8) * (TC_STAGE_BYTES / 4) + lane * 32;
                float og[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 o = *reinterpret_cast<const float4*>(other + ((c ^ (lane & 7)) << 2));
                    og[4 * c] = o.x; og[4 * c + 1] = o.y; og[4 * c + 2] = o.z; og[4 * c + 3] = o.w;
                }
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");     // staging blocks are reused by sweep 2
#pragma unroll
                for (int i = 0; i < 32; ++i) gm[i] = fminf(gm[i], og[31 - i]);
                bitonic_layer32<32, 16>(gm); bitonic_layer32<32, 8>(gm); bitonic_layer32<32, 4>(gm);
                bitonic_layer32<32, 2>(gm); bitonic_layer32<32, 1>(gm);
                float t = -INFINITY;   // sorted ascending: kk-th smallest = max of the first kk (no indexed register access)
#pragma unroll
                for (int i = 0; i < TC_NCLS; ++i) t = fmaxf(t, i < kk ? gm[i] : -INFINITY);
                const float cmax = __int_as_float(__ldg(norm_max_bits + 2 * b)), rmax = __int_as_float(__ldg(norm_max_bits + 2 * b + 1));
                const int q = row_ok ? qrow : N - 1;
                tau = t + 2.f * tc_row_err(__ldg(cnorm + cloud0 + q), cmax, __ldg(sqnorm + cloud0 + q), rmax);
                // strictly above tau: the hit test below is the sign bit of (score - tau)
                tau = tau + fmaxf(fabsf(tau) * 2.4e-7f, 1e-37f);
            }
            mbar_wait(smem_u32(acc_full + a), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // This thread's 32 scores are pulled into registers and the accumulator buffer is handed back to the MMA
            // warp BEFORE the scores are processed.
            float v[32];
            tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(TC_TMEM_ACC0 + (a * TC_QT + u) * TC_NB + cb * 32), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + a));
            if (jb + 32 > N) {         // last tile: columns beyond the cloud are padding / the next cloud
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = (jb + e < N) ? v[e] : INFINITY;
            }
            if (diag_zero && jb == q_row0 + u * TC_M + w4 * 32) {      // warp-uniform: this block holds the warp's diagonal
                // the reference forces d(i,i) = 0 (general_utils.py:52): the query itself always survives
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = (jb + e == qrow) ? -FLT_MAX : v[e];
            }
            if (it < T) {
#pragma unroll
                for (int e = 0; e < 32; ++e) gm[e] = fminf(gm[e], v[e]);
            } else {
                // sweep 2: hit mask from sign bits (one FADD + one funnel shift per score), scores staged in shared
                // memory (16-byte chunks XOR-swizzled by lane so the 128-bit stores are conflict-free), then each
                // lane walks its own few hits
                unsigned hits = 0;
#pragma unroll
                for (int e = 31; e >= 0; --e) hits = __funnelshift_l(__float_as_uint(v[e] - tau), hits, 1);
                if (!row_ok) hits = 0;
                if (__any_sync(FS_FULL_MASK, hits != 0)) {
                    __syncwarp();
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(stage + ((c ^ (lane & 7)) << 2)) =
                            make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
                    while (hits) {
                        const int e = __ffs(hits) - 1;
                        hits &= hits - 1;
                        const float dv = stage[(((e >> 2) ^ (lane & 7)) << 2) + (e & 3)];
                        if (cnt < TC_CAP / TC_HALVES) {
                            cand_j[out_base + cnt] = jb + e;
                            cand_d[out_base + cnt] = dv;
                        }
                        ++cnt;
                    }
                }
            }
        }
        if (row_ok) cand_n[(cloud0 + qrow) * TC_HALVES + cb] = cnt;
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_WARP_MMA) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
    }
}

// ----------------------------------------------------------------------------------------------- finalize
// One warp per query: sort the survivors by approximate distance, decide everything further than 2 err
// from the kk-th by the approximation, re-evaluate the rest exactly (reference FP32 arithmetic).
// H = entries per lane: 1 when the row has at most 32 survivors (one 32-wide bitonic sort), else 2.
template <int H>
__device__ __forceinline__ void tc_finalize_row(const float* __restrict__ x, int ldx, long long cloud0, int q, long long row,
                                                int k, int kk, int skip, int diag_zero, int n, int n0,
                                                const int32_t* __restrict__ cand_j, const float* __restrict__ cand_d,
                                                const float* __restrict__ sqnorm, float err, float qq,
                                                float* qd, int* qi, float* xq, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    float sd[H];
    int sj[H];
    // survivors of the two column halves sit in [0, n0) and [TC_CAP/2, TC_CAP/2 + n - n0)
    auto phys = [&](int slot) { return slot < n0 ? slot : TC_CAP / TC_HALVES + slot - n0; };
    if (H <= 2) {
        // one or two survivors per lane: bitonic network on packed (distance, index) keys
        unsigned long long key[H <= 2 ? H : 1];
#pragma unroll
        for (int h = 0; h < (H <= 2 ? H : 1); ++h) {
            const int slot = h * 32 + lane;
            const bool valid = slot < n;
            key[h] = fs_pack_key(valid ? __ldg(cand_d + row * TC_CAP + phys(slot)) : INFINITY,
                                 valid ? __ldg(cand_j + row * TC_CAP + phys(slot)) : FS_IDX_PAD);
        }
        fs_warp_bitonic_sort_keys<(H <= 2 ? H : 1)>(key, lane);
#pragma unroll
        for (int h = 0; h < (H <= 2 ? H : 1); ++h) fs_unpack_key(key[h], sd[h], sj[h]);
    } else {
        FsWarpSelect<H> sel;
        sel.init(qd, qi, 32 * H);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int slot = h * 32 + lane;
            const bool valid = slot < n;
            sel.offer(valid ? __ldg(cand_d + row * TC_CAP + phys(slot)) : INFINITY, valid ? __ldg(cand_j + row * TC_CAP + phys(slot)) : FS_IDX_PAD, valid);
        }
        sel.finish();
#pragma unroll
        for (int h = 0; h < H; ++h) { sd[h] = sel.d[h]; sj[h] = sel.i[h]; }
    }
    const int r_thr = kk - 1;
    // kk <= 24 < 32: the kk-th smallest always sits in slot 0 of the ascending order
    const float thr = __shfl_sync(FS_FULL_MASK, sd[0], r_thr & 31);
    const float lo = thr - 2.f * err, hi = thr + 2.f * err;
    bool in_[H], amb[H];
    int n_in = 0, n_amb = 0;
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const bool valid = sj[h] != FS_IDX_PAD;
        in_[h] = valid && sd[h] < lo;
        amb[h] = valid && !in_[h] && sd[h] <= hi;
        n_in += __popc(__ballot_sync(FS_FULL_MASK, in_[h]));
        n_amb += __popc(__ballot_sync(FS_FULL_MASK, amb[h]));
    }
    const int slots = kk - n_in;
    if (n_amb == slots) {
        // the approximation alone decides the set: ranks [0, kk) in approximate order
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int r = h * 32 + lane;
            if (r >= skip && r < kk) out[r - skip] = sj[h];
        }
        return;
    }
    // ambiguous boundary: exact distances (same arithmetic as knn_feat_kernel) for the ambiguous entries
    xq[lane] = __ldg(x + row * ldx + lane);
    xq[lane + 32] = __ldg(x + row * ldx + lane + 32);
    __syncwarp();
    float ex[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        ex[h] = INFINITY;
        if (amb[h]) {
            const int j = sj[h];
            const float* xr = x + (cloud0 + j) * ldx;
            float acc = 0.f;
#pragma unroll
            for (int c4 = 0; c4 < TC_C / 4; ++c4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + c4);
                acc = fmaf(xq[4 * c4], v.x, acc);
                acc = fmaf(xq[4 * c4 + 1], v.y, acc);
                acc = fmaf(xq[4 * c4 + 2], v.z, acc);
                acc = fmaf(xq[4 * c4 + 3], v.w, acc);
            }
            const float nj = __ldg(sqnorm + cloud0 + j);
            float d = diag_zero ? __fadd_rn(__fsub_rn(qq, 2.0f * acc), nj) : __fadd_rn(__fsub_rn(nj, 2.0f * acc), qq);
            if (diag_zero && j == q) d = 0.f;
            ex[h] = d;
        }
    }
    // rank of every ambiguous entry among the ambiguous ones by exact (distance, index)
    int rank[H];
#pragma unroll
    for (int h = 0; h < H; ++h) rank[h] = 0;
#pragma unroll
    for (int sh = 0; sh < H; ++sh) {
        unsigned todo = __ballot_sync(FS_FULL_MASK, amb[sh]);      // only the ambiguous lanes are broadcast
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1;
            const float xd = __shfl_sync(FS_FULL_MASK, ex[sh], sl);
            const int xj = __shfl_sync(FS_FULL_MASK, sj[sh], sl);
#pragma unroll
            for (int h = 0; h < H; ++h)
                if (amb[h] && fs_pair_less(xd, xj, ex[h], sj[h])) ++rank[h];
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) {
        const int r = h * 32 + lane;
        int pos = -1;
        if (in_[h]) pos = r;                                   // sure-in entries occupy ranks [0, n_in)
        else if (amb[h] && rank[h] < slots) pos = n_in + rank[h];
        if (pos >= skip && pos < kk) out[pos - skip] = sj[h];
    }
}

__global__ void __launch_bounds__(256)
knn_tc_finalize_kernel(const float* __restrict__ x, int ldx, int N, long long P, int k, int self_loop, int diag_zero,
                       const int32_t* __restrict__ cand_j, const float* __restrict__ cand_d,
                       const int32_t* __restrict__ cand_n, const float* __restrict__ sqnorm,
                       const float* __restrict__ cnorm, const int* __restrict__ norm_max_bits,
                       int32_t* __restrict__ idx, uint8_t* __restrict__ redo) {
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
    const int skip = self_loop ? 0 : 1;
    const int n0 = __ldg(cand_n + row * TC_HALVES), n1 = __ldg(cand_n + row * TC_HALVES + 1);
    const int n = n0 + n1;
    const float cmax = __int_as_float(__ldg(norm_max_bits + 2 * b)), rmax = __int_as_float(__ldg(norm_max_bits + 2 * b + 1));
    const float qq = __ldg(sqnorm + row);
    const float err = tc_row_err(__ldg(cnorm + row), cmax, qq, rmax);
    // overflow of a half's list, too few survivors or a NaN row: the exact kernel redoes this query
    if (n0 > TC_CAP / TC_HALVES || n1 > TC_CAP / TC_HALVES || n < kk || !(err == err)) {
        if (lane == 0) redo[row] = 1;
        return;
    }
    if (lane == 0) redo[row] = 0;
    if (n <= 32)
        tc_finalize_row<1>(x, ldx, cloud0, q, row, k, kk, skip, diag_zero, n, n0, cand_j, cand_d, sqnorm, err, qq,
                           qd_all + warp * 64, qi_all + warp * 64, xq[warp], idx + row * k);
    else if (n <= 64)
        tc_finalize_row<2>(x, ldx, cloud0, q, row, k, kk, skip, diag_zero, n, n0, cand_j, cand_d, sqnorm, err, qq,
                           qd_all + warp * 64, qi_all + warp * 64, xq[warp], idx + row * k);
    else
        tc_finalize_row<4>(x, ldx, cloud0, q, row, k, kk, skip, diag_zero, n, n0, cand_j, cand_d, sqnorm, err, qq,
                           qd_all + warp * 64, qi_all + warp * 64, xq[warp], idx + row * k);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_operand_map(EncodeTiledFn fn, CUtensorMap* map, void* base, long long rows, int box_rows) {
    const cuuint64_t dims[2] = {(cuuint64_t)TC_KROW, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)TC_KROW * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace

// Defined in knn.cu.
int fs_knn_feat_masked(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                       int diag_zero, int32_t* idx, float* dist2, const float* sqnorm, const uint8_t* redo);
int fs_knn_feat_exact(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                      int diag_zero, int32_t* idx, float* dist2, float* sqnorm_ws);

extern "C" size_t fs_knn_feat_tc_workspace_bytes(int B, int N, int C, int k) {
    (void)C; (void)k;
    const size_t P = (size_t)B * N;
    size_t bytes = 0;
    bytes += align_up(P * TC_KROW * 2, 256) * 2;        // A', B'
    bytes += align_up(P * TC_CAP * 4, 256) * 2;         // survivor indices, distances
    bytes += align_up(P * 4 * TC_HALVES, 256);          // survivor counts (one per column half)
    bytes += align_up(P * 4, 256) * 2;                  // raw norms, centred norms
    bytes += align_up((size_t)B * TC_C * 4, 256);       // channel sums
    bytes += align_up((size_t)B * 8, 256);              // max centred / raw norm per cloud
    bytes += align_up(P, 256);                          // redo flags
    return bytes;
}

extern "C" size_t fs_knn_feat_tc_redo_offset(int B, int N, int C, int k) {
    return fs_knn_feat_tc_workspace_bytes(B, N, C, k) - align_up((size_t)B * N, 256);
}

extern "C" int fs_knn_feat_tc_supported(int B, int N, int C, int k, int self_loop) {
    const int kk = k + (self_loop ? 0 : 1);
    return (C == TC_C && kk <= TC_MAX_KK && N >= 2 * TC_NCLS && N <= 32768 && (long long)B * N <= 0x7fffffff / TC_KROW) ? 1 : 0;
}

extern "C" int fs_knn_feat_tc(int device, fs_stream_t stream_, const float* x, int ldx, int B, int N, int C, int k,
                              int self_loop, int diag_zero, int32_t* idx, float* dist2, void* workspace,
                              size_t workspace_bytes) {
    if (B < 0 || N <= 0 || k <= 0 || ldx < C) return FS_ERR_BAD_ARG;
    if (B == 0) return FS_OK;
    if (!x || !idx || !workspace) return FS_ERR_BAD_ARG;
    const int kk = k + (self_loop ? 0 : 1);
    if (kk > N || kk > FS_MAX_K + 1) return FS_ERR_BAD_ARG;
    if (workspace_bytes < fs_knn_feat_tc_workspace_bytes(B, N, C, k)) return FS_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(x) & 15) || (ldx & 3) || (reinterpret_cast<uintptr_t>(workspace) & 255)) return FS_ERR_ALIGNMENT;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;

    uint8_t* ws = static_cast<uint8_t*>(workspace);
    __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(ws); ws += align_up((size_t)P * TC_KROW * 2, 256);
    __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(ws); ws += align_up((size_t)P * TC_KROW * 2, 256);
    int32_t* cand_j = reinterpret_cast<int32_t*>(ws); ws += align_up((size_t)P * TC_CAP * 4, 256);
    float* cand_d = reinterpret_cast<float*>(ws); ws += align_up((size_t)P * TC_CAP * 4, 256);
    int32_t* cand_n = reinterpret_cast<int32_t*>(ws); ws += align_up((size_t)P * 4 * TC_HALVES, 256);
    float* sqnorm = reinterpret_cast<float*>(ws); ws += align_up((size_t)P * 4, 256);
    float* cnorm = reinterpret_cast<float*>(ws); ws += align_up((size_t)P * 4, 256);
    float* sums = reinterpret_cast<float*>(ws); ws += align_up((size_t)B * TC_C * 4, 256);
    int* nmax = reinterpret_cast<int*>(ws); ws += align_up((size_t)B * 8, 256);
    uint8_t* redo = ws;

    // distances requested (public knn(..., return_dist=True)) or shape outside the tensor-core path: exact kernel
    if (dist2 || !fs_knn_feat_tc_supported(B, N, C, k, self_loop))
        return fs_knn_feat_exact(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, dist2, sqnorm);

    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        FS_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    CUtensorMap map_b;
    int e = make_operand_map(encode, &map_b, Bm, P, TC_NB);
    if (e) return e;

    // 1. prep
    FS_CUDA_TRY(cudaMemsetAsync(sums, 0, align_up((size_t)B * TC_C * 4, 256) + (size_t)B * 8, stream));
    tc_colsum_kernel<<<dim3(16, B), 256, 0, stream>>>(x, ldx, N, sums);
    tc_split_kernel<<<fs_div_up(P, 8 * TC_SPLIT_ROWS), 256, 0, stream>>>(x, ldx, P, N, sums, A, Bm, cnorm, sqnorm, nmax);
    FS_RETURN_IF_LAUNCH_FAILED();

    // 2. tensor-core sweeps
    FS_CUDA_TRY(cudaFuncSetAttribute(knn_tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
    dim3 grid(fs_div_up(N, TC_QT * TC_M), B);
    knn_tc_select_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, stream>>>(A, P, map_b, N, kk, diag_zero, cnorm, sqnorm, nmax,
                                                                      cand_j, cand_d, cand_n);
    FS_RETURN_IF_LAUNCH_FAILED();

    // 3. finalize, then the exact kernel on rows whose survivor list overflowed
    knn_tc_finalize_kernel<<<fs_div_up(P, 8), 256, 0, stream>>>(x, ldx, N, P, k, self_loop, diag_zero, cand_j, cand_d, cand_n,
                                                                sqnorm, cnorm, nmax, idx, redo);
    FS_RETURN_IF_LAUNCH_FAILED();
    return fs_knn_feat_masked(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, nullptr, sqnorm, redo);
}
