// kNN graph build on the 5th-generation tensor cores (tcgen05 + TMEM + TMA): feature space (C = 64 / 128 / 256) and,
// through the same kernels, 3-D coordinates (C = 3).
//
// The reference builds every graph from a dense N x N distance matrix (utils/general_utils.py:43-53 + :315-327, called
// from models/dgcnn.py:26; models/dgcnn_opensrc.py:34-40). Here the -2 X X^T contraction runs on tcgen05.mma and the
// N x N scores never leave the SM:
//
//   1. prep     : per cloud, centre the features and scale them into [-1, 1]; write two fp16 operand tables A', B'
//                 such that ONE tensor-core product gives the approximate squared distance
//                     A'_i . B'_j = |h_i|^2 + |h_j|^2 - 2 h_i.h_j  (- T_i in sweep 2)       h = fp16(centred, scaled x)
//                 C >= 64: one fp16 term per channel (K = C) + one K = 16 step of extras (norms, threshold);
//                 C == 3 : hi/lo split, hi*hi + lo*hi + hi*lo and the extras share a single K = 16 step.
//                 In sqrt space the representation error is additive: | |h_i-h_j| - |x_i-x_j| | <= e_i, with
//                 e_i = u (|x_i| + max|x|), u = 2^-11 (2^-21 with the hi/lo split) - proportional to the DISTANCE, not to
//                 the norms, so near neighbours are resolved far better than a norm-relative bound suggests.
//   2. select   : one CTA per (cloud, 128 queries), two CTAs per SM. TMA stages 128B-swizzled candidate tiles, one thread
//                 chosen with elect.sync issues tcgen05.mma (M = 128 queries x N = 64 candidates, the query operand lives
//                 in TMEM: TS mode) into 2-4 TMEM accumulator buffers; 8 epilogue warps read the scores with tcgen05.ld,
//                 one query row per TMEM lane and one warp per 32-column half of the tile. Two sweeps over the cloud:
//                 sweep 1 keeps the minimum of 32 interleaved column classes per thread (one FMNMX per score) = 64 disjoint
//                 classes per row; the kk-th smallest class minimum bounds the kk-th smallest distance. The threshold T_i
//                 (bound + error margin) is then written INTO the query operand in TMEM (a B-side extras column holds 1,
//                 the A side -T_i), so in sweep 2 a survivor is simply a negative score: one funnel shift per score builds
//                 the hit mask, and the ~1.2 kk survivors per row go to two per-half lists.
//   3. finalize : one warp per query sorts its survivors (packed 64-bit keys, bitonic network). Entries provably inside /
//                 outside the kk nearest (sqrt-space margin 2 e_i around the kk-th) are decided by the approximation;
//                 the few in between are re-evaluated exactly in the reference's FP32 arithmetic (same code as knn.cu), so
//                 the neighbour SET equals the exact kernel's. Rows whose lists overflowed, or that hold NaN / Inf, are
//                 recomputed by the exact SIMT kernel (row mask). For C == 3 the final k are also ORDERED exactly.
#include <cuda.h>
#include <cuda_fp16.h>

#include "warp_select.cuh"

namespace {

constexpr int TC_M = 128;            // queries per CTA (UMMA M, TMEM lanes)
constexpr int TC_NB = 64;            // candidates per tile (UMMA N)
constexpr int TC_BOX_BYTES = TC_NB * 128;       // one TMA box: 64 rows x 64 fp16 (128 B, SWIZZLE_128B)
constexpr int TC_HALVES = 2;         // 32-column halves of a candidate tile, one epilogue warp each
constexpr int TC_EPI_WARPS = 4 * TC_HALVES;     // warp w: TMEM lane quarter w & 3, column half w >> 2
constexpr int TC_WARP_TMA = TC_EPI_WARPS;
constexpr int TC_WARP_MMA = TC_EPI_WARPS + 1;
constexpr int TC_THREADS = (TC_EPI_WARPS + 2) * 32;
constexpr int TC_STAGE_BYTES = 32 * 32 * 4;     // per-warp staging of one 32 x 32 score block
constexpr int TC_CAP = 128;          // survivor slots per query, TC_CAP / 2 per column half
constexpr int TC_MAX_KK = 64;        // 64 class minima per row
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_ACC = 4;
constexpr int TC_EXTRAS = 6;         // extras halfs: A = [1, 1, n_h, n_l, -T_h, -T_l], B = [n_h, n_l, 1, 1, 1, 1]

// Error model (scaled, centred units; see the header comment and DESIGN.md section 4):
constexpr float TC_U_1TERM = 4.93e-4f;     // 2^-11 (1 + 2^-6): fp16 unit roundoff, one term per channel
constexpr float TC_U_SPLIT = 4.9e-7f;      // 2^-21 (1 + 2^-6): hi + lo split
constexpr float TC_G_LIN = 7.63e-6f;       // 2^-17 (|h_i|^2 + max |h|^2): fp32 accumulation in the tensor core, 2-term norms / T
constexpr float TC_G_RAW = 1.0e-6f;        // rounding of the exact FP32 expansion form, per raw squared norm

struct TcShape {
    int C;            // channels
    int split;        // 1: hi/lo split packed into one K step (C == 3)
    int nboxes;       // 128-byte boxes per operand row
    int krow;         // fp16 elements per operand row = nboxes * 64
    int ksteps;       // K = 16 MMA steps
    int xstep;        // K step that holds the extras
    int xhalf;        // first extras half inside that step
    int acc0;         // first accumulator column in TMEM (query operand occupies [0, ksteps * 8))
    int nacc;         // accumulator buffers (64 columns each)
    int tmem_cols;    // allocation (256: two CTAs per SM, 512: one)
    int stages;       // smem stages of candidate tiles
};

__host__ __device__ inline int tc_step_box(const TcShape& s, int st) { return (s.split || st < s.xstep) ? st / 4 : s.nboxes - 1; }
__host__ __device__ inline int tc_step_off(const TcShape& s, int st) { return (s.split || st < s.xstep) ? (st % 4) * 32 : 0; }

bool tc_make_shape(int C, TcShape* out) {
    TcShape s{};
    s.C = C;
    if (C == 3) {
        s.split = 1; s.nboxes = 1; s.ksteps = 1; s.xstep = 0; s.xhalf = 9;
    } else if (C == 64 || C == 128 || C == 256) {
        s.split = 0; s.nboxes = C / 64 + 1; s.ksteps = C / 16 + 1; s.xstep = C / 16; s.xhalf = 0;
    } else {
        return false;
    }
    s.krow = s.nboxes * 64;
    const int a_cols = s.ksteps * 8;
    s.acc0 = (a_cols + 63) / 64 * 64;
    s.tmem_cols = C == 256 ? 512 : 256;
    s.nacc = (s.tmem_cols - s.acc0) / TC_NB;
    if (s.nacc > TC_MAX_ACC) s.nacc = TC_MAX_ACC;
    s.stages = C == 3 ? 8 : (C == 64 ? 4 : 3);      // 8 / 16 / 24 / 40 KB per stage
    *out = s;
    return true;
}
size_t tc_smem_bytes(const TcShape& s) {
    return (size_t)s.stages * s.nboxes * TC_BOX_BYTES + 1024 /*align*/ + 256 /*barriers*/ + (size_t)TC_EPI_WARPS * TC_STAGE_BYTES;
}

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
// Wait on the phase parity. try_wait suspends the thread in hardware until the phase completes or the time hint (about
// 1 ms) runs out, so a waiting warp does not burn issue slots of the other CTA on the SM. The wait is bounded: a protocol
// bug traps after 2^12 time-outs (seconds; the launch fails with an error the host sees) instead of hanging the GPU.
// -DFS_TC_UNBOUNDED_WAIT removes the counter.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000u)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
#ifndef FS_TC_UNBOUNDED_WAIT
    uint32_t polls = 0;
#endif
    while (!mbar_try_wait(bar, parity)) {
#ifndef FS_TC_UNBOUNDED_WAIT
        if (++polls > (1u << 12)) __trap();
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
// kind::f16 instruction descriptor: D = F32, A = B = F16 (format 0), both K-major, N = 64, M = 128.
constexpr uint32_t TC_IDESC = (1u << 4) | ((uint32_t)(TC_NB >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

// tcgen05.mma with the A operand (queries) in TMEM: D[tmem] (+)= A[tmem] * B[smem].
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(TC_IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& lo, const uint4& hi) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
                 : "memory");
}

// ----------------------------------------------------------------------------------------------- prep
// Per-cloud statistics: stats[b] = { column sums [C] | max |x| bits | max centred scaled norm bits | max raw norm bits | - }.
__host__ __device__ inline int tc_stats_stride(int C) { return C + 4; }

// element (b, r, c) at x[b * batch_stride + r * row_stride + c * chan_stride]: point-major feature tables
// (batch_stride = N * ld, row_stride = ld, chan_stride = 1) and the reference's B x C x N coordinate layout alike
__global__ void tc_cloud_stats_kernel(const float* __restrict__ x, long long batch_stride, long long row_stride,
                                      long long chan_stride, int N, int C, float* __restrict__ stats) {
    // grid (chunks, B); thread t owns channel t % C, rows (t / C) + i * (blockDim / C)
    const int b = blockIdx.y;
    const int c = threadIdx.x % C;
    const int rg = threadIdx.x / C, nrg = blockDim.x / C;
    const float* xb = x + (long long)b * batch_stride + (long long)c * chan_stride;
    float acc = 0.f;
    int amax = 0;
    if (rg < nrg) {
        for (int r = blockIdx.x * nrg + rg; r < N; r += gridDim.x * nrg) {
            const float v = __ldg(xb + (long long)r * row_stride);
            acc += v;
            amax = max(amax, __float_as_int(fabsf(v)));       // int order: finite < Inf < NaN
        }
    }
    __shared__ float red[256];
    __shared__ int s_amax;
    if (threadIdx.x == 0) s_amax = 0;
    red[threadIdx.x] = acc;
    __syncthreads();
    atomicMax(&s_amax, amax);
    __syncthreads();
    float* st = stats + (long long)b * tc_stats_stride(C);
    if (threadIdx.x < C) {
        float s = 0.f;
        for (int g = 0; g < nrg; ++g) s += red[g * C + threadIdx.x];
        atomicAdd(st + threadIdx.x, s);
    }
    if (threadIdx.x == 0) atomicMax(reinterpret_cast<int*>(st + C), s_amax);
}

__device__ __forceinline__ void split2h(float v, __half& h, __half& l) {
    h = __float2half_rn(v);
    l = __float2half_rn(v - __half2float(h));
}
// extras of one operand row: A = [1, 1, n_h, n_l, 0, 0], B = [n_h, n_l, 1, 1, 1, 1]
__device__ __forceinline__ void tc_extras(float nrm, __half* ea, __half* eb) {
    __half nh, nl;
    split2h(nrm, nh, nl);
    const __half one = __float2half_rn(1.0f), zero = __float2half_rn(0.0f);
    ea[0] = one; ea[1] = one; ea[2] = nh; ea[3] = nl; ea[4] = zero; ea[5] = zero;
    eb[0] = nh; eb[1] = nl; eb[2] = one; eb[3] = one; eb[4] = one; eb[5] = one;
}
// scale of a cloud: centred values lie in [-2 amax, 2 amax] -> [-1, 1]
__device__ __forceinline__ float tc_inv_scale(float amax) { return amax > 0.f ? 0.5f / amax : 1.0f; }

// Feature rows (C = 64 CH, CH = 1 / 2 / 4): one warp per row, ROWS = 4 / CH rows per warp with all their loads in
// flight together. The raw squared norm is summed exactly like row_sqnorm_kernel (knn.cu) so the exact re-evaluation
// matches the exact kernel bit for bit.
template <int CH>
__global__ void __launch_bounds__(256)
tc_split_feat_kernel(const float* __restrict__ x, int ldx, long long P, int N, TcShape sh, float* __restrict__ stats,
                     __half* __restrict__ A, __half* __restrict__ Bm, float* __restrict__ cnorm, float* __restrict__ sqnorm) {
    constexpr int ROWS = 4 / CH;
    constexpr int C = 64 * CH;
    __shared__ int s_max[2];
    const long long blk_row0 = (long long)blockIdx.x * 8 * ROWS;
    const long long blk_row1 = blk_row0 + 8 * ROWS - 1 < P - 1 ? blk_row0 + 8 * ROWS - 1 : P - 1;
    const bool one_cloud = blk_row0 / N == blk_row1 / N;
    if (threadIdx.x < 2) s_max[threadIdx.x] = 0;
    __syncthreads();
    const long long row0 = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * ROWS;
    const int lane = threadIdx.x & 31;
    const float inv_n = 1.0f / (float)N;
    constexpr int sstride = C + 4;
    const __half2 m2 = __float2half2_rn(-2.0f);
    if (row0 < P) {
        float2 xv[ROWS][CH];
        float ra[ROWS][CH], rb[ROWS][CH];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const long long row = row0 + r < P ? row0 + r : P - 1;
#pragma unroll
            for (int h = 0; h < CH; ++h) {
                xv[r][h] = __ldg(reinterpret_cast<const float2*>(x + row * ldx + 64 * h + 2 * lane));
                ra[r][h] = __ldg(x + row * ldx + 64 * h + lane);
                rb[r][h] = __ldg(x + row * ldx + 64 * h + lane + 32);
            }
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const long long row = row0 + r;
            if (row >= P) break;                                    // warp-uniform
            const int b = (int)(row / N);
            const float* st = stats + (long long)b * sstride;
            const float inv_s = tc_inv_scale(__int_as_float(__ldg(reinterpret_cast<const int*>(st + C))));
            float nrm = 0.f, raw = 0.f;
            __half2* Ar = reinterpret_cast<__half2*>(A + row * sh.krow);
            __half2* Br = reinterpret_cast<__half2*>(Bm + row * sh.krow);
#pragma unroll
            for (int h = 0; h < CH; ++h) {
                raw = __fadd_rn(__fadd_rn(raw, __fmul_rn(ra[r][h], ra[r][h])), __fmul_rn(rb[r][h], rb[r][h]));   // channels lane, lane+32, ...
                const float v0 = (xv[r][h].x - __ldg(st + 64 * h + 2 * lane) * inv_n) * inv_s;
                const float v1 = (xv[r][h].y - __ldg(st + 64 * h + 2 * lane + 1) * inv_n) * inv_s;
                const __half2 hh = __floats2half2_rn(v0, v1);
                const float2 e = __half22float2(hh);
                nrm = fmaf(e.x, e.x, fmaf(e.y, e.y, nrm));
                Ar[32 * h + lane] = hh;
                Br[32 * h + lane] = __hmul2(hh, m2);
            }
            nrm = fs_warp_sum(nrm);
            raw = fs_warp_sum(raw);
            // extras box: the whole 128-byte box is written (16 halfs are read by the MMA, 6 of them meaningful)
            {
                __half ea[8], eb[8];
                tc_extras(nrm, ea, eb);
                ea[6] = ea[7] = eb[6] = eb[7] = __float2half_rn(0.f);
                __half a0 = __float2half_rn(0.f), a1 = a0, b0 = a0, b1 = a0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (lane == i) { a0 = ea[2 * i]; a1 = ea[2 * i + 1]; b0 = eb[2 * i]; b1 = eb[2 * i + 1]; }
                Ar[(C >> 1) + lane] = __halves2half2(a0, a1);
                Br[(C >> 1) + lane] = __halves2half2(b0, b1);
            }
            if (lane == 0) {
                cnorm[row] = nrm;
                sqnorm[row] = raw;
                if (one_cloud) {
                    atomicMax(&s_max[0], __float_as_int(nrm));
                    atomicMax(&s_max[1], __float_as_int(raw));
                } else {
                    atomicMax(reinterpret_cast<int*>(stats + (long long)b * sstride + C + 1), __float_as_int(nrm));
                    atomicMax(reinterpret_cast<int*>(stats + (long long)b * sstride + C + 2), __float_as_int(raw));
                }
            }
        }
    }
    __syncthreads();
    if (one_cloud && threadIdx.x < 2)
        atomicMax(reinterpret_cast<int*>(stats + (blk_row0 / N) * sstride + C + 1 + threadIdx.x), s_max[threadIdx.x]);
}

__device__ __forceinline__ float tc_sqnorm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));      // knn.cu sqnorm3
}

// Coordinates (C == 3): one thread per point. Writes the 64-half operand rows (16 halfs used: hi*hi | lo*hi | hi*lo |
// extras; the caller zero-fills the rest) and a point-major fp32 copy xyz_pm[row] = (x, y, z, |p|^2) for the exact
// re-evaluation and the exact kernel.
__global__ void __launch_bounds__(256)
tc_split_xyz_kernel(const float* __restrict__ coords, long long batch_stride, long long chan_stride, long long point_stride,
                    long long P, int N, TcShape sh, float* __restrict__ stats, __half* __restrict__ A, __half* __restrict__ Bm,
                    float* __restrict__ cnorm, float* __restrict__ sqnorm, float4* __restrict__ xyz_pm) {
    const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= P) return;
    const int b = (int)(row / N);
    const int n = (int)(row - (long long)b * N);
    const float* cb = coords + (long long)b * batch_stride + (long long)n * point_stride;
    const float px = __ldg(cb), py = __ldg(cb + chan_stride), pz = __ldg(cb + 2 * chan_stride);
    const float raw = tc_sqnorm3(px, py, pz);
    xyz_pm[row] = make_float4(px, py, pz, raw);
    float* st = stats + (long long)b * tc_stats_stride(3);
    const float inv_s = tc_inv_scale(__int_as_float(__ldg(reinterpret_cast<const int*>(st + 3))));
    const float inv_n = 1.0f / (float)N;
    const float v[3] = {(px - __ldg(st) * inv_n) * inv_s, (py - __ldg(st + 1) * inv_n) * inv_s, (pz - __ldg(st + 2) * inv_n) * inv_s};
    __half h[3], l[3];
    float nrm = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        split2h(v[i], h[i], l[i]);
        const float e = __half2float(h[i]) + __half2float(l[i]);
        nrm = fmaf(e, e, nrm);
    }
    __half ea[TC_EXTRAS], eb[TC_EXTRAS];
    tc_extras(nrm, ea, eb);
    const __half zero = __float2half_rn(0.f);
    const __half m2 = __float2half_rn(-2.0f);
    __align__(16) __half ra[16];
    __align__(16) __half rb[16];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ra[i] = h[i];     rb[i] = __hmul(h[i], m2);       // hi * hi
        ra[3 + i] = l[i]; rb[3 + i] = __hmul(h[i], m2);   // lo * hi
        ra[6 + i] = h[i]; rb[6 + i] = __hmul(l[i], m2);   // hi * lo
    }
#pragma unroll
    for (int i = 0; i < TC_EXTRAS; ++i) { ra[9 + i] = ea[i]; rb[9 + i] = eb[i]; }
    ra[15] = zero; rb[15] = zero;
    // the whole 128-byte operand row is written (TMA reads whole rows; only the first 16 halfs enter the MMA)
    uint4* Ar = reinterpret_cast<uint4*>(A + row * sh.krow);
    uint4* Br = reinterpret_cast<uint4*>(Bm + row * sh.krow);
    const uint4 z4 = make_uint4(0, 0, 0, 0);
    Ar[0] = *reinterpret_cast<uint4*>(ra); Ar[1] = *reinterpret_cast<uint4*>(ra + 8);
    Br[0] = *reinterpret_cast<uint4*>(rb); Br[1] = *reinterpret_cast<uint4*>(rb + 8);
#pragma unroll
    for (int i = 2; i < 8; ++i) { Ar[i] = z4; Br[i] = z4; }
    cnorm[row] = nrm;
    sqnorm[row] = raw;
    // per-cloud maxima: one atomic per warp when the warp's rows share a cloud (per-thread atomics serialise)
    const unsigned act = __activemask();
    const int b0 = __shfl_sync(act, b, __ffs(act) - 1);
    if (__all_sync(act, b == b0)) {
        int mn = __float_as_int(nrm), mr = __float_as_int(raw);
        mn = __reduce_max_sync(act, mn);
        mr = __reduce_max_sync(act, mr);
        if ((int)(threadIdx.x & 31) == __ffs(act) - 1) {
            atomicMax(reinterpret_cast<int*>(st + 4), mn);
            atomicMax(reinterpret_cast<int*>(st + 5), mr);
        }
    } else {
        atomicMax(reinterpret_cast<int*>(st + 4), __float_as_int(nrm));
        atomicMax(reinterpret_cast<int*>(st + 5), __float_as_int(raw));
    }
}

// Per-row error terms (scaled, centred units): e = sqrt-space margin, g = linear margin.
__device__ __forceinline__ void tc_row_err(const TcShape& sh, float cn, float cmax, float rn, float rmax, float amax,
                                           float& e, float& g) {
    const float u = sh.split ? TC_U_SPLIT : TC_U_1TERM;
    const float inv_s = tc_inv_scale(amax);
    // absolute floor: fp16 subnormals (2^-25 per channel) and the fp32 rounding of (x - mean) / s (2^-24 per channel)
    const float e_abs = 2.0f * sqrtf((float)sh.C) * 9.0e-8f;
    e = u * (sqrtf(cn) + sqrtf(cmax)) + e_abs;
    g = TC_G_LIN * (cn + cmax) + TC_G_RAW * (rn + rmax) * inv_s * inv_s + 1e-30f;
}

// Ascending bitonic sorting network on 32 registers; every index is a compile-time constant.
template <int SIZE, int STRIDE>
__device__ __forceinline__ void bitonic_layer32(float (&g)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
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
// Shared-memory bandwidth is the scarce resource of the MMA side (an SS-mode MMA re-reads its 128-row A slab for every
// 64 candidates), so the query operand lives in TMEM (TS mode): each epilogue thread copies its own query row (TMEM
// lane = row) from global memory with tcgen05.st once, and shared memory only carries the streamed candidate tiles
// plus the sweep-2 staging blocks. Two CTAs share an SM (256 TMEM columns each): while the epilogue warps of one CTA
// wait on an accumulator barrier, the other CTA's warps issue.
__global__ void __launch_bounds__(TC_THREADS, 2)
knn_tc_select_kernel(const __half* __restrict__ a_rows, long long P, const __grid_constant__ CUtensorMap map_b,
                     int N, int kk, int diag_zero, const TcShape sh, const float* __restrict__ cnorm,
                     const float* __restrict__ sqnorm, const float* __restrict__ stats, int32_t* __restrict__ cand_j,
                     float* __restrict__ cand_d, int32_t* __restrict__ cand_n, float* __restrict__ row_T) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_b = smem;
    const int stage_bytes = sh.nboxes * TC_BOX_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + sh.stages * stage_bytes);
    // bars: b_full[8] | b_empty[8] | acc_full[4] | acc_empty[4] | a_full | a2_full
    uint64_t* b_full = bars;
    uint64_t* b_empty = bars + TC_MAX_STAGES;
    uint64_t* acc_full = bars + 2 * TC_MAX_STAGES;
    uint64_t* acc_empty = acc_full + TC_MAX_ACC;
    uint64_t* a_full = acc_empty + TC_MAX_ACC;
    uint64_t* a2_full = a_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a2_full + 1);
    float* stage_all = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int T = (N + TC_NB - 1) / TC_NB;        // candidate tiles per sweep
    const int b = blockIdx.y;
    const long long cloud0 = (long long)b * N;
    const int q_row0 = blockIdx.x * TC_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_MAX_STAGES; ++s) { mbar_init(smem_u32(b_full + s), 1); mbar_init(smem_u32(b_empty + s), 1); }
        for (int a = 0; a < TC_MAX_ACC; ++a) { mbar_init(smem_u32(acc_full + a), 1); mbar_init(smem_u32(acc_empty + a), TC_EPI_WARPS); }
        mbar_init(smem_u32(a_full), TC_EPI_WARPS);
        mbar_init(smem_u32(a2_full), TC_EPI_WARPS / TC_HALVES);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == TC_WARP_MMA) {
        if (sh.tmem_cols == 512)
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        else
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
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
            // A stage is free once the MMAs of the tile that used it have completed - exactly what that tile's acc_full
            // barrier signals - so the issuer does not spend a second tcgen05.commit per tile on a stage-empty barrier
            // (a commit costs the single issuing thread ~200 cycles, and that thread is the serial bottleneck).
            int s = 0, tile = 0, ra = 0;
            uint32_t rph = 0;
            for (int it = 0; it < 2 * T; ++it) {
                if (it >= sh.stages) {                 // tile it - stages used this stage: wait for its accumulator-full
                    mbar_wait(smem_u32(acc_full + ra), rph);
                    if (++ra == sh.nacc) { ra = 0; rph ^= 1; }
                }
                mbar_expect_tx(smem_u32(b_full + s), stage_bytes);
                const int row = (int)(cloud0 + tile * TC_NB);
                for (int bx = 0; bx < sh.nboxes; ++bx)
                    tma_load_2d(smem_u32(smem_b + s * stage_bytes + bx * TC_BOX_BYTES), &map_b, smem_u32(b_full + s), bx * 64, row);
                if (++s == sh.stages) s = 0;
                if (++tile == T) tile = 0;
            }
        }
    } else if (warp == TC_WARP_MMA) {
        // ===================== MMA issuer: one elected thread =====================
        mbar_wait(smem_u32(a_full), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t db0 = umma_desc_sw128(smem_u32(smem_b));
        int s = 0, a = 0;
        uint32_t ph_s = 0, ph_a = 0;
        for (int it = 0; it < 2 * T; ++it) {
            if (it == T) {
                // sweep 2 multiplies the thresholds the epilogue wrote into the query operand: wait for all rows
                mbar_wait(smem_u32(a2_full), 0);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait(smem_u32(b_full + s), ph_s);            // operands landed
            mbar_wait(smem_u32(acc_empty + a), ph_a ^ 1);     // accumulator buffer drained by the epilogue
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one_sync()) {
                const uint64_t db_s = db0 + (uint64_t)((s * stage_bytes) >> 4);
                const uint32_t acc_a = tmem_base + sh.acc0 + a * TC_NB;
                for (int st = 0; st < sh.ksteps; ++st) {
                    const uint64_t db = db_s + (uint64_t)((tc_step_box(sh, st) * TC_BOX_BYTES + tc_step_off(sh, st)) >> 4);
                    umma_f16_ts(acc_a, tmem_base + st * 8, db, st ? 1u : 0u);
                }
                umma_commit(smem_u32(acc_full + a));        // accumulator ready
            }
            __syncwarp();
            if (++s == sh.stages) { s = 0; ph_s ^= 1; }
            if (++a == sh.nacc) { a = 0; ph_a ^= 1; }
        }
    } else {
        // ===================== epilogue: TMEM lane = query row; two warps share a row, one per 32-column half ======
        const int w4 = warp & 3;                       // TMEM lane quarter this warp may access
        const int cb = warp >> 2;                      // column half of every candidate tile
        const int qrow = q_row0 + w4 * 32 + lane;
        const bool row_ok = qrow < N;
        const uint32_t lane_base = (uint32_t)(w4 * 32) << 16;
        const long long grow = cloud0 + qrow;
        const bool in_table = grow < P;
        {
            // A operand: this thread's query row -> TMEM columns [0, ksteps * 8), the K steps split between the two warps
            // of the row; rows past the table are zero
            const uint8_t* src = reinterpret_cast<const uint8_t*>(a_rows + (in_table ? grow : 0) * sh.krow);
            const int half_steps = (sh.ksteps + 1) / 2;
            for (int st = cb ? half_steps : 0; st < (cb ? sh.ksteps : half_steps); ++st) {
                uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
                if (in_table) {
                    const uint4* p = reinterpret_cast<const uint4*>(src + tc_step_box(sh, st) * 128 + tc_step_off(sh, st));
                    lo = __ldg(p); hi = __ldg(p + 1);
                }
                tmem_st8(tmem_base + lane_base + (uint32_t)(st * 8), lo, hi);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(a_full));
        }
        float* stage = stage_all + warp * (TC_STAGE_BYTES / 4) + lane * 32;   // this thread's 32 staged scores
        float gm[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) gm[e] = INFINITY;
        // One 128-slot survivor list per query row, filled from both ends: the warp of column half 0 appends upwards from
        // slot 0, the warp of half 1 downwards from the last slot, so a row overflows only when BOTH halves together exceed
        // the list. (With 64 slots per half a Morton-sorted cloud overflowed ~1 % of the rows at k = 40 - the survivors of
        // a row are runs of consecutive indices, i.e. whole 32-column blocks, and two of three blocks belong to one half -
        // and every such row costs a pass of the exact kernel.)
        int cnt = 0;
        int32_t* const out_j = cand_j + (cloud0 + (row_ok ? qrow : 0)) * TC_CAP;
        float* const out_d = cand_d + (cloud0 + (row_ok ? qrow : 0)) * TC_CAP;
        const int diag_jb = diag_zero ? q_row0 + w4 * 32 : -1;       // candidate block that holds this warp's diagonal
        int a = 0, jb = cb * 32;
        uint32_t ph = 0;
        for (int it = 0; it < 2 * T; ++it) {
            if (it == T) {
                // between the sweeps: the two column halves of a row hold 32 class minima each over DISJOINT candidate
                // sets = 64 classes. Each thread sorts its own 32, the pair exchanges them through the staging blocks
                // (named barrier per warp pair); min / max(own[i], other[31-i]) is the (bitonic) lower / upper half of the
                // union, one bitonic merge sorts it: tau = kk-th smallest of the 64 class minima.
                const int pair_bar = 1 + w4;
                bitonic_sort32(gm);
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    *reinterpret_cast<float4*>(stage + ((c ^ (lane & 7)) << 2)) = make_float4(gm[4 * c], gm[4 * c + 1], gm[4 * c + 2], gm[4 * c + 3]);
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                const float* other = stage_all + (warp ^ 4) * (TC_STAGE_BYTES / 4) + lane * 32;
                float og[32];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 o = *reinterpret_cast<const float4*>(other + ((c ^ (lane & 7)) << 2));
                    og[4 * c] = o.x; og[4 * c + 1] = o.y; og[4 * c + 2] = o.z; og[4 * c + 3] = o.w;
                }
                asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");     // staging blocks are reused by sweep 2
                const bool low = kk <= 32;                                       // warp-uniform
#pragma unroll
                for (int i = 0; i < 32; ++i) gm[i] = low ? fminf(gm[i], og[31 - i]) : fmaxf(gm[i], og[31 - i]);
                bitonic_layer32<32, 16>(gm); bitonic_layer32<32, 8>(gm); bitonic_layer32<32, 4>(gm);
                bitonic_layer32<32, 2>(gm); bitonic_layer32<32, 1>(gm);
                const int want = low ? kk : kk - 32;
                float t = -INFINITY;   // sorted ascending: want-th smallest = max of the first `want` (no indexed register access)
#pragma unroll
                for (int i = 0; i < 32; ++i) t = fmaxf(t, i < want ? gm[i] : -INFINITY);
                if (cb == 0) {
                    // threshold T_i = (sqrt(tau + g) + 2 e)^2 + g, slightly enlarged, as an fp16 hi + lo pair in the
                    // extras of the query operand (the A side holds -T, the B side 1): score = d~ - T
                    const float* st_c = stats + (long long)b * tc_stats_stride(sh.C) + sh.C;
                    const float amax = __int_as_float(__ldg(reinterpret_cast<const int*>(st_c)));
                    const float cmax = __int_as_float(__ldg(reinterpret_cast<const int*>(st_c + 1)));
                    const float rmax = __int_as_float(__ldg(reinterpret_cast<const int*>(st_c + 2)));
                    const int q = row_ok ? qrow : N - 1;
                    float e, g;
                    tc_row_err(sh, __ldg(cnorm + cloud0 + q), cmax, __ldg(sqnorm + cloud0 + q), rmax, amax, e, g);
                    const float rt = sqrtf(fmaxf(t, 0.f) + g) + 2.f * e;
                    float Tv = fmaf(rt, rt, g);
                    // the fp16 hi + lo pair must not fall below T (relative 2^-21, or a subnormal ulp of the lo part for
                    // small T): enlarge, represent, and compensate once if the representation came out short
                    const float Tmin = Tv;
                    Tv = Tv * (1.0f + 3.9e-6f) + 1.3e-7f;
                    __half th, tl;
                    split2h(Tv, th, tl);
                    if (__half2float(th) + __half2float(tl) < Tmin) split2h(Tv + 2.f * (Tmin - (__half2float(th) + __half2float(tl))) + 1.3e-7f, th, tl);
                    if (row_ok) row_T[cloud0 + qrow] = __half2float(th) + __half2float(tl);
                    // rewrite the extras K step of this row: reload its 32 bytes, patch the two threshold halfs
                    uint4 lohi[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                    if (in_table) {
                        const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(a_rows + grow * sh.krow) +
                                                                        tc_step_box(sh, sh.xstep) * 128 + tc_step_off(sh, sh.xstep));
                        lohi[0] = __ldg(p); lohi[1] = __ldg(p + 1);
                    }
                    const uint32_t nth = (uint32_t)__half_as_ushort(__hneg(th)), ntl = (uint32_t)__half_as_ushort(__hneg(tl));
                    if (sh.split) {            // extras start at half 9: -T_h = half 13 (word 6, high), -T_l = half 14 (word 7, low)
                        lohi[1].z = (lohi[1].z & 0x0000ffffu) | (nth << 16);
                        lohi[1].w = (lohi[1].w & 0xffff0000u) | ntl;
                    } else {                   // extras start at half 0: halfs 4, 5 = word 2
                        lohi[0].z = nth | (ntl << 16);
                    }
                    tmem_st8(tmem_base + lane_base + (uint32_t)(sh.xstep * 8), lohi[0], lohi[1]);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(a2_full));
                }
            }
            mbar_wait(smem_u32(acc_full + a), ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // This thread's 32 scores are pulled into registers and the accumulator buffer is handed back to the MMA
            // warp BEFORE the scores are processed.
            float v[32];
            tmem_ld32_nowait(tmem_base + lane_base + (uint32_t)(sh.acc0 + a * TC_NB + cb * 32), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(acc_empty + a));
            if (jb + 32 > N) {         // last tile: columns beyond the cloud are padding / the next cloud
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = (jb + e < N) ? v[e] : INFINITY;
            }
            if (jb == diag_jb) {      // warp-uniform: this block holds the warp's diagonal
                // the reference forces d(i,i) = 0 (general_utils.py:52): the query itself always survives
#pragma unroll
                for (int e = 0; e < 32; ++e) v[e] = (jb + e == qrow) ? -FLT_MAX : v[e];
            }
            if (it < T) {
#pragma unroll
                for (int e = 0; e < 32; ++e) gm[e] = fminf(gm[e], v[e]);
            } else {
                // sweep 2: a survivor is a negative score (the threshold is part of the product): one funnel shift per
                // score collects the sign bits; the scores are staged in shared memory (16-byte chunks XOR-swizzled by
                // lane so the 128-bit stores are conflict-free), then each lane walks its own few hits
                unsigned hits = 0;
#pragma unroll
                for (int e = 31; e >= 0; --e) hits = __funnelshift_l(__float_as_uint(v[e]), hits, 1);
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
                        if (cnt < TC_CAP) {          // a collision with the other end implies n0 + n1 > TC_CAP: the row is redone
                            const int pos = cb ? TC_CAP - 1 - cnt : cnt;
                            out_j[pos] = jb + e;
                            out_d[pos] = dv;
                        }
                        ++cnt;
                    }
                }
            }
            if (++a == sh.nacc) { a = 0; ph ^= 1; }
            jb += TC_NB;
            if (it == T - 1) jb = cb * 32;
        }
        if (row_ok) cand_n[(cloud0 + qrow) * TC_HALVES + cb] = cnt;
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == TC_WARP_MMA) {
        if (sh.tmem_cols == 512)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base));
    }
}

// ----------------------------------------------------------------------------------------------- finalize
// Exact squared distance in the reference's FP32 arithmetic: the feature form is knn_feat_kernel's (knn.cu), the
// coordinate form knn3d's.
struct TcExactFeat {
    const float* x; int ldx; int C; const float* sqnorm;
    __device__ __forceinline__ float operator()(const float* xq, float qq, long long cloud0, int q, int j, int diag_zero) const {
        const float* xr = x + (cloud0 + j) * ldx;
        float acc = 0.f;
#pragma unroll 8
        for (int c4 = 0; c4 < C / 4; ++c4) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(xr) + c4);
            acc = fmaf(xq[4 * c4], v.x, acc);
            acc = fmaf(xq[4 * c4 + 1], v.y, acc);
            acc = fmaf(xq[4 * c4 + 2], v.z, acc);
            acc = fmaf(xq[4 * c4 + 3], v.w, acc);
        }
        const float nj = __ldg(sqnorm + cloud0 + j);
        float d = diag_zero ? __fadd_rn(__fsub_rn(qq, 2.0f * acc), nj) : __fadd_rn(__fsub_rn(nj, 2.0f * acc), qq);
        if (diag_zero && j == q) d = 0.f;
        return d;
    }
};
struct TcExactXyz {
    const float4* pm;
    __device__ __forceinline__ float operator()(const float* xq, float qq, long long cloud0, int q, int j, int diag_zero) const {
        const float4 p = __ldg(pm + cloud0 + j);
        const float dot = fmaf(xq[2], p.z, fmaf(xq[1], p.y, __fmul_rn(xq[0], p.x)));
        float d = diag_zero ? __fadd_rn(__fsub_rn(qq, 2.0f * dot), p.w) : __fadd_rn(__fsub_rn(p.w, 2.0f * dot), qq);
        if (diag_zero && j == q) d = 0.f;
        return d;
    }
};

// H = entries per lane: 1 when the row has at most 32 survivors, 2 up to 64, 4 up to 128.
template <int H, bool ORDER, typename Exact>
__device__ __forceinline__ void tc_finalize_row(const Exact& exact, long long cloud0, int q, long long row, int kk,
                                                int skip, int diag_zero, int n, int n0, const int32_t* __restrict__ cand_j,
                                                const float* __restrict__ cand_d, float Trow, float e, float g, float qq,
                                                const float* xq, int32_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    float sd[H];
    int sj[H];
    // survivors of the two column halves sit in [0, n0) and, filled downwards, in (TC_CAP - 1 - (n - n0), TC_CAP - 1]
    auto phys = [&](int slot) { return slot < n0 ? slot : TC_CAP - 1 - (slot - n0); };
    {
        unsigned long long key[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int slot = h * 32 + lane;
            const bool valid = slot < n;
            key[h] = fs_pack_key(valid ? __ldg(cand_d + row * TC_CAP + phys(slot)) : INFINITY,
                                 valid ? __ldg(cand_j + row * TC_CAP + phys(slot)) : FS_IDX_PAD);
        }
        fs_warp_bitonic_sort_keys<H>(key, lane);
#pragma unroll
        for (int h = 0; h < H; ++h) fs_unpack_key(key[h], sd[h], sj[h]);
    }
    // kk-th smallest score (scores are d~ - T): element kk-1 of the ascending order
    const int r_thr = kk - 1;
    float thr_s = sd[0];
#pragma unroll
    for (int h = 1; h < H; ++h) thr_s = (r_thr >> 5) == h ? sd[h] : thr_s;
    thr_s = __shfl_sync(FS_FULL_MASK, thr_s, r_thr & 31);
    // sqrt-space classification around the kk-th approximate distance d~_k = thr_s + T
    const float dk = fmaxf(thr_s + Trow, 0.f);
    const float r_lo = fmaxf(sqrtf(fmaxf(dk - g, 0.f)) - 2.f * e, 0.f);
    const float r_hi = sqrtf(dk + g) + 2.f * e;
    const float lo = fmaf(r_lo, r_lo, -g) - Trow;        // back to score units
    const float hi = fmaf(r_hi, r_hi, g) - Trow;
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
    if (!ORDER && n_amb == slots) {
        // the approximation alone decides the set: ranks [0, kk) in approximate order
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const int r = h * 32 + lane;
            if (r >= skip && r < kk) out[r - skip] = sj[h];
        }
        return;
    }
    // exact distances for the ambiguous entries (ORDER: for every entry that can be part of the result)
    float ex[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        ex[h] = INFINITY;
        if (amb[h] || (ORDER && in_[h])) ex[h] = exact(xq, qq, cloud0, q, sj[h], diag_zero);
    }
    if (ORDER) {
        // coordinates: select AND order by exact (distance, index) among sure-in + ambiguous entries: the result is the
        // exact kernel's, order included
        unsigned long long key[H];
#pragma unroll
        for (int h = 0; h < H; ++h) key[h] = fs_pack_key(ex[h], (amb[h] || in_[h]) ? sj[h] : FS_IDX_PAD);
        fs_warp_bitonic_sort_keys<H>(key, lane);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float dd; int jj;
            fs_unpack_key(key[h], dd, jj);
            const int r = h * 32 + lane;
            if (r >= skip && r < kk) out[r - skip] = jj;
        }
        return;
    }
    // rank of every ambiguous entry among the ambiguous ones by exact (distance, index)
    int rank[H];
#pragma unroll
    for (int h = 0; h < H; ++h) rank[h] = 0;
#pragma unroll
    for (int sh2 = 0; sh2 < H; ++sh2) {
        unsigned todo = __ballot_sync(FS_FULL_MASK, amb[sh2]);      // only the ambiguous lanes are broadcast
        while (todo) {
            const int sl = __ffs(todo) - 1;
            todo &= todo - 1;
            const float xd = __shfl_sync(FS_FULL_MASK, ex[sh2], sl);
            const int xj = __shfl_sync(FS_FULL_MASK, sj[sh2], sl);
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

template <bool XYZ>
__global__ void __launch_bounds__(256)
knn_tc_finalize_kernel(const float* __restrict__ x, int ldx, const float4* __restrict__ xyz_pm, int N, long long P, int k,
                       int self_loop, int diag_zero, const TcShape sh, const int32_t* __restrict__ cand_j,
                       const float* __restrict__ cand_d, const int32_t* __restrict__ cand_n, const float* __restrict__ row_T,
                       const float* __restrict__ sqnorm, const float* __restrict__ cnorm, const float* __restrict__ stats,
                       int32_t* __restrict__ idx, uint8_t* __restrict__ redo) {
    __shared__ float xq_all[8][XYZ ? 4 : 256];
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
    const float* st_c = stats + (long long)b * tc_stats_stride(sh.C) + sh.C;
    const float amax = __int_as_float(__ldg(reinterpret_cast<const int*>(st_c)));
    const float cmax = __int_as_float(__ldg(reinterpret_cast<const int*>(st_c + 1)));
    const float rmax = __int_as_float(__ldg(reinterpret_cast<const int*>(st_c + 2)));
    const float qq = __ldg(sqnorm + row);
    float e, g;
    tc_row_err(sh, __ldg(cnorm + row), cmax, qq, rmax, amax, e, g);
    const float Trow = __ldg(row_T + row);
    // overflow of the row's list, too few survivors, or a cloud with NaN / Inf (amax, the norms and with them e, g and
    // T are then not finite): the exact kernel redoes this query
    const bool finite = (e + g + Trow + amax) < INFINITY;          // false for NaN and Inf
    if (n > TC_CAP || n < kk || !finite) {
        if (lane == 0) redo[row] = 1;
        return;
    }
    if (lane == 0) redo[row] = 0;
    float* xq = xq_all[warp];
    if (XYZ) {
        if (lane == 0) {
            const float4 p = __ldg(xyz_pm + row);
            xq[0] = p.x; xq[1] = p.y; xq[2] = p.z; xq[3] = p.w;
        }
    } else {
        for (int c = lane; c < sh.C; c += 32) xq[c] = __ldg(x + row * ldx + c);
    }
    __syncwarp();
    int32_t* out = idx + row * k;
    if (XYZ) {
        const TcExactXyz ex{xyz_pm};
        if (n <= 32) tc_finalize_row<1, true>(ex, cloud0, q, row, kk, skip, diag_zero, n, n0, cand_j, cand_d, Trow, e, g, qq, xq, out);
        else if (n <= 64) tc_finalize_row<2, true>(ex, cloud0, q, row, kk, skip, diag_zero, n, n0, cand_j, cand_d, Trow, e, g, qq, xq, out);
        else tc_finalize_row<4, true>(ex, cloud0, q, row, kk, skip, diag_zero, n, n0, cand_j, cand_d, Trow, e, g, qq, xq, out);
    } else {
        const TcExactFeat ex{x, ldx, sh.C, sqnorm};
        if (n <= 32) tc_finalize_row<1, false>(ex, cloud0, q, row, kk, skip, diag_zero, n, n0, cand_j, cand_d, Trow, e, g, qq, xq, out);
        else if (n <= 64) tc_finalize_row<2, false>(ex, cloud0, q, row, kk, skip, diag_zero, n, n0, cand_j, cand_d, Trow, e, g, qq, xq, out);
        else tc_finalize_row<4, false>(ex, cloud0, q, row, kk, skip, diag_zero, n, n0, cand_j, cand_d, Trow, e, g, qq, xq, out);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_operand_map(EncodeTiledFn fn, CUtensorMap* map, void* base, long long rows, int krow) {
    const cuuint64_t dims[2] = {(cuuint64_t)krow, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)krow * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)TC_NB};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

// The driver entry point is resolved per call (cudaGetDriverEntryPoint is a table lookup): the library keeps no mutable
// global state, as include/fissure_b200.h promises.
int get_encode_fn(EncodeTiledFn* out) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return (int)e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
    *out = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct TcWorkspace {
    __half* A; __half* Bm; int32_t* cand_j; float* cand_d; int32_t* cand_n; float* sqnorm; float* cnorm; float* row_T;
    float* stats; float4* xyz_pm; uint8_t* redo; size_t stats_bytes; size_t table_bytes; size_t total;
};
TcWorkspace tc_carve(void* workspace, int B, int N, const TcShape& sh) {
    const size_t P = (size_t)B * N;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    size_t off = 0;
    TcWorkspace w{};
    auto take = [&](size_t bytes) { uint8_t* p = ws ? ws + off : nullptr; off += align_up(bytes, 256); return p; };
    w.table_bytes = P * sh.krow * 2;
    w.A = reinterpret_cast<__half*>(take(w.table_bytes));
    w.Bm = reinterpret_cast<__half*>(take(w.table_bytes));
    w.cand_j = reinterpret_cast<int32_t*>(take(P * TC_CAP * 4));
    w.cand_d = reinterpret_cast<float*>(take(P * TC_CAP * 4));
    w.cand_n = reinterpret_cast<int32_t*>(take(P * 4 * TC_HALVES));
    w.sqnorm = reinterpret_cast<float*>(take(P * 4));
    w.cnorm = reinterpret_cast<float*>(take(P * 4));
    w.row_T = reinterpret_cast<float*>(take(P * 4));
    w.stats_bytes = align_up((size_t)B * tc_stats_stride(sh.C) * 4, 256);
    w.stats = reinterpret_cast<float*>(take(w.stats_bytes));
    w.xyz_pm = reinterpret_cast<float4*>(take(sh.C == 3 ? P * 16 : 0));
    w.redo = take(P);                                   // last block: fs_knn_feat_tc_redo_offset
    w.total = off;
    return w;
}

bool tc_supported(int B, int N, int C, int k, int self_loop, TcShape* sh) {
    const int kk = k + (self_loop ? 0 : 1);
    TcShape s;
    if (!tc_make_shape(C, &s)) return false;
    if (kk > TC_MAX_KK || N < 64 || N > 32768 || (long long)B * N > 0x7fffffff / s.krow || B > 65535) return false;
    if (sh) *sh = s;
    return true;
}

// select -> finalize (the operand tables are already in the workspace)
template <bool XYZ>
int tc_run(cudaStream_t stream, const TcShape& sh, const TcWorkspace& w, const float* x, int ldx, int B, int N, int k,
           int self_loop, int diag_zero, int32_t* idx) {
    const long long P = (long long)B * N;
    const int kk = k + (self_loop ? 0 : 1);
    EncodeTiledFn encode = nullptr;
    int e = get_encode_fn(&encode);
    if (e) return e;
    CUtensorMap map_b;
    e = make_operand_map(encode, &map_b, w.Bm, P, sh.krow);
    if (e) return e;
    const size_t smem = tc_smem_bytes(sh);
    FS_CUDA_TRY(cudaFuncSetAttribute(knn_tc_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(fs_div_up(N, TC_M), B);
    knn_tc_select_kernel<<<grid, TC_THREADS, smem, stream>>>(w.A, P, map_b, N, kk, diag_zero, sh, w.cnorm, w.sqnorm, w.stats,
                                                              w.cand_j, w.cand_d, w.cand_n, w.row_T);
    FS_RETURN_IF_LAUNCH_FAILED();
    knn_tc_finalize_kernel<XYZ><<<fs_div_up(P, 8), 256, 0, stream>>>(x, ldx, w.xyz_pm, N, P, k, self_loop, diag_zero, sh, w.cand_j,
                                                                     w.cand_d, w.cand_n, w.row_T, w.sqnorm, w.cnorm, w.stats,
                                                                     idx, w.redo);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

}  // namespace

// Defined in knn.cu.
int fs_knn_feat_masked(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                       int diag_zero, int32_t* idx, float* dist2, const float* sqnorm, const uint8_t* redo);
int fs_knn_feat_exact(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                      int diag_zero, int32_t* idx, float* dist2, float* sqnorm_ws);

extern "C" size_t fs_knn_feat_tc_workspace_bytes(int B, int N, int C, int k) {
    (void)k;
    TcShape sh;
    if (!tc_make_shape(C, &sh)) {
        // shapes outside the tensor-core path only need the row norms of the exact kernel (+ the flags block)
        return align_up((size_t)B * N * 4, 256) + align_up((size_t)B * N, 256);
    }
    return tc_carve(nullptr, B, N, sh).total;
}

extern "C" size_t fs_knn_feat_tc_redo_offset(int B, int N, int C, int k) {
    return fs_knn_feat_tc_workspace_bytes(B, N, C, k) - align_up((size_t)B * N, 256);
}

extern "C" int fs_knn_feat_tc_supported(int B, int N, int C, int k, int self_loop) {
    return (C != 3 && tc_supported(B, N, C, k, self_loop, nullptr)) ? 1 : 0;
}

extern "C" int fs_knn3d_tc_supported(int B, int N, int k, int self_loop) {
    return tc_supported(B, N, 3, k, self_loop, nullptr) ? 1 : 0;
}

extern "C" size_t fs_knn3d_tc_workspace_bytes(int B, int N, int k) { return fs_knn_feat_tc_workspace_bytes(B, N, 3, k); }

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

    TcShape sh;
    // distances requested (public knn(..., return_dist=True)) or shape outside the tensor-core path: exact kernel
    if (dist2 || C == 3 || !tc_supported(B, N, C, k, self_loop, &sh))
        return fs_knn_feat_exact(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, dist2, static_cast<float*>(workspace));
    const TcWorkspace w = tc_carve(workspace, B, N, sh);

    // 1. prep: per-cloud column sums and absolute maximum, then the fp16 operand tables
    FS_CUDA_TRY(cudaMemsetAsync(w.stats, 0, w.stats_bytes, stream));
    tc_cloud_stats_kernel<<<dim3(32, B), 256, 0, stream>>>(x, (long long)N * ldx, ldx, 1, N, C, w.stats);
    if (C == 64) tc_split_feat_kernel<1><<<fs_div_up(P, 32), 256, 0, stream>>>(x, ldx, P, N, sh, w.stats, w.A, w.Bm, w.cnorm, w.sqnorm);
    else if (C == 128) tc_split_feat_kernel<2><<<fs_div_up(P, 16), 256, 0, stream>>>(x, ldx, P, N, sh, w.stats, w.A, w.Bm, w.cnorm, w.sqnorm);
    else tc_split_feat_kernel<4><<<fs_div_up(P, 8), 256, 0, stream>>>(x, ldx, P, N, sh, w.stats, w.A, w.Bm, w.cnorm, w.sqnorm);
    FS_RETURN_IF_LAUNCH_FAILED();
    // 2. tensor-core sweeps, 3. finalize
    int e = tc_run<false>(stream, sh, w, x, ldx, B, N, k, self_loop, diag_zero, idx);
    if (e) return e;
    // 4. the exact kernel on rows without a certificate
    return fs_knn_feat_masked(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, nullptr, w.sqnorm, w.redo);
}

// Coordinate kNN (fs_knn3d contract, include/fissure_b200.h) through the tensor-core kernels. The neighbour indices come
// out in exact ascending (distance, index) order, like the SIMT kernels'; distances are not produced.
extern "C" int fs_knn3d_tc(int device, fs_stream_t stream_, const float* coords, long long batch_stride, long long chan_stride,
                           long long point_stride, int B, int N, int k, int self_loop, int diag_zero, int32_t* idx,
                           void* workspace, size_t workspace_bytes) {
    if (B < 0 || N <= 0 || k <= 0) return FS_ERR_BAD_ARG;
    if (B == 0) return FS_OK;
    if (!coords || !idx || !workspace) return FS_ERR_BAD_ARG;
    const int kk = k + (self_loop ? 0 : 1);
    if (kk > N || kk > FS_MAX_K + 1) return FS_ERR_BAD_ARG;
    TcShape sh;
    if (!tc_supported(B, N, 3, k, self_loop, &sh)) return FS_ERR_UNSUPPORTED;
    if (workspace_bytes < fs_knn3d_tc_workspace_bytes(B, N, k)) return FS_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(workspace) & 255) return FS_ERR_ALIGNMENT;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    const TcWorkspace w = tc_carve(workspace, B, N, sh);
    FS_CUDA_TRY(cudaMemsetAsync(w.stats, 0, w.stats_bytes, stream));
    tc_cloud_stats_kernel<<<dim3(16, B), 192, 0, stream>>>(coords, batch_stride, point_stride, chan_stride, N, 3, w.stats);
    tc_split_xyz_kernel<<<fs_div_up(P, 256), 256, 0, stream>>>(coords, batch_stride, chan_stride, point_stride, P, N, sh, w.stats,
                                                              w.A, w.Bm, w.cnorm, w.sqnorm, w.xyz_pm);
    FS_RETURN_IF_LAUNCH_FAILED();
    int e = tc_run<true>(stream, sh, w, nullptr, 0, B, N, k, self_loop, diag_zero, idx);
    if (e) return e;
    // rows without a certificate: the exact feature kernel on the point-major copy (x, y, z, |p|^2): C = 3, ld = 4 - the
    // same dot-product chain and the same norms as the coordinate kernels of knn.cu
    return fs_knn_feat_masked(stream, reinterpret_cast<const float*>(w.xyz_pm), 4, B, N, 3, k, self_loop, diag_zero, idx, nullptr,
                              w.sqnorm, w.redo);
}
