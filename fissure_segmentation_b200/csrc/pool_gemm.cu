// Global feature forward on tcgen05: out[b, c] = max_r (+-) (X W^T)[b N + r, c] with its arg-max row, WITHOUT writing
// the (B*N x C) product (models/dgcnn.py:123-126, 156: Conv1d + BatchNorm + LeakyReLU + AdaptiveMaxPool1d; the monotone
// BN + LeakyReLU is applied to the B x C selected values afterwards, fs_bn_act_apply).
//
// Transposed product Z^T = W' X^T: TMEM lane = output channel, TMEM column = point, so the max over the points of a
// cloud is a per-thread running maximum over the columns a thread reads - no cross-lane traffic. W' = sign(gamma) W
// (folded on the host side) turns "max where gamma >= 0, min where gamma < 0" into a pure max of keys.
//   * persistent CTAs over (cloud, 128-point tile) items; per item the X tile (128 x K bf16) is loaded once by TMA
//     (128B swizzle, K / 64 boxes) and multiplied against the C / 128 channel tiles of W' streamed through two stages;
//   * tcgen05.mma cta_group::1 kind::f16, M = 128 channels, N = 128 points, K / 16 steps, four 128-column fp32
//     accumulators in TMEM (all 512 columns): the MMA warp runs up to three channel tiles ahead of the epilogue;
//   * 8 epilogue warps (TMEM lane quarter x column half): tcgen05.ld 2 x 32 columns, running (key, row) maximum,
//     one 64-bit atomicMax per thread and channel tile into packed[b, c] = (ordered key << 32) | ~row (ties -> lower row,
//     like a sequential arg-max; same encoding as pool_reduce_kernel, decoded by fs_pool_decode).
// BatchNorm statistics of the product do not need the product either: sum y = W colsum(X), sum y^2 = diag(W (X^T X) W^T)
// (fs_pool_stats_from_gram) - the Gram matrix is needed by the backward anyway (csrc/heads.cu).
#include <cuda.h>

#include "fs_common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace tcx;

constexpr int PG_M = 128;               // channels per tile (TMEM lanes)
constexpr int PG_N = 128;               // points per tile (TMEM columns)
constexpr int PG_EPI_WARPS = 8;
constexpr int PG_WARP_TMA = 8;
constexpr int PG_WARP_MMA = 9;
constexpr int PG_THREADS = 320;
constexpr int PG_BOX_BYTES = 64 * 2 * 128;      // one TMA box: 64 bf16 (128 bytes) x 128 rows
constexpr int PG_ACC = 4;

__device__ __forceinline__ unsigned int pg_ordered_u32(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int KB>       // K = 64 * KB
__global__ void __launch_bounds__(PG_THREADS, 1)
pool_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, int B, int N, int C,
                 unsigned long long* __restrict__ packed) {
    constexpr int TILE_BYTES = KB * PG_BOX_BYTES;
    extern __shared__ uint8_t pg_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(pg_smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* x_s = smem;                            // 2 stages
    uint8_t* w_s = smem + 2 * TILE_BYTES;           // 2 stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * TILE_BYTES);
    uint64_t* x_full = bars;                        // [2]
    uint64_t* x_empty = bars + 2;                   // [2]
    uint64_t* w_full = bars + 4;                    // [2]
    uint64_t* acc_full = bars + 6;                  // [4]
    uint64_t* acc_empty = bars + 10;                // [4]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_cloud = (N + PG_N - 1) / PG_N;
    const long long n_items = (long long)B * tiles_per_cloud;
    const int n_ct = C / PG_M;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(x_full + s), 1);
            mbar_init(smem_u32(x_empty + s), 1);
            mbar_init(smem_u32(w_full + s), 1);
        }
        for (int a = 0; a < PG_ACC; ++a) {
            mbar_init(smem_u32(acc_full + a), 1);
            mbar_init(smem_u32(acc_empty + a), PG_EPI_WARPS);
        }
        mbar_init_fence();
    }
    if (warp == PG_WARP_MMA) tmem_alloc512(smem_u32(tmem_slot));
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == PG_WARP_TMA) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            long long g = 0;                        // (item, channel tile) pairs issued by this CTA
            int il = 0;                             // items issued by this CTA
            for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++il) {
                const int xs = il & 1;
                if (il >= 2) mbar_wait(smem_u32(x_empty + xs), (uint32_t)(((il >> 1) - 1) & 1));
                const int b = (int)(item / tiles_per_cloud), t = (int)(item - (long long)b * tiles_per_cloud);
                const int row0 = b * N + t * PG_N;
                mbar_expect_tx(smem_u32(x_full + xs), TILE_BYTES);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(smem_u32(x_s + xs * TILE_BYTES + kb * PG_BOX_BYTES), &map_x, smem_u32(x_full + xs), kb * 64, row0);
                for (int ct = 0; ct < n_ct; ++ct, ++g) {
                    const int ws = (int)(g & 1);
                    // the stage is free once the MMAs of pair g - 2 have completed = that pair's accumulator-full barrier
                    // (pair g + 2, the next user of the same accumulator, cannot have been issued yet: it needs this stage)
                    if (g >= 2) mbar_wait(smem_u32(acc_full + ((g - 2) & 3)), (uint32_t)(((g - 2) >> 2) & 1));
                    mbar_expect_tx(smem_u32(w_full + ws), TILE_BYTES);
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_2d(smem_u32(w_s + ws * TILE_BYTES + kb * PG_BOX_BYTES), &map_w, smem_u32(w_full + ws), kb * 64, ct * PG_M);
                }
            }
        }
    } else if (warp == PG_WARP_MMA) {
        // ===================== MMA issuer =====================
        constexpr uint32_t IDESC = instr_desc_f16(PG_M, PG_N, 1, 0, 0);
        long long g = 0;
        int il = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x, ++il) {
            const int xs = il & 1;
            mbar_wait(smem_u32(x_full + xs), (uint32_t)((il >> 1) & 1));
            const uint32_t xb = smem_u32(x_s + xs * TILE_BYTES);
            for (int ct = 0; ct < n_ct; ++ct, ++g) {
                const int ws = (int)(g & 1), a = (int)(g & 3);
                mbar_wait(smem_u32(w_full + ws), (uint32_t)((g >> 1) & 1));
                mbar_wait(smem_u32(acc_empty + a), (uint32_t)(((g >> 2) & 1) ^ 1));
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t wb = smem_u32(w_s + ws * TILE_BYTES);
#pragma unroll
                    for (int ks = 0; ks < 4 * KB; ++ks) {
                        const uint64_t da = smem_desc_sw128(wb + (ks >> 2) * PG_BOX_BYTES + (ks & 3) * 32, 16, 1024);
                        const uint64_t db = smem_desc_sw128(xb + (ks >> 2) * PG_BOX_BYTES + (ks & 3) * 32, 16, 1024);
                        umma_ss(tmem_base + a * PG_N, da, db, IDESC, ks ? 1u : 0u);
                    }
                    umma_commit(smem_u32(acc_full + a));
                    if (ct == n_ct - 1) umma_commit(smem_u32(x_empty + xs));
                }
                __syncwarp();
            }
        }
    } else {
        // ===================== epilogue: TMEM lane = channel, column = point =====================
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = warp >> 2;                   // 64-column half of every accumulator
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        long long g = 0;
        for (long long item = blockIdx.x; item < n_items; item += gridDim.x) {
            const int b = (int)(item / tiles_per_cloud), t = (int)(item - (long long)b * tiles_per_cloud);
            const int pt0 = t * PG_N + half * 64;      // row (within the cloud) of this thread's first column
            const bool ragged = t * PG_N + PG_N > N;   // warp-uniform: columns past the cloud hold the next cloud / zero fill
            for (int ct = 0; ct < n_ct; ++ct, ++g) {
                const int a = (int)(g & 3);
                mbar_wait(smem_u32(acc_full + a), (uint32_t)((g >> 2) & 1));
                tc_fence_after();
                float v[64];
                const uint32_t col = tmem_base + lane_base + (uint32_t)(a * PG_N + half * 64);
                tmem_ld32_nowait(col, v);
                tmem_ld32_nowait(col + 32, v + 32);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(acc_empty + a));      // handed back before the scores are processed
                float best = -INFINITY;
                int barg = -1;
                if (!ragged) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const bool better = v[j] > best;
                        best = better ? v[j] : best;
                        barg = better ? j : barg;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 64; ++j) {
                        const bool better = (pt0 + j < N) && v[j] > best;
                        best = better ? v[j] : best;
                        barg = better ? j : barg;
                    }
                }
                if (barg >= 0) {
                    const int c = ct * PG_M + quarter * 32 + lane;
                    const unsigned long long pk = ((unsigned long long)pg_ordered_u32(best) << 32) |
                                                  (unsigned long long)(0xffffffffu - (unsigned)(pt0 + barg));
                    atomicMax(packed + (long long)b * C + c, pk);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PG_WARP_MMA) tmem_dealloc512(tmem_base);
}

// stats (fs_stats_commit layout, pivot 0): sum y = w_c . colsum, sum y^2 = (W G)_c . w_c, fp64 accumulation. One warp
// per channel.
template <typename T>
__global__ void __launch_bounds__(256)
pool_stats_from_gram_kernel(const T* __restrict__ w, int ldw, const float* __restrict__ wg, const float* __restrict__ colsum,
                            int C, int K, double* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    double s1 = 0.0, s2 = 0.0;
    for (int k = lane; k < K; k += 32) {
        const float wv = sizeof(T) == 2 ? __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(w + (long long)c * ldw + k))
                                        : *reinterpret_cast<const float*>(w + (long long)c * ldw + k);
        s1 += (double)wv * (double)__ldg(colsum + k);
        s2 += (double)wv * (double)__ldg(wg + (long long)c * K + k);
    }
    s1 = fs_warp_sum(s1);
    s2 = fs_warp_sum(s2);
    if (lane == 0) {
        stats[c] = s1;
        stats[C + c] = s2 > 0.0 ? s2 : 0.0;
        stats[2 * C + c] = 0.0;
    }
}

typedef CUresult (*PgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bf16 [rows, K] table with row pitch ld (elements): boxes of 64 columns x 128 rows, 128-byte swizzle, zero fill outside
int pg_make_map(CUtensorMap* map, const void* base, long long rows, int K, int ld) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess) return (int)e;
    if (!fn || qres != cudaDriverEntryPointSuccess) return (int)cudaErrorNotSupported;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<PgEncodeFn>(fn)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                                                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int KB>
int pg_launch(cudaStream_t stream, const CUtensorMap& mx, const CUtensorMap& mw, int B, int N, int C, unsigned long long* packed) {
    const size_t smem = (size_t)4 * KB * PG_BOX_BYTES + 1024 + 256;
    FS_CUDA_TRY(cudaFuncSetAttribute(pool_gemm_kernel<KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long items = (long long)B * ((N + PG_N - 1) / PG_N);
    const int grid = (int)(items < FS_NUM_SMS ? items : FS_NUM_SMS);
    pool_gemm_kernel<KB><<<grid, PG_THREADS, smem, stream>>>(mx, mw, B, N, C, packed);
    return FS_OK;
}

}  // namespace

extern "C" int fs_pool_gemm_supported(int B, int N, int C, int K) {
    return B > 0 && N > 0 && (long long)B * N < (1ll << 31) && C >= PG_M && C % PG_M == 0 && (K == 64 || K == 128 || K == 192);
}

extern "C" int fs_pool_gemm(int device, fs_stream_t stream_, const void* x, int ldx, const void* w_signed, int B, int N, int C,
                            int K, unsigned long long* packed) {
    if (!x || !w_signed || !packed || ldx < K) return FS_ERR_BAD_ARG;
    if (!fs_pool_gemm_supported(B, N, C, K) || ldx % 8 || ((uintptr_t)x & 15) || ((uintptr_t)w_signed & 15)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    CUtensorMap mx, mw;
    int rc = pg_make_map(&mx, x, (long long)B * N, K, ldx);
    if (rc) return rc;
    rc = pg_make_map(&mw, w_signed, C, K, K);
    if (rc) return rc;
    if (K == 64) rc = pg_launch<1>(stream, mx, mw, B, N, C, packed);
    else if (K == 128) rc = pg_launch<2>(stream, mx, mw, B, N, C, packed);
    else rc = pg_launch<3>(stream, mx, mw, B, N, C, packed);
    if (rc != FS_OK) return rc;
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_pool_stats_from_gram(int device, fs_stream_t stream_, const void* w, int dtype, int ldw, const float* wg,
                                       const float* colsum, int C, int K, double* stats) {
    if (!w || !wg || !colsum || !stats || C <= 0 || K <= 0 || ldw < K) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = (C + 7) / 8;
    if (dtype == FS_BF16)
        pool_stats_from_gram_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)w, ldw, wg, colsum, C, K, stats);
    else
        pool_stats_from_gram_kernel<<<grid, 256, 0, stream>>>((const float*)w, ldw, wg, colsum, C, K, stats);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
