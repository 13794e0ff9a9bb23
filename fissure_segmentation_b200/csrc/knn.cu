// kNN graph build: 3-D coordinates (batched and offset-segmented) and exact FP32 feature space.
// Distances follow the reference's arithmetic (utils/general_utils.py:43-53): squared norms are
// summed without FMA contraction, the dot product is an FMA chain, and the expansion is evaluated
// as (xx_i - 2*dot) + xx_j with the diagonal forced to 0.
#include "warp_select.cuh"

namespace {

constexpr int KNN_WARPS = 8;
constexpr int KNN_THREADS = KNN_WARPS * 32;
constexpr int KNN3D_QUERIES_PER_WARP = 8;
constexpr int KNN3D_TILE_Q = KNN_WARPS * KNN3D_QUERIES_PER_WARP;  // queries per CTA
constexpr int KNN3D_MAX_CHUNK = 8192;                             // staged candidates per pass

__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// ---------------------------------------------------------------------------------------------
// Batched 3-D kNN. grid = (ceil(N / TILE_Q), B). The cloud (or an 8192-point chunk of it) is staged
// in shared memory as SoA x|y|z|norm; each warp streams all candidates for one query at a time.
// ---------------------------------------------------------------------------------------------
template <int KPL>
__global__ void __launch_bounds__(KNN_THREADS)
knn3d_batch_kernel(const float* __restrict__ coords, long long batch_stride, long long chan_stride,
                   long long point_stride, int N, int k, int self_loop, int diag_zero, int chunk,
                   int32_t* __restrict__ idx, float* __restrict__ dist2) {
    extern __shared__ float smem[];
    float* sx = smem;
    float* sy = sx + chunk;
    float* sz = sy + chunk;
    float* sn = sz + chunk;
    float* qd_all = sn + chunk;
    int* qi_all = reinterpret_cast<int*>(qd_all + KNN_WARPS * 64);

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const float* cb = coords + (long long)b * batch_stride;
    const int kk = k + (self_loop ? 0 : 1);
    const int nchunks = (N + chunk - 1) / chunk;

    FsWarpSelect<KPL> sel;

    for (int r = 0; r < KNN3D_QUERIES_PER_WARP; ++r) {
        const int q = blockIdx.x * KNN3D_TILE_Q + r * KNN_WARPS + warp;
        const bool active = q < N;
        float qx = 0.f, qy = 0.f, qz = 0.f;
        if (active) {
            qx = __ldg(cb + (long long)q * point_stride);
            qy = __ldg(cb + chan_stride + (long long)q * point_stride);
            qz = __ldg(cb + 2 * chan_stride + (long long)q * point_stride);
        }
        const float qq = sqnorm3(qx, qy, qz);
        sel.init(qd_all + warp * 64, qi_all + warp * 64, kk);

        for (int c = 0; c < nchunks; ++c) {
            const int c0 = c * chunk;
            const int cn = min(chunk, N - c0);
            if (nchunks > 1 || r == 0) {
                __syncthreads();
                for (int j = threadIdx.x; j < cn; j += KNN_THREADS) {
                    const long long o = (long long)(c0 + j) * point_stride;
                    float x = __ldg(cb + o), y = __ldg(cb + chan_stride + o), z = __ldg(cb + 2 * chan_stride + o);
                    sx[j] = x; sy[j] = y; sz[j] = z; sn[j] = sqnorm3(x, y, z);
                }
                __syncthreads();
            }
            if (active) {
                for (int base = 0; base < cn; base += 32) {
                    const int j = base + lane;
                    const bool valid = j < cn;
                    float d = INFINITY;
                    if (valid) {
                        float dot = fmaf(qz, sz[j], fmaf(qy, sy[j], __fmul_rn(qx, sx[j])));
                        // general_utils.py:51 evaluates (xx_i - 2 x.y) + xx_j; dgcnn_opensrc.py:37 negates
                        // (-xx_j + 2 x.y) - xx_i, i.e. the same sum associated the other way round
                        d = diag_zero ? __fadd_rn(__fsub_rn(qq, 2.0f * dot), sn[j])
                                      : __fadd_rn(__fsub_rn(sn[j], 2.0f * dot), qq);
                        if (diag_zero && (c0 + j) == q) d = 0.f;
                    }
                    sel.offer(d, c0 + j, valid);
                }
            }
        }
        if (active) {
            sel.finish();
            const long long row = ((long long)b * N + q) * k;
            sel.store(self_loop ? 0 : 1, idx + row, dist2 ? dist2 + row : nullptr, 0, 0, INFINITY);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Batched 3-D kNN, register-resident variant for clouds of up to 32*NPL points (N <= 2048).
// One warp per query, lane l owns candidates j = i*32 + l and keeps ALL its NPL distances in registers.
//   sweep    : distances + running minimum of CPL interleaved classes per lane (one FMNMX per candidate,
//              no ballot, no branch);
//   bound    : tau = kk-th smallest of the 32*CPL class minima — kk distinct candidates lie below it, so it
//              bounds the kk-th smallest distance; its expected rank is M ln(M/(M-kk)), M = 32*CPL;
//   collect  : lanes append their few entries <= tau to a shared queue (about 1.2 kk entries);
//   select   : the warp-select of warp_select.cuh orders the survivors by (distance, index).
// Same total order as the streaming kernel, hence the same result. A query whose survivors overflow the
// queue falls back to streaming selection over the staged cloud.
// ---------------------------------------------------------------------------------------------
constexpr int KNN3R_WARPS = 8;
constexpr int KNN3R_THREADS = KNN3R_WARPS * 32;
constexpr int KNN3R_QPW = 4;                                  // queries per warp (sequential)
constexpr int KNN3R_TILE_Q = KNN3R_WARPS * KNN3R_QPW;

template <int NPL, int CPL, int KPL>
__global__ void __launch_bounds__(KNN3R_THREADS, 2)
knn3d_regs_kernel(const float* __restrict__ coords, long long batch_stride, long long chan_stride,
                  long long point_stride, int N, int k, int self_loop, int diag_zero,
                  int32_t* __restrict__ idx, float* __restrict__ dist2) {
    constexpr int CAP = 32 * CPL * 2;                         // survivor queue entries per warp
    extern __shared__ float smem[];
    constexpr int chunk = NPL * 32;
    float4* sp = reinterpret_cast<float4*>(smem);             // staged cloud: (x, y, z, |p|^2) per point, one LDS.128
    float* qd_all = smem + 4 * chunk;                         // [warps][64] warp-select queue
    int* qi_all = reinterpret_cast<int*>(qd_all + KNN3R_WARPS * 64);
    float* hd_all = reinterpret_cast<float*>(qi_all + KNN3R_WARPS * 64);   // [warps][CAP] survivors
    int* hi_all = reinterpret_cast<int*>(hd_all + KNN3R_WARPS * CAP);

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const float* cb = coords + (long long)b * batch_stride;
    const int kk = k + (self_loop ? 0 : 1);

    for (int j = threadIdx.x; j < chunk; j += KNN3R_THREADS) {
        float4 p = make_float4(0.f, 0.f, 0.f, INFINITY);      // padding beyond the cloud: distance +inf, no per-candidate bound check
        if (j < N) {
            const long long o = (long long)j * point_stride;
            p.x = __ldg(cb + o); p.y = __ldg(cb + chan_stride + o); p.z = __ldg(cb + 2 * chan_stride + o);
            p.w = sqnorm3(p.x, p.y, p.z);
        }
        sp[j] = p;
    }
    __syncthreads();

    float* hd = hd_all + warp * CAP;
    int* hi = hi_all + warp * CAP;
    FsWarpSelect<KPL> sel;

    for (int r = 0; r < KNN3R_QPW; ++r) {
        const int q = blockIdx.x * KNN3R_TILE_Q + r * KNN3R_WARPS + warp;
        if (q >= N) break;                                    // warp-uniform
        const float4 qp = sp[q];
        const float qx = qp.x, qy = qp.y, qz = qp.z, qq = qp.w;
        const int self_i = diag_zero ? (q >> 5) : -1, self_lane = q & 31;
        float dist[NPL];
        float cm[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) cm[c] = INFINITY;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const float4 p = sp[i * 32 + lane];
            const float dot = fmaf(qz, p.z, fmaf(qy, p.y, __fmul_rn(qx, p.x)));
            float d = diag_zero ? __fadd_rn(__fsub_rn(qq, 2.0f * dot), p.w) : __fadd_rn(__fsub_rn(p.w, 2.0f * dot), qq);
            if (i == self_i && lane == self_lane) d = 0.f;    // general_utils.py:52
            dist[i] = d;
            cm[i % CPL] = fminf(cm[i % CPL], d);
        }
        // tau = kk-th smallest of the 32*CPL class minima
        float tau;
        if (CPL == 1) {
            float td = cm[0];
            int ti = lane;
            fs_warp_bitonic_sort(td, ti, lane);
            tau = __shfl_sync(FS_FULL_MASK, td, kk - 1);
        } else {
            sel.init(qd_all + warp * 64, qi_all + warp * 64, kk);
#pragma unroll
            for (int c = 0; c < CPL; ++c) sel.offer(cm[c], c * 32 + lane, true);
            sel.finish();
            int tj;
            sel.get(kk - 1, tau, tj);
        }
        // collect survivors without atomics: per-lane hit count, exclusive scan over the lanes, private writes
        int mine = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) mine += (dist[i] <= tau) ? 1 : 0;
        int incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FS_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        const int n = __shfl_sync(FS_FULL_MASK, incl, 31);
        if (n <= CAP) {
            int pos = incl - mine;
#pragma unroll
            for (int i = 0; i < NPL; ++i) {
                if (dist[i] <= tau) { hd[pos] = dist[i]; hi[pos] = i * 32 + lane; ++pos; }
            }
        }
        __syncwarp();
        const long long row = ((long long)b * N + q) * k;
        if (n <= 32 && n >= kk && tau < INFINITY) {
            // common case: one 32-wide bitonic sort of the survivors gives the answer
            float sd = lane < n ? hd[lane] : INFINITY;
            int sj = lane < n ? hi[lane] : FS_IDX_PAD;
            fs_warp_bitonic_sort(sd, sj, lane);
            const int skip = self_loop ? 0 : 1;
            if (lane >= skip && lane < kk) {
                idx[row + lane - skip] = sj;
                if (dist2) dist2[row + lane - skip] = sd;
            }
            __syncwarp();
            continue;
        }
        sel.init(qd_all + warp * 64, qi_all + warp * 64, kk);
        if (n <= CAP && tau < INFINITY) {
            for (int base = 0; base < n; base += 32) {
                const int s = base + lane;
                const bool valid = s < n;
                sel.offer(valid ? hd[s] : INFINITY, valid ? hi[s] : FS_IDX_PAD, valid);
            }
        } else {
            // overflow (heavy ties) or fewer than kk finite distances: stream every candidate
            // (distances recomputed from the staged cloud: a dynamic loop must not index the register array)
#pragma unroll 1
            for (int i = 0; i < NPL; ++i) {
                const int j = i * 32 + lane;
                const float4 p = sp[j];
                const float dot = fmaf(qz, p.z, fmaf(qy, p.y, __fmul_rn(qx, p.x)));
                float d = diag_zero ? __fadd_rn(__fsub_rn(qq, 2.0f * dot), p.w) : __fadd_rn(__fsub_rn(p.w, 2.0f * dot), qq);
                if (diag_zero && j == q) d = 0.f;
                sel.offer(d, j, j < N);
            }
        }
        sel.finish();
        sel.store(self_loop ? 0 : 1, idx + row, dist2 ? dist2 + row : nullptr, 0, 0, INFINITY);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// Offset-segmented kNN (pointops knnquery): one warp per query, candidates read through L1.
// ---------------------------------------------------------------------------------------------
template <int KPL>
__global__ void __launch_bounds__(KNN_THREADS)
knnquery_seg_kernel(int m, int nsample, const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                    const int32_t* __restrict__ offset, const int32_t* __restrict__ new_offset, int nseg,
                    int32_t* __restrict__ idx, float* __restrict__ dist2) {
    __shared__ float qd_all[KNN_WARPS * 64];
    __shared__ int qi_all[KNN_WARPS * 64];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q = blockIdx.x * KNN_WARPS + warp;
    if (q >= m) return;

    // segment of this query: first s with q < new_offset[s]
    int lo = 0, hi = nseg - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (q < __ldg(new_offset + mid)) hi = mid; else lo = mid + 1;
    }
    const int start = lo == 0 ? 0 : __ldg(offset + lo - 1);
    const int end = __ldg(offset + lo);

    const float qx = __ldg(new_xyz + 3ll * q), qy = __ldg(new_xyz + 3ll * q + 1), qz = __ldg(new_xyz + 3ll * q + 2);
    FsWarpSelect<KPL> sel;
    sel.init(qd_all + warp * 64, qi_all + warp * 64, nsample);
    for (int base = start; base < end; base += 32) {
        const int j = base + lane;
        const bool valid = j < end;
        float d = INFINITY;
        if (valid) {
            float dx = qx - __ldg(xyz + 3ll * j), dy = qy - __ldg(xyz + 3ll * j + 1), dz = qz - __ldg(xyz + 3ll * j + 2);
            d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        }
        sel.offer(d, j, valid);
    }
    sel.finish();
    const long long row = (long long)q * nsample;
    sel.store(0, idx + row, dist2 + row, 0, start, 1e10f);
}

// ---------------------------------------------------------------------------------------------
// Exact FP32 feature-space kNN on a point-major table x[P, ldx].
// CTA = 8 warps, 32 queries (4 per warp, evaluated together so a candidate value loaded from
// shared memory feeds 4 FMAs). Candidates are staged channel-major in chunks of TC points.
// ---------------------------------------------------------------------------------------------
__global__ void row_sqnorm_kernel(const float* __restrict__ x, int ldx, long long P, int C, float* __restrict__ out) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= P) return;
    const int lane = threadIdx.x & 31;
    // sequential-in-channel sum per lane stripe, then a fixed shuffle tree: deterministic
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
        float v = __ldg(x + row * ldx + c);
        acc = __fadd_rn(acc, __fmul_rn(v, v));
    }
    acc = fs_warp_sum(acc);
    if (lane == 0) out[row] = acc;
}

constexpr int KF_QPW = 4;                      // queries per warp
constexpr int KF_TILE_Q = KNN_WARPS * KF_QPW;  // 32 queries per CTA

template <int KPL>
__global__ void __launch_bounds__(KNN_THREADS)
knn_feat_kernel(const float* __restrict__ x, int ldx, int N, int C, int k, int self_loop, int diag_zero,
                int tc, const float* __restrict__ sqnorm, int32_t* __restrict__ idx, float* __restrict__ dist2,
                const uint8_t* __restrict__ redo) {
    extern __shared__ float smem[];
    if (redo) {
        // masked mode (after the tensor-core path): only rows without certificate are recomputed
        const int q = blockIdx.x * KF_TILE_Q + (threadIdx.x & 31);
        const bool mine = (threadIdx.x < 32) && q < N && redo[(long long)blockIdx.y * N + q];
        if (!__syncthreads_or(mine)) return;
    }
    const int tcp = tc + 1;                       // padded candidate stride (bank-conflict-free staging)
    float* sq = smem;                             // [C][KF_TILE_Q] query tile, channel-major
    float* sc = sq + C * KF_TILE_Q;               // [C][tcp] candidate chunk, channel-major
    float* sn = sc + C * tcp;                     // [tc] candidate norms
    float* qd_all = sn + tc;                      // [warps][QPW][64]
    int* qi_all = reinterpret_cast<int*>(qd_all + KNN_WARPS * KF_QPW * 64);

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const long long cloud0 = (long long)b * N;
    const int q0 = blockIdx.x * KF_TILE_Q;
    const int kk = k + (self_loop ? 0 : 1);

    for (int e = threadIdx.x; e < KF_TILE_Q * C; e += KNN_THREADS) {
        const int qi = e / C, c = e - qi * C;
        const int q = q0 + qi;
        sq[c * KF_TILE_Q + qi] = q < N ? __ldg(x + (cloud0 + q) * ldx + c) : 0.f;
    }

    FsWarpSelect<KPL> sel[KF_QPW];
    float qq[KF_QPW];
#pragma unroll
    for (int u = 0; u < KF_QPW; ++u) {
        const int q = q0 + warp * KF_QPW + u;
        qq[u] = q < N ? __ldg(sqnorm + cloud0 + q) : 0.f;
        sel[u].init(qd_all + (warp * KF_QPW + u) * 64, qi_all + (warp * KF_QPW + u) * 64, kk);
    }

    for (int c0 = 0; c0 < N; c0 += tc) {
        const int cn = min(tc, N - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < cn * C; e += KNN_THREADS) {
            const int j = e / C, c = e - j * C;
            sc[c * tcp + j] = __ldg(x + (cloud0 + c0 + j) * ldx + c);
        }
        for (int j = threadIdx.x; j < cn; j += KNN_THREADS) sn[j] = __ldg(sqnorm + cloud0 + c0 + j);
        __syncthreads();

        for (int base = 0; base < cn; base += 32) {
            const int j = base + lane;
            const bool valid = j < cn;
            const int jj = valid ? j : 0;
            float acc[KF_QPW];
#pragma unroll
            for (int u = 0; u < KF_QPW; ++u) acc[u] = 0.f;
            const float* qrow = sq + warp * KF_QPW;
#pragma unroll 4
            for (int c = 0; c < C; ++c) {
                const float cv = sc[c * tcp + jj];
                const float4 qv = *reinterpret_cast<const float4*>(qrow + c * KF_TILE_Q);
                acc[0] = fmaf(qv.x, cv, acc[0]);
                acc[1] = fmaf(qv.y, cv, acc[1]);
                acc[2] = fmaf(qv.z, cv, acc[2]);
                acc[3] = fmaf(qv.w, cv, acc[3]);
            }
            const float nj = sn[jj];
#pragma unroll
            for (int u = 0; u < KF_QPW; ++u) {
                const int q = q0 + warp * KF_QPW + u;
                float d = diag_zero ? __fadd_rn(__fsub_rn(qq[u], 2.0f * acc[u]), nj)
                                    : __fadd_rn(__fsub_rn(nj, 2.0f * acc[u]), qq[u]);
                if (diag_zero && (c0 + j) == q) d = 0.f;
                sel[u].offer(d, c0 + j, valid && q < N);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < KF_QPW; ++u) {
        const int q = q0 + warp * KF_QPW + u;
        if (q < N && (!redo || redo[cloud0 + q])) {
            sel[u].finish();
            const long long row = (cloud0 + q) * k;
            sel[u].store(self_loop ? 0 : 1, idx + row, dist2 ? dist2 + row : nullptr, 0, 0, INFINITY);
        }
    }
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

}  // namespace

extern "C" int fs_knn3d(int device, fs_stream_t stream_, const float* coords, long long batch_stride,
                        long long chan_stride, long long point_stride, int B, int N, int k, int self_loop,
                        int diag_zero, int32_t* idx, float* dist2) {
    if (B < 0 || N < 0 || k <= 0) return FS_ERR_BAD_ARG;
    if (B == 0) return FS_OK;  // empty batch: nothing to do, pointers may be null
    if (!coords || !idx) return FS_ERR_BAD_ARG;
    const int kk = k + (self_loop ? 0 : 1);
    if (kk > N || kk > FS_MAX_K + 1) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (N <= 2048 && kk <= 64) {
        // register-resident variant: every lane keeps its N/32 distances
        dim3 rgrid(fs_div_up(N, KNN3R_TILE_Q), B);
#define LAUNCH3R(NPL, CPL, KPL)                                                                                  \
    do {                                                                                                         \
        const size_t rs = (size_t)(NPL) * 32 * 4 * sizeof(float) + KNN3R_WARPS * 64 * 8 +                        \
                          (size_t)KNN3R_WARPS * (32 * (CPL) * 2) * 8;                                            \
        int e = set_smem(knn3d_regs_kernel<NPL, CPL, KPL>, rs);                                                  \
        if (e) return e;                                                                                         \
        knn3d_regs_kernel<NPL, CPL, KPL><<<rgrid, KNN3R_THREADS, rs, stream>>>(                                  \
            coords, batch_stride, chan_stride, point_stride, N, k, self_loop, diag_zero, idx, dist2);            \
    } while (0)
        if (kk <= 22) {          // 32 class minima: tau near rank 32 ln(32 / (32 - kk)) <= 37
            if (N <= 1024) LAUNCH3R(32, 1, 1); else LAUNCH3R(64, 1, 1);
        } else if (kk <= 32) {
            if (N <= 1024) LAUNCH3R(32, 2, 1); else LAUNCH3R(64, 2, 1);
        } else {
            if (N <= 1024) LAUNCH3R(32, 4, 2); else LAUNCH3R(64, 4, 2);
        }
#undef LAUNCH3R
        FS_RETURN_IF_LAUNCH_FAILED();
        return FS_OK;
    }
    const int chunk = N < KNN3D_MAX_CHUNK ? ((N + 31) / 32) * 32 : KNN3D_MAX_CHUNK;
    const size_t smem = (size_t)chunk * 4 * sizeof(float) + KNN_WARPS * 64 * 8;
    dim3 grid(fs_div_up(N, KNN3D_TILE_Q), B);
#define LAUNCH3D(KPL)                                                                              \
    do {                                                                                           \
        int e = set_smem(knn3d_batch_kernel<KPL>, smem);                                           \
        if (e) return e;                                                                           \
        knn3d_batch_kernel<KPL><<<grid, KNN_THREADS, smem, stream>>>(                              \
            coords, batch_stride, chan_stride, point_stride, N, k, self_loop, diag_zero, chunk,    \
            idx, dist2);                                                                           \
    } while (0)
    if (kk <= 32) LAUNCH3D(1); else if (kk <= 64) LAUNCH3D(2); else LAUNCH3D(4);
#undef LAUNCH3D
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_knnquery(int device, fs_stream_t stream_, int m, int nsample, const float* xyz,
                           const float* new_xyz, const int32_t* offset, const int32_t* new_offset, int b,
                           int32_t* idx, float* dist2) {
    if (!xyz || !new_xyz || !offset || !new_offset || !idx || !dist2 || m < 0 || b <= 0 || nsample <= 0)
        return FS_ERR_BAD_ARG;
    if (nsample > FS_MAX_K + 1) return FS_ERR_BAD_ARG;
    if (m == 0) return FS_OK;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int grid = fs_div_up(m, KNN_WARPS);
    if (nsample <= 32)
        knnquery_seg_kernel<1><<<grid, KNN_THREADS, 0, stream>>>(m, nsample, xyz, new_xyz, offset, new_offset, b, idx, dist2);
    else if (nsample <= 64)
        knnquery_seg_kernel<2><<<grid, KNN_THREADS, 0, stream>>>(m, nsample, xyz, new_xyz, offset, new_offset, b, idx, dist2);
    else
        knnquery_seg_kernel<4><<<grid, KNN_THREADS, 0, stream>>>(m, nsample, xyz, new_xyz, offset, new_offset, b, idx, dist2);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

// Shared launcher of the exact kernel; `redo` (nullable) restricts it to flagged rows.
static int launch_knn_feat(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                           int diag_zero, int32_t* idx, float* dist2, const float* sqnorm, const uint8_t* redo) {
    const int kk = k + (self_loop ? 0 : 1);
    // candidate chunk: keep the staged chunk near 64 KB
    int tc = (16384 / C) / 32 * 32;
    if (tc < 32) tc = 32;
    if (tc > 256) tc = 256;
    const size_t smem = ((size_t)C * KF_TILE_Q + (size_t)C * (tc + 1) + tc) * sizeof(float) +
                        (size_t)KNN_WARPS * KF_QPW * 64 * 8;
    if (smem > 220 * 1024) return FS_ERR_UNSUPPORTED;
    dim3 grid(fs_div_up(N, KF_TILE_Q), B);
#define LAUNCHF(KPL)                                                                               \
    do {                                                                                           \
        int e = set_smem(knn_feat_kernel<KPL>, smem);                                              \
        if (e) return e;                                                                           \
        knn_feat_kernel<KPL><<<grid, KNN_THREADS, smem, stream>>>(x, ldx, N, C, k, self_loop,      \
                                                                  diag_zero, tc, sqnorm, idx, dist2, redo); \
    } while (0)
    if (kk <= 32) LAUNCHF(1); else if (kk <= 64) LAUNCHF(2); else LAUNCHF(4);
#undef LAUNCHF
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

// Used by knn_tc.cu.
void fs_row_sqnorm(cudaStream_t stream, const float* x, int ldx, long long P, int C, float* out) {
    row_sqnorm_kernel<<<fs_div_up(P, 8), 256, 0, stream>>>(x, ldx, P, C, out);
}
int fs_knn_feat_masked(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                       int diag_zero, int32_t* idx, float* dist2, const float* sqnorm, const uint8_t* redo) {
    return launch_knn_feat(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, dist2, sqnorm, redo);
}
int fs_knn_feat_exact(cudaStream_t stream, const float* x, int ldx, int B, int N, int C, int k, int self_loop,
                      int diag_zero, int32_t* idx, float* dist2, float* sqnorm_ws) {
    row_sqnorm_kernel<<<fs_div_up((long long)B * N, 8), 256, 0, stream>>>(x, ldx, (long long)B * N, C, sqnorm_ws);
    FS_RETURN_IF_LAUNCH_FAILED();
    return launch_knn_feat(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, dist2, sqnorm_ws, nullptr);
}

extern "C" int fs_knn_feat(int device, fs_stream_t stream_, const float* x, int ldx, int B, int N, int C, int k,
                           int self_loop, int diag_zero, int32_t* idx, float* dist2, float* sqnorm_ws) {
    if (B < 0 || N < 0 || C <= 0 || k <= 0 || ldx < C) return FS_ERR_BAD_ARG;
    if (B == 0) return FS_OK;
    if (!x || !idx || !sqnorm_ws) return FS_ERR_BAD_ARG;
    const int kk = k + (self_loop ? 0 : 1);
    if (kk > N || kk > FS_MAX_K + 1) return FS_ERR_BAD_ARG;
    if (C > 1024) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long P = (long long)B * N;
    row_sqnorm_kernel<<<fs_div_up(P, 8), 256, 0, stream>>>(x, ldx, P, C, sqnorm_ws);
    FS_RETURN_IF_LAUNCH_FAILED();
    return launch_knn_feat(stream, x, ldx, B, N, C, k, self_loop, diag_zero, idx, dist2, sqnorm_ws, nullptr);
}
