// Backward of the pooled global-feature layer without the dense (B*N x C_out) gradient, and the feature-table glue.
//
// Layer (models/dgcnn.py:123-126, 156): y = X W^T (P x C, P = B*N rows, K input channels), z = BN(y), out[b, c] =
// max_r LeakyReLU(z). The gradient of y is   dy[r, c] = S[r, c] + a_c + b_c y[r, c]   where S has ONE non-zero per
// (cloud, channel) - the arg-max row - and a, b come from the two BatchNorm reductions (which only need the B x C
// selected values). With y = X W^T the dense part collapses onto K x K matrices:
//     dX = S W + 1 (a^T W) + X (W^T diag(b) W)                  one P x K x K GEMM instead of P x C x K
//     dW = S^T X + a colsum(X)^T + diag(b) W (X^T X)            one K x P x K Gram instead of C x P x K
// so the P x C gradient (134 MB at P = 65536, C = 1024) is neither written nor read and y is not kept for backward.
// The GEMMs stay library calls (host side, ops.py); the kernels here are the sparse parts and the coefficient vectors.
//
// Also in this file, the glue of the dense heads that used to be dozens of small ATen launches per step:
//   cat_cast / split_cast        [x1 | x2 | x3] in the compute dtype in one pass, and the gradient back
//   colsum_*                     deterministic column sums
//   final_linear_fwd / bwd       last head layer (-> num_classes) fused with the B x classes x N output in caller order
//   logits_out (+ adjoint)       the output transform alone
//   edge_weight_table (+ adj.)   [W1 ; W2 - W1] from the EdgeConv weight
//   sum_leading                  sum of the partial products of a row-chunked weight-gradient GEMM
//   multi_copy                   a gradient bucket into the flat buffer in one launch
#include "fs_common.cuh"

namespace {

// a_c, b_c and the routed, scaled gradient sp[b, c] = scale_c * g[b, c] * LeakyReLU'(z_sel[b, c]).
// coef = [mean | inv_std | scale = gamma * inv_std | beta] (fs_bn_finalize layout), dgb = [dbeta | dgamma] (double).
__global__ void pool_lin_prep_kernel(const float* __restrict__ g, const float* __restrict__ sel, const float* __restrict__ coef,
                                     float slope, const double* __restrict__ dgb, double count, int train_stats, int B, int C,
                                     float* __restrict__ a, float* __restrict__ bvec, float* __restrict__ sp) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)B * C) return;
    const int c = (int)(e % C);
    const float mu = __ldg(coef + c), sc = __ldg(coef + 2 * C + c);
    const float zs = fmaf(sc, __ldg(sel + e) - mu, __ldg(coef + 3 * C + c));
    const float gv = __ldg(g + e);
    sp[e] = sc * (zs > 0.f ? gv : slope * gv);
    if (e < C) {
        const float mb = train_stats ? (float)(dgb[c] / count) : 0.f;
        const float mg = train_stats ? (float)(dgb[C + c] / count) * __ldg(coef + C + c) : 0.f;
        a[c] = sc * (mg * mu - mb);
        bvec[c] = -sc * mg;
    }
}

template <typename T> __device__ __forceinline__ float ld_as_float(const T* p);
template <> __device__ __forceinline__ float ld_as_float<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_from_float(T* p, float v);
template <> __device__ __forceinline__ void st_from_float<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_from_float<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ void load_quad(const float* p, float* f) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load_quad(const __nv_bfloat16* p, float* f) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&v);
    const float2 a = __bfloat1622float2(hh[0]), b = __bfloat1622float2(hh[1]);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}
__device__ __forceinline__ void store_quad(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
__device__ __forceinline__ void store_quad(__nv_bfloat16* p, const float* f) {
    uint2 v;
    __nv_bfloat162* hh = reinterpret_cast<__nv_bfloat162*>(&v);
    hh[0] = __floats2bfloat162_rn(f[0], f[1]);
    hh[1] = __floats2bfloat162_rn(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = v;
}

// 16-byte chunk of a row: 8 bf16 or 4 fp32 values
template <typename T> struct Chunk;
template <> struct Chunk<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ __forceinline__ static void decode(const uint4& r, float* f) { fs_bf16x8_to_float(r, f); }
    __device__ __forceinline__ static void red_add(__nv_bfloat16* p, const float* f) {
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(reinterpret_cast<__nv_bfloat162*>(p) + i, __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]));
    }
    __device__ __forceinline__ static void load(const __nv_bfloat16* p, float* f) { fs_bf16x8_to_float(*reinterpret_cast<const uint4*>(p), f); }
    __device__ __forceinline__ static void store(__nv_bfloat16* p, const float* f) {
        uint4 v;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        *reinterpret_cast<uint4*>(p) = v;
    }
};
template <> struct Chunk<float> {
    static constexpr int N = 4;
    __device__ __forceinline__ static void decode(const uint4& r, float* f) {
        f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y); f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
    }
    // fire-and-forget add (RED): the caller guarantees ONE add per element, so the result does not depend on any order
    __device__ __forceinline__ static void red_add(float* p, const float* f) {
#pragma unroll
        for (int i = 0; i < 4; ++i) atomicAdd(p + i, f[i]);
    }
    __device__ __forceinline__ static void load(const float* p, float* f) {
        const float4 v = *reinterpret_cast<const float4*>(p);
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
    }
    __device__ __forceinline__ static void store(float* p, const float* f) { *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]); }
};

__device__ __forceinline__ void red_add2(float* p, float a, float b) { atomicAdd(p, a); atomicAdd(p + 1, b); }
__device__ __forceinline__ void red_add2(__nv_bfloat16* p, float a, float b) {
    atomicAdd(reinterpret_cast<__nv_bfloat162*>(p), __floats2bfloat162_rn(a, b));
}

// dX[row] += sum over the channels c' whose arg-max is `row` of sp[b, c'] W[c', :], in three small kernels.
// The arg-max rows of a cloud are very unevenly used (a few "critical" points win hundreds of channels), so the work is
// balanced by ENTRY, not by row: (1) one block per cloud sorts its C (row, channel) pairs (bitonic, shared memory);
// (2) every warp of the grid walks 32 consecutive sorted entries, lanes across the row in 16-byte chunks (KV chunks per
// lane, eight entries' W chunks in flight). A run of equal rows inside a warp is added to dX by that warp alone; the runs
// touching a warp's first / last entry go to a partial buffer; (3) one block per cloud merges those partials in warp
// order. Every element of dX receives exactly ONE add (issued as a fire-and-forget RED so that no warp waits for a row)
// of a sum formed in a fixed order: deterministic. (One kernel doing all three on 32 SMs took 41 us; instruction-bound.)
__global__ void __launch_bounds__(1024)
pool_lin_sort_kernel(const int32_t* __restrict__ arg, int N, int C, int log2C, unsigned* __restrict__ keys) {
    extern __shared__ unsigned key_s[];                                        // [C]
    const int b = blockIdx.x, t = threadIdx.x;
    const int row = __ldg(arg + (long long)b * C + t);
    key_s[t] = (row >= 0 && row < N) ? ((unsigned)row << log2C) | (unsigned)t : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= C; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int p = t ^ j;
            if (p > t) {
                const unsigned x0 = key_s[t], x1 = key_s[p];
                const bool up = (t & k) == 0;
                if ((x0 > x1) == up) { key_s[t] = x1; key_s[p] = x0; }
            }
            __syncthreads();
        }
    }
    keys[(long long)b * C + t] = key_s[t];
}

template <typename T, int KV>
__global__ void __launch_bounds__(128)
pool_lin_rows_kernel(const unsigned* __restrict__ keys, const float* __restrict__ sp, const T* __restrict__ w, int ldw, int N,
                     int C, int log2C, int K, T* __restrict__ dx, int ld_dx, float* __restrict__ part, int* __restrict__ rows_ws) {
    constexpr int V = Chunk<T>::N;
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int wi = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);        // warp of the cloud: entries [32 wi, 32 wi + 32)
    const int nw = C >> 5;
    if (wi >= nw) return;
    const int chunks = K / V;                              // 16-byte chunks per row; lane owns chunks lane, lane + 32, ...
    const unsigned* kb = keys + (long long)b * C + wi * 32;
    const float* spb = sp + (long long)b * C;
    T* dxb = dx + (long long)b * N * ld_dx;
    float* pw = part + ((long long)b * nw + wi) * 2 * K;
    float acc[KV][V];
#pragma unroll
    for (int q = 0; q < KV; ++q)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[q][i] = 0.f;
    int cur = -1;
    bool first = true;
    int head_row = -1, tail_row = -1;
    auto emit = [&](bool last) {
        // the run `cur` is complete (or the warp's entries are): first run -> head slot, last -> tail slot, else global
        if (cur < 0) return;
#pragma unroll
        for (int q = 0; q < KV; ++q) {
            const int ch = lane + 32 * q;
            if (ch < chunks) {
                if (first || last) {
                    float* ps = pw + (first ? 0 : K) + ch * V;
#pragma unroll
                    for (int i = 0; i < V; ++i) ps[i] = acc[q][i];
                } else {
                    Chunk<T>::red_add(dxb + (long long)cur * ld_dx + ch * V, acc[q]);
                }
            }
#pragma unroll
            for (int i = 0; i < V; ++i) acc[q][i] = 0.f;
        }
        if (first) head_row = cur; else if (last) tail_row = cur;
        first = false;
    };
    const unsigned cmask = (unsigned)C - 1u;
    for (int i0 = 0; i0 < 32; i0 += 8) {
        // eight entries per round: their W chunks are all requested before the first one is used
        unsigned ks[8];
        float scs[8];
        uint4 raw[8][KV];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            ks[u] = __ldg(kb + i0 + u);                                       // same address for the whole warp
            const int c2 = ks[u] == 0xffffffffu ? 0 : (int)(ks[u] & cmask);
            scs[u] = __ldg(spb + c2);
            const T* wr = w + (long long)c2 * ldw;
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const int ch = lane + 32 * q;
                raw[u][q] = ch < chunks ? __ldg(reinterpret_cast<const uint4*>(wr + ch * V)) : make_uint4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (ks[u] != 0xffffffffu) {                                      // invalid entries sort last (warp-uniform)
                const int row = (int)(ks[u] >> log2C);
                if (row != cur) { emit(false); cur = row; }
#pragma unroll
                for (int q = 0; q < KV; ++q) {
                    float f[V];
                    Chunk<T>::decode(raw[u][q], f);
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[q][j] = fmaf(scs[u], f[j], acc[q][j]);
                }
            }
        }
    }
    emit(true);
    if (lane == 0) {
        rows_ws[((long long)b * nw + wi) * 2] = head_row;
        rows_ws[((long long)b * nw + wi) * 2 + 1] = tail_row;
    }
}

// merge the boundary runs of a cloud in warp order: thread = column pair (one add per element)
template <typename T>
__global__ void pool_lin_merge_kernel(const float* __restrict__ part, const int* __restrict__ rows_ws, int N, int C, int K,
                                      T* __restrict__ dx, int ld_dx) {
    extern __shared__ int mrows_s[];                                           // [2 nw]
    const int b = blockIdx.x;
    const int nw = C >> 5;
    for (int i = threadIdx.x; i < 2 * nw; i += blockDim.x) mrows_s[i] = rows_ws[(long long)b * nw * 2 + i];
    __syncthreads();
    const float* pb = part + (long long)b * nw * 2 * K;
    T* dxb = dx + (long long)b * N * ld_dx;
    for (int col = 2 * threadIdx.x; col < K; col += 2 * blockDim.x) {
        int r0 = -1;
        float v0 = 0.f, v1 = 0.f;
        for (int q0 = 0; q0 < 2 * nw; q0 += 8) {       // 2 nw is a multiple of 8 (C >= 128) or smaller than 8 (C = 64)
            float2 pv[8];                                // eight partials in flight (unused slots hold stale but valid memory)
#pragma unroll
            for (int u = 0; u < 8; ++u)
                pv[u] = q0 + u < 2 * nw ? *reinterpret_cast<const float2*>(pb + (long long)(q0 + u) * K + col) : make_float2(0.f, 0.f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int r = q0 + u < 2 * nw ? mrows_s[q0 + u] : -1;
                if (r < 0) continue;
                if (r != r0) {
                    if (r0 >= 0) red_add2(dxb + (long long)r0 * ld_dx + col, v0, v1);
                    r0 = r;
                    v0 = v1 = 0.f;
                }
                v0 += pv[u].x;
                v1 += pv[u].y;
            }
        }
        if (r0 >= 0) red_add2(dxb + (long long)r0 * ld_dx + col, v0, v1);
    }
}

// dW[c, :] = sum_b sp[b, c] X[b N + arg[b, c], :] + a_c colsum(X) + b_c (W G)[c, :]. One warp per output channel, lanes
// across the row in 16-byte chunks; (row, scale) of 32 clouds are fetched by the lanes at once and broadcast.
template <typename T, int KV>
__global__ void __launch_bounds__(256)
pool_lin_dw_kernel(const float* __restrict__ sp, const int32_t* __restrict__ arg, const T* __restrict__ x, int ldx, int B,
                   int N, int C, int K, const float* __restrict__ a, const float* __restrict__ bvec,
                   const float* __restrict__ colsum, const float* __restrict__ wg, float* __restrict__ dw) {
    constexpr int V = Chunk<T>::N;
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= C) return;
    const int chunks = K / V;
    float acc[KV][V];
#pragma unroll
    for (int q = 0; q < KV; ++q)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[q][i] = 0.f;
    for (int b0 = 0; b0 < B; b0 += 32) {
        const int mb = b0 + lane;
        int row_l = mb < B ? __ldg(arg + (long long)mb * C + c) : -1;
        if (row_l >= N) row_l = -1;
        const float s_l = row_l >= 0 ? __ldg(sp + (long long)mb * C + c) : 0.f;
        const int nb = B - b0 < 32 ? B - b0 : 32;
        for (int u0 = 0; u0 < nb; u0 += 8) {
            float scs[8];
            uint4 raw[8][KV];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int row = __shfl_sync(FS_FULL_MASK, row_l, (u0 + u) & 31);
                const float sc = __shfl_sync(FS_FULL_MASK, s_l, (u0 + u) & 31);
                const bool ok = u0 + u < nb && row >= 0;
                scs[u] = ok ? sc : 0.f;
                const T* xr = x + ((long long)(ok ? b0 + u0 + u : 0) * N + (ok ? row : 0)) * ldx;
#pragma unroll
                for (int q = 0; q < KV; ++q) {
                    const int ch = lane + 32 * q;
                    raw[u][q] = ch < chunks ? __ldg(reinterpret_cast<const uint4*>(xr + ch * V)) : make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int q = 0; q < KV; ++q) {
                    float f[V];
                    Chunk<T>::decode(raw[u][q], f);
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[q][j] = fmaf(scs[u], f[j], acc[q][j]);
                }
            }
        }
    }
    const float ac = a ? __ldg(a + c) : 0.f, bc = bvec ? __ldg(bvec + c) : 0.f;
#pragma unroll
    for (int q = 0; q < KV; ++q) {
        const int ch = lane + 32 * q;
        if (ch < chunks) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const int col = ch * V + j;
                float v = acc[q][j];
                if (colsum) v = fmaf(ac, __ldg(colsum + col), v);
                if (wg) v = fmaf(bc, __ldg(wg + (long long)c * K + col), v);
                dw[(long long)c * K + col] = v;
            }
        }
    }
}

// Column sums of a (rows, K) table: per-block partials (fixed chunking) and a second pass over the partials, so the
// result is deterministic. blockDim 256; V values per thread, K / V threads per row.
template <typename T, int V>
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const T* __restrict__ x, int ld, long long rows, int K, float* __restrict__ partial) {
    extern __shared__ float cs_s[];            // [rows_per_pass][K]
    const int tpr = K / V;
    const int rpp = 256 / tpr;
    const int r = threadIdx.x / tpr, v0 = (threadIdx.x - r * tpr) * V;
    const long long chunk = (rows + gridDim.x - 1) / gridDim.x;
    const long long r_begin = (long long)blockIdx.x * chunk;
    const long long r_end = r_begin + chunk < rows ? r_begin + chunk : rows;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    if (r < rpp) {
        for (long long rr = r_begin + r; rr < r_end; rr += 8 * rpp) {
            float f[8][V];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const long long q = rr + (long long)u * rpp;
                const T* p = x + (q < r_end ? q : rr) * ld + v0;
                if (V == 8) {
                    const uint4 pk = __ldg(reinterpret_cast<const uint4*>(p));
                    fs_bf16x8_to_float(pk, f[u]);
                } else {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                    f[u][0] = t.x; f[u][1] = t.y; f[u][2] = t.z; f[u][3] = t.w;
                }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (rr + (long long)u * rpp < r_end) {
#pragma unroll
                    for (int i = 0; i < V; ++i) acc[i] += f[u][i];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) cs_s[r * K + v0 + i] = acc[i];
    }
    __syncthreads();
    for (int col = threadIdx.x; col < K; col += blockDim.x) {
        float t = 0.f;
        for (int q = 0; q < rpp; ++q) t += cs_s[q * K + col];
        partial[(long long)blockIdx.x * K + col] = t;
    }
}
// block = 32 columns x 32 partial groups; fixed order within a group, fixed order over the groups
__global__ void __launch_bounds__(1024)
colsum_final_kernel(const float* __restrict__ partial, int G, int K, float* __restrict__ out) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + lane;
    float t = 0.f;
    if (col < K) {
#pragma unroll 4
        for (int g = grp; g < G; g += 32) t += __ldg(partial + (long long)g * K + col);
    }
    red[grp][lane] = t;
    __syncthreads();
    if (grp == 0 && col < K) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) v += red[q][lane];
        out[col] = v;
    }
}

// ---- network output: per-point logits of the (Morton-)sorted cloud, point-major (B*N, C) in the compute dtype ->
// (B, C, N) fp32 in the CALLER's point order: out[b, c, perm[b, n]] = logits[b*N + n, c] (perm nullable = identity),
// and the reverse for the gradient. Replaces permute + cast + scatter (and gather + permute + cast on the way back).
template <typename T>
__global__ void logits_out_kernel(const T* __restrict__ logits, int ld, const long long* __restrict__ perm, int B, int N, int C,
                                  float* __restrict__ out) {
    const long long total = (long long)B * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / N;
        const long long dst = perm ? perm[e] : e - b * N;
        const T* row = logits + e * ld;
        float* ob = out + b * C * N + dst;
        for (int c = 0; c < C; ++c) ob[(long long)c * N] = ld_as_float(row + c);
    }
}
template <typename T>
__global__ void logits_out_bwd_kernel(const float* __restrict__ g, const long long* __restrict__ perm, int B, int N, int C,
                                      T* __restrict__ dlogits, int ld) {
    const long long total = (long long)B * N;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / N;
        const long long src = perm ? perm[e] : e - b * N;
        const float* gb = g + b * C * N + src;
        T* row = dlogits + e * ld;
        for (int c = 0; c < C; ++c) st_from_float(row + c, gb[(long long)c * N]);
    }
}

// ---- last layer of the segmentation head fused with the network output (models/dgcnn.py:137, 160-162): a 1x1 conv to
// `num_classes` (<= 8) channels with bias, no BatchNorm, written straight as (B, classes, N) fp32 in the caller's point
// order. As library calls this is three GEMMs with a 4-wide dimension (16 + 12 + 16 us at P = 65536, C_in = 128) plus
// the un-sort / transpose / cast passes; here one pass over H each way.
// Forward: a block stages 128 rows of H in shared memory with coalesced 16-byte loads (rows padded by one chunk, so the
// per-point reads below are bank-conflict-free), then one THREAD per point forms its NOUT sums - no cross-lane reduction;
// the weights are shared-memory broadcasts.
template <typename T, int NOUT, int CIN>
__global__ void __launch_bounds__(128)
final_linear_fwd_kernel(const T* __restrict__ h, int ld, const float* __restrict__ w, const float* __restrict__ bias,
                        const long long* __restrict__ perm, int B, int N, float* __restrict__ out) {
    constexpr int V = Chunk<T>::N;              // values per 16-byte chunk
    constexpr int RS = CIN + V;                 // padded row stride (elements)
    extern __shared__ __align__(16) unsigned char ff_smem[];
    T* h_s = reinterpret_cast<T*>(ff_smem);                                         // [128][RS]
    float* w_s = reinterpret_cast<float*>(ff_smem + (size_t)128 * RS * sizeof(T));   // [NOUT][CIN]
    const long long total = (long long)B * N;
    const long long e0 = (long long)blockIdx.x * 128;
    const int rows = (int)(total - e0 < 128 ? total - e0 : 128);
    for (int i = threadIdx.x; i < NOUT * CIN; i += blockDim.x) w_s[i] = __ldg(w + i);
    {   // all of a thread's chunks are requested before the first one is stored (no load -> store -> load chain)
        constexpr int PER = 128 * (CIN / V) / 128;
        uint4 v[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = threadIdx.x + u * 128;
            const int p = i / (CIN / V), q = i - p * (CIN / V);
            v[u] = p < rows ? __ldg(reinterpret_cast<const uint4*>(h + (e0 + p) * ld + q * V)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = threadIdx.x + u * 128;
            const int p = i / (CIN / V), q = i - p * (CIN / V);
            *reinterpret_cast<uint4*>(h_s + p * RS + q * V) = v[u];
        }
    }
    __syncthreads();
    const int p = threadIdx.x;
    if (p >= rows) return;
    float acc[NOUT];
#pragma unroll
    for (int c = 0; c < NOUT; ++c) acc[c] = bias ? __ldg(bias + c) : 0.f;
#pragma unroll 4
    for (int q = 0; q < CIN / V; ++q) {
        float hv[V];
        Chunk<T>::load(h_s + p * RS + q * V, hv);
#pragma unroll
        for (int c = 0; c < NOUT; ++c)
#pragma unroll
            for (int i = 0; i < V; ++i) acc[c] = fmaf(w_s[c * CIN + q * V + i], hv[i], acc[c]);
    }
    const long long e = e0 + p;
    const long long b = e / N;
    const long long dst = perm ? perm[e] : e - b * N;
#pragma unroll
    for (int c = 0; c < NOUT; ++c) out[(b * NOUT + c) * N + dst] = acc[c];
}

// Backward: a block stages TILE rows of H and their (gathered) output gradients in shared memory; dH = G W is written per
// point, dW = G^T H and dbias = sum G leave as per-block partials (summed by final_linear_red_kernel in block order:
// deterministic).
constexpr int FL_TILE = 128;
template <typename T, int NOUT, int CIN>
__global__ void __launch_bounds__(256)
final_linear_bwd_kernel(const T* __restrict__ h, int ld, const float* __restrict__ w, const float* __restrict__ g,
                        const long long* __restrict__ perm, int B, int N, T* __restrict__ dh, int ld_dh,
                        float* __restrict__ part /* [grid][NOUT * CIN + NOUT] */) {
    constexpr int V = Chunk<T>::N;
    extern __shared__ __align__(16) unsigned char fl_smem[];
    T* h_s = reinterpret_cast<T*>(fl_smem);                                         // [TILE][CIN]
    float* w_s = reinterpret_cast<float*>(fl_smem + (size_t)FL_TILE * CIN * sizeof(T));   // [NOUT][CIN]
    float* g_s = w_s + NOUT * CIN;                                                  // [TILE][NOUT]
    const long long total = (long long)B * N;
    const long long e0 = (long long)blockIdx.x * FL_TILE;
    const int rows = (int)(total - e0 < FL_TILE ? total - e0 : FL_TILE);
    for (int i = threadIdx.x; i < NOUT * CIN; i += blockDim.x) w_s[i] = __ldg(w + i);
    {   // all of a thread's chunks are requested before the first one is stored (no load -> store -> load chain)
        constexpr int PER = FL_TILE * (CIN / V) / 256;
        uint4 v[PER];
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = threadIdx.x + u * 256;
            const int p = i / (CIN / V), q = i - p * (CIN / V);
            v[u] = p < rows ? __ldg(reinterpret_cast<const uint4*>(h + (e0 + p) * ld + q * V)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = threadIdx.x + u * 256;
            const int p = i / (CIN / V), q = i - p * (CIN / V);
            *reinterpret_cast<uint4*>(h_s + p * CIN + q * V) = v[u];
        }
    }
    for (int i = threadIdx.x; i < FL_TILE * NOUT; i += blockDim.x) {
        const int p = i / NOUT, c = i - p * NOUT;
        float v = 0.f;
        if (p < rows) {
            const long long e = e0 + p, b = e / N;
            const long long src = perm ? perm[e] : e - b * N;
            v = __ldg(g + (b * NOUT + c) * N + src);
        }
        g_s[i] = v;
    }
    __syncthreads();
    // dH: one 16-byte chunk per thread and step, consecutive threads on consecutive chunks of a row (coalesced stores)
    for (int i = threadIdx.x; i < rows * (CIN / V); i += blockDim.x) {
        const int p = i / (CIN / V), q = i - p * (CIN / V);
        float dv[V];
#pragma unroll
        for (int j = 0; j < V; ++j) dv[j] = 0.f;
#pragma unroll
        for (int c = 0; c < NOUT; ++c) {
            const float gc = g_s[p * NOUT + c];
#pragma unroll
            for (int j = 0; j < V; ++j) dv[j] = fmaf(gc, w_s[c * CIN + q * V + j], dv[j]);
        }
        Chunk<T>::store(dh + (e0 + p) * ld_dh + q * V, dv);
    }
    // dW, dbias partials of this tile: a thread owns two adjacent input channels for ALL classes over a quarter (CIN 128)
    // or half (CIN 256) of the tile's points - one 4- or 8-byte shared-memory read of H per point feeds 2 * NOUT FMAs -
    // then the point groups are summed in a fixed order through shared memory (the H tile is no longer needed).
    constexpr int PAIRS = CIN / 2, PG = 256 / PAIRS, PPG = FL_TILE / PG;
    const int ip = threadIdx.x % PAIRS, pg = threadIdx.x / PAIRS;
    float acc[NOUT][2];
#pragma unroll
    for (int c = 0; c < NOUT; ++c) { acc[c][0] = 0.f; acc[c][1] = 0.f; }
#pragma unroll 4
    for (int p = pg * PPG; p < (pg + 1) * PPG; ++p) {
        float h0, h1;
        if (sizeof(T) == 2) {
            const float2 hv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(h_s + p * CIN + 2 * ip));
            h0 = hv.x; h1 = hv.y;
        } else {
            const float2 hv = *reinterpret_cast<const float2*>(h_s + p * CIN + 2 * ip);
            h0 = hv.x; h1 = hv.y;
        }
#pragma unroll
        for (int c = 0; c < NOUT; ++c) {
            const float gc = g_s[p * NOUT + c];
            acc[c][0] = fmaf(gc, h0, acc[c][0]);
            acc[c][1] = fmaf(gc, h1, acc[c][1]);
        }
    }
    __syncthreads();                                   // every read of the H tile is done: reuse it as scratch
    float* red = reinterpret_cast<float*>(fl_smem);    // [PG][NOUT][CIN]  (PG * NOUT * CIN * 4 <= TILE * CIN * sizeof(T))
#pragma unroll
    for (int c = 0; c < NOUT; ++c) {
        red[(pg * NOUT + c) * CIN + 2 * ip] = acc[c][0];
        red[(pg * NOUT + c) * CIN + 2 * ip + 1] = acc[c][1];
    }
    __syncthreads();
    float* pb = part + (long long)blockIdx.x * (NOUT * CIN + NOUT);
    for (int o = threadIdx.x; o < NOUT * CIN; o += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < PG; ++q) v += red[q * NOUT * CIN + o];
        pb[o] = v;
    }
    if (threadIdx.x < NOUT) {
        float a = 0.f;
        for (int p = 0; p < FL_TILE; ++p) a += g_s[p * NOUT + threadIdx.x];
        pb[NOUT * CIN + threadIdx.x] = a;
    }
}
// out[t] = sum over the G partial rows, block = 32 outputs x 32 partial groups (fixed order)
__global__ void __launch_bounds__(1024)
final_linear_red_kernel(const float* __restrict__ part, int G, int n, float* __restrict__ out) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int t = blockIdx.x * 32 + lane;
    float v = 0.f;
    if (t < n) {
#pragma unroll 4
        for (int q = grp; q < G; q += 32) v += __ldg(part + (long long)q * n + t);
    }
    red[grp][lane] = v;
    __syncthreads();
    if (grp == 0 && t < n) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) a += red[q][lane];
        out[t] = a;
    }
}

// ---- EdgeConv weight table: [W1 ; W2 - W1] (2Cp x C) from the conv weight W = [W1 | W2] (Cp x 2C), and the adjoint
__global__ void edge_weight_table_kernel(const float* __restrict__ w, int Cp, int C, float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * Cp * C) return;
    const int r = e / C, c = e - r * C;
    out[e] = r < Cp ? w[r * 2 * C + c] : w[(r - Cp) * 2 * C + C + c] - w[(r - Cp) * 2 * C + c];
}
__global__ void edge_weight_table_bwd_kernel(const float* __restrict__ g, int Cp, int C, float* __restrict__ dw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= 2 * Cp * C) return;
    const int r = e / (2 * C), c = e - r * 2 * C;
    dw[e] = c < C ? g[r * C + c] - g[(Cp + r) * C + c] : g[(Cp + r) * C + (c - C)];
}

// ---- out[t] = sum_s part[s, t]: the partial results of a chunked (batched) weight-gradient GEMM, fixed order over s.
// Threads across t (coalesced), S / SG partials per thread in SG interleaved groups, combined through shared memory.
template <int SG>
__global__ void __launch_bounds__(256)
sum_leading_kernel(const float* __restrict__ part, int S, long long n, float* __restrict__ out) {
    __shared__ float red[SG][256 / SG];
    constexpr int TPB = 256 / SG;                     // outputs per block
    const int lane_t = threadIdx.x % TPB, grp = threadIdx.x / TPB;
    const long long t = (long long)blockIdx.x * TPB + lane_t;
    float v = 0.f;
    if (t < n) {
#pragma unroll 4
        for (int q = grp; q < S; q += SG) v += __ldg(part + (long long)q * n + t);
    }
    red[grp][lane_t] = v;
    __syncthreads();
    if (grp == 0 && t < n) {
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < SG; ++q) a += red[q][lane_t];
        out[t] = a;
    }
}

// ---- gradient bucket: up to 32 fp32 tensors copied into their slices of a flat buffer by ONE launch. The table of
// (source, destination, element count) travels as a kernel parameter, so nothing is staged through device memory.
struct MultiCopy {
    const float* src[32];
    float* dst[32];
    unsigned first_chunk[33];       // prefix sums of the 1024-element chunk counts
    unsigned count[32];
    int n;
};
__global__ void __launch_bounds__(256)
multi_copy_kernel(const __grid_constant__ MultiCopy t) {
    const unsigned chunk = blockIdx.x;
    int i = 0;
#pragma unroll 1
    while (i + 1 < t.n && chunk >= t.first_chunk[i + 1]) ++i;
    const unsigned base = (chunk - t.first_chunk[i]) * 1024u;
    const float* __restrict__ s = t.src[i];
    float* __restrict__ d = t.dst[i];
    const unsigned n = t.count[i];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const unsigned e = base + q * 256u + threadIdx.x;
        if (e < n) d[e] = __ldg(s + e);
    }
}

// ---- feature table: [x1 | x2 | ...] (fp32, point-major) -> one (P, sum C_i) table in the compute dtype, and back
struct CatSrc {
    const float* p[4];
    float* q[4];
    int ld[4];
    int c0[5];        // prefix sums of the widths; c0[n] = total
    int n;
};

// Both kernels keep two independent groups of loads in flight per thread (U = 2): with one 32-byte load per thread the
// 75 MB pass ran at 3.5 TB/s.
template <typename OT, int V>
__global__ void __launch_bounds__(256)
cat_cast_kernel(CatSrc s, unsigned rows, OT* __restrict__ out, int ld_out) {
    constexpr int U = 2;
    const unsigned groups = (unsigned)s.c0[s.n] / V;
    const unsigned total = rows * groups;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += U * stride) {
        float f[U][V];
        unsigned r[U];
        int col[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned e = e0 + u * stride < total ? e0 + u * stride : e0;
            r[u] = e / groups;
            col[u] = (int)(e - r[u] * groups) * V;
            int i = 0;
#pragma unroll
            for (int t = 1; t < 4; ++t) i += (t < s.n && col[u] >= s.c0[t]) ? 1 : 0;
            const float* sp = s.p[i] + (long long)r[u] * s.ld[i] + (col[u] - s.c0[i]);
#pragma unroll
            for (int q = 0; q < V / 4; ++q) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(sp) + q);
                f[u][4 * q] = v.x; f[u][4 * q + 1] = v.y; f[u][4 * q + 2] = v.z; f[u][4 * q + 3] = v.w;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u > 0 && e0 + u * stride >= total) break;
            OT* o = out + (long long)r[u] * ld_out + col[u];
            if (sizeof(OT) == 2) {
                __nv_bfloat162 h[V / 2];
#pragma unroll
                for (int q = 0; q < V / 2; ++q) h[q] = __floats2bfloat162_rn(f[u][2 * q], f[u][2 * q + 1]);
                if (V == 8) *reinterpret_cast<uint4*>(o) = *reinterpret_cast<uint4*>(h);
                else *reinterpret_cast<uint2*>(o) = *reinterpret_cast<uint2*>(h);
            } else {
#pragma unroll
                for (int q = 0; q < V / 4; ++q)
                    reinterpret_cast<float4*>(o)[q] = make_float4(f[u][4 * q], f[u][4 * q + 1], f[u][4 * q + 2], f[u][4 * q + 3]);
            }
        }
    }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
split_cast_kernel(CatSrc s, unsigned rows, const T* __restrict__ g, int ld_g) {
    constexpr int U = 2;
    const unsigned groups = (unsigned)s.c0[s.n] / V;
    const unsigned total = rows * groups;
    const unsigned stride = gridDim.x * blockDim.x;
    for (unsigned e0 = blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += U * stride) {
        float f[U][V];
        unsigned r[U];
        int col[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned e = e0 + u * stride < total ? e0 + u * stride : e0;
            r[u] = e / groups;
            col[u] = (int)(e - r[u] * groups) * V;
            const T* gp = g + (long long)r[u] * ld_g + col[u];
            if (sizeof(T) == 2) {
                if (V == 8) {
                    const uint4 pk = __ldg(reinterpret_cast<const uint4*>(gp));
                    fs_bf16x8_to_float(pk, f[u]);
                } else {
                    const uint2 pk = __ldg(reinterpret_cast<const uint2*>(gp));
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
                    const float2 lo = __bfloat1622float2(h[0]), hi = __bfloat1622float2(h[1]);
                    f[u][0] = lo.x; f[u][1] = lo.y; f[u][2] = hi.x; f[u][3] = hi.y;
                }
            } else {
#pragma unroll
                for (int q = 0; q < V / 4; ++q) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(gp) + q);
                    f[u][4 * q] = v.x; f[u][4 * q + 1] = v.y; f[u][4 * q + 2] = v.z; f[u][4 * q + 3] = v.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (u > 0 && e0 + u * stride >= total) break;
            int i = 0;
#pragma unroll
            for (int t = 1; t < 4; ++t) i += (t < s.n && col[u] >= s.c0[t]) ? 1 : 0;
            float* o = s.q[i] + (long long)r[u] * s.ld[i] + (col[u] - s.c0[i]);
#pragma unroll
            for (int q = 0; q < V / 4; ++q)
                reinterpret_cast<float4*>(o)[q] = make_float4(f[u][4 * q], f[u][4 * q + 1], f[u][4 * q + 2], f[u][4 * q + 3]);
        }
    }
}

bool fill_cat(CatSrc& s, int n, const void* const* ptrs, const int* widths, const int* lds, bool writable) {
    if (n < 1 || n > 4 || !ptrs || !widths || !lds) return false;
    s.n = n;
    s.c0[0] = 0;
    for (int i = 0; i < 4; ++i) { s.p[i] = nullptr; s.q[i] = nullptr; s.ld[i] = 0; }
    for (int i = 0; i < n; ++i) {
        if (!ptrs[i] || widths[i] <= 0 || widths[i] % 4 || lds[i] < widths[i] || lds[i] % 4) return false;
        if (writable) s.q[i] = (float*)ptrs[i]; else s.p[i] = (const float*)ptrs[i];
        s.ld[i] = lds[i];
        s.c0[i + 1] = s.c0[i] + widths[i];
    }
    for (int i = n + 1; i < 5; ++i) s.c0[i] = s.c0[n];
    return true;
}

}  // namespace

extern "C" int fs_colsum_partials(void);

extern "C" int fs_pool_lin_bwd_prep(int device, fs_stream_t stream_, const float* g, const float* sel, const float* coef,
                                    float slope, const double* dgb, double count, int train_stats, int B, int C, float* a,
                                    float* bvec, float* sp) {
    if (!g || !sel || !coef || !a || !bvec || !sp || B <= 0 || C <= 0) return FS_ERR_BAD_ARG;
    if (train_stats && (!dgb || count <= 0)) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long total = (long long)B * C;
    pool_lin_prep_kernel<<<fs_div_up(total, 256), 256, 0, stream>>>(g, sel, coef, slope, dgb, count, train_stats, B, C, a, bvec, sp);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

namespace {
// KV = 16-byte chunks per lane: K * sizeof(T) / 16 chunks over 32 lanes. ws: keys [B, C] u32 | rows [B, C/32, 2] i32 |
// partials [B, C/32, 2, K] f32
size_t sparse_dx_ws_bytes(int B, int C, int K) {
    return (size_t)B * C * 4 + (size_t)B * (C / 32) * 2 * 4 + (size_t)B * (C / 32) * 2 * K * 4;
}
template <typename T>
int launch_sparse_dx(cudaStream_t stream, const float* sp, const int32_t* arg, const T* w, int ldw, int B, int N, int C, int K,
                     T* dx, int ld_dx, void* ws) {
    int log2C = 0;
    while ((1 << log2C) < C) ++log2C;
    const int nw = C / 32;
    unsigned* keys = (unsigned*)ws;
    int* rows_ws = (int*)(keys + (size_t)B * C);
    float* part = (float*)(rows_ws + (size_t)B * nw * 2);
    pool_lin_sort_kernel<<<B, C, (size_t)C * 4, stream>>>(arg, N, C, log2C, keys);
    const int kv = (K / Chunk<T>::N + 31) / 32;
    const dim3 grid((nw + 3) / 4, B);
    if (kv == 1) pool_lin_rows_kernel<T, 1><<<grid, 128, 0, stream>>>(keys, sp, w, ldw, N, C, log2C, K, dx, ld_dx, part, rows_ws);
    else if (kv == 2) pool_lin_rows_kernel<T, 2><<<grid, 128, 0, stream>>>(keys, sp, w, ldw, N, C, log2C, K, dx, ld_dx, part, rows_ws);
    else if (kv <= 4) pool_lin_rows_kernel<T, 4><<<grid, 128, 0, stream>>>(keys, sp, w, ldw, N, C, log2C, K, dx, ld_dx, part, rows_ws);
    else return FS_ERR_UNSUPPORTED;
    int mt = (K / 2 + 31) / 32 * 32;
    if (mt > 256) mt = 256;
    pool_lin_merge_kernel<T><<<B, mt, (size_t)2 * nw * 4, stream>>>(part, rows_ws, N, C, K, dx, ld_dx);
    return FS_OK;
}
template <typename T>
int launch_dw(cudaStream_t stream, const float* sp, const int32_t* arg, const T* x, int ldx, int B, int N, int C, int K,
              const float* a, const float* bvec, const float* colsum, const float* wg, float* dw) {
    const int kv = (K / Chunk<T>::N + 31) / 32;
    const int grid = (C + 7) / 8;
    if (kv == 1) pool_lin_dw_kernel<T, 1><<<grid, 256, 0, stream>>>(sp, arg, x, ldx, B, N, C, K, a, bvec, colsum, wg, dw);
    else if (kv == 2) pool_lin_dw_kernel<T, 2><<<grid, 256, 0, stream>>>(sp, arg, x, ldx, B, N, C, K, a, bvec, colsum, wg, dw);
    else if (kv <= 4) pool_lin_dw_kernel<T, 4><<<grid, 256, 0, stream>>>(sp, arg, x, ldx, B, N, C, K, a, bvec, colsum, wg, dw);
    else return FS_ERR_UNSUPPORTED;
    return FS_OK;
}
}  // namespace

extern "C" size_t fs_pool_lin_bwd_ws_bytes(int B, int C, int K) {
    if (B <= 0 || C <= 0 || K <= 0) return 0;
    return sparse_dx_ws_bytes(B, C, K);
}

extern "C" int fs_pool_lin_bwd_dx_sparse(int device, fs_stream_t stream_, const float* sp, const int32_t* arg, const void* w,
                                         int dtype, int ldw, int B, int N, int C, int K, void* dx, int ld_dx, void* ws) {
    if (!sp || !arg || !w || !dx || !ws || B <= 0 || N <= 0 || C <= 0 || ldw < K || ld_dx < K) return FS_ERR_BAD_ARG;
    const int V = dtype == FS_BF16 ? 8 : 4;
    if (K % V || K > 512 || ldw % V || ld_dx % V || C < 32 || C > 1024 || (C & (C - 1)) || (long long)N * C >= (1ll << 32) - 1 ||
        B > 65535 || ((uintptr_t)w & 15) || ((uintptr_t)dx & 15) || ((uintptr_t)ws & 15))
        return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rc = dtype == FS_BF16
        ? launch_sparse_dx(stream, sp, arg, (const __nv_bfloat16*)w, ldw, B, N, C, K, (__nv_bfloat16*)dx, ld_dx, ws)
        : launch_sparse_dx(stream, sp, arg, (const float*)w, ldw, B, N, C, K, (float*)dx, ld_dx, ws);
    if (rc != FS_OK) return rc;
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_pool_lin_bwd_dw(int device, fs_stream_t stream_, const float* sp, const int32_t* arg, const void* x, int dtype,
                                  int ldx, int B, int N, int C, int K, const float* a, const float* bvec, const float* colsum,
                                  const float* wg, float* dw) {
    if (!sp || !arg || !x || !dw || B <= 0 || N <= 0 || C <= 0 || ldx < K) return FS_ERR_BAD_ARG;
    const int V = dtype == FS_BF16 ? 8 : 4;
    if (K % V || K > 512 || ldx % V || ((uintptr_t)x & 15)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int rc = dtype == FS_BF16
        ? launch_dw(stream, sp, arg, (const __nv_bfloat16*)x, ldx, B, N, C, K, a, bvec, colsum, wg, dw)
        : launch_dw(stream, sp, arg, (const float*)x, ldx, B, N, C, K, a, bvec, colsum, wg, dw);
    if (rc != FS_OK) return rc;
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

namespace {
bool cat_wide(const CatSrc& s, int ld_other) {        // eight columns per thread when every width and stride allows it
    bool ok = ld_other % 8 == 0;
    for (int i = 0; i < s.n; ++i) ok = ok && (s.c0[i + 1] - s.c0[i]) % 8 == 0 && s.ld[i] % 8 == 0;
    return ok;
}
bool cat_aligned(const void* const* ptrs, int n, const void* other) {
    bool ok = ((uintptr_t)other & 15) == 0;
    for (int i = 0; i < n; ++i) ok = ok && ((uintptr_t)ptrs[i] & 15) == 0;
    return ok;
}
}  // namespace

extern "C" int fs_colsum(int device, fs_stream_t stream_, const void* x, int dtype, int ld, long long rows, int K,
                         float* partial_ws, float* out) {
    const int V = dtype == FS_BF16 ? 8 : 4;
    if (!x || !partial_ws || !out || rows <= 0 || K <= 0 || ld < K) return FS_ERR_BAD_ARG;
    if (K % V || K / V > 256 || ld % V || ((uintptr_t)x & 15)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int G = fs_colsum_partials();
    const int rpp = 256 / (K / V);
    const size_t smem = (size_t)rpp * K * sizeof(float);
    if (dtype == FS_BF16) colsum_partial_kernel<__nv_bfloat16, 8><<<G, 256, smem, stream>>>((const __nv_bfloat16*)x, ld, rows, K, partial_ws);
    else colsum_partial_kernel<float, 4><<<G, 256, smem, stream>>>((const float*)x, ld, rows, K, partial_ws);
    FS_RETURN_IF_LAUNCH_FAILED();
    colsum_final_kernel<<<(K + 31) / 32, 1024, 0, stream>>>(partial_ws, G, K, out);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_colsum_partials(void) { return 4 * FS_NUM_SMS; }

extern "C" int fs_cat_cast(int device, fs_stream_t stream_, int n, const void* const* srcs, const int* widths, const int* lds,
                           long long rows, void* out, int out_dtype, int ld_out) {
    CatSrc s;
    if (!out || rows <= 0 || !fill_cat(s, n, srcs, widths, lds, false) || ld_out < s.c0[n] || ld_out % 4) return FS_ERR_BAD_ARG;
    if (!cat_aligned(srcs, n, out) || rows * (long long)(s.c0[n] / 4) >= (1ll << 31)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool wide = cat_wide(s, ld_out);
    const long long total = rows * (s.c0[n] / (wide ? 8 : 4));
    const int grid = (int)(fs_div_up(total, 256) < (long long)FS_NUM_SMS * 16 ? fs_div_up(total, 256) : (long long)FS_NUM_SMS * 16);
    if (out_dtype == FS_BF16) {
        if (wide) cat_cast_kernel<__nv_bfloat16, 8><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (__nv_bfloat16*)out, ld_out);
        else cat_cast_kernel<__nv_bfloat16, 4><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (__nv_bfloat16*)out, ld_out);
    } else {
        if (wide) cat_cast_kernel<float, 8><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (float*)out, ld_out);
        else cat_cast_kernel<float, 4><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (float*)out, ld_out);
    }
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_split_cast(int device, fs_stream_t stream_, int n, void* const* dsts, const int* widths, const int* lds,
                             long long rows, const void* g, int g_dtype, int ld_g) {
    CatSrc s;
    if (!g || rows <= 0 || !fill_cat(s, n, (const void* const*)dsts, widths, lds, true) || ld_g < s.c0[n] || ld_g % 4) return FS_ERR_BAD_ARG;
    if (!cat_aligned((const void* const*)dsts, n, g) || rows * (long long)(s.c0[n] / 4) >= (1ll << 31)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const bool wide = cat_wide(s, ld_g);
    const long long total = rows * (s.c0[n] / (wide ? 8 : 4));
    const int grid = (int)(fs_div_up(total, 256) < (long long)FS_NUM_SMS * 16 ? fs_div_up(total, 256) : (long long)FS_NUM_SMS * 16);
    if (g_dtype == FS_BF16) {
        if (wide) split_cast_kernel<__nv_bfloat16, 8><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (const __nv_bfloat16*)g, ld_g);
        else split_cast_kernel<__nv_bfloat16, 4><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (const __nv_bfloat16*)g, ld_g);
    } else {
        if (wide) split_cast_kernel<float, 8><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (const float*)g, ld_g);
        else split_cast_kernel<float, 4><<<grid, 256, 0, stream>>>(s, (unsigned)rows, (const float*)g, ld_g);
    }
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_multi_copy_f32(int device, fs_stream_t stream_, int n, const void* const* srcs, void* const* dsts,
                                 const long long* counts) {
    if (n < 0 || (n > 0 && (!srcs || !dsts || !counts))) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    for (int i0 = 0; i0 < n; i0 += 32) {
        MultiCopy t;
        t.n = n - i0 < 32 ? n - i0 : 32;
        unsigned chunks = 0;
        for (int i = 0; i < t.n; ++i) {
            const long long c = counts[i0 + i];
            if (c < 0 || c >= (1ll << 31) || (c > 0 && (!srcs[i0 + i] || !dsts[i0 + i]))) return FS_ERR_BAD_ARG;
            t.src[i] = (const float*)srcs[i0 + i];
            t.dst[i] = (float*)dsts[i0 + i];
            t.count[i] = (unsigned)c;
            t.first_chunk[i] = chunks;
            chunks += (unsigned)((c + 1023) / 1024);
        }
        t.first_chunk[t.n] = chunks;
        for (int i = t.n; i < 32; ++i) { t.src[i] = nullptr; t.dst[i] = nullptr; t.count[i] = 0; t.first_chunk[i + 1] = chunks; }
        if (chunks == 0) continue;
        multi_copy_kernel<<<chunks, 256, 0, stream>>>(t);
        FS_RETURN_IF_LAUNCH_FAILED();
    }
    return FS_OK;
}

extern "C" int fs_logits_out(int device, fs_stream_t stream_, const void* logits, int dtype, int ld, const long long* perm, int B,
                             int N, int C, float* out) {
    if (!logits || !out || B <= 0 || N <= 0 || C <= 0 || ld < C) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long total = (long long)B * N;
    const int grid = (int)(fs_div_up(total, 256) < (long long)FS_NUM_SMS * 8 ? fs_div_up(total, 256) : (long long)FS_NUM_SMS * 8);
    if (dtype == FS_BF16) logits_out_kernel<<<grid, 256, 0, stream>>>((const __nv_bfloat16*)logits, ld, perm, B, N, C, out);
    else logits_out_kernel<<<grid, 256, 0, stream>>>((const float*)logits, ld, perm, B, N, C, out);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_logits_out_bwd(int device, fs_stream_t stream_, const float* g, const long long* perm, int B, int N, int C,
                                 void* dlogits, int dtype, int ld) {
    if (!g || !dlogits || B <= 0 || N <= 0 || C <= 0 || ld < C) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long total = (long long)B * N;
    const int grid = (int)(fs_div_up(total, 256) < (long long)FS_NUM_SMS * 8 ? fs_div_up(total, 256) : (long long)FS_NUM_SMS * 8);
    if (dtype == FS_BF16) logits_out_bwd_kernel<<<grid, 256, 0, stream>>>(g, perm, B, N, C, (__nv_bfloat16*)dlogits, ld);
    else logits_out_bwd_kernel<<<grid, 256, 0, stream>>>(g, perm, B, N, C, (float*)dlogits, ld);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

namespace {
template <typename T, int NOUT, int CIN>
int final_fwd_t(cudaStream_t stream, const T* h, int ld, const float* w, const float* bias, const long long* perm, int B, int N,
                float* out) {
    const long long total = (long long)B * N;
    const size_t smem = (size_t)128 * (CIN + Chunk<T>::N) * sizeof(T) + (size_t)NOUT * CIN * 4;
    FS_CUDA_TRY(cudaFuncSetAttribute(final_linear_fwd_kernel<T, NOUT, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    final_linear_fwd_kernel<T, NOUT, CIN><<<(unsigned)fs_div_up(total, 128), 128, smem, stream>>>(h, ld, w, bias, perm, B, N, out);
    return FS_OK;
}
template <typename T, int NOUT, int CIN>
int final_bwd_t(cudaStream_t stream, const T* h, int ld, const float* w, const float* g, const long long* perm, int B, int N,
                T* dh, int ld_dh, float* part, float* dwb) {
    const long long total = (long long)B * N;
    const int grid = (int)fs_div_up(total, FL_TILE);
    const size_t smem = (size_t)FL_TILE * CIN * sizeof(T) + (size_t)NOUT * CIN * 4 + (size_t)FL_TILE * NOUT * 4;
    FS_CUDA_TRY(cudaFuncSetAttribute(final_linear_bwd_kernel<T, NOUT, CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    final_linear_bwd_kernel<T, NOUT, CIN><<<grid, 256, smem, stream>>>(h, ld, w, g, perm, B, N, dh, ld_dh, part);
    const int n = NOUT * CIN + NOUT;
    final_linear_red_kernel<<<(n + 31) / 32, 1024, 0, stream>>>(part, grid, n, dwb);
    return FS_OK;
}
bool final_ok(int C_in, int C_out) {
    return (C_in == 128 && (C_out == 2 || C_out == 4 || C_out == 8)) || (C_in == 256 && (C_out == 2 || C_out == 4));
}
}  // namespace

// one dispatch for both directions: FL_GO(fn, args...) instantiates fn<T, NOUT, CIN>
#define FL_DISPATCH(FN, ...)                                                                                        \
    do {                                                                                                            \
        if (C_in == 128 && C_out == 2) rc = FN<TT, 2, 128>(__VA_ARGS__);                                             \
        else if (C_in == 128 && C_out == 4) rc = FN<TT, 4, 128>(__VA_ARGS__);                                        \
        else if (C_in == 128 && C_out == 8) rc = FN<TT, 8, 128>(__VA_ARGS__);                                        \
        else if (C_in == 256 && C_out == 2) rc = FN<TT, 2, 256>(__VA_ARGS__);                                        \
        else rc = FN<TT, 4, 256>(__VA_ARGS__);                                                                       \
    } while (0)

extern "C" int fs_final_linear_supported(int C_in, int C_out) { return final_ok(C_in, C_out) ? 1 : 0; }
extern "C" size_t fs_final_linear_ws_floats(long long rows, int C_in, int C_out) {
    if (rows <= 0) return 0;
    return (size_t)fs_div_up(rows, FL_TILE) * (size_t)(C_out * C_in + C_out);
}

extern "C" int fs_final_linear_fwd(int device, fs_stream_t stream_, const void* h, int dtype, int ld, const float* w,
                                   const float* bias, const long long* perm, int B, int N, int C_in, int C_out, float* out) {
    if (!h || !w || !out || B <= 0 || N <= 0 || ld < C_in) return FS_ERR_BAD_ARG;
    if (!final_ok(C_in, C_out) || ld % 8 || ((uintptr_t)h & 15)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = FS_OK;
    if (dtype == FS_BF16) {
        using TT = __nv_bfloat16;
        FL_DISPATCH(final_fwd_t, stream, (const TT*)h, ld, w, bias, perm, B, N, out);
    } else {
        using TT = float;
        FL_DISPATCH(final_fwd_t, stream, (const TT*)h, ld, w, bias, perm, B, N, out);
    }
    if (rc != FS_OK) return rc;
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_final_linear_bwd(int device, fs_stream_t stream_, const void* h, int dtype, int ld, const float* w, const float* g,
                                   const long long* perm, int B, int N, int C_in, int C_out, void* dh, int ld_dh, float* ws,
                                   float* dw_db /* [C_out * C_in] dW, then [C_out] dbias */) {
    if (!h || !w || !g || !dh || !ws || !dw_db || B <= 0 || N <= 0 || ld < C_in || ld_dh < C_in) return FS_ERR_BAD_ARG;
    if (!final_ok(C_in, C_out) || ld % 8 || ld_dh % 8 || ((uintptr_t)h & 15) || ((uintptr_t)dh & 15)) return FS_ERR_UNSUPPORTED;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = FS_OK;
    if (dtype == FS_BF16) {
        using TT = __nv_bfloat16;
        FL_DISPATCH(final_bwd_t, stream, (const TT*)h, ld, w, g, perm, B, N, (TT*)dh, ld_dh, ws, dw_db);
    } else {
        using TT = float;
        FL_DISPATCH(final_bwd_t, stream, (const TT*)h, ld, w, g, perm, B, N, (TT*)dh, ld_dh, ws, dw_db);
    }
    if (rc != FS_OK) return rc;
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_edge_weight_table(int device, fs_stream_t stream_, const float* w, int Cp, int C, float* out) {
    if (!w || !out || Cp <= 0 || C <= 0) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    edge_weight_table_kernel<<<(2 * Cp * C + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(w, Cp, C, out);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
extern "C" int fs_edge_weight_table_bwd(int device, fs_stream_t stream_, const float* g, int Cp, int C, float* dw) {
    if (!g || !dw || Cp <= 0 || C <= 0) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    edge_weight_table_bwd_kernel<<<(2 * Cp * C + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(g, Cp, C, dw);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_sum_leading(int device, fs_stream_t stream_, const float* part, int S, long long n, float* out) {
    if (!part || !out || S <= 0 || n <= 0) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    cudaStream_t stream = (cudaStream_t)stream_;
    // small outputs: more partial groups per output so that the grid still covers the chip
    if (n >= 32768) sum_leading_kernel<4><<<(unsigned)fs_div_up(n, 64), 256, 0, stream>>>(part, S, n, out);
    else sum_leading_kernel<8><<<(unsigned)fs_div_up(n, 32), 256, 0, stream>>>(part, S, n, out);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
