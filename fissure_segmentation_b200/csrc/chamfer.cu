// Chamfer distance building block: brute-force nearest neighbour (K = 1) with running minimum in
// registers, and its gradient. Follows losses/chamfer_loss.py:9-20, i.e. pytorch3d's
// chamfer_distance defaults: squared L2 from direct differences, mean over points, sum of both
// directions, mean over the batch.
#include "fs_common.cuh"

namespace {

constexpr int NN_THREADS = 256;
constexpr int NN_QPT = 2;      // queries per thread (shares every staged candidate load)
constexpr int NN_TILE = 2048;  // staged candidates per pass (24 KB)

__global__ void __launch_bounds__(NN_THREADS)
nn_points_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int M,
                 float* __restrict__ nn_d2, int32_t* __restrict__ nn_idx) {
    __shared__ float sx[NN_TILE], sy[NN_TILE], sz[NN_TILE];
    const int b = blockIdx.y;
    const float* xb = x + (long long)b * N * 3;
    const float* yb = y + (long long)b * M * 3;
    float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], best[NN_QPT];
    int bi[NN_QPT];
    int q[NN_QPT];
#pragma unroll
    for (int u = 0; u < NN_QPT; ++u) {
        q[u] = (blockIdx.x * NN_QPT + u) * NN_THREADS + threadIdx.x;
        const int qq = q[u] < N ? q[u] : N - 1;
        qx[u] = __ldg(xb + 3ll * qq); qy[u] = __ldg(xb + 3ll * qq + 1); qz[u] = __ldg(xb + 3ll * qq + 2);
        best[u] = INFINITY; bi[u] = 0;
    }
    for (int m0 = 0; m0 < M; m0 += NN_TILE) {
        const int mn = min(NN_TILE, M - m0);
        __syncthreads();
        for (int j = threadIdx.x; j < mn; j += NN_THREADS) {
            sx[j] = __ldg(yb + 3ll * (m0 + j));
            sy[j] = __ldg(yb + 3ll * (m0 + j) + 1);
            sz[j] = __ldg(yb + 3ll * (m0 + j) + 2);
        }
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < mn; ++j) {
            const float cx = sx[j], cy = sy[j], cz = sz[j];
#pragma unroll
            for (int u = 0; u < NN_QPT; ++u) {
                const float dx = qx[u] - cx, dy = qy[u] - cy, dz = qz[u] - cz;
                const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (d < best[u]) { best[u] = d; bi[u] = m0 + j; }
            }
        }
    }
#pragma unroll
    for (int u = 0; u < NN_QPT; ++u) {
        if (q[u] < N) {
            nn_d2[(long long)b * N + q[u]] = best[u];
            nn_idx[(long long)b * N + q[u]] = bi[u];
        }
    }
}

__global__ void __launch_bounds__(256)
chamfer_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y, const int32_t* __restrict__ nn_idx,
                   long long total, int N, int M, float weight, const float* __restrict__ w_ptr,
                   float* __restrict__ gx, float* __restrict__ gy) {
    const float w = 2.f * weight * (w_ptr ? __ldg(w_ptr) : 1.f);
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const long long b = e / N;
        const int j = __ldg(nn_idx + e);
        const float* xp = x + 3 * e;
        const float* yp = y + 3 * (b * M + j);
        const float gx0 = w * (xp[0] - yp[0]), gx1 = w * (xp[1] - yp[1]), gx2 = w * (xp[2] - yp[2]);
        atomicAdd(gx + 3 * e, gx0);
        atomicAdd(gx + 3 * e + 1, gx1);
        atomicAdd(gx + 3 * e + 2, gx2);
        if (gy) {
            float* gp = gy + 3 * (b * M + j);
            atomicAdd(gp, -gx0);
            atomicAdd(gp + 1, -gx1);
            atomicAdd(gp + 2, -gx2);
        }
    }
}

}  // namespace

extern "C" int fs_nn_points(int device, fs_stream_t stream_, const float* x, const float* y, int B, int N, int M,
                            float* nn_d2, int32_t* nn_idx) {
    if (!x || !y || !nn_d2 || !nn_idx || B < 0 || N < 0 || M <= 0) return FS_ERR_BAD_ARG;
    if (B == 0 || N == 0) return FS_OK;
    FS_ENTER(device);
    dim3 grid(fs_div_up(N, NN_THREADS * NN_QPT), B);
    nn_points_kernel<<<grid, NN_THREADS, 0, (cudaStream_t)stream_>>>(x, y, N, M, nn_d2, nn_idx);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}

extern "C" int fs_chamfer_bwd(int device, fs_stream_t stream_, const float* x, const float* y, const int32_t* nn_idx,
                              int B, int N, int M, float weight, const float* w_ptr, float* gx, float* gy) {
    if (!x || !y || !nn_idx || !gx || B < 0 || N < 0 || M <= 0) return FS_ERR_BAD_ARG;
    if (B == 0 || N == 0) return FS_OK;
    FS_ENTER(device);
    const long long total = (long long)B * N;
    int grid = fs_div_up(total, 256);
    if (grid > FS_NUM_SMS * 8) grid = FS_NUM_SMS * 8;
    chamfer_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream_>>>(x, y, nn_idx, total, N, M, weight, w_ptr, gx, gy);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
