// Ensemble inference helpers: PointSegmentationModelBase.predict_full_pointcloud (models/point_seg_net.py:21-48)
// runs >= 50 eval-mode forwards on random subsets of one large cloud and accumulates the class probabilities per
// point. With the subset forwards batched into one launch, the accumulation
//     softmax_accumulation[..., perm] += softmax(self(pc[..., perm]))          (point_seg_net.py:27-29, :43)
// of all R runs is this one kernel.
#include "fs_common.cuh"

namespace {

constexpr int INFER_MAX_CLASSES = 32;

// One thread per (run, subset position): softmax over the classes (stride S in the B x classes x N logits), then one
// atomic per class into acc[c][sub[r][s]]. Within a run the subset indices are unique (randperm), across runs a point
// receives up to R contributions; fp32 atomics make the summation ORDER over the runs non-deterministic (differences
// of one ulp of the sum), documented in DESIGN.md.
__global__ void __launch_bounds__(256)
softmax_scatter_add_kernel(const float* __restrict__ logits, const long long* __restrict__ sub, int R, int classes, int S,
                           int n_total, float* __restrict__ acc) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)R * S) return;
    const int r = (int)(t / S), s = (int)(t - (long long)r * S);
    const long long p = sub[t];
    if (p < 0 || p >= n_total) return;
    const float* lg = logits + (long long)r * classes * S + s;
    float v[INFER_MAX_CLASSES];
    float m = -INFINITY;
#pragma unroll
    for (int c = 0; c < INFER_MAX_CLASSES; ++c) {
        if (c < classes) {
            v[c] = __ldg(lg + (long long)c * S);
            m = fmaxf(m, v[c]);
        }
    }
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < INFER_MAX_CLASSES; ++c) {
        if (c < classes) {
            v[c] = expf(v[c] - m);
            sum += v[c];
        }
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int c = 0; c < INFER_MAX_CLASSES; ++c)
        if (c < classes) atomicAdd(acc + (long long)c * n_total + p, v[c] * inv);
}

}  // namespace

extern "C" int fs_softmax_scatter_add(int device, fs_stream_t stream_, const float* logits, const long long* sub, int R,
                                      int classes, int S, int n_total, float* acc) {
    if (R < 0 || S < 0 || classes <= 0 || n_total <= 0) return FS_ERR_BAD_ARG;
    if (classes > INFER_MAX_CLASSES) return FS_ERR_UNSUPPORTED;
    if (R == 0 || S == 0) return FS_OK;
    if (!logits || !sub || !acc) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    const long long total = (long long)R * S;
    softmax_scatter_add_kernel<<<fs_div_up(total, 256), 256, 0, (cudaStream_t)stream_>>>(logits, sub, R, classes, S, n_total, acc);
    FS_RETURN_IF_LAUNCH_FAILED();
    return FS_OK;
}
