// FP32 FMA throughput micro-benchmark: the roofline denominator of the CUDA-core kernels (3-D kNN, Chamfer), which
// MEASURED_PEAKS.json does not hold ("report % of measured FP32 FMA peak (micro-benchmark)").
#include "fs_common.cuh"

namespace {

// 8 independent FMA chains per thread (ILP 8), 256 threads, 8 CTAs per SM: enough to saturate the FP32 pipes.
__global__ void __launch_bounds__(256)
fma_peak_kernel(int iters, float a, float b, float* __restrict__ out) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = (float)(threadIdx.x + i) * 1e-3f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    if (s == 123.456f) out[0] = s;      // never true; keeps the chains alive
}

}  // namespace

// Launches grid = 148 SMs x 8 CTAs of 256 threads, each thread executing iters * 64 FMAs.
// flops per launch = 2 * 64 * iters * 256 * 148 * 8 (returned through *flops_out on the host).
extern "C" int fs_fma_microbench(int device, fs_stream_t stream_, int iters, float* out, double* flops_out) {
    if (iters <= 0 || !out) return FS_ERR_BAD_ARG;
    FS_ENTER(device);
    const int ctas = FS_NUM_SMS * 8;
    fma_peak_kernel<<<ctas, 256, 0, (cudaStream_t)stream_>>>(iters, 0.999f, 1e-4f, out);
    FS_RETURN_IF_LAUNCH_FAILED();
    if (flops_out) *flops_out = 2.0 * 64.0 * (double)iters * 256.0 * (double)ctas;
    return FS_OK;
}
