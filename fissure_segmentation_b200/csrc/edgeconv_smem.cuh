// Internal interface between edgeconv.cu (C-ABI entry points) and edgeconv_smem.cu (shared-memory-resident
// gather kernels). Returns 0, a positive cudaError_t, or FS_SMEM_GATHER_UNSUPPORTED when the shape does not fit
// (the caller then launches the global-memory gather kernel).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define FS_SMEM_GATHER_UNSUPPORTED (-1000)

int fs_gather_smem_train(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k,
                         int CP, const float* gamma, const int32_t* rev_ptr, float* sel, uint8_t* arg, float* sy,
                         double* stats);
int fs_gather_smem_eval(cudaStream_t stream, const float* table, int ld, const int32_t* idx, int B, int N, int k,
                        int CP, const float* coef, void* out, int out_bf16, int ld_out, uint8_t* arg);
