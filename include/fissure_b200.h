/*
 * fissure_b200 — C ABI of the B200-native DGCNN EdgeConv hot path.
 *
 * Every entry point is what a binding for the reference (kaftanski/fissure-segmentation) would
 * call in place of the PyTorch expression cited beside it. Citations are file:line in the
 * reference tree.
 *
 * Conventions
 *  - All pointers are DEVICE pointers owned by the caller (torch tensors in practice). The library
 *    never allocates, frees or retains device memory and keeps no mutable global state.
 *  - `device` is the CUDA ordinal the buffers live on; `stream` is a cudaStream_t of that device.
 *    Calls are asynchronous with respect to the host and re-entrant (autograd worker threads).
 *  - Return value: 0 = success, < 0 = fs_status (bad arguments, nothing launched),
 *    > 0 = cudaError_t of the failed launch. No exceptions, no exit().
 *  - Point-major tables: a "table" is a row-major matrix with one row per point (row = b*N + n),
 *    `ld` = row stride in elements. dtype codes: FS_F32 = 0, FS_BF16 = 1.
 *  - Neighbour indices are int32, local to the cloud (0..N-1), one row of k per point.
 */
#ifndef FISSURE_B200_H
#define FISSURE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fs_stream_t; /* cudaStream_t */

enum fs_status {
    FS_OK = 0,
    FS_ERR_BAD_ARG = -1,     /* null pointer, negative size, k out of range ... */
    FS_ERR_UNSUPPORTED = -2, /* shape or dtype outside the compiled set */
    FS_ERR_ALIGNMENT = -3    /* pointer / leading dimension not 16-byte aligned */
};

enum fs_dtype { FS_F32 = 0, FS_BF16 = 1 };

#define FS_MAX_K 127 /* largest k (k+1 when the self match is dropped must be <= 128) */

int fs_version(void);
/*
 * Number of doubles of a per-channel statistics buffer for C channels (the `stats` / `dgb` arguments below):
 * [0,2C) final sums | [2C,3C) pivot | slot partials | ticket. The caller zero-fills it before each use.
 */
size_t fs_stats_buffer_doubles(int C);
/* Static string for a negative fs_status; cudaGetErrorString for positive codes. */
const char* fs_error_string(int code);

/* ---------------------------------------------------------------- kNN graph build ---------- */

/*
 * Batched k-nearest-neighbour search on 3-D coordinates (FP32 FMA + warp-shuffle bitonic top-k).
 * Replaces utils/general_utils.py:315-327 knn() on top of :43-53 pairwise_dist() for C == 3 and
 * models/dgcnn_opensrc.py:34-40 knn().
 *   coords           element (b, c, n) at coords[b*batch_stride + c*chan_stride + n*point_stride]
 *                    (B x C x N reference layout: chan_stride = N, point_stride = 1)
 *   self_loop        1: the query itself may be returned (knn(..., self_loop=True));
 *                    0: k+1 are selected and rank 0 is dropped (general_utils.py:317-322)
 *   diag_zero        1: d(i,i) is forced to 0 (general_utils.py:52); 0: dgcnn_opensrc.knn
 *   idx  [B*N*k]     int32 neighbour indices, ascending distance, ties -> lower index
 *   dist2 [B*N*k]    squared distances in the reference's expansion form (nullable)
 */
int fs_knn3d(int device, fs_stream_t stream, const float* coords, long long batch_stride,
             long long chan_stride, long long point_stride, int B, int N, int k, int self_loop,
             int diag_zero, int32_t* idx, float* dist2);

/*
 * Exact FP32 k-nearest-neighbour search in C-dimensional feature space on a point-major table.
 * Replaces the same reference functions for C > 3 (models/dgcnn.py:26 on 64-channel features).
 *   x [B*N, ldx]     fp32 features, point-major
 *   sqnorm_ws [B*N]  workspace for the row squared norms
 */
int fs_knn_feat(int device, fs_stream_t stream, const float* x, int ldx, int B, int N, int C, int k,
                int self_loop, int diag_zero, int32_t* idx, float* dist2, float* sqnorm_ws);

/*
 * Same contract as fs_knn_feat on the 5th-generation tensor cores, for C = 64, 128 or 256 features: the -2 X X^T
 * contraction runs as tcgen05.mma on TMA-staged fp16 operand tiles (centred and scaled per cloud; accumulators and the
 * query operand in TMEM). Two sweeps select ~1.2 k candidates per query (the threshold of the second sweep is part of
 * the tensor-core product), a finalize kernel decides everything further than the representation error from the
 * k-th by the approximation and re-evaluates the rest in the reference's FP32 arithmetic. Rows whose candidate
 * set cannot be certified complete (list overflow, NaN / Inf in the cloud) are recomputed by the exact kernel, so the
 * neighbour SET equals fs_knn_feat's; the order inside the set follows the approximate distances. dist2 != NULL
 * selects the exact kernel for every row.
 *   workspace        >= fs_knn_feat_tc_workspace_bytes(B, N, C, k) bytes, 256-byte aligned; contents are scratch
 * fs_knn_feat_tc_supported returns 1 when the shape is handled by the tensor-core kernels
 * (C in {64, 128, 256}, k + !self_loop <= 64, 64 <= N <= 32768); other shapes run the exact kernel.
 */
size_t fs_knn_feat_tc_workspace_bytes(int B, int N, int C, int k);
/* Byte offset, inside the workspace, of the B*N per-row flags (uint8) that are 1 for every query the tensor-core path
 * handed to the exact kernel (survivor-list overflow, too few survivors, NaN/Inf row). Diagnostics only. */
size_t fs_knn_feat_tc_redo_offset(int B, int N, int C, int k);
int fs_knn_feat_tc_supported(int B, int N, int C, int k, int self_loop);
int fs_knn_feat_tc(int device, fs_stream_t stream, const float* x, int ldx, int B, int N, int C, int k,
                   int self_loop, int diag_zero, int32_t* idx, float* dist2, void* workspace,
                   size_t workspace_bytes);

/*
 * fs_knn3d (same arguments, same neighbour order: ascending (distance, index)) through the tensor-core kernels: the
 * hi/lo fp16 split of the three centred coordinates and the norms share ONE K = 16 MMA step, so the tensor core
 * produces all N x N distances and the CUDA cores only select. Any N in [64, 32768], k + !self_loop <= 64 (the
 * register-resident SIMT kernel stops at N = 2048). Distances are not produced: use fs_knn3d when dist2 is wanted.
 * The redo flags sit at fs_knn_feat_tc_redo_offset(B, N, 3, k).
 */
size_t fs_knn3d_tc_workspace_bytes(int B, int N, int k);
int fs_knn3d_tc_supported(int B, int N, int k, int self_loop);
int fs_knn3d_tc(int device, fs_stream_t stream, const float* coords, long long batch_stride,
                long long chan_stride, long long point_stride, int B, int N, int k, int self_loop,
                int diag_zero, int32_t* idx, void* workspace, size_t workspace_bytes);

/*
 * Offset-segmented kNN of `new_xyz` in `xyz` (direct squared differences).
 * Replaces pointops_cuda.knnquery_cuda at models/pointtransformer/pointops.py:59.
 *   xyz [n,3], new_xyz [m,3], offset [b], new_offset [b]: cumulative int32 segment ends
 *   idx [m,nsample] int32 GLOBAL row indices into xyz; dist2 [m,nsample] squared distances.
 *   Segments with fewer than nsample points are padded with (segment start, 1e10).
 */
int fs_knnquery(int device, fs_stream_t stream, int m, int nsample, const float* xyz,
                const float* new_xyz, const int32_t* offset, const int32_t* new_offset, int b,
                int32_t* idx, float* dist2);

/*
 * Farthest point sampling per segment. Replaces pointops_cuda.furthestsampling_cuda at
 * models/pointtransformer/pointops.py:35. `tmp` [n] must be pre-filled with 1e10 by the caller
 * (pointops.py:32); idx [m_total] receives global row indices, first pick = segment start.
 * n_max = size of the largest segment (the second argument of the upstream call, pointops.py:25-27, 35); when
 * 0 < n_max <= 8192 every thread keeps its points and running distances in registers and the segment's coordinates
 * sit in shared memory (one block barrier per pick); n_max = 0 selects the global-memory kernel.
 */
int fs_furthestsampling(int device, fs_stream_t stream, int b, int n_max, const float* xyz,
                        const int32_t* offset, const int32_t* new_offset, float* tmp, int32_t* idx);

/*
 * 30-bit Morton (Z-order) code of the xyz channels of each point (box [-2,2)^3, 10 bits per axis). The
 * modules sort every cloud by this code before the EdgeConvs so that the neighbour rows of consecutive
 * points overlap (L1/L2 locality of the gathers); all operators are permutation-equivariant, the logits are
 * un-permuted at the end. Same addressing as fs_knn3d.
 */
int fs_morton_codes(int device, fs_stream_t stream, const float* coords, long long batch_stride,
                    long long chan_stride, long long point_stride, int B, int N, int32_t* codes);

/* ---------------------------------------------------------------- fused EdgeConv ----------- */

/*
 * Pass 1 of the fused single-layer EdgeConv (models/dgcnn.py:226-243 with one SharedFullyConnected,
 * i.e. create_neighbor_features :15-36 + Conv2d 1x1 + BatchNorm2d + LeakyReLU(0.2) + max over k).
 * The per-point table T = X * [W1 ; W2-W1]^T holds a = T[:, :Cp] and b = T[:, Cp:2Cp]; the edge
 * pre-activation is y(i,j) = a_j + b_i. One gather pass over the k neighbour rows produces
 *   sel [P,Cp]  f32   max_j a_j where gamma >= 0, min_j a_j where gamma < 0
 *   arg [P,Cp]  u8    slot (0..k-1) of the selected neighbour
 *   sy  [P,Cp]  f32   sum_j y(i,j)                      (nullable: eval / no-grad)
 *   stats f64 [fs_stats_buffer_doubles(Cp)], zeroed by the caller: sum(y-p), sum((y-p)^2) per channel,
 *                     then the pivot p (nullable: eval mode, no batch statistics)
 */
int fs_edgeconv_gather(int device, fs_stream_t stream, const void* table, int dtype, int ld,
                       const int32_t* idx, int B, int N, int k, int Cp, const float* gamma,
                       const int32_t* rev_ptr /* in-degrees from fs_reverse_graph; required with stats */,
                       float* sel, uint8_t* arg, float* sy, double* stats);

/*
 * BatchNorm2d batch statistics -> per-channel affine (torch.nn.BatchNorm2d, models/dgcnn.py:307).
 *   stats           from fs_edgeconv_gather; count = number of edges B*N*k
 *   coef [4*Cp] f32 out: mu, invstd, scale = gamma*invstd, shift = beta
 *   running_mean / running_var / num_batches_tracked updated in place when non-null
 *   (momentum 0.1, unbiased variance).
 */
int fs_bn_finalize(int device, fs_stream_t stream, const double* stats, double count, int Cp,
                   const float* gamma, const float* beta, float eps, float momentum, float* coef,
                   float* running_mean, float* running_var, long long* num_batches_tracked);

/* Eval-mode coefficients from running statistics (same coef layout). */
int fs_bn_coef_eval(int device, fs_stream_t stream, int Cp, const float* gamma, const float* beta,
                    const float* running_mean, const float* running_var, float eps, float* coef);

/*
 * Pass 2: out = LeakyReLU_0.2(scale*(sel + b - mu) + beta), written to a point-major table
 * (fp32 or bf16, leading dimension ld_out — typically a slice of the 192-wide concat buffer).
 * table == NULL means b = 0 (second layer of a two-layer EdgeConv, where sel comes from
 * fs_edge_reduce); the same holds for fs_edgeconv_bwd_reduce.
 */
int fs_edgeconv_apply(int device, fs_stream_t stream, const float* sel, const void* table,
                      int dtype, int ld, long long P, int Cp, const float* coef, void* out,
                      int out_dtype, int ld_out);
/* The same with fs_bn_finalize folded in (training): see fs_bn_act_apply_fin. */
int fs_edgeconv_apply_fin(int device, fs_stream_t stream, const float* sel, const void* table, int dtype, int ld,
                          long long P, int Cp, const double* stats, double count, const float* gamma,
                          const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                          long long* num_batches_tracked, float* coef_out, void* out, int out_dtype, int ld_out);

/*
 * Eval-mode single pass: gather + select + folded BN + LeakyReLU straight to `out`.
 */
int fs_edgeconv_fused_eval(int device, fs_stream_t stream, const void* table, int dtype, int ld,
                           const int32_t* idx, int B, int N, int k, int Cp, const float* coef,
                           void* out, int out_dtype, int ld_out, uint8_t* arg /* nullable */);

/*
 * Reverse (incoming-edge) graph of a kNN graph, per cloud: counting sort of the B*N*k edges by
 * target. rev_ptr [B*N+1] global exclusive offsets; rev_src [B*N*k] global source rows (order inside a
 * target's list unspecified). Every idx entry must be a valid row of its cloud (0 <= idx < N), as the kNN
 * entry points produce them; three small launches (histogram, per-cloud scan, fill).
 */
int fs_reverse_graph(int device, fs_stream_t stream, const int32_t* idx, int B, int N, int k,
                     int32_t* rev_ptr, int32_t* rev_src);

/*
 * Backward, step 1: d = g * LeakyReLU'(z), dbeta = sum d, dgamma = sum d*yhat.
 *   g [P, ldg] fp32/bf16 upstream gradient; d [P,Cp] f32 out; dgb [2*Cp] f64 (zeroed by caller).
 */
int fs_edgeconv_bwd_reduce(int device, fs_stream_t stream, const void* g, int g_dtype, int ldg,
                           const float* sel, const void* table, int dtype, int ld, long long P,
                           int Cp, const float* coef, float* d, double* dgb);

/*
 * Backward, step 2: BatchNorm-coupled gradient of the per-point table, dT [P, 2*Cp] f32:
 *   db_i = s*( d_i - k*dbeta/M - (dgamma/M)*invstd*(sy_i - k*mu) )
 *   da_j = -s*( indeg_j*dbeta/M + (dgamma/M)*invstd*(indeg_j*(a_j-mu) + sum_{i->j} b_i) )
 * (train_stats = 0: eval-mode BN, da_j = 0 and db_i = s*d_i.)
 * Also writes dgamma/dbeta as fp32 [2*Cp] (dgamma first) when dgamma_dbeta is non-null.
 */
int fs_edgeconv_bwd_point(int device, fs_stream_t stream, const float* d, const float* sy,
                          const void* table, int dtype, int ld, const int32_t* rev_ptr,
                          const int32_t* rev_src, long long P, int k, int Cp, const float* coef,
                          const double* dgb, double count, int train_stats, float* dT,
                          float* dgamma_dbeta);

/*
 * Backward, step 3: argmax-routed scatter  dT[idx[i, arg[i,c]], c] += s_c * d[i,c]
 * (warp-aggregated atomics on the a-half of dT).
 */
int fs_edgeconv_bwd_route(int device, fs_stream_t stream, const float* d, const uint8_t* arg,
                          const int32_t* idx, int B, int N, int k, int Cp, const float* coef,
                          float* dT);

/* ---------------------------------------------------------------- edge tensors (2-layer) --- */

/*
 * Y[(i,t), :] = a[idx[i,t]] + b[i] for all edges (first layer of a two-layer EdgeConv,
 * models/dgcnn.py:119 ec1 and :251). Y [P*k, Cp] fp32 or bf16.
 */
int fs_edge_build(int device, fs_stream_t stream, const void* table, int dtype, int ld,
                  const int32_t* idx, int B, int N, int k, int Cp, void* y, int y_dtype);

/* Backward of fs_edge_build: dT[:, Cp:] = sum_t dY ; dT[idx[i,t], :Cp] += dY (atomics). dT f32 zeroed by caller. */
int fs_edge_build_bwd(int device, fs_stream_t stream, const void* dy, int dy_dtype,
                      const int32_t* idx, int B, int N, int k, int Cp, float* dT);

/*
 * Statistics + max/min over k of a materialised edge tensor Z [P*k, Cp] (second layer output):
 * same outputs as fs_edgeconv_gather with y := z.
 */
int fs_edge_reduce(int device, fs_stream_t stream, const void* z, int z_dtype, long long P, int k,
                   int Cp, const float* gamma, float* sel, uint8_t* arg, float* sy, double* stats);

/*
 * Backward of (BatchNorm2d + LeakyReLU + max over k) on a materialised edge tensor:
 *   dZ[(i,t),c] = s*( [arg[i,c]==t]*d[i,c] - dbeta/M - (dgamma/M)*invstd*(z[(i,t),c]-mu) )
 * written dense, fp32 or bf16.
 */
int fs_edge_reduce_bwd(int device, fs_stream_t stream, const void* z, int z_dtype, const float* d,
                       const uint8_t* arg, long long P, int k, int Cp, const float* coef,
                       const double* dgb, double count, int train_stats, void* dz, int dz_dtype);

/* ---------------------------------------------------------------- first EdgeConv layer on 3-D inputs - */

/*
 * Conv2d(6 -> Cp, 1x1) + BatchNorm2d + LeakyReLU(0.2) on the edge features [x_j - x_i, x_i] of a 3-channel input
 * (first shared layer of ec1 and of the spatial transformer's EdgeConv, models/dgcnn.py:119, 251, 15-36):
 *   fs_edge3_bn_coef  batch statistics of all Cp channels from the 27 moments of the 6-D edge vectors
 *                     (moments: zero-filled f64 workspace of fs_edge3_moment_doubles() entries) -> coef [4*Cp],
 *                     running statistics updated like fs_bn_finalize
 *   fs_edge3_hidden   H [B*N*k, Cp] (fp32 / bf16) = LeakyReLU(scale (w.e - mu) + beta)
 *   fs_edge3_bwd      ONE pass over dH: dgb (zeroed stats buffer) = (sum d, sum d*yhat), acc_ws = sum_e d_e (x) e; the
 *                     BatchNorm-coupled dW [Cp, 6] follows in closed form from the moments (the 3-channel input
 *                     itself gets no gradient)
 *   x [B*N, ldx] fp32 point-major (first 3 channels), w [Cp, 6] fp32 row-major, idx as everywhere.
 */
size_t fs_edge3_moment_doubles(void);
int fs_edge3_bn_coef(int device, fs_stream_t stream, const float* x, int ldx, const int32_t* idx, int B, int N,
                     int k, const float* w, int Cp, const float* gamma, const float* beta, float eps,
                     float momentum, double* moments, float* coef, float* running_mean, float* running_var,
                     long long* num_batches_tracked);
int fs_edge3_hidden(int device, fs_stream_t stream, const float* x, int ldx, const int32_t* idx, int B, int N,
                    int k, const float* w, int Cp, const float* coef, void* h, int h_dtype);
int fs_edge3_bwd(int device, fs_stream_t stream, const float* x, int ldx, const int32_t* idx, int B, int N, int k,
                 const float* w, int Cp, const float* coef, const void* dh, int dh_dtype, int train_stats,
                 const double* moments /* from fs_edge3_bn_coef */, double* dgb,
                 float* acc_ws /* [Cp*6] zeroed */, float* dw /* [Cp*6] out */);

/* ---------------------------------------------------------------- dense layers ------------- */

/*
 * BatchNorm + LeakyReLU around the 1x1-conv GEMMs of SharedFullyConnected(dim=1) (models/dgcnn.py:282-323,
 * used at :123-137) on point-major tables x [rows, ld] (fp32 or bf16, C a power of two >= 64).
 * `rowbias` (nullable) is a per-cloud bias [rows/N, C] fp32 added to x on load: segmentation[0] on
 * [local | broadcast global] (models/dgcnn.py:159) = local GEMM + one bias row per cloud.
 *   fs_colstats      stats [3*C] f64 += per-column sum(x-p), sum((x-p)^2); writes pivot p (zero stats first);
 *                    feed fs_bn_finalize with count = rows.
 *   fs_bn_act_apply  out = LeakyReLU_slope(scale*(x - mu) + beta)
 *   fs_bn_act_bwd    dgb [2*C] f64 += (sum d, sum d*xhat) with d = g*LeakyReLU'(z); then (dx non-null)
 *                    dx = scale*(d - dbeta/M - xhat*dgamma/M)   (train_stats = 0: dx = scale*d)
 */
int fs_colstats(int device, fs_stream_t stream, const void* x, int dtype, int ld, long long rows, int C,
                const float* rowbias, int N, double* stats);
int fs_bn_act_apply(int device, fs_stream_t stream, const void* x, int dtype, int ld, long long rows,
                    int C, const float* rowbias, int N, const float* coef, float slope, void* out,
                    int out_dtype, int ld_out);
/* fs_bn_act_apply with fs_bn_finalize folded in (training): coefficients from `stats` (fs_colstats layout), published to
 * coef_out [4C] for the backward, running statistics / num_batches_tracked updated - one launch instead of two. */
int fs_bn_act_apply_fin(int device, fs_stream_t stream, const void* x, int dtype, int ld, long long rows, int C,
                        const float* rowbias, int N, const double* stats, double count, const float* gamma,
                        const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                        long long* num_batches_tracked, float* coef_out, float slope, void* out, int out_dtype,
                        int ld_out);
int fs_bn_act_bwd(int device, fs_stream_t stream, const void* g, int g_dtype, int ldg, const void* x,
                  int dtype, int ld, long long rows, int C, const float* rowbias, int N, const float* coef,
                  float slope, double* dgb, double count, int train_stats, void* dx, int dx_dtype,
                  int ld_dx);

/*
 * Global max-pool fused with BatchNorm statistics (global_feature = Conv1d + BN + LeakyReLU +
 * AdaptiveMaxPool1d, models/dgcnn.py:123-126, 156): LeakyReLU(BN(.)) is monotone, so only the per-cloud
 * max (gamma >= 0) / min (gamma < 0) of the GEMM output x [B*N, ld] is needed.
 *   fs_pool_reduce  sel [B,C] f32, arg [B,C] i32 (row within the cloud), stats as fs_colstats (nullable)
 *   fs_pool_bwd     dX[r,c] = scale*([r==arg[b,c]]*g[b,c]*LeakyReLU'(z_sel) - dbeta/M - xhat*dgamma/M)
 */
int fs_pool_reduce(int device, fs_stream_t stream, const void* x, int dtype, int ld, int B, int N, int C,
                   const float* gamma, float* sel, int32_t* arg, double* stats,
                   unsigned long long* packed_ws /* [B*C], zero-filled */);
int fs_pool_bwd(int device, fs_stream_t stream, const void* x, int dtype, int ld, int B, int N, int C,
                const float* g, const float* sel, const int32_t* arg, const float* coef, float slope,
                const double* dgb, double count, int train_stats, void* dx, int dx_dtype, int ld_dx);

/*
 * Backward of the pooled layer  out[b,c] = max_r LeakyReLU(BN(X W^T))[r,c]  (global_feature, models/dgcnn.py:123-126,
 * 156) WITHOUT the dense [B*N, C] gradient: dy[r,c] = S[r,c] + a_c + b_c y[r,c] with one non-zero of S per (cloud,
 * channel), hence
 *     dX = S W + 1 (a^T W) + X (W^T diag(b) W)          dW = S^T X + a colsum(X)^T + diag(b) W (X^T X)
 * The K x K products are library GEMMs on the caller's side; these entry points are the rest:
 *   fs_pool_lin_bwd_prep       a [C], bvec [C], sp [B,C] = scale * g * LeakyReLU'(z_sel) from g [B,C], sel [B,C],
 *                              coef [4C] (fs_bn_finalize layout) and dgb [2C] (double; fs_bn_act_bwd on the B x C
 *                              selected values)
 *   fs_pool_lin_bwd_dx_sparse  dx[b*N + arg[b,c], :] += sp[b,c] * w[c, :]   (w [C,K], dx [B*N,K], same dtype; C a power of
 *                              two <= 1024; every element receives ONE add of a sum formed in a fixed order:
 *                              deterministic; three launches: per-cloud sort, row sums, boundary merge)
 *   fs_pool_lin_bwd_dw         dw [C,K] f32 = S^T X + a colsum^T + diag(bvec) wg  (a / bvec / colsum / wg nullable:
 *                              eval-mode statistics have no dense part); K % 32 == 0, K <= 512
 */
/*
 * Global feature forward on tcgen05 (csrc/pool_gemm.cu): packed[b,c] = max over the N rows of cloud b of
 * (x w_signed^T)[., c] with its row, encoded (ordered key << 32) | ~row like fs_pool_reduce; the [B*N, C] product is never
 * written. x [B*N, K] bf16 (row pitch ldx elements, 16-byte aligned rows), w_signed [C, K] bf16 contiguous = sign(gamma) * W
 * (gamma >= 0 counts as +), packed [B*C] zero-filled. Supported: K in {64, 128, 192}, C % 128 == 0 (fs_pool_gemm_supported).
 *   fs_pool_decode            sel [B,C] f32 (sign restored), arg [B,C] i32 (row within the cloud, -1: no finite value)
 *   fs_pool_stats_from_gram   BatchNorm sums of the product without the product: stats[0:C] = W colsum(X),
 *                             stats[C:2C] = rowwise (W G) . W, G = X^T X (fs_colstats layout, pivot 0); wg = W G [C,K] f32
 */
int fs_pool_gemm_supported(int B, int N, int C, int K);
int fs_pool_gemm(int device, fs_stream_t stream, const void* x, int ldx, const void* w_signed, int B, int N, int C, int K,
                 unsigned long long* packed);
int fs_pool_decode(int device, fs_stream_t stream, const unsigned long long* packed, const float* gamma, int B, int C,
                   float* sel, int32_t* arg);
int fs_pool_stats_from_gram(int device, fs_stream_t stream, const void* w, int dtype, int ldw, const float* wg,
                            const float* colsum, int C, int K, double* stats);

/* Column sums of x [rows, K] (fp32 or bf16; K % 8 == 0 (bf16) / % 4 (fp32), 16-byte aligned rows) -> out [K] f32, through
 * fs_colsum_partials() x K fp32 partials in partial_ws (fixed chunking: deterministic). */
int fs_colsum_partials(void);
int fs_colsum(int device, fs_stream_t stream, const void* x, int dtype, int ld, long long rows, int K,
              float* partial_ws, float* out);
int fs_pool_lin_bwd_prep(int device, fs_stream_t stream, const float* g, const float* sel, const float* coef,
                         float slope, const double* dgb, double count, int train_stats, int B, int C, float* a,
                         float* bvec, float* sp);
size_t fs_pool_lin_bwd_ws_bytes(int B, int C, int K);
int fs_pool_lin_bwd_dx_sparse(int device, fs_stream_t stream, const float* sp, const int32_t* arg, const void* w,
                              int dtype, int ldw, int B, int N, int C, int K, void* dx, int ld_dx,
                              void* ws /* fs_pool_lin_bwd_ws_bytes(B, C, K) bytes, 16-byte aligned */);
int fs_pool_lin_bwd_dw(int device, fs_stream_t stream, const float* sp, const int32_t* arg, const void* x, int dtype,
                       int ldx, int B, int N, int C, int K, const float* a, const float* bvec, const float* colsum,
                       const float* wg, float* dw);

/*
 * Feature table of the heads: torch.cat((x1, x2, x3), dim=1) (models/dgcnn.py:154, 200) of n <= 4 fp32 point-major
 * tables [rows, widths[i]] (row strides lds[i], widths % 4 == 0) written once in the compute dtype, and the reverse for
 * the gradient (g [rows, sum widths] -> n contiguous fp32 tables). srcs / dsts / widths / lds are HOST arrays of n.
 */
int fs_cat_cast(int device, fs_stream_t stream, int n, const void* const* srcs, const int* widths, const int* lds,
                long long rows, void* out, int out_dtype, int ld_out);
int fs_split_cast(int device, fs_stream_t stream, int n, void* const* dsts, const int* widths, const int* lds,
                  long long rows, const void* g, int g_dtype, int ld_g);

/*
 * Gradient bucket -> flat buffer (ddp.py): dsts[i][0..counts[i]) = srcs[i][...] for n fp32 tensors in one launch per 32
 * tensors; srcs / dsts / counts are HOST arrays (the table is passed to the kernel by value).
 */
int fs_multi_copy_f32(int device, fs_stream_t stream, int n, const void* const* srcs, void* const* dsts,
                      const long long* counts);

/*
 * Network output: point-major logits [B*N, C] (fp32 / bf16, row pitch ld) of the internally re-ordered cloud ->
 * out [B, C, N] fp32 in the caller's point order, out[b, c, perm[b,n]] = logits[b*N + n, c] (perm [B,N] int64, nullable =
 * identity): the `B x classes x N` tensor the reference's forward returns (models/dgcnn.py:162). _bwd is the adjoint.
 */
int fs_logits_out(int device, fs_stream_t stream, const void* logits, int dtype, int ld, const long long* perm, int B, int N,
                  int C, float* out);
int fs_logits_out_bwd(int device, fs_stream_t stream, const float* g, const long long* perm, int B, int N, int C,
                      void* dlogits, int dtype, int ld);

/*
 * Last layer of the segmentation head fused with the network output (models/dgcnn.py:137, 160-162): 1x1 conv to
 * C_out = num_classes (2, 4 or 8) channels with bias, no BatchNorm; h [B*N, C_in] (C_in 128, or 256 with C_out <= 4; fp32 / bf16), w [C_out,
 * C_in] f32, out [B, C_out, N] f32 in the caller's point order (perm as fs_logits_out). Backward: dh [B*N, C_in] in h's dtype,
 * dw_db [C_out*C_in + C_out] f32 = dW [C_out, C_in] followed by dbias [C_out]; ws = fs_final_linear_ws_floats(B*N, ...)
 * floats of scratch (per-block partials, summed in block order: deterministic).
 */
int fs_final_linear_supported(int C_in, int C_out);
size_t fs_final_linear_ws_floats(long long rows, int C_in, int C_out);
int fs_final_linear_fwd(int device, fs_stream_t stream, const void* h, int dtype, int ld, const float* w, const float* bias,
                        const long long* perm, int B, int N, int C_in, int C_out, float* out);
int fs_final_linear_bwd(int device, fs_stream_t stream, const void* h, int dtype, int ld, const float* w, const float* g,
                        const long long* perm, int B, int N, int C_in, int C_out, void* dh, int ld_dh, float* ws,
                        float* dw_db);

/*
 * EdgeConv weight table: out [2Cp, C] f32 = [W1 ; W2 - W1] from the conv weight w [Cp, 2C] f32 = [W1 | W2] (column order of
 * models/dgcnn.py:36: neighbour difference first, centre second) - the right-hand side of the per-point table GEMM - and
 * the adjoint dw [Cp, 2C] from g [2Cp, C].
 */
int fs_edge_weight_table(int device, fs_stream_t stream, const float* w, int Cp, int C, float* out);
int fs_edge_weight_table_bwd(int device, fs_stream_t stream, const float* g, int Cp, int C, float* dw);

/* out [n] f32 = sum over s of part [S, n] f32 (fixed order): the reduction of the partial products of a row-chunked
 * weight-gradient GEMM (dW = dY^T X issued as a batched GEMM over row chunks). */
int fs_sum_leading(int device, fs_stream_t stream, const float* part, int S, long long n, float* out);

/* ---------------------------------------------------------------- Chamfer ------------------ */

/*
 * Nearest neighbour of every x_i among y (squared L2, direct differences), per batch element.
 * Replaces pytorch3d.ops.knn_points(K=1) inside chamfer_distance (losses/chamfer_loss.py:19).
 *   x [B,N,3], y [B,M,3] contiguous fp32; nn_d2 [B,N], nn_idx [B,N].
 */
int fs_nn_points(int device, fs_stream_t stream, const float* x, const float* y, int B, int N,
                 int M, float* nn_d2, int32_t* nn_idx);

/*
 * Chamfer gradient: gx_i += w * 2 (x_i - y_nn(i)), gy_nn(i) -= w * 2 (x_i - y_nn(i)) (atomics).
 * gx / gy must be zero-initialised by the caller (gy nullable). w_ptr: device scalar upstream grad.
 */
int fs_chamfer_bwd(int device, fs_stream_t stream, const float* x, const float* y,
                   const int32_t* nn_idx, int B, int N, int M, float weight, const float* w_ptr,
                   float* gx, float* gy);

/* ---------------------------------------------------------------- pointops (gather family) - */

/* grouping: out[m,s,:] = in[idx[m,s],:]  (pointops.py:78) and its backward (:94). */
int fs_grouping_fwd(int device, fs_stream_t stream, int m, int nsample, int c, const float* in,
                    const int32_t* idx, float* out);
int fs_grouping_bwd(int device, fs_stream_t stream, int m, int nsample, int c,
                    const float* grad_out, const int32_t* idx, float* grad_in);
/* interpolation: out[n,:] = sum_k w[n,k] * in[idx[n,k],:]  (pointops.py:236) and backward (:252). */
int fs_interpolation_fwd(int device, fs_stream_t stream, int n, int c, int k, const float* in,
                         const int32_t* idx, const float* weight, float* out);
int fs_interpolation_bwd(int device, fs_stream_t stream, int n, int c, int k,
                         const float* grad_out, const int32_t* idx, const float* weight,
                         float* grad_in);
/* subtraction: out[n,s,:] = in1[n,:] - in2[idx[n,s],:]  (pointops.py:139) and backward (:155). */
int fs_subtraction_fwd(int device, fs_stream_t stream, int n, int nsample, int c, const float* in1,
                       const float* in2, const int32_t* idx, float* out);
int fs_subtraction_bwd(int device, fs_stream_t stream, int n, int nsample, int c,
                       const int32_t* idx, const float* grad_out, float* grad_in1, float* grad_in2);
/* aggregation: out[n,c] = sum_s (in[idx[n,s],c] + pos[n,s,c]) * w[n,s,c % w_c]  (pointops.py:174, :192). */
int fs_aggregation_fwd(int device, fs_stream_t stream, int n, int nsample, int c, int w_c,
                       const float* in, const float* pos, const float* weight, const int32_t* idx,
                       float* out);
int fs_aggregation_bwd(int device, fs_stream_t stream, int n, int nsample, int c, int w_c,
                       const float* in, const float* pos, const float* weight, const int32_t* idx,
                       const float* grad_out, float* grad_in, float* grad_pos, float* grad_weight);

/* ---------------------------------------------------------------- optimiser ---------------- */

/*
 * Fused Adam step over a flat parameter buffer (torch.optim.Adam semantics with L2 weight decay,
 * model_trainer.py:57). `grad_scale` multiplies the gradient (1/world_size averaging); `step` is the 1-based step count. `dyn_step_lr` (nullable) is a device
 * array [step, lr] that overrides the host values, so a captured CUDA graph can be replayed.
 */
int fs_adam_step(int device, fs_stream_t stream, float* param, const float* grad, float* exp_avg,
                 float* exp_avg_sq, long long n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int step, float grad_scale, const float* dyn_step_lr);

/*
 * Data-parallel tail: all-reduce (sum over ranks) + Adam in one kernel over PEER memory. peer_grad_ptrs is a DEVICE array
 * of `world` base pointers of the ranks' symmetric gradient buffers (mapped into this process, e.g.
 * torch.distributed._symmetric_memory: handle.buffer_ptrs_dev); elements [elem_offset, elem_offset + n) of every buffer are
 * summed in rank order (bit-identical on all ranks) and applied to param / exp_avg / exp_avg_sq (local, length n).
 * grad_sum_out (nullable, local, length n) receives the sum. The caller provides the barriers around the launch.
 */
int fs_adam_step_peers(int device, fs_stream_t stream, float* param, const unsigned long long* peer_grad_ptrs, int world,
                       long long elem_offset, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
                       float beta2, float eps, float weight_decay, int step, float grad_scale,
                       const float* dyn_step_lr, float* grad_sum_out);

/* ---------------------------------------------------------------- fused two-layer EdgeConv --- */

/*
 * Two-layer EdgeConv on raw coordinates, forward (models/dgcnn.py:119 `ec1 = EdgeConv(in_features, [64, 64])` with the
 * 3-channel coordinate input; loop at :237-241): neither the hidden edge tensor H (P*k x 64) nor the layer-2
 * pre-activation Z is written. Layer 1 (Conv 6->64 + BatchNorm + LeakyReLU) is recomputed from the coordinates with
 * the coefficients of fs_edge3_bn_coef; layer 2 (64 -> C2) runs as tcgen05.mma on bf16 tiles of H staged in shared
 * memory, transposed so that a TMEM lane is an output channel and the max over the k edges of a point is a
 * per-thread reduction.
 *   w1 [64,6], coef1 [4*64] (mu | 1/sigma | gamma/sigma | beta), w2 [C2,64] fp32, gamma2 [C2] (sign selects max / min)
 *   sel [P,C2] fp32 selected pre-activation, arg [P,C2] uint8 slot of the selected edge
 *   stats (nullable = eval): fs_stats_buffer_doubles(C2) doubles: receives sum z, sum z^2 over all P*k edges, derived
 *   from gram [64*64] = sum_e h_e h_e^T (accumulated on the tensor cores) and hsum [64] = sum_e h_e (both required with
 *   stats, zero-filled), which the BatchNorm coupling of the backward pass needs as well
 * fs_edge2_supported: C1 == 64, C2 == 64, k in {8, 12, 16, 20, 40}.
 */
int fs_edge2_supported(int k, int C1, int C2);
int fs_edge2_fwd(int device, fs_stream_t stream, const float* x, int ldx, const int32_t* idx, int B, int N, int k,
                 const float* w1, const float* coef1, const float* w2, int C2, const float* gamma2, float* sel,
                 uint8_t* arg, double* stats, float* gram, float* hsum);
/*
 * Backward of fs_edge2_fwd: dz_e = R_e + alpha + beta' (.) z_e (routed gradient + BatchNorm-2 coupling) gives
 *   dh_e = [W2^T | Gm] [R_e ; h_e] + a0     (one tcgen05 product per tile, Gm = W2^T diag(beta') W2, a0 = W2^T alpha)
 * with H recomputed from the coordinates; the epilogue applies LeakyReLU' of layer 1 and accumulates the sums
 * fs_edge3_bwd takes from a materialised dH; sum_e R_e h_e^T accumulates on the tensor cores (MN-major operands).
 * Three launches: coupling coefficients (one CTA), the tile kernel, dW2 = R^T H + alpha hsum^T + diag(beta') W2 S.
 *   coef2 [4*64] BatchNorm-2 coefficients (mu | 1/sigma | gamma/sigma | beta), dgb2 = statistics buffer of
 *   fs_edgeconv_bwd_reduce (sum d | sum d zhat), train_stats = 1 in training mode, gram / hsum from the forward,
 *   d2 [P,64] = upstream gradient times LeakyReLU' of layer 2, arg [P,64]
 *   scratch: fs_edge2_bwd_scratch_floats() floats, zero-filled
 *   outputs: dgb1 (fs_stats_buffer_doubles(64), zero-filled: sum t | sum t * yhat of layer 1), acc1 [64,6] zero-filled
 *   (sum t (x) e, input of fs_edge3_dw), dw2 [64,64]
 */
size_t fs_edge2_bwd_scratch_floats(void);
int fs_edge2_bwd(int device, fs_stream_t stream, const float* x, int ldx, const int32_t* idx, int B, int N, int k,
                 const float* w1, const float* coef1, const float* w2, int C2, const float* coef2, const double* dgb2,
                 int train_stats, const float* gram, const float* hsum, const float* d2, const uint8_t* arg,
                 float* scratch, double* dgb1, float* acc1, float* dw2);
/* dW1 [Cp,6] from the sums of a layer-1 backward pass (the closed-form BatchNorm coupling of fs_edge3_bwd). */
int fs_edge3_dw(int device, fs_stream_t stream, const float* acc, const double* dgb, const double* moments,
                double count, const float* w, const float* coef, int Cp, int train_stats, float* dw);

/* ---------------------------------------------------------------- ensemble inference ------- */

/*
 * acc[c][sub[r][s]] += softmax_c(logits[r][:, s]) for all R runs of
 * models/point_seg_net.py:27-29 and :43 (`softmax_accumulation[..., perm] += output_activation(self(pc[..., perm]))`)
 * in one launch, the R subset forwards having been batched into one B = R forward.
 *   logits [R, classes, S]  fp32, the network output (B x classes x N layout)
 *   sub    [R, S]           int64 point indices into the full cloud (unique within a run)
 *   acc    [classes, n_total] fp32, accumulated in place (fp32 atomics: order over the runs is not fixed)
 * classes <= 32.
 */
int fs_softmax_scatter_add(int device, fs_stream_t stream, const float* logits, const long long* sub, int R,
                           int classes, int S, int n_total, float* acc);

/* ---------------------------------------------------------------- measurement -------------- */

/*
 * FP32 FMA throughput micro-benchmark (one launch of 148 x 8 CTAs x 256 threads x iters x 64 dependent-chain FMAs,
 * ILP 8): the measured roofline denominator for the CUDA-core kernels (fs_knn3d, fs_nn_points). `out` is a device
 * scalar that is never written in practice; *flops_out (HOST pointer, nullable) receives the flops of the launch.
 */
int fs_fma_microbench(int device, fs_stream_t stream, int iters, float* out, double* flops_out);

#ifdef __cplusplus
}
#endif
#endif /* FISSURE_B200_H */
