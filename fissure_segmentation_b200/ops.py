"""Host-side operators over the C ABI: kNN graph build, fused EdgeConv (autograd), nearest-neighbour
reduction for Chamfer. Tensors are torch CUDA tensors; the arithmetic runs in libfissure_b200.so.

Layout convention: "point-major" tables of shape (P, C) with P = B*N rows (row = b*N + n) and unit
stride along C. The reference's (B, C, N) tensors are converted at the module boundary only.
"""
import ctypes
import os

import torch

from . import _lib

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
# feature-space kNN (C in {64, 128, 256}): tcgen05 candidates + exact re-rank. FS_KNN_TC=0 selects the exact SIMT kernels
# (development switch: both are CUDA paths of this library)
USE_TENSOR_CORE_KNN = os.environ.get("FS_KNN_TC", "1") != "0"


class KnnGraph:
    """A kNN graph of B clouds with N points: int32 neighbour indices (B, N, k), local to the cloud,
    plus the lazily built reverse (incoming-edge) graph used by the EdgeConv backward."""

    __slots__ = ("idx", "B", "N", "k", "_rev")

    def __init__(self, idx):
        assert idx.dtype == torch.int32 and idx.dim() == 3 and idx.is_contiguous()
        self.idx = idx
        self.B, self.N, self.k = idx.shape
        self._rev = None

    @classmethod
    def from_reference(cls, idx64, validate=True):
        """Accept the reference's (B, N, k) int64 graph (models/dgcnn.py:28-29 fixed_knn_graph). The kernels index
        the tables with these values unchecked, so a user-supplied graph is range-checked here (one min/max
        reduction; pass validate=False for graphs produced by this package)."""
        if idx64.dim() != 3 or idx64.dtype not in (torch.int64, torch.int32):
            raise ValueError("fixed_knn_graph must be an integer (B, N, k) tensor, got %s %s"
                             % (idx64.dtype, tuple(idx64.shape)))
        if validate and idx64.numel():
            lo, hi = torch.aminmax(idx64)
            if int(lo) < 0 or int(hi) >= idx64.shape[1]:
                raise IndexError("fixed_knn_graph holds indices outside [0, N=%d): min %d, max %d"
                                 % (idx64.shape[1], int(lo), int(hi)))
        return cls(idx64.to(torch.int32).contiguous())

    def reverse(self):
        if self._rev is None:
            P = self.B * self.N
            rev_ptr = torch.empty(P + 1, dtype=torch.int32, device=self.idx.device)
            rev_src = torch.empty(P * self.k, dtype=torch.int32, device=self.idx.device)
            _lib.call("fs_reverse_graph", self.idx, self.idx, self.B, self.N, self.k, rev_ptr, rev_src)
            self._rev = (rev_ptr, rev_src)
        return self._rev


# ------------------------------------------------------------------------------------------- kNN

def _check_k(k, self_loop, N):
    kk = k + (0 if self_loop else 1)
    if kk > N:
        # torch.topk raises the same way in the reference (general_utils.py:320)
        raise RuntimeError("selected index k out of range: k=%d (+%d) > N=%d" % (k, kk - k, N))
    if kk > _lib.FS_MAX_K + 1:
        raise RuntimeError("k=%d exceeds the supported maximum of %d" % (k, _lib.FS_MAX_K))


USE_TENSOR_CORE_KNN3D = os.environ.get("FS_KNN3D_TC", os.environ.get("FS_KNN_TC", "1")) != "0"  # coordinate kNN through the same kernels
_workspaces = {}              # (device, stream, tag) -> uint8 scratch kept across calls (no 50-130 MB allocation per graph build)


def _workspace(nbytes, device, tag):
    """Scratch buffer of at least nbytes, reused across calls on the same stream (the kernels of one call finish
    before the next call on that stream starts). Inside a CUDA-graph capture the buffer comes from the graph's pool
    and is not cached."""
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if torch.cuda.is_current_stream_capturing():
        return torch.empty(nbytes, dtype=torch.uint8, device=device)
    key = (device, torch.cuda.current_stream(device).cuda_stream, tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _report_tc(lib, ws, B, N, C, k):
    if knn_tc_report is not None:
        off = lib.fs_knn_feat_tc_redo_offset(B, N, C, k)
        knn_tc_report["rows"] = knn_tc_report.get("rows", 0) + B * N
        knn_tc_report["redo_rows"] = knn_tc_report.get("redo_rows", 0) + int(ws[off:off + B * N].sum())
        knn_tc_report["calls"] = knn_tc_report.get("calls", 0) + 1
        knn_tc_report.setdefault("channels", []).append(C)


def knn_coords(x, k, self_loop=False, diag_zero=True, return_dist=False):
    """kNN on the first three channels of x (B, C>=3, N), any strides. Returns int32 (B, N, k)."""
    B, _, N = x.shape
    _check_k(k, self_loop, N)
    if x.dtype != torch.float32:
        x = x.float()
    idx = torch.empty(B, N, k, dtype=torch.int32, device=x.device)
    dist = torch.empty(B, N, k, dtype=torch.float32, device=x.device) if return_dist else None
    if B == 0:
        return (idx, dist) if return_dist else idx
    lib = _lib.load()
    if USE_TENSOR_CORE_KNN3D and not return_dist and x.is_cuda and lib.fs_knn3d_tc_supported(B, N, k, int(self_loop)):
        # tcgen05 distances + SIMT selection, exact re-rank and exact ordering (same result as fs_knn3d)
        nbytes = lib.fs_knn3d_tc_workspace_bytes(B, N, k)
        ws = _workspace(nbytes, x.device, "knn")
        _lib.call("fs_knn3d_tc", x, x, x.stride(0), x.stride(1), x.stride(2), B, N, k, int(self_loop), int(diag_zero),
                  idx, ws, nbytes)
        _report_tc(lib, ws, B, N, 3, k)
        return idx
    _lib.call("fs_knn3d", x, x, x.stride(0), x.stride(1), x.stride(2), B, N, k, int(self_loop), int(diag_zero),
              idx, dist)
    return (idx, dist) if return_dist else idx


knn_tc_report = None      # set to a dict to collect {"rows", "redo_rows", "calls"} of every tensor-core kNN call (tests; syncs)


def knn_features(feat, B, N, k, self_loop=False, diag_zero=True, return_dist=False):
    """Exact FP32 kNN on a point-major feature table (B*N, C) (row stride arbitrary)."""
    _check_k(k, self_loop, N)
    if feat.dtype != torch.float32:
        feat = feat.float()
    if feat.stride(1) != 1:
        feat = feat.contiguous()
    C = feat.shape[1]
    idx = torch.empty(B, N, k, dtype=torch.int32, device=feat.device)
    dist = torch.empty(B, N, k, dtype=torch.float32, device=feat.device) if return_dist else None
    if B == 0:
        return (idx, dist) if return_dist else idx
    lib = _lib.load()
    if (USE_TENSOR_CORE_KNN and not return_dist and feat.stride(0) % 4 == 0 and feat.data_ptr() % 16 == 0
            and lib.fs_knn_feat_tc_supported(B, N, C, k, int(self_loop))):
        # tcgen05 candidate search + exact FP32 re-rank (same result as the exact kernel)
        nbytes = lib.fs_knn_feat_tc_workspace_bytes(B, N, C, k)
        ws = _workspace(nbytes, feat.device, "knn")
        _lib.call("fs_knn_feat_tc", feat, feat, feat.stride(0), B, N, C, k, int(self_loop), int(diag_zero), idx, dist,
                  ws, nbytes)
        _report_tc(lib, ws, B, N, C, k)
        return (idx, dist) if return_dist else idx
    ws = torch.empty(B * N, dtype=torch.float32, device=feat.device)
    _lib.call("fs_knn_feat", feat, feat, feat.stride(0), B, N, C, k, int(self_loop), int(diag_zero), idx, dist, ws)
    return (idx, dist) if return_dist else idx


def spatial_order(x):
    """Permutation (B, N) int64 that sorts every cloud of x (B, C>=3, N) along the Morton curve of its xyz."""
    B, _, N = x.shape
    codes = torch.empty(B, N, dtype=torch.int32, device=x.device)
    xf = x if x.dtype == torch.float32 else x.float()
    _lib.call("fs_morton_codes", xf, xf, xf.stride(0), xf.stride(1), xf.stride(2), B, N, codes)
    return torch.sort(codes, dim=1, stable=True)[1]


def to_point_major(x):
    """(B, C, N) -> contiguous (B*N, C)."""
    B, C, N = x.shape
    return x.transpose(1, 2).reshape(B * N, C)


def knn_any(x, k, self_loop=False, diag_zero=True, return_dist=False):
    """kNN over all channels of x (B, C, N): 3-D kernel for C == 3, feature kernel otherwise."""
    B, C, N = x.shape
    if C == 3:
        return knn_coords(x, k, self_loop, diag_zero, return_dist)
    return knn_features(to_point_major(x.float()), B, N, k, self_loop, diag_zero, return_dist)


# ------------------------------------------------------------------------------------------- EdgeConv

class _ZeroArena:
    """Zero-filled fp64 scratch shared by all per-layer statistics buffers of one training step: ONE memset per
    step instead of one fill kernel per BatchNorm layer and direction (16 launches in DGCNNSeg). The networks
    call `begin_step` at the top of forward(); the capacity follows the demand of the previous step, anything
    beyond it (first step, stand-alone layers) falls back to an individual torch.zeros."""

    def __init__(self):
        self.buf, self.off, self.cap, self.demand = None, 0, 0, 0
        self.capturing = False       # the buffer belongs to a CUDA graph's private pool

    def begin_step(self, device):
        self.cap = max(self.cap, self.demand)
        self.demand, self.off = 0, 0
        self.capturing = device.type == "cuda" and torch.cuda.is_current_stream_capturing()
        self.buf = torch.zeros(self.cap, dtype=torch.float64, device=device) if self.cap else None

    def take(self, n, device):
        n = (n + 1) & ~1                      # keep 16-byte alignment of every slice
        self.demand += n
        if self.buf is not None and self.buf.device == device and self.off + n <= self.cap:
            out = self.buf[self.off:self.off + n]
            self.off += n
            return out
        return torch.zeros(n, dtype=torch.float64, device=device)


_arenas = {}          # (device, stream id) -> arena of the step that is running on that stream


def _arena_key(device):
    device = torch.device(device)
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device, torch.cuda.current_stream(device).cuda_stream


def begin_step(device, owner=None):
    """Start a new zero-filled statistics arena for the step that runs on the CURRENT stream of `device` (called by
    the networks' forward). Arenas are keyed by (device, stream), so two models driven from different streams never
    share a memset; `owner` (the model) keeps the capacity estimate per model."""
    device = torch.device(device)
    if device.type != "cuda":
        return None
    key = _arena_key(device)
    arena = getattr(owner, "_fs_arena", None) if owner is not None else _arenas.get(key)
    if arena is None:
        arena = _ZeroArena()
        if owner is not None:
            owner._fs_arena = arena
    arena.begin_step(key[0])
    _arenas[key] = arena
    return arena


def end_step(device):
    """Detach the arena of the current stream: stand-alone layers (and anything that runs after a CUDA-graph
    capture, whose arena belongs to the graph's private pool) fall back to individual torch.zeros buffers."""
    device = torch.device(device)
    if device.type != "cuda":
        return
    arena = _arenas.pop(_arena_key(device), None)
    if arena is not None:
        arena.buf = None


def _zeros64(n, device):
    """n zero-filled 8-byte words (fp64 view) from the step arena of the current stream."""
    device = torch.device(device)
    arena = _arenas.get(_arena_key(device)) if device.type == "cuda" else None
    if arena is not None and arena.buf is not None and arena.capturing != torch.cuda.is_current_stream_capturing():
        # a buffer handed out during a capture lives in the graph's pool and is memset on every replay: never
        # slice it from eager code (and the other way round)
        arena = None
    if arena is None or arena.buf is None:
        if arena is not None:
            arena.demand += (n + 1) & ~1
        return torch.zeros(n, dtype=torch.float64, device=device)
    return arena.take(n, arena.buf.device)


def _stats_buffer(C, device):
    """Zero-filled fp64 statistics buffer (final sums | pivot | slot partials | ticket) for C channels."""
    return _zeros64(_lib.load().fs_stats_buffer_doubles(C), device)


def _bn_coef(table_ref, stats, count, gamma, beta, running_mean, running_var, nbt, training, Cp, eps, momentum):
    coef = torch.empty(4 * Cp, dtype=torch.float32, device=table_ref.device)
    if training:
        _lib.call("fs_bn_finalize", table_ref, stats, float(count), Cp, gamma, beta, eps, momentum, coef,
                  running_mean, running_var, nbt)
    else:
        _lib.call("fs_bn_coef_eval", table_ref, Cp, gamma, beta, running_mean, running_var, eps, coef)
    return coef


def _bn_then_edgeconv_apply(ref, sel, table, dt, ld, P, Cp, stats, count, gamma, beta, running_mean, running_var, nbt,
                            training, eps, momentum, out):
    """BatchNorm coefficients + fs_edgeconv_apply. Training: ONE launch (the finalisation of the batch statistics is
    folded into the apply kernel, which also publishes the coefficients for the backward); eval: running statistics."""
    if training:
        coef = torch.empty(4 * Cp, dtype=torch.float32, device=ref.device)
        _lib.call("fs_edgeconv_apply_fin", ref, sel, table, dt, ld, P, Cp, stats, float(count), gamma, beta, eps, momentum,
                  running_mean, running_var, nbt, coef, out, _lib.dtype_code(out), out.stride(0))
        return coef
    coef = _bn_coef(ref, None, 1, gamma, beta, running_mean, running_var, nbt, False, Cp, eps, momentum)
    _lib.call("fs_edgeconv_apply", ref, sel, table, dt, ld, P, Cp, coef, out, _lib.dtype_code(out), out.stride(0))
    return coef


def _bn_then_act_apply(ref, src, src_dt, ld, rows, C, rowbias, N, stats, count, gamma, beta, running_mean, running_var, nbt,
                       training, eps, momentum, slope, out):
    """The same for fs_bn_act_apply (dense heads, pooled values)."""
    if training:
        coef = torch.empty(4 * C, dtype=torch.float32, device=ref.device)
        _lib.call("fs_bn_act_apply_fin", ref, src, src_dt, ld, rows, C, rowbias, N, stats, float(count), gamma, beta, eps,
                  momentum, running_mean, running_var, nbt, coef, float(slope), out, _lib.dtype_code(out), out.stride(0))
        return coef
    coef = _bn_coef(ref, None, 1, gamma, beta, running_mean, running_var, nbt, False, C, eps, momentum)
    _lib.call("fs_bn_act_apply", ref, src, src_dt, ld, rows, C, rowbias, N, coef, float(slope), out, _lib.dtype_code(out),
              out.stride(0))
    return coef


class _EdgeConvFn(torch.autograd.Function):
    """Single-layer EdgeConv on the per-point table T = [a | b]:
    out_i = LeakyReLU(BN(max_j (a_j + b_i))), models/dgcnn.py:226-243 with a one-layer shared MLP."""

    @staticmethod
    def forward(ctx, table, gamma, beta, graph, running_mean, running_var, nbt, training, eps, momentum):
        idx = graph.idx
        B, N, k = graph.B, graph.N, graph.k
        P, Cp = B * N, table.shape[1] // 2
        dev = table.device
        dt = _lib.dtype_code(table)
        gamma32, beta32 = gamma.detach().float(), beta.detach().float()
        sel = torch.empty(P, Cp, dtype=torch.float32, device=dev)
        arg = torch.empty(P, Cp, dtype=torch.uint8, device=dev)
        need_grad = any(ctx.needs_input_grad[:3])
        sy = torch.empty(P, Cp, dtype=torch.float32, device=dev) if (training and need_grad) else None
        stats = _stats_buffer(Cp, dev) if training else None
        rev_ptr = graph.reverse()[0] if training else None      # in-degrees for the batch statistics
        _lib.call("fs_edgeconv_gather", table, table, dt, table.stride(0), idx, B, N, k, Cp, gamma32, rev_ptr, sel, arg,
                  sy, stats)
        out = torch.empty(P, Cp, dtype=table.dtype, device=dev)
        coef = _bn_then_edgeconv_apply(table, sel, table, dt, table.stride(0), P, Cp, stats, P * k, gamma32, beta32,
                                       running_mean, running_var, nbt, training, eps, momentum, out)
        ctx.graph = graph
        ctx.training = training
        ctx.save_for_backward(table, sel, arg, sy, coef)
        return out

    @staticmethod
    def backward(ctx, g):
        table, sel, arg, sy, coef = ctx.saved_tensors
        graph = ctx.graph
        B, N, k = graph.B, graph.N, graph.k
        P, Cp = B * N, table.shape[1] // 2
        dev = table.device
        dt = _lib.dtype_code(table)
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        if g.stride(1) != 1:
            g = g.contiguous()
        d = torch.empty(P, Cp, dtype=torch.float32, device=dev)
        dgb = _stats_buffer(Cp, dev)
        _lib.call("fs_edgeconv_bwd_reduce", table, g, _lib.dtype_code(g), g.stride(0), sel, table, dt, table.stride(0),
                  P, Cp, coef, d, dgb)
        dT = torch.empty(P, 2 * Cp, dtype=torch.float32, device=dev)
        dgam_dbet = torch.empty(2 * Cp, dtype=torch.float32, device=dev)
        if ctx.training:
            rev_ptr, rev_src = graph.reverse()
        else:
            rev_ptr = rev_src = None
        _lib.call("fs_edgeconv_bwd_point", table, d, sy, table, dt, table.stride(0), rev_ptr, rev_src, P, k, Cp, coef,
                  dgb, float(P * k), int(ctx.training), dT, dgam_dbet)
        _lib.call("fs_edgeconv_bwd_route", table, d, arg, graph.idx, B, N, k, Cp, coef, dT)
        if dT.dtype != table.dtype:
            dT = dT.to(table.dtype)
        return dT, dgam_dbet[:Cp], dgam_dbet[Cp:], None, None, None, None, None, None, None


def edgeconv_fused(table, gamma, beta, graph, running_mean, running_var, nbt, training, out=None,
                   eps=BN_EPS, momentum=BN_MOMENTUM):
    """table (P, 2*Cp) fp32/bf16 contiguous rows; returns (P, Cp) in table.dtype (or writes `out`)."""
    assert table.stride(1) == 1
    if not torch.is_grad_enabled() and not training:
        # inference: single fused pass, nothing saved
        B, N, k = graph.B, graph.N, graph.k
        P, Cp = B * N, table.shape[1] // 2
        coef = _bn_coef(table, None, 1, gamma.detach().float(), beta.detach().float(), running_mean, running_var,
                        None, False, Cp, eps, momentum)
        if out is None:
            out = torch.empty(P, Cp, dtype=table.dtype, device=table.device)
        _lib.call("fs_edgeconv_fused_eval", table, table, _lib.dtype_code(table), table.stride(0), graph.idx, B, N, k,
                  Cp, coef, out, _lib.dtype_code(out), out.stride(0), None)
        return out
    return _EdgeConvFn.apply(table, gamma, beta, graph, running_mean, running_var, nbt, training, eps, momentum)


class _EdgeBuildFn(torch.autograd.Function):
    """Y[(i,t)] = a[idx[i,t]] + b[i]: first layer of a two-layer EdgeConv as a materialised edge tensor."""

    @staticmethod
    def forward(ctx, table, graph, out_dtype):
        B, N, k = graph.B, graph.N, graph.k
        P, Cp = B * N, table.shape[1] // 2
        y = torch.empty(P * k, Cp, dtype=out_dtype, device=table.device)
        _lib.call("fs_edge_build", table, table, _lib.dtype_code(table), table.stride(0), graph.idx, B, N, k, Cp, y,
                  _lib.dtype_code(y))
        ctx.graph = graph
        ctx.table_dtype = table.dtype
        ctx.Cp = Cp
        return y

    @staticmethod
    def backward(ctx, dy):
        graph = ctx.graph
        B, N, k = graph.B, graph.N, graph.k
        if dy.dtype not in (torch.float32, torch.bfloat16):
            dy = dy.float()
        dy = dy.contiguous()
        dT = torch.zeros(B * N, 2 * ctx.Cp, dtype=torch.float32, device=dy.device)
        _lib.call("fs_edge_build_bwd", dy, dy, _lib.dtype_code(dy), graph.idx, B, N, k, ctx.Cp, dT)
        return dT.to(ctx.table_dtype), None, None


def edge_build(table, graph, out_dtype=None):
    return _EdgeBuildFn.apply(table, graph, out_dtype or table.dtype)


class _EdgeReduceFn(torch.autograd.Function):
    """out_i = LeakyReLU(BN(max_t Z[(i,t)])) on a materialised edge tensor Z (P*k, Cp): BatchNorm2d batch
    statistics over all edges, LeakyReLU(0.2) and max over k (models/dgcnn.py:237-241) in one reduction."""

    @staticmethod
    def forward(ctx, z, gamma, beta, k, running_mean, running_var, nbt, training, eps, momentum):
        P, Cp = z.shape[0] // k, z.shape[1]
        dev = z.device
        gamma32, beta32 = gamma.detach().float(), beta.detach().float()
        sel = torch.empty(P, Cp, dtype=torch.float32, device=dev)
        arg = torch.empty(P, Cp, dtype=torch.uint8, device=dev)
        stats = _stats_buffer(Cp, dev) if training else None
        _lib.call("fs_edge_reduce", z, z, _lib.dtype_code(z), P, k, Cp, gamma32, sel, arg, None, stats)
        coef = _bn_coef(z, stats, P * k, gamma32, beta32, running_mean, running_var, nbt, training, Cp, eps, momentum)
        # fp32 output in every precision mode: it is the next layer's kNN / table input (no bf16 round trip, no cast)
        out = torch.empty(P, Cp, dtype=torch.float32, device=dev)
        _lib.call("fs_edgeconv_apply", z, sel, None, 0, 0, P, Cp, coef, out, _lib.dtype_code(out), out.stride(0))
        ctx.k = k
        ctx.training = training
        ctx.save_for_backward(z, sel, arg, coef)
        return out

    @staticmethod
    def backward(ctx, g):
        z, sel, arg, coef = ctx.saved_tensors
        k = ctx.k
        P, Cp = z.shape[0] // k, z.shape[1]
        dev = z.device
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        if g.stride(1) != 1:
            g = g.contiguous()
        d = torch.empty(P, Cp, dtype=torch.float32, device=dev)
        dgb = _stats_buffer(Cp, dev)
        _lib.call("fs_edgeconv_bwd_reduce", z, g, _lib.dtype_code(g), g.stride(0), sel, None, 0, 0, P, Cp, coef, d, dgb)
        dz = torch.empty_like(z)
        _lib.call("fs_edge_reduce_bwd", z, z, _lib.dtype_code(z), d, arg, P, k, Cp, coef, dgb, float(P * k),
                  int(ctx.training), dz, _lib.dtype_code(dz))
        dgb32 = dgb[:2 * Cp].float()
        return dz, dgb32[Cp:], dgb32[:Cp], None, None, None, None, None, None, None


def edge_reduce(z, gamma, beta, k, running_mean, running_var, nbt, training, eps=BN_EPS, momentum=BN_MOMENTUM):
    return _EdgeReduceFn.apply(z, gamma, beta, k, running_mean, running_var, nbt, training, eps, momentum)


# ------------------------------------------------------------------------------------------- Chamfer

class _ChamferFn(torch.autograd.Function):
    """chamfer_distance(x, y)[0] with pytorch3d defaults (losses/chamfer_loss.py:19):
    mean_b [ mean_i min_j |x_i - y_j|^2 + mean_j min_i |x_i - y_j|^2 ]."""

    @staticmethod
    def forward(ctx, x, y):
        B, N, _ = x.shape
        M = y.shape[1]
        dev = x.device
        dxy = torch.empty(B, N, dtype=torch.float32, device=dev)
        ixy = torch.empty(B, N, dtype=torch.int32, device=dev)
        dyx = torch.empty(B, M, dtype=torch.float32, device=dev)
        iyx = torch.empty(B, M, dtype=torch.int32, device=dev)
        _lib.call("fs_nn_points", x, x, y, B, N, M, dxy, ixy)
        _lib.call("fs_nn_points", x, y, x, B, M, N, dyx, iyx)
        loss = (dxy.sum(dtype=torch.float64) / (N * B) + dyx.sum(dtype=torch.float64) / (M * B)).float()
        ctx.save_for_backward(x, y, ixy, iyx)
        return loss

    @staticmethod
    def backward(ctx, g):
        x, y, ixy, iyx = ctx.saved_tensors
        B, N, _ = x.shape
        M = y.shape[1]
        gx = torch.zeros_like(x)
        gy = torch.zeros_like(y)
        g = g.float().contiguous()
        _lib.call("fs_chamfer_bwd", x, x, y, ixy, B, N, M, 1.0 / (N * B), g, gx, gy)
        _lib.call("fs_chamfer_bwd", x, y, x, iyx, B, M, N, 1.0 / (M * B), g, gy, gx)
        return gx, gy


def chamfer_distance(x, y):
    """x (B, N, 3), y (B, M, 3) float32 CUDA tensors -> scalar loss."""
    return _ChamferFn.apply(x.float().contiguous(), y.float().contiguous())


def nn_points(x, y):
    """Nearest neighbour in y of every point of x: (squared distance (B, N), index int32 (B, N))."""
    x, y = x.float().contiguous(), y.float().contiguous()
    B, N, _ = x.shape
    d = torch.empty(B, N, dtype=torch.float32, device=x.device)
    i = torch.empty(B, N, dtype=torch.int32, device=x.device)
    _lib.call("fs_nn_points", x, x, y, B, N, y.shape[1], d, i)
    return d, i


# ------------------------------------------------------------------------------------------- dense layers

class _BnActFn(torch.autograd.Function):
    """y = LeakyReLU(BatchNorm(x + rowbias[cloud])) on a point-major table (SharedFullyConnected, dim=1:
    models/dgcnn.py:306-315 after the 1x1 conv). Statistics over all rows, fp64 accumulation."""

    @staticmethod
    def forward(ctx, x, rowbias, gamma, beta, running_mean, running_var, nbt, training, eps, momentum, slope, N):
        rows, C = x.shape
        gamma32, beta32 = gamma.detach().float(), beta.detach().float()
        rb = rowbias.detach().float().contiguous() if rowbias is not None else None
        stats = None
        if training:
            stats = _stats_buffer(C, x.device)
            _lib.call("fs_colstats", x, x, _lib.dtype_code(x), x.stride(0), rows, C, rb, N, stats)
        out = torch.empty(rows, C, dtype=x.dtype, device=x.device)
        coef = _bn_then_act_apply(x, x, _lib.dtype_code(x), x.stride(0), rows, C, rb, N, stats, rows, gamma32, beta32,
                                  running_mean, running_var, nbt, training, eps, momentum, slope, out)
        ctx.save_for_backward(x, rb, coef)
        ctx.training, ctx.slope, ctx.N = training, slope, N
        return out

    @staticmethod
    def backward(ctx, g):
        x, rb, coef = ctx.saved_tensors
        rows, C = x.shape
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        if g.stride(1) != 1:
            g = g.contiguous()
        dgb = _stats_buffer(C, x.device)
        dx = torch.empty(rows, C, dtype=x.dtype, device=x.device)
        _lib.call("fs_bn_act_bwd", x, g, _lib.dtype_code(g), g.stride(0), x, _lib.dtype_code(x), x.stride(0), rows, C, rb,
                  ctx.N, coef, float(ctx.slope), dgb, float(rows), int(ctx.training), dx, _lib.dtype_code(dx),
                  dx.stride(0))
        dgb32 = dgb[:2 * C].float()
        drb = _cloudsum_f32(dx, rows // ctx.N, ctx.N) if rb is not None else None
        return dx, drb, dgb32[C:], dgb32[:C], None, None, None, None, None, None, None, None


def bn_act(x, bn, slope, rowbias=None, N=1):
    """BatchNorm(+LeakyReLU) over the rows of x (rows, C); bn is the nn.BatchNorm module holding the state."""
    assert x.stride(1) == 1
    return _BnActFn.apply(x, rowbias, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                          bn.training, bn.eps, bn.momentum, slope, N)


class _PoolBnActFn(torch.autograd.Function):
    """(B, C) = max over the N rows of each cloud of LeakyReLU(BatchNorm(x)) without writing the activation:
    Conv1d + BN + LeakyReLU + AdaptiveMaxPool1d of the global feature (models/dgcnn.py:123-126, 156)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, training, eps, momentum, slope, B, N):
        C = x.shape[1]
        dev = x.device
        gamma32, beta32 = gamma.detach().float(), beta.detach().float()
        sel = torch.empty(B, C, dtype=torch.float32, device=dev)
        arg = torch.empty(B, C, dtype=torch.int32, device=dev)
        stats = _stats_buffer(C, dev) if training else None
        packed = _zeros64(B * C, dev).view(torch.int64)
        _lib.call("fs_pool_reduce", x, x, _lib.dtype_code(x), x.stride(0), B, N, C, gamma32, sel, arg, stats, packed)
        out = torch.empty(B, C, dtype=x.dtype, device=dev)
        coef = _bn_then_act_apply(x, sel, 0, C, B, C, None, 1, stats, B * N, gamma32, beta32, running_mean, running_var, nbt,
                                  training, eps, momentum, slope, out)
        ctx.save_for_backward(x, sel, arg, coef)
        ctx.training, ctx.slope, ctx.B, ctx.N = training, slope, B, N
        return out

    @staticmethod
    def backward(ctx, g):
        x, sel, arg, coef = ctx.saved_tensors
        B, N, C = ctx.B, ctx.N, x.shape[1]
        g32 = g.float().contiguous()
        dgb = _stats_buffer(C, x.device)
        _lib.call("fs_bn_act_bwd", x, g32, 0, C, sel, 0, C, B, C, None, 1, coef, float(ctx.slope), dgb, float(B * N),
                  int(ctx.training), None, 0, C)
        dx = torch.empty(B * N, C, dtype=x.dtype, device=x.device)
        _lib.call("fs_pool_bwd", x, x, _lib.dtype_code(x), x.stride(0), B, N, C, g32, sel, arg, coef, float(ctx.slope), dgb,
                  float(B * N), int(ctx.training), dx, _lib.dtype_code(dx), C)
        dgb32 = dgb[:2 * C].float()
        return dx, dgb32[C:], dgb32[:C], None, None, None, None, None, None, None, None, None


def pool_bn_act(x, bn, slope, B, N):
    assert x.stride(1) == 1
    return _PoolBnActFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                              bn.training, bn.eps, bn.momentum, slope, B, N)


USE_POOL_GRAM = os.environ.get("FS_POOL_GRAM", "1") != "0"


def _gram_f32(x):
    """X^T X with fp32 accumulation AND fp32 output (a bf16 output would lose the covariance to rounding)."""
    P = x.shape[0]
    S = 1
    while S < 32 and P % (2 * S) == 0 and P // (2 * S) >= 1024:
        S *= 2
    if S > 1 and x.is_contiguous():
        # rows-long reduction, tiny output: batched over row chunks (fp32 partials), then summed - the library's own
        # split-K of the single GEMM runs at a fraction of the rate
        xv = x.view(S, P // S, -1)
        part = torch.bmm(xv.transpose(1, 2), xv) if x.dtype == torch.float32 else torch.bmm(xv.transpose(1, 2), xv, out_dtype=torch.float32)
        return _sum_chunks(part)
    if x.dtype == torch.float32:
        return x.t() @ x
    return torch.mm(x.t(), x, out_dtype=torch.float32)


USE_POOL_GEMM = os.environ.get("FS_POOL_GEMM", "1") != "0"


def _colsum_f32(x):
    rows, K = x.shape
    out = torch.empty(K, dtype=torch.float32, device=x.device)
    partial = _workspace(4 * K * _lib.load().fs_colsum_partials(), x.device, "colsum")
    _lib.call("fs_colsum", x, x, _lib.dtype_code(x), x.stride(0), rows, K, partial, out)
    return out


class _PoolLinearFn(torch.autograd.Function):
    """out (B, C) = max over the N rows of each cloud of LeakyReLU(BatchNorm(X W^T)): Conv1d + BN + LeakyReLU +
    AdaptiveMaxPool1d of the global feature (models/dgcnn.py:123-126, 156).

    Forward, bf16 tables with K in {64, 128, 192}: ONE tcgen05 kernel forms the product tile by tile in TMEM and keeps
    only per-cloud max/min + arg (csrc/pool_gemm.cu) - the (B*N, C) product is never written; the BatchNorm sums come from
    the Gram matrix (sum y = W colsum(X), sum y^2 = diag(W X^T X W^T)). Otherwise: library GEMM, then one pass that keeps
    max/min + statistics. Either way nothing of size B*N x C is kept for backward.
    Backward: dy = S + a + b*y with one non-zero of S per (cloud, channel), so with y = X W^T
        dX = S W + 1 (a^T W) + X (W^T diag(b) W),      dW = S^T X + a colsum(X)^T + diag(b) W (X^T X)
    - K x K products instead of the P x C gradient (csrc/heads.cu)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, running_mean, running_var, nbt, training, eps, momentum, slope, B, N):
        wc = w.detach().to(x.dtype).contiguous()              # (C, K)
        C, K = wc.shape
        dev = x.device
        lib = _lib.load()
        gamma32, beta32 = gamma.detach().float(), beta.detach().float()
        sel = torch.empty(B, C, dtype=torch.float32, device=dev)
        arg = torch.empty(B, C, dtype=torch.int32, device=dev)
        stats = _stats_buffer(C, dev) if training else None
        packed = _zeros64(B * C, dev).view(torch.int64)
        colsum = wg = None
        fused = (USE_POOL_GEMM and x.dtype == torch.bfloat16 and lib.fs_pool_gemm_supported(B, N, C, K)
                 and x.stride(0) % 8 == 0 and x.data_ptr() % 16 == 0)
        if fused:
            if training:
                with _matmul_tf32(False):
                    colsum = _colsum_f32(x)
                    wg = (wc.float() @ _gram_f32(x)).contiguous()                    # W (X^T X)       (C, K)
                _lib.call("fs_pool_stats_from_gram", x, wc, _lib.dtype_code(wc), wc.stride(0), wg, colsum, C, K, stats)
            w_signed = wc * torch.where(gamma32 >= 0, 1.0, -1.0).to(wc.dtype).unsqueeze(1)
            _lib.call("fs_pool_gemm", x, x, x.stride(0), w_signed, B, N, C, K, packed)
            _lib.call("fs_pool_decode", x, packed, gamma32, B, C, sel, arg)
            ref = x
        else:
            y = x @ wc.t()
            _lib.call("fs_pool_reduce", y, y, _lib.dtype_code(y), y.stride(0), B, N, C, gamma32, sel, arg, stats, packed)
            ref = y
        out = torch.empty(B, C, dtype=x.dtype, device=dev)
        coef = _bn_then_act_apply(ref, sel, 0, C, B, C, None, 1, stats, B * N, gamma32, beta32, running_mean, running_var,
                                  nbt, training, eps, momentum, slope, out)
        ctx.save_for_backward(x, wc, sel, arg, coef, colsum, wg)
        ctx.training, ctx.slope, ctx.B, ctx.N, ctx.w_dtype = training, slope, B, N, w.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        x, wc, sel, arg, coef, colsum, wg = ctx.saved_tensors
        B, N, C, K = ctx.B, ctx.N, wc.shape[0], wc.shape[1]
        dev = x.device
        g32 = g.float().contiguous()
        dgb = _stats_buffer(C, dev)
        _lib.call("fs_bn_act_bwd", x, g32, 0, C, sel, 0, C, B, C, None, 1, coef, float(ctx.slope), dgb, float(B * N),
                  int(ctx.training), None, 0, C)
        vec = torch.empty(2 * C + B * C, dtype=torch.float32, device=dev)
        a, bvec, sp = vec[:C], vec[C:2 * C], vec[2 * C:]
        _lib.call("fs_pool_lin_bwd_prep", x, g32, sel, coef, float(ctx.slope), dgb, float(B * N), int(ctx.training), B, C,
                  a, bvec, sp)
        w32 = wc.float()
        dx = dw = None
        if ctx.training:
            with _matmul_tf32(False):     # the K x K factors carry a covariance (cancellation): full fp32 products
                if ctx.needs_input_grad[0]:
                    m = (w32 * bvec.unsqueeze(1)).t() @ w32                              # W^T diag(b) W   (K, K)
                    r = a @ w32                                                          # a^T W           (K,)
                    dx = torch.addmm(r.to(x.dtype), x, m.to(x.dtype))
                if ctx.needs_input_grad[1] and wg is None:
                    colsum = _colsum_f32(x)
                    wg = (w32 @ _gram_f32(x)).contiguous()                               # W (X^T X)       (C, K)
        elif ctx.needs_input_grad[0]:
            dx = torch.zeros_like(x)
        if dx is not None:
            ws = _workspace(_lib.load().fs_pool_lin_bwd_ws_bytes(B, C, K), dev, "pool_lin")
            _lib.call("fs_pool_lin_bwd_dx_sparse", x, sp, arg, wc, _lib.dtype_code(wc), wc.stride(0), B, N, C, K, dx,
                      dx.stride(0), ws)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(C, K, dtype=torch.float32, device=dev)
            tr = ctx.training
            _lib.call("fs_pool_lin_bwd_dw", x, sp, arg, x, _lib.dtype_code(x), x.stride(0), B, N, C, K,
                      a if tr else None, bvec if tr else None, colsum if tr else None, wg if tr else None, dw)
            dw = dw.to(ctx.w_dtype)
        dgb32 = dgb[:2 * C].float()
        return dx, dw, dgb32[C:], dgb32[:C], None, None, None, None, None, None, None, None, None


def pool_linear_supported(x, w):
    K = x.shape[1]
    return (USE_POOL_GRAM and x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and x.is_contiguous()
            and K % 32 == 0 and K <= 512 and w.shape[1] == K and 64 <= w.shape[0] <= 1024
            and (w.shape[0] & (w.shape[0] - 1)) == 0)


def pool_linear_bn_act(x, w, bn, slope, B, N):
    """x (B*N, K) contiguous, w (C, K): Conv1d(k=1, bias=False) + BN + LeakyReLU + max over the points of each cloud."""
    return _PoolLinearFn.apply(x, w, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked,
                               bn.training, bn.eps, bn.momentum, slope, B, N)


def _host_arrays(tensors):
    n = len(tensors)
    ptrs = (ctypes.c_void_p * 4)(*([t.data_ptr() for t in tensors] + [None] * (4 - n)))
    widths = (ctypes.c_int * 4)(*([t.shape[1] for t in tensors] + [0] * (4 - n)))
    lds = (ctypes.c_int * 4)(*([t.stride(0) for t in tensors] + [0] * (4 - n)))
    return ptrs, widths, lds


class _CatCastFn(torch.autograd.Function):
    """torch.cat((x1, x2, ...), dim=1).to(dtype) of fp32 point-major tables in one pass (models/dgcnn.py:154, 200);
    the gradient comes back as contiguous fp32 tables, one pass as well."""

    @staticmethod
    def forward(ctx, dtype, *xs):
        rows = xs[0].shape[0]
        total = sum(t.shape[1] for t in xs)
        out = torch.empty(rows, total, dtype=dtype, device=xs[0].device)
        ptrs, widths, lds = _host_arrays(xs)
        _lib.call("fs_cat_cast", out, len(xs), ptrs, widths, lds, rows, out, _lib.dtype_code(out), out.stride(0))
        ctx.widths = [t.shape[1] for t in xs]
        return out

    @staticmethod
    def backward(ctx, g):
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        if g.stride(1) != 1:
            g = g.contiguous()
        rows = g.shape[0]
        outs = [torch.empty(rows, c, dtype=torch.float32, device=g.device) for c in ctx.widths]
        ptrs, widths, lds = _host_arrays(outs)
        _lib.call("fs_split_cast", g, len(outs), ptrs, widths, lds, rows, g, _lib.dtype_code(g), g.stride(0))
        return (None, *outs)


def cat_cast(xs, dtype):
    """[x1 | x2 | ...] of fp32 (rows, C_i) tables as one (rows, sum C_i) table in `dtype` (fp32 or bf16)."""
    ok = (1 <= len(xs) <= 4 and all(t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1
                                    and t.shape[1] % 4 == 0 and t.stride(0) % 4 == 0 for t in xs)
          and dtype in (torch.float32, torch.bfloat16))
    if not ok:
        return torch.cat(list(xs), dim=1).to(dtype)
    return _CatCastFn.apply(dtype, *xs)


# ------------------------------------------------------------------------------------------- per-point GEMM

class _matmul_tf32:
    """Scoped torch.backends.cuda.matmul.allow_tf32: the product decides per call whether an fp32 GEMM may round its
    operands to TF32, instead of inheriting whatever the process set globally."""

    def __init__(self, allow):
        self.allow = bool(allow)

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.allow

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


class _TableGemmFn(torch.autograd.Function):
    """T = X W^T for a tall-skinny X (P x C, P ~ 1e5) and a small W (2Cp x C). The weight gradient
    dW = dT^T X has a P-long reduction dimension and a tiny output; as one GEMM it runs on a few dozen CTAs,
    so it is issued as a batched GEMM over row chunks followed by a sum over the chunks.
    tf32: operands rounded to TF32 (10-bit mantissa) on the tensor cores, fp32 accumulation and output - the 'bf16'
    precision mode; the 'fp32' mode keeps full fp32 products."""

    @staticmethod
    def forward(ctx, x, w, tf32):
        ctx.save_for_backward(x, w)
        ctx.tf32 = tf32
        with _matmul_tf32(tf32):
            return x @ w.t()

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = None
        with _matmul_tf32(ctx.tf32):
            if ctx.needs_input_grad[0]:
                gx = g @ w
            if ctx.needs_input_grad[1]:
                P = x.shape[0]
                S = 1
                while S < 128 and P % (2 * S) == 0 and P // (2 * S) >= 256:
                    S *= 2
                if S > 1 and g.is_contiguous() and x.is_contiguous():
                    gw = _sum_chunks(torch.bmm(g.view(S, P // S, -1).transpose(1, 2), x.view(S, P // S, -1)))
                else:
                    gw = g.t() @ x
        return gx, gw, None


class _LogitsOutFn(torch.autograd.Function):
    """(B*N, C) point-major logits of the internally sorted cloud -> (B, C, N) fp32 in the caller's order, one kernel each
    way (instead of permute + cast + scatter and gather + permute + cast)."""

    @staticmethod
    def forward(ctx, logits, perm, B, N):
        C = logits.shape[1]
        out = torch.empty(B, C, N, dtype=torch.float32, device=logits.device)
        _lib.call("fs_logits_out", logits, logits, _lib.dtype_code(logits), logits.stride(0), perm, B, N, C, out)
        ctx.perm, ctx.shape, ctx.dtype = perm, (B, N, C), logits.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        B, N, C = ctx.shape
        g = g.float().contiguous()
        d = torch.empty(B * N, C, dtype=ctx.dtype, device=g.device)
        _lib.call("fs_logits_out_bwd", g, g, ctx.perm, B, N, C, d, _lib.dtype_code(d), d.stride(0))
        return d, None, None, None


def logits_out(logits_pm, perm, B, N):
    """logits_pm (B*N, C) fp32 / bf16 with unit channel stride; perm (B, N) int64 or None."""
    if not logits_pm.is_cuda or logits_pm.dtype not in (torch.float32, torch.bfloat16) or logits_pm.stride(1) != 1:
        y = logits_pm.view(B, N, -1).permute(0, 2, 1).float()
        if perm is None:
            return y
        return torch.empty_like(y).scatter(2, perm.unsqueeze(1).expand_as(y), y)
    return _LogitsOutFn.apply(logits_pm, perm.contiguous() if perm is not None else None, B, N)


class _FinalLinearFn(torch.autograd.Function):
    """Last layer of the segmentation head (1x1 conv to num_classes channels + bias, no BatchNorm) fused with the network
    output: (B*N, C_in) -> (B, classes, N) fp32 in the caller's point order, one pass over H each way (csrc/heads.cu);
    as library calls it is three GEMMs with a 4-wide dimension plus the un-sort / transpose / cast passes."""

    @staticmethod
    def forward(ctx, h, w, bias, perm, B, N):
        w32 = w.detach().float().contiguous()
        b32 = bias.detach().float().contiguous() if bias is not None else None
        C_out, C_in = w32.shape
        out = torch.empty(B, C_out, N, dtype=torch.float32, device=h.device)
        _lib.call("fs_final_linear_fwd", h, h, _lib.dtype_code(h), h.stride(0), w32, b32, perm, B, N, C_in, C_out, out)
        ctx.save_for_backward(h, w32)
        ctx.perm, ctx.dims, ctx.w_dtype, ctx.has_bias = perm, (B, N, C_in, C_out), w.dtype, bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        h, w32 = ctx.saved_tensors
        B, N, C_in, C_out = ctx.dims
        g = g.float().contiguous()
        dh = torch.empty(B * N, C_in, dtype=h.dtype, device=h.device)
        ws = _workspace(4 * _lib.load().fs_final_linear_ws_floats(B * N, C_in, C_out), h.device, "final_linear")
        dwdb = torch.empty(C_out * C_in + C_out, dtype=torch.float32, device=h.device)
        _lib.call("fs_final_linear_bwd", h, h, _lib.dtype_code(h), h.stride(0), w32, g, ctx.perm, B, N, C_in, C_out, dh,
                  dh.stride(0), ws, dwdb)
        dw = dwdb[:C_out * C_in].view(C_out, C_in).to(ctx.w_dtype)
        db = dwdb[C_out * C_in:].to(ctx.w_dtype) if ctx.has_bias else None
        return dh, dw, db, None, None, None


def final_linear_supported(h, w):
    return (h.is_cuda and h.dtype in (torch.float32, torch.bfloat16) and h.dim() == 2 and h.stride(1) == 1
            and h.stride(0) % 8 == 0 and h.data_ptr() % 16 == 0 and w.dim() == 2 and w.shape[1] == h.shape[1]
            and bool(_lib.load().fs_final_linear_supported(w.shape[1], w.shape[0])))


def final_linear_out(h, w, bias, perm, B, N):
    """h (B*N, C_in), w (classes, C_in), bias (classes,) or None, perm (B, N) int64 or None -> (B, classes, N) fp32."""
    return _FinalLinearFn.apply(h, w, bias, perm.contiguous() if perm is not None else None, B, N)


def _sum_chunks(part):
    """(S, M, N) fp32 partial products -> (M, N): one coalesced pass (ATen's reduction over a short leading dimension of a
    small tensor takes 8 us for 4 MB)."""
    if not part.is_cuda or part.dtype != torch.float32 or not part.is_contiguous():
        return part.sum(dim=0)
    S, M, N = part.shape
    out = torch.empty(M, N, dtype=torch.float32, device=part.device)
    _lib.call("fs_sum_leading", part, part, S, M * N, out)
    return out


_ones_cache = {}


def _ones(n, dtype, device):
    key = (n, dtype, torch.device(device))
    t = _ones_cache.get(key)
    if t is None:
        t = torch.ones(n, dtype=dtype, device=device)
        _ones_cache[key] = t
    return t


def _rowsum_f32(g):
    """Column sums of a tall (rows, C) table as a GEMM with a row of ones (fp32 result): ATen's reduction over the long
    dimension of a narrow table runs at a few percent of the memory rate."""
    if not g.is_cuda or g.dtype == torch.float32 or not g.is_contiguous():
        return g.float().sum(dim=0)
    return torch.mm(_ones(g.shape[0], g.dtype, g.device).unsqueeze(0), g, out_dtype=torch.float32).squeeze(0)


def _cloudsum_f32(dx, B, N):
    """Per-cloud column sums of a (B*N, C) table -> (B, C) fp32, as a batched GEMM with a row of ones."""
    C = dx.shape[1]
    if not dx.is_cuda or dx.dtype == torch.float32 or not dx.is_contiguous():
        return dx.view(B, N, C).sum(dim=1, dtype=torch.float32)
    ones = _ones(N, dx.dtype, dx.device).view(1, 1, N).expand(B, 1, N)
    return torch.bmm(ones, dx.view(B, N, C), out_dtype=torch.float32).squeeze(1)


class _LinearFn(torch.autograd.Function):
    """y = X W^T for a tall X (rows ~ 1e5) and a small fp32 parameter W (C_out x C_in), computed in X's dtype (1x1 conv of
    the dense heads, models/dgcnn.py:127-137). The weight gradient dW = dY^T X has a rows-long reduction and a tiny
    output; as one GEMM the library runs it split-K at ~15 % of the tensor peak, so it is issued as a batched GEMM over
    row chunks with fp32 partial outputs followed by a sum over the chunks. dW comes out in fp32 (the parameter's dtype):
    no bf16 round trip and no cast kernel."""

    @staticmethod
    def forward(ctx, x, w, bias=None):
        wc = w.detach().to(x.dtype)
        ctx.save_for_backward(x, wc)
        ctx.w_dtype = w.dtype
        ctx.has_bias = bias is not None
        if bias is not None:
            return torch.addmm(bias.detach().to(x.dtype), x, wc.t())
        return x @ wc.t()

    @staticmethod
    def backward(ctx, g):
        x, wc = ctx.saved_tensors
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = g @ wc
        if ctx.needs_input_grad[1]:
            P = x.shape[0]
            S = 1
            while S < 16 and P % (2 * S) == 0 and P // (2 * S) >= 1024:
                S *= 2
            if S > 1 and g.is_contiguous() and x.is_contiguous() and g.is_cuda and g.shape[1] * x.shape[1] >= 16384:
                gt, xv = g.view(S, P // S, -1).transpose(1, 2), x.view(S, P // S, -1)
                part = torch.bmm(gt, xv) if g.dtype == torch.float32 else torch.bmm(gt, xv, out_dtype=torch.float32)
                gw = _sum_chunks(part)
            elif g.dtype == torch.float32 or not g.is_cuda:
                gw = g.float().t() @ x.float()
            else:
                gw = torch.mm(g.t(), x, out_dtype=torch.float32)      # small output: one GEMM, fp32 result, no casts of X
            gw = gw.to(ctx.w_dtype)
        gb = None
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _rowsum_f32(g).to(ctx.w_dtype)
        return gx, gw, gb


def linear_pm(x, w, bias=None):
    """x (rows, C_in) @ w (C_out, C_in)^T (+ bias) with w, bias fp32 parameters (or views of one); see _LinearFn."""
    return _LinearFn.apply(x, w, bias)


def table_gemm(x, w, tf32=False):
    return _TableGemmFn.apply(x, w, tf32)


class _EdgeWeightFn(torch.autograd.Function):
    """[W1 ; W2 - W1] (2Cp, C) from the EdgeConv weight W = [W1 | W2] (Cp, 2C) (models/dgcnn.py:36 column order).
    Written as `cat([w[:, :C], w[:, C:] - w[:, :C]])` autograd replays slices, a subtraction and a concatenation:
    4 small kernels forward and 9 backward per layer; here it is 1 + 1."""

    @staticmethod
    def forward(ctx, w, C):
        Cp = w.shape[0]
        wf = w if w.dtype == torch.float32 else w.float()
        out = torch.empty(2 * Cp, C, dtype=torch.float32, device=w.device)
        if wf.is_cuda and wf.is_contiguous():
            _lib.call("fs_edge_weight_table", wf, wf, Cp, C, out)
        else:
            out[:Cp].copy_(wf[:, :C])
            torch.sub(wf[:, C:], wf[:, :C], out=out[Cp:])
        ctx.C = C
        ctx.w_dtype = w.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        C = ctx.C
        Cp = g.shape[0] // 2
        g = g.float()
        dw = torch.empty(Cp, 2 * C, dtype=torch.float32, device=g.device)
        if g.is_cuda and g.is_contiguous():
            _lib.call("fs_edge_weight_table_bwd", g, g, Cp, C, dw)
        else:
            torch.sub(g[:Cp], g[Cp:], out=dw[:, :C])        # dW1 = g_a - g_b
            dw[:, C:].copy_(g[Cp:])                         # dW2 = g_b
        return dw.to(ctx.w_dtype), None


def edge_weight_table(w, C):
    """w (Cp, 2C) -> fp32 (2Cp, C) = [W1 ; W2 - W1], the right-hand side of the per-point table GEMM."""
    return _EdgeWeightFn.apply(w, C)


# ------------------------------------------------------------------------------------------- first layer, 3-D input

class _EdgeFirst3Fn(torch.autograd.Function):
    """H = LeakyReLU(BatchNorm2d(Conv2d(6 -> Cp)([x_j - x_i, x_i]))) for a 3-channel input x (P, 3): first shared
    layer of a multi-layer EdgeConv on coordinates (models/dgcnn.py:119, 251). Batch statistics come from the
    moments of the 6-D edge vectors; H (P*k, Cp) is the only tensor written, the weight gradient is accumulated
    directly from dH (no scatter, x gets no gradient)."""

    @staticmethod
    def forward(ctx, x, w, gamma, beta, graph, running_mean, running_var, nbt, training, eps, momentum, out_dtype):
        B, N, k = graph.B, graph.N, graph.k
        Cp = w.shape[0]
        dev = x.device
        w32 = w.detach().reshape(Cp, 6).float().contiguous()
        gamma32, beta32 = gamma.detach().float(), beta.detach().float()
        coef = torch.empty(4 * Cp, dtype=torch.float32, device=dev)
        mom = None
        if training:
            mom = _zeros64(_lib.load().fs_edge3_moment_doubles(), dev)
            _lib.call("fs_edge3_bn_coef", x, x, x.stride(0), graph.idx, B, N, k, w32, Cp, gamma32, beta32, eps, momentum,
                      mom, coef, running_mean, running_var, nbt)
        else:
            _lib.call("fs_bn_coef_eval", x, Cp, gamma32, beta32, running_mean, running_var, eps, coef)
        h = torch.empty(B * N * k, Cp, dtype=out_dtype, device=dev)
        _lib.call("fs_edge3_hidden", x, x, x.stride(0), graph.idx, B, N, k, w32, Cp, coef, h, _lib.dtype_code(h))
        ctx.graph, ctx.training = graph, training
        ctx.w_shape = w.shape
        ctx.save_for_backward(x, w32, coef, mom)
        return h

    @staticmethod
    def backward(ctx, dh):
        x, w32, coef, mom = ctx.saved_tensors
        graph = ctx.graph
        B, N, k = graph.B, graph.N, graph.k
        Cp = w32.shape[0]
        if dh.dtype not in (torch.float32, torch.bfloat16):
            dh = dh.float()
        dh = dh.contiguous()
        dgb = _stats_buffer(Cp, x.device)
        acc = _zeros64(Cp * 3, x.device).view(torch.float32).view(Cp, 6)
        dw = torch.empty(Cp, 6, dtype=torch.float32, device=x.device)
        _lib.call("fs_edge3_bwd", x, x, x.stride(0), graph.idx, B, N, k, w32, Cp, coef, dh, _lib.dtype_code(dh),
                  int(ctx.training), mom, dgb, acc, dw)
        dgb32 = dgb[:2 * Cp].float()
        return None, dw.view(ctx.w_shape), dgb32[Cp:], dgb32[:Cp], None, None, None, None, None, None, None, None


def edge_first3(x, conv_weight, bn, graph, out_dtype):
    """x (P, >=3) fp32 point-major (first three channels are used), conv_weight (Cp, 6, 1, 1)."""
    assert x.dtype == torch.float32 and x.stride(1) == 1
    return _EdgeFirst3Fn.apply(x, conv_weight, bn.weight, bn.bias, graph, bn.running_mean, bn.running_var,
                               bn.num_batches_tracked, bn.training, bn.eps, bn.momentum, out_dtype)


# ------------------------------------------------------------------------------------------- fused two-layer EdgeConv

USE_FUSED_EDGE2 = os.environ.get("FS_EDGE2", "1") != "0"


def edge2_supported(k, C1, C2):
    return bool(USE_FUSED_EDGE2 and _lib.load().fs_edge2_supported(int(k), int(C1), int(C2)))


class _EdgeConv2Fn(torch.autograd.Function):
    """out_i = LeakyReLU(BN2(max_j W2 LeakyReLU(BN1(W1 [x_j - x_i, x_i])))) for 3-channel coordinates x: the two-layer
    EdgeConv of models/dgcnn.py:119, 237-241 in two kernels forward / backward that never write an edge tensor
    (csrc/edge2.cu). bf16 tensor-core operands, fp32 accumulation."""

    @staticmethod
    def forward(ctx, x, w1, g1, b1, w2, g2, b2, graph, rm1, rv1, nbt1, rm2, rv2, nbt2, training, eps, momentum):
        B, N, k = graph.B, graph.N, graph.k
        P = B * N
        dev = x.device
        C1, C2 = w1.shape[0], w2.shape[0]
        w1m = w1.detach().reshape(C1, 6).float().contiguous()
        w2m = w2.detach().reshape(C2, C1).float().contiguous()
        g1f, b1f, g2f, b2f = (t.detach().float() for t in (g1, b1, g2, b2))
        coef1 = torch.empty(4 * C1, dtype=torch.float32, device=dev)
        mom = None
        if training:
            mom = _zeros64(_lib.load().fs_edge3_moment_doubles(), dev)
            _lib.call("fs_edge3_bn_coef", x, x, x.stride(0), graph.idx, B, N, k, w1m, C1, g1f, b1f, eps, momentum, mom, coef1,
                      rm1, rv1, nbt1)
        else:
            _lib.call("fs_bn_coef_eval", x, C1, g1f, b1f, rm1, rv1, eps, coef1)
        sel = torch.empty(P, C2, dtype=torch.float32, device=dev)
        arg = torch.empty(P, C2, dtype=torch.uint8, device=dev)
        stats = gram = hsum = None
        if training:
            stats = _stats_buffer(C2, dev)
            gh = _zeros64((C1 * C1 + C1 + 1) // 2 + 1, dev).view(torch.float32)
            gram, hsum = gh[:C1 * C1].view(C1, C1), gh[C1 * C1:C1 * C1 + C1]
        _lib.call("fs_edge2_fwd", x, x, x.stride(0), graph.idx, B, N, k, w1m, coef1, w2m, C2, g2f, sel, arg, stats, gram, hsum)
        out = torch.empty(P, C2, dtype=torch.float32, device=dev)
        coef2 = _bn_then_edgeconv_apply(x, sel, None, 0, 0, P, C2, stats, P * k, g2f, b2f, rm2, rv2, nbt2, training, eps,
                                        momentum, out)
        ctx.graph, ctx.training = graph, training
        ctx.shapes = (w1.shape, w2.shape)
        ctx.save_for_backward(x, w1m, w2m, coef1, coef2, sel, arg, mom, gram, hsum)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w1m, w2m, coef1, coef2, sel, arg, mom, gram, hsum = ctx.saved_tensors
        graph = ctx.graph
        B, N, k = graph.B, graph.N, graph.k
        P = B * N
        C1, C2 = w1m.shape[0], w2m.shape[0]
        dev = x.device
        if g.dtype not in (torch.float32, torch.bfloat16):
            g = g.float()
        if g.stride(1) != 1:
            g = g.contiguous()
        d = torch.empty(P, C2, dtype=torch.float32, device=dev)
        dgb2 = _stats_buffer(C2, dev)
        _lib.call("fs_edgeconv_bwd_reduce", x, g, _lib.dtype_code(g), g.stride(0), sel, None, 0, 0, P, C2, coef2, d, dgb2)
        dgb1 = _stats_buffer(C1, dev)
        nscr = _lib.load().fs_edge2_bwd_scratch_floats()
        scratch = _zeros64((nscr + C1 * 6 + 1) // 2 + 1, dev).view(torch.float32)
        acc1 = scratch[nscr:nscr + C1 * 6]
        dw2 = torch.empty(C2, C1, dtype=torch.float32, device=dev)
        _lib.call("fs_edge2_bwd", x, x, x.stride(0), graph.idx, B, N, k, w1m, coef1, w2m, C2, coef2, dgb2, int(ctx.training),
                  gram, hsum, d, arg, scratch, dgb1, acc1, dw2)
        dw1 = torch.empty(C1, 6, dtype=torch.float32, device=dev)
        _lib.call("fs_edge3_dw", x, acc1, dgb1, mom, float(P * k), w1m, coef1, C1, int(ctx.training), dw1)
        dgb1f, dgb2f = dgb1[:2 * C1].float(), dgb2[:2 * C2].float()
        return (None, dw1.view(ctx.shapes[0]), dgb1f[C1:], dgb1f[:C1], dw2.view(ctx.shapes[1]), dgb2f[C2:], dgb2f[:C2],
                None, None, None, None, None, None, None, None, None, None)


def edgeconv2_fused(x, layer1, layer2, graph):
    """x (P, >=3) fp32 point-major coordinates; layer1 / layer2: the two SharedFullyConnected blocks of the EdgeConv."""
    bn1, bn2 = layer1.norm, layer2.norm
    return _EdgeConv2Fn.apply(x, layer1.conv.weight, bn1.weight, bn1.bias, layer2.conv.weight, bn2.weight, bn2.bias, graph,
                              bn1.running_mean, bn1.running_var, bn1.num_batches_tracked, bn2.running_mean, bn2.running_var,
                              bn2.num_batches_tracked, bn1.training, bn1.eps, bn1.momentum)
