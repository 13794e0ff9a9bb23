"""BatchNorm+LeakyReLU and fused global max-pool kernels vs plain PyTorch fp32 (floating-point kernels:
rtol 1e-4 / atol 1e-5 on outputs and gradients, rtol 1e-4 on running statistics)."""
import pytest
import torch
import torch.nn.functional as F

from fissure_segmentation_b200 import ops
from parity import assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _bn(C, seed):
    gen = torch.Generator().manual_seed(seed)
    bn = torch.nn.BatchNorm1d(C)
    with torch.no_grad():
        bn.weight.copy_((0.5 + torch.rand(C, generator=gen)) * torch.where(torch.rand(C, generator=gen) < 0.3, -1.0, 1.0))
        bn.bias.copy_(0.2 * torch.randn(C, generator=gen))
        bn.running_mean.copy_(0.1 * torch.randn(C, generator=gen))
        bn.running_var.copy_(0.5 + torch.rand(C, generator=gen))
    return bn


@pytest.mark.parametrize("rows,C,N,bias,training", [(4096, 256, 1024, True, True), (4096, 64, 1, False, True),
                                                    (3000, 1024, 1, False, True), (2048, 128, 512, True, False)])
def test_bn_act_matches_torch(lib, rows, C, N, bias, training):
    gen = torch.Generator().manual_seed(rows + C)
    x = (torch.randn(rows, C, generator=gen) * 2 + 0.7)
    rb = torch.randn(rows // N, C, generator=gen) if bias else None
    gout = torch.randn(rows, C, generator=gen)
    bn_ref, bn_gpu = _bn(C, 3), _bn(C, 3).to(DEV)
    bn_ref.train(training); bn_gpu.train(training)

    xr = x.clone().requires_grad_(True)
    rbr = rb.clone().requires_grad_(True) if bias else None
    xin = xr if not bias else (xr.view(-1, N, C) + rbr.unsqueeze(1)).view(rows, C)
    ref = F.leaky_relu(bn_ref(xin), 0.2)
    ref.backward(gout)

    xg = x.to(DEV).requires_grad_(True)
    rbg = rb.to(DEV).requires_grad_(True) if bias else None
    out = ops.bn_act(xg, bn_gpu, 0.2, rbg, N)
    out.backward(gout.to(DEV))
    assert_close(out, ref, 1e-4, 1e-5, "forward")
    assert rel_err(xg.grad, xr.grad) < 1e-4
    assert_close(xg.grad, xr.grad, 1e-3, 1e-4 * float(xr.grad.abs().max()), "dx")
    assert rel_err(bn_gpu.weight.grad, bn_ref.weight.grad) < 1e-4
    assert rel_err(bn_gpu.bias.grad, bn_ref.bias.grad) < 1e-4
    if bias:
        assert rel_err(rbg.grad, rbr.grad) < 1e-4
    if training:
        assert_close(bn_gpu.running_mean, bn_ref.running_mean, 1e-4, 1e-6, "running_mean")
        assert_close(bn_gpu.running_var, bn_ref.running_var, 1e-4, 1e-6, "running_var")
        assert int(bn_gpu.num_batches_tracked) == 1


@pytest.mark.parametrize("B,N,C,training,dtype", [(4, 1000, 1024, True, torch.float32), (2, 2048, 128, False, torch.float32),
                                                  (3, 512, 1024, True, torch.bfloat16)])
def test_pool_bn_act_matches_torch(lib, B, N, C, training, dtype):
    gen = torch.Generator().manual_seed(B * N + C)
    x = (torch.randn(B * N, C, generator=gen) * 1.5 - 0.3).to(dtype)
    gout = torch.randn(B, C, generator=gen)
    bn_ref, bn_gpu = _bn(C, 5), _bn(C, 5).to(DEV)
    bn_ref.train(training); bn_gpu.train(training)
    xr = x.float().clone().requires_grad_(True)
    # AdaptiveMaxPool1d like the reference (models/dgcnn.py:125): one arg-max per (cloud, channel), first index on ties
    ref = F.adaptive_max_pool1d(F.leaky_relu(bn_ref(xr), 0.2).view(B, N, C).transpose(1, 2), 1).squeeze(-1)
    ref.backward(gout)
    xg = x.to(DEV).requires_grad_(True)
    out = ops.pool_bn_act(xg, bn_gpu, 0.2, B, N)
    out.float().backward(gout.to(DEV))
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert_close(out, ref, tol, tol, "pooled output")
    gtol = 1e-4 if dtype == torch.float32 else 1e-2     # bf16 output => the incoming gradient is bf16-rounded
    assert rel_err(xg.grad, xr.grad) < gtol
    assert rel_err(bn_gpu.weight.grad, bn_ref.weight.grad) < gtol
    assert rel_err(bn_gpu.bias.grad, bn_ref.bias.grad) < gtol
    if training:
        assert_close(bn_gpu.running_var, bn_ref.running_var, 1e-4, 1e-6, "running_var")


@pytest.mark.parametrize("B,N,K,C,training,dtype", [(4, 1000, 192, 1024, True, torch.float32),
                                                    (2, 2048, 128, 1024, False, torch.float32),
                                                    (3, 777, 512, 256, True, torch.float32),
                                                    (4, 1024, 192, 1024, True, torch.bfloat16),
                                                    (2, 1000, 128, 1024, False, torch.bfloat16),
                                                    (3, 777, 64, 256, True, torch.bfloat16),
                                                    (32, 2048, 192, 1024, True, torch.bfloat16)])
def test_pool_linear_matches_torch(lib, B, N, K, C, training, dtype):
    """Conv1d(k=1) + BN + LeakyReLU + max over points as ONE node (ops.pool_linear_bn_act): the backward works on
    K x K products (no P x C gradient). Reference: the same layer in plain PyTorch, fp64 so that the comparison measures
    the kernel and not the reference's own fp32 noise. bf16 tables (K <= 192) take the tcgen05 forward (csrc/pool_gemm.cu:
    bf16 operands, fp32 accumulation, the product is never rounded or written), so the same reference applies to the
    bf16-rounded operands; ragged N exercises the tile that straddles two clouds."""
    gen = torch.Generator().manual_seed(B * N + C + K)
    x = torch.relu(torch.randn(B * N, K, generator=gen) + 0.3) + 0.05 * torch.randn(B * N, K, generator=gen)
    w = torch.randn(C, K, generator=gen) / K ** 0.5
    if dtype == torch.bfloat16:
        x, w = x.bfloat16().float(), w.bfloat16().float()
    gout = torch.randn(B, C, generator=gen)
    bn_ref, bn_gpu = _bn(C, 5).double(), _bn(C, 5).to(DEV)
    bn_ref.train(training); bn_gpu.train(training)
    xr, wr = x.double().requires_grad_(True), w.double().requires_grad_(True)
    yr = xr @ wr.t()
    ref = F.adaptive_max_pool1d(F.leaky_relu(bn_ref(yr), 0.2).view(B, N, C).transpose(1, 2), 1).squeeze(-1)
    ref.backward(gout.double())
    xg = x.to(DEV).to(dtype).requires_grad_(True)
    wg = w.to(DEV).requires_grad_(True)
    assert ops.pool_linear_supported(xg, wg)
    out = ops.pool_linear_bn_act(xg, wg, bn_gpu, 0.2, B, N)
    out.float().backward(gout.to(DEV))
    # bf16: the pooled output and dX are stored in bf16 (2^-9 relative each)
    tol = 2e-4 if dtype == torch.float32 else 1e-2
    assert_close(out.float(), ref.float(), tol, tol, "pooled output")
    gtol = 5e-4 if dtype == torch.float32 else 1e-2
    assert rel_err(xg.grad.float(), xr.grad.float()) < gtol, rel_err(xg.grad.float(), xr.grad.float())
    assert rel_err(wg.grad, wr.grad.float()) < gtol, rel_err(wg.grad, wr.grad.float())
    assert rel_err(bn_gpu.weight.grad, bn_ref.weight.grad.float()) < gtol
    assert rel_err(bn_gpu.bias.grad, bn_ref.bias.grad.float()) < gtol
    if training:
        assert_close(bn_gpu.running_var, bn_ref.running_var.float(), 1e-3 if dtype == torch.float32 else 1e-2, 1e-6, "running_var")


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
def test_pool_linear_equals_materialised_path(lib, dtype, tol, monkeypatch):
    """Same inputs through the dense-gradient kernels and the K x K formulation: gradients agree to fp32 noise (fp32) /
    to the bf16 rounding of the dense gradient that the materialised path stores (bf16). Both sides use the library-GEMM
    forward here so that they see the same (bf16-rounded) product and pick the same arg-max rows; the tcgen05 forward is
    checked against fp64 in test_pool_linear_matches_torch."""
    monkeypatch.setattr(ops, "USE_POOL_GEMM", False)
    B, N, K, C = 4, 1024, 192, 1024
    gen = torch.Generator().manual_seed(9)
    x = torch.relu(torch.randn(B * N, K, generator=gen)).to(DEV).to(dtype)
    w = (torch.randn(C, K, generator=gen) / K ** 0.5).to(DEV).to(dtype).float()
    gout = torch.randn(B, C, generator=gen).to(DEV)
    res = []
    for fused in (False, True):
        bn = _bn(C, 5).to(DEV)
        xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        out = ops.pool_linear_bn_act(xg, wg, bn, 0.2, B, N) if fused else ops.pool_bn_act(xg @ wg.to(dtype).t(), bn, 0.2, B, N)
        out.float().backward(gout)
        res.append((out.detach().float(), xg.grad.float(), wg.grad, bn.weight.grad, bn.bias.grad))
    for a, b, name in zip(res[0], res[1], ("out", "dx", "dw", "dgamma", "dbeta")):
        assert rel_err(b, a) < tol, (name, rel_err(b, a))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_cat_cast_matches_torch(lib, dtype):
    gen = torch.Generator().manual_seed(4)
    rows = 5000
    xs = [torch.randn(rows, c, generator=gen).to(DEV).requires_grad_(True) for c in (64, 64, 128, 256)]
    wide = torch.randn(rows, 80, generator=gen).to(DEV)
    xs[1] = wide[:, 8:72].detach().requires_grad_(True)                       # strided source (ld 80)
    out = ops.cat_cast(xs, dtype)
    ref = torch.cat([t.detach() for t in xs], dim=1).to(dtype)
    assert torch.equal(out, ref)
    g = torch.randn(rows, 512, generator=gen).to(DEV).to(dtype)
    out.backward(g)
    off = 0
    for t in xs:
        assert torch.equal(t.grad, g[:, off:off + t.shape[1]].float())
        off += t.shape[1]


def test_pool_gemm_ties_and_non_finite_rows(lib, monkeypatch):
    """tcgen05 pooled forward on awkward inputs, eval mode (running statistics): identical rows (every channel ties over all
    points -> the lowest row must win, like a sequential arg-max), NaN / Inf rows (a NaN never wins; +Inf wins), a cloud
    length that is not a multiple of the 128-point tile. Compared with the library-GEMM + reduction path on the same
    bf16 operands: same arg-max rows, outputs equal up to the bf16 rounding of the materialised product."""
    from fissure_segmentation_b200 import _lib
    B, N, K, C = 3, 300, 192, 256
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B * N, K, generator=gen)
    x[:N] = x[0]                                           # cloud 0: all rows identical
    x[N + 17] = float("nan")                               # cloud 1: one NaN row, one +Inf entry elsewhere
    x[N + 40, 5] = float("inf")
    x = x.to(DEV).bfloat16()
    w = (torch.randn(C, K, generator=gen) / K ** 0.5).to(DEV)
    w[:, 5] = w[:, 5].abs() + 0.1                          # +Inf * positive weight = +Inf for every channel
    outs = []
    for fused in (True, False):
        monkeypatch.setattr(ops, "USE_POOL_GEMM", fused)
        bn = _bn(C, 5).to(DEV)
        with torch.no_grad():
            bn.weight.abs_()                               # gamma >= 0: max pooling on every channel
        bn.eval()
        with torch.no_grad():
            outs.append(ops.pool_linear_bn_act(x, w, bn, 0.2, B, N).float())
    a, b = outs
    assert torch.equal(torch.isnan(a), torch.isnan(b)) and torch.equal(torch.isinf(a), torch.isinf(b))
    fin = torch.isfinite(a)
    assert torch.isinf(a[1]).all() and (a[1] > 0).all()    # the +Inf row wins every channel of cloud 1, the NaN row none
    assert torch.allclose(a[fin], b[fin], rtol=2e-2, atol=2e-2)
    # the tie: arg-max of cloud 0 is row 0 for every channel
    wc = w.to(torch.bfloat16).contiguous()
    packed = torch.zeros(B * C, dtype=torch.int64, device=DEV)
    _lib.call("fs_pool_gemm", x, x, x.stride(0), wc, B, N, C, K, packed)
    sel = torch.empty(B, C, device=DEV)
    arg = torch.empty(B, C, dtype=torch.int32, device=DEV)
    _lib.call("fs_pool_decode", x, packed, torch.ones(C, device=DEV), B, C, sel, arg)
    assert int(arg[0].max()) == 0 and int(arg[0].min()) == 0
    assert (arg[1] == 40).all()
    assert int(arg[2].min()) >= 0 and int(arg[2].max()) < N


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_logits_out_matches_permute_scatter(lib, dtype):
    B, N, C = 3, 1000, 4
    gen = torch.Generator().manual_seed(6)
    logits = torch.randn(B * N, C, generator=gen).to(DEV).to(dtype)
    perm = torch.stack([torch.randperm(N, generator=gen) for _ in range(B)]).to(DEV)
    g = torch.randn(B, C, N, generator=gen).to(DEV)
    for p in (perm, None):
        a = logits.clone().requires_grad_(True)
        out = ops.logits_out(a, p, B, N)
        out.backward(g)
        b = logits.clone().requires_grad_(True)
        y = b.view(B, N, C).permute(0, 2, 1).float()
        ref = y if p is None else torch.empty_like(y).scatter(2, p.unsqueeze(1).expand_as(y), y)
        ref.backward(g)
        assert torch.equal(out, ref)
        assert torch.equal(a.grad, b.grad)


@pytest.mark.parametrize("dtype,C_in,C_out,N", [(torch.float32, 128, 4, 1000), (torch.bfloat16, 128, 4, 2048),
                                               (torch.float32, 256, 2, 300), (torch.bfloat16, 128, 8, 777)])
def test_final_linear_out_matches_torch(lib, dtype, C_in, C_out, N):
    """Last head layer fused with the output (classes-wide conv + bias -> (B, classes, N) fp32 in the caller's order)."""
    B = 3
    gen = torch.Generator().manual_seed(C_in + C_out + N)
    h = torch.randn(B * N, C_in, generator=gen).to(DEV).to(dtype)
    w = (torch.randn(C_out, C_in, generator=gen) / C_in ** 0.5).to(DEV)
    bias = torch.randn(C_out, generator=gen).to(DEV)
    perm = torch.stack([torch.randperm(N, generator=gen) for _ in range(B)]).to(DEV)
    g = torch.randn(B, C_out, N, generator=gen).to(DEV)
    for p in (perm, None):
        a, wa, ba = h.clone().requires_grad_(True), w.clone().requires_grad_(True), bias.clone().requires_grad_(True)
        assert ops.final_linear_supported(a, wa)
        out = ops.final_linear_out(a, wa, ba, p, B, N)
        out.backward(g)
        b, wb, bb = h.double().requires_grad_(True), w.double().requires_grad_(True), bias.double().requires_grad_(True)
        y = (b @ wb.t() + bb).view(B, N, C_out).permute(0, 2, 1)
        ref = y if p is None else torch.empty_like(y).scatter(2, p.unsqueeze(1).expand_as(y), y)
        ref.backward(g.double())
        assert_close(out, ref.float(), 1e-5, 1e-5, "logits")
        gt = 1e-5 if dtype == torch.float32 else 1e-2          # dH is stored in h's dtype
        assert rel_err(a.grad.float(), b.grad.float()) < gt
        assert rel_err(wa.grad, wb.grad.float()) < 1e-5
        assert rel_err(ba.grad, bb.grad.float()) < 1e-5


def test_small_head_helpers_match_torch(lib):
    """fs_sum_leading, fs_edge_weight_table (+ adjoint) and the folded BatchNorm finalisation against their torch forms."""
    from fissure_segmentation_b200 import _lib
    gen = torch.Generator().manual_seed(12)
    for S, M, N in ((16, 256, 256), (128, 128, 64), (3, 5, 7)):
        part = torch.randn(S, M, N, generator=gen).to(DEV)
        assert_close(ops._sum_chunks(part), part.double().sum(0).float(), 1e-6, 1e-5, "sum over chunks")
    Cp, C = 64, 64
    w = torch.randn(Cp, 2 * C, generator=gen).to(DEV).requires_grad_(True)
    t = ops.edge_weight_table(w, C)
    ref = torch.cat([w[:, :C], w[:, C:] - w[:, :C]], dim=0)
    assert torch.equal(t, ref.detach())
    g = torch.randn(2 * Cp, C, generator=gen).to(DEV)
    (gw,) = torch.autograd.grad(t, w, g)
    (gr,) = torch.autograd.grad(ref, w, g)
    assert torch.equal(gw, gr)
    # folded finalisation == fs_bn_finalize followed by the plain apply (coefficients, output, running statistics)
    rows, Cn = 4096, 256
    x = (torch.randn(rows, Cn, generator=gen) * 2 + 0.5).to(DEV).bfloat16()
    res = []
    for fused in (True, False):
        bn = _bn(Cn, 7).to(DEV).train()
        stats = torch.zeros(int(_lib.load().fs_stats_buffer_doubles(Cn)), dtype=torch.float64, device=DEV)
        _lib.call("fs_colstats", x, x, 1, Cn, rows, Cn, None, 1, stats)
        out = torch.empty_like(x)
        coef = torch.empty(4 * Cn, device=DEV)
        g32, b32 = bn.weight.detach().float(), bn.bias.detach().float()
        if fused:
            _lib.call("fs_bn_act_apply_fin", x, x, 1, Cn, rows, Cn, None, 1, stats, float(rows), g32, b32, bn.eps, bn.momentum,
                      bn.running_mean, bn.running_var, bn.num_batches_tracked, coef, 0.2, out, 1, Cn)
        else:
            _lib.call("fs_bn_finalize", x, stats, float(rows), Cn, g32, b32, bn.eps, bn.momentum, coef, bn.running_mean,
                      bn.running_var, bn.num_batches_tracked)
            _lib.call("fs_bn_act_apply", x, x, 1, Cn, rows, Cn, None, 1, coef, 0.2, out, 1, Cn)
        res.append((out.float(), coef.clone(), bn.running_mean.clone(), bn.running_var.clone(), int(bn.num_batches_tracked)))
    for a, b, name in zip(res[0][:4], res[1][:4], ("out", "coef", "running_mean", "running_var")):
        assert_close(a, b, 1e-6 if name != "out" else 1e-2, 1e-6 if name != "out" else 1e-2, name)
    assert res[0][4] == res[1][4] == 1
