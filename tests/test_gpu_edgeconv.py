"""Fused EdgeConv forward/backward vs the reference goldens and the oracle, teacher-forced graphs."""
import pytest
import torch

import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import ops
from oracle import dgcnn_oracle as O
from parity import assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _shapes(cin2, widths):
    shapes = []
    for i, w in enumerate(widths):
        shapes.append((f"shared_mlp.{i}.layers.0.weight", (w, cin2, 1, 1)))
        for nm in ("weight", "bias", "running_mean", "running_var"):
            shapes.append((f"shared_mlp.{i}.layers.1.{nm}", (w,)))
        shapes.append((f"shared_mlp.{i}.layers.1.num_batches_tracked", ()))
        cin2 = w
    return shapes


@pytest.mark.parametrize("tag,widths", [("ec_single", [64]), ("ec_double", [64, 128])])
def test_edgeconv_matches_reference_golden(golden, lib, tag, widths):
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden[tag]
    gen = torch.Generator().manual_seed(23)
    xin = torch.randn(2, 64, 256, generator=gen)
    ec = fs.EdgeConv(64, widths, 8).to(DEV)
    ec.load_state_dict(O.make_params(_shapes(128, widths), 29))
    ec.train()
    xr = xin.to(DEV).requires_grad_(True)
    out = ec(xr, g["graph"].long().to(DEV))
    assert_close(out, g["out"], 1e-4, 1e-5, tag + " forward")
    gen2 = torch.Generator().manual_seed(31)
    out.backward(torch.randn(out.shape, generator=gen2).to(DEV))
    assert_close(xr.grad, g["dx"], 1e-4, 1e-5, tag + " dx")
    for n, gr in g["grads"].items():
        got = dict(ec.named_parameters())[n].grad
        assert rel_err(got, gr) < 1e-4, (n, rel_err(got, gr))
        assert_close(got, gr, 1e-3, 1e-4 * float(gr.abs().max()), tag + " " + n)
    for n, v in g["running"].items():
        assert_close(ec.state_dict()[n], v, 1e-4, 1e-6, tag + " " + n)
    assert int(ec.state_dict()["shared_mlp.0.layers.1.num_batches_tracked"]) == 1


@pytest.mark.parametrize("C,Cp,N,k,training", [(64, 64, 2048, 20, True), (64, 128, 512, 40, True), (128, 256, 300, 16, True),
                                               (3, 64, 1024, 20, True), (64, 64, 777, 20, False)])
def test_single_layer_edgeconv_vs_oracle(lib, C, Cp, N, k, training):
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 2
    gen = torch.Generator().manual_seed(C + Cp + N)
    x = torch.randn(B, C, N, generator=gen)
    graph = O.knn(x, k, self_loop=True)
    p = O.make_params(_shapes(2 * C, [Cp]), 41)
    ec = fs.EdgeConv(C, [Cp], k).to(DEV)
    ec.load_state_dict(p)
    ec.train(training)
    xr = x.to(DEV).requires_grad_(True)
    out = ec(xr, graph.to(DEV))
    po = {"ec." + n: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in n else v)
          for n, v in p.items()}
    xo = x.clone().requires_grad_(True)
    stats = {}
    ref = O.edgeconv(xo, po, "ec", 1, k, graph, False, training, stats)
    assert_close(out, ref, 1e-4, 2e-5, "forward")
    gout = torch.randn(ref.shape, generator=gen)
    out.backward(gout.to(DEV))
    ref.backward(gout)
    assert rel_err(xr.grad, xo.grad) < 1e-4
    assert_close(xr.grad, xo.grad, 1e-3, 1e-4 * float(xo.grad.abs().max()), "dx")
    for n, q in ec.named_parameters():
        assert rel_err(q.grad, po["ec." + n].grad) < 2e-4, (n, rel_err(q.grad, po["ec." + n].grad))
    if training:
        for n, v in stats.items():
            assert_close(ec.state_dict()[n[3:]], v, 1e-4, 1e-6, n)


def test_edgeconv_bf16_mode_no_worse_than_reference_amp(lib):
    """bf16 mode (bf16 edge tensors and second-layer GEMM of a two-layer EdgeConv; per-point tables stay
    fp32). Stated tolerance: deviation from the fp32 oracle at most 2x the deviation of the reference's
    own mixed-precision path (the oracle under torch.autocast(bfloat16) on the same GPU), for the output
    (max abs) and for the input gradient (1 - cosine)."""
    B, C, Cp, N, k = 2, 64, 128, 1024, 20
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(B, C, N, generator=gen)
    graph = O.knn(x, k, self_loop=True)
    p = O.make_params(_shapes(2 * C, [64, Cp]), 43)
    ec = fs.EdgeConv(C, [64, Cp], k).to(DEV)
    ec.load_state_dict(p)
    ec.precision = "bf16"
    ec.train()
    xr = x.to(DEV).requires_grad_(True)
    out = ec(xr, graph.to(DEV))
    gout = torch.randn(out.shape, generator=gen)
    out.backward(gout.to(DEV))

    def oracle(device, amp):
        po = {"ec." + n: (v.clone().to(device).requires_grad_(True) if v.dtype.is_floating_point and "running" not in n
                          else v.clone().to(device)) for n, v in p.items()}
        xo = x.clone().to(device).requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            ref = O.edgeconv(xo, po, "ec", 2, k, graph.to(device), False, True, None)
        ref.float().backward(gout.to(device))
        return ref.float().detach().cpu(), xo.grad.float().cpu()

    ref32, g32 = oracle("cpu", False)
    amp, gamp = oracle(DEV, True)
    cos = lambda a, b: float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))   # noqa: E731
    err_ours, err_amp = float((out.cpu() - ref32).abs().max()), float((amp - ref32).abs().max())
    dcos_ours, dcos_amp = 1 - cos(xr.grad.cpu(), g32), 1 - cos(gamp, g32)
    print("bf16 EdgeConv: max|out-ref32| ours %.3e, reference AMP %.3e; 1-cos(dx) ours %.2e, AMP %.2e"
          % (err_ours, err_amp, dcos_ours, dcos_amp))
    assert err_ours <= 2 * err_amp + 1e-3
    assert dcos_ours <= 2 * dcos_amp + 1e-4


def test_edgeconv_is_permutation_equivariant(lib):
    """Size-independent property at the full bench size: relabelling the points permutes the output."""
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, k = 4, 2048, 20
    x = torch.randn(B, 64, N, device=DEV)
    ec = fs.EdgeConv(64, [64], k).to(DEV).eval()
    perm = torch.randperm(N, device=DEV)
    with torch.no_grad():
        a = ec(x)
        b = ec(x[:, :, perm])
    assert torch.allclose(a[:, :, perm], b, rtol=1e-4, atol=1e-5)


def test_reverse_graph_is_consistent(lib):
    B, N, k = 3, 500, 12
    idx = torch.stack([torch.stack([torch.randperm(N)[:k] for _ in range(N)]) for _ in range(B)]).to(torch.int32).to(DEV)
    g = ops.KnnGraph(idx)
    rev_ptr, rev_src = g.reverse()
    rev_ptr, rev_src = rev_ptr.cpu().long(), rev_src.cpu().long()
    assert int(rev_ptr[0]) == 0 and int(rev_ptr[-1]) == B * N * k
    indeg = torch.zeros(B * N, dtype=torch.long)
    glob = (idx.cpu().long() + (torch.arange(B) * N).view(B, 1, 1)).reshape(-1)
    indeg.index_add_(0, glob, torch.ones_like(glob))
    assert torch.equal(rev_ptr[1:] - rev_ptr[:-1], indeg)
    # every (src -> tgt) edge appears exactly once in tgt's list
    src_of_edge = torch.arange(B * N).repeat_interleave(k)
    fwd = torch.stack([glob, src_of_edge], 1)
    tgt_of_slot = torch.repeat_interleave(torch.arange(B * N), indeg)
    rev = torch.stack([tgt_of_slot, rev_src], 1)
    key = lambda t: (t[:, 0] * (B * N) + t[:, 1]).sort()[0]   # noqa: E731
    assert torch.equal(key(fwd), key(rev))


@pytest.mark.parametrize("widths,N,k", [([64, 64], 2048, 20), ([64, 128], 700, 16)])
def test_two_layer_edgeconv_on_coordinates_vs_oracle(lib, widths, N, k):
    """ec1 / spatial-transformer EdgeConv on raw xyz (csrc/edge3.cu: statistics from the moments of the 6-D edge
    vectors, weight gradient without scatter): strict layer-wise tolerances against the oracle."""
    torch.backends.cuda.matmul.allow_tf32 = False
    from fissure_segmentation_b200 import synth
    B = 2
    x, _ = synth.make_batch(B, N, seed=77, jitter=True)
    graph = O.knn(x, k, self_loop=True)
    p = O.make_params(_shapes(6, widths), 47)
    ec = fs.EdgeConv(3, widths, k, first_layer=True).to(DEV)
    ec.load_state_dict(p)
    ec.train()
    out = ec(x.to(DEV), graph.to(DEV))
    po = {"ec." + n: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in n else v)
          for n, v in p.items()}
    stats = {}
    ref = O.edgeconv(x, po, "ec", 2, k, graph, True, True, stats)
    assert_close(out, ref, 1e-4, 2e-5, "forward")
    gen = torch.Generator().manual_seed(3)
    gout = torch.randn(ref.shape, generator=gen)
    out.backward(gout.to(DEV))
    ref.backward(gout)
    for n, q in ec.named_parameters():
        assert rel_err(q.grad, po["ec." + n].grad) < 2e-4, (n, rel_err(q.grad, po["ec." + n].grad))
    for n, v in stats.items():
        assert_close(ec.state_dict()[n[3:]], v, 1e-4, 1e-6, n)
    assert int(ec.state_dict()["shared_mlp.0.layers.1.num_batches_tracked"]) == 1


@pytest.mark.parametrize("B,N,k,Cp", [(2, 256, 8, 64), (32, 256, 8, 64), (2, 2048, 20, 64), (3, 1000, 20, 64),
                                      (2, 4096, 40, 64), (1, 8192, 40, 64), (3, 700, 7, 128), (2, 300, 33, 256)])
def test_smem_gather_is_bit_identical_to_global_gather(lib, B, N, k, Cp):
    """The shared-memory-resident gather (edgeconv_smem.cu: channel-sliced table per CTA, cp.async index staging,
    small-batch point split, N = 8192 falls back) must reproduce the global-memory kernel: sel / arg / eval outputs
    bit for bit, the edge sums and the fp64 statistics up to summation order."""
    import os
    from fissure_segmentation_b200 import _lib
    torch.manual_seed(B * 1000 + N + k)
    P = B * N
    idx = torch.stack([torch.stack([torch.randperm(N, device=DEV)[:k] for _ in range(N)]) for _ in range(B)]).int().contiguous()
    rev_ptr, _ = ops.KnnGraph(idx).reverse()
    table = torch.randn(P, 2 * Cp, device=DEV)
    gamma = torch.randn(Cp, device=DEV)             # both signs: max and min channels
    coef = torch.randn(4 * Cp, device=DEV)
    res = {}
    try:
        for mode in ("global", "smem"):
            os.environ["FS_GATHER"] = mode
            sel = torch.empty(P, Cp, device=DEV)
            arg = torch.empty(P, Cp, dtype=torch.uint8, device=DEV)
            sy = torch.empty(P, Cp, device=DEV)
            st = torch.zeros(_lib.load().fs_stats_buffer_doubles(Cp), dtype=torch.float64, device=DEV)
            _lib.call("fs_edgeconv_gather", table, table, 0, table.stride(0), idx, B, N, k, Cp, gamma, rev_ptr, sel, arg, sy, st)
            out = torch.empty(P, Cp, device=DEV)
            out16 = torch.empty(P, Cp, device=DEV, dtype=torch.bfloat16)
            arg2 = torch.empty_like(arg)
            _lib.call("fs_edgeconv_fused_eval", table, table, 0, table.stride(0), idx, B, N, k, Cp, coef, out, 0, out.stride(0), arg2)
            _lib.call("fs_edgeconv_fused_eval", table, table, 0, table.stride(0), idx, B, N, k, Cp, coef, out16, 1, out16.stride(0), None)
            torch.cuda.synchronize()
            res[mode] = (sel, arg, sy, st[:3 * Cp].clone(), out, arg2, out16)
    finally:
        os.environ.pop("FS_GATHER", None)
    a, b = res["global"], res["smem"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])                    # selected value and slot
    assert torch.equal(a[4], b[4]) and torch.equal(a[5], b[5]) and torch.equal(a[6], b[6])   # eval mode
    assert torch.allclose(a[2], b[2], rtol=1e-5, atol=1e-5)                       # sum over the k neighbours
    assert torch.allclose(a[3], b[3], rtol=1e-6, atol=1e-6)                       # fp64 batch statistics + pivots
