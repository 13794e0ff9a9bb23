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


def test_edgeconv_bf16_tables_within_stated_tolerance(lib):
    """bf16 mode (bf16 edge tensors / second-layer GEMM of a two-layer EdgeConv; the per-point tables stay
    fp32) vs the fp32 oracle: rtol 3e-2 / atol 3e-2 on outputs, cosine >= 0.999 on gradients."""
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
    xo = x.clone().requires_grad_(True)
    po = {"ec." + n: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in n else v)
          for n, v in p.items()}
    ref = O.edgeconv(xo, po, "ec", 2, k, graph, False, True, None)
    assert_close(out, ref, 3e-2, 3e-2, "bf16 forward")
    gout = torch.randn(ref.shape, generator=gen)
    out.backward(gout.to(DEV))
    ref.backward(gout)
    cos = torch.nn.functional.cosine_similarity(xr.grad.cpu().flatten(), xo.grad.flatten(), dim=0)
    assert float(cos) >= 0.999, float(cos)


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
