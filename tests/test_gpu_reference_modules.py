"""The UNMODIFIED reference modules (staged under baseline/_ref by tools/stage_reference.py) running on the GPU box
next to this package: the reference's own `pointops.py` drives this package's `pointops_cuda` kernels,
the reference PointTransformer runs end to end on them (BASELINE configs[3] shape), and the reference DGCNNSeg on
CUDA is compared directly with the B200 module."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import pointops_cuda as pc
from fissure_segmentation_b200 import synth
from oracle import pointops_oracle as PO
from oracle import reference_shim
from parity import assert_close

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not reference_shim.available(), reason="reference files not staged (baseline/_ref)")]
DEV = "cuda:0"


def _segments(B, n_each, seed):
    gen = torch.Generator().manual_seed(seed)
    xyz = torch.rand(sum(n_each), 3, generator=gen)
    offset = torch.tensor(np.cumsum(n_each), dtype=torch.int32)
    return xyz, offset


def test_reference_pointops_py_runs_on_our_pointops_cuda(lib):
    """models/pointtransformer/pointops.py:16-62 (FurthestSampling, KNNQuery) unchanged, `import pointops_cuda`
    resolving to this package: results equal the numpy restatement of the upstream operators."""
    ref_pointops, _ = reference_shim.load_pointtransformer()
    import pointops_cuda as top
    assert top.knnquery_cuda is pc.knnquery_cuda
    xyz, offset = _segments(3, [700, 1024, 333], 3)
    new_offset = torch.tensor(np.cumsum([175, 256, 83]), dtype=torch.int32)
    idx = ref_pointops.furthestsampling(xyz.to(DEV), offset.to(DEV), new_offset.to(DEV))
    want = PO.furthestsampling(xyz.numpy(), offset.numpy(), new_offset.numpy())
    assert np.array_equal(idx.cpu().numpy(), want)
    new_xyz = xyz[idx.cpu().long()]
    pc.clear_knn_cache()
    got_i, got_d = ref_pointops.knnquery(16, xyz.to(DEV), new_xyz.to(DEV), offset.to(DEV), new_offset.to(DEV))
    wi, wd = PO.knnquery(16, xyz.numpy(), new_xyz.numpy(), offset.numpy(), new_offset.numpy())
    assert np.array_equal(np.sort(got_i.cpu().numpy(), 1), np.sort(wi, 1))
    assert np.allclose(got_d.cpu().numpy(), np.sqrt(wd), rtol=1e-5, atol=1e-6)      # pointops.py:60 returns sqrt


def test_reference_point_transformer_end_to_end_with_knn_cache(lib):
    """The reference's PointTransformerCompatibility (models/pointtransformer/seg_model.py) forward + backward on our
    kernels; the per-level kNN cache (seg_model.py:38-39 issues every query twice) must not change any value."""
    _, ref_seg = reference_shim.load_pointtransformer()
    x, y = synth.make_batch(2, 1024, seed=12, jitter=True)
    x, y = x.to(DEV), y.to(DEV)
    outs, grads = [], []
    for slots in (0, 8):
        pc.KNN_CACHE_SLOTS = slots
        pc.clear_knn_cache()
        torch.manual_seed(0)
        model = ref_seg.PointTransformerCompatibility(in_features=3, num_classes=4).to(DEV).train()
        out = model(x)
        assert out.shape == (2, 4, 1024) and bool(torch.isfinite(out).all())
        F.cross_entropy(out, y).backward()
        outs.append(out.detach().clone())
        grads.append(torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]))
        stats = dict(pc.knn_cache_stats)
    pc.KNN_CACHE_SLOTS = 8
    print("PointTransformer kNN cache:", stats)
    assert stats["hits"] > stats["misses"] > 0
    assert torch.equal(outs[0], outs[1])
    assert float((grads[0] - grads[1]).norm() / grads[0].norm()) < 1e-5      # atomics in the backward scatter


@pytest.mark.parametrize("dynamic", [False, True])
def test_b200_module_against_the_reference_module_on_the_same_gpu(lib, dynamic):
    """Same state_dict loaded into the reference's DGCNNSeg (PyTorch eager, CUDA, fp32) and into the B200 module, same
    input: logits rtol 1e-4 (static) / reported (dynamic), running statistics, reference checkpoint round trip."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ref_dgcnn, _, _ = reference_shim.load()
    torch.manual_seed(3)
    ref = ref_dgcnn.DGCNNSeg(k=20, in_features=3, num_classes=4, dynamic=dynamic).to(DEV).train()
    ours = fs.DGCNNSeg(k=20, in_features=3, num_classes=4, dynamic=dynamic).to(DEV).train()
    ours.load_state_dict(ref.state_dict())                      # reference checkpoints load unchanged
    ours.precision = "fp32"
    x, y = synth.make_batch(4, 2048, seed=8, jitter=True)
    x, y = x.to(DEV), y.to(DEV)
    lr = ref(x)
    lo = ours(x)
    diff = (lo - lr).abs()
    print("vs reference module on GPU (dynamic=%s): max |dlogit| %.3e" % (dynamic, float(diff.max())))
    if not dynamic:
        assert_close(lo, lr, 1e-4, 1e-4, "logits vs reference module")
        for (n, a), (_, b) in zip(ours.state_dict().items(), ref.state_dict().items()):
            if "running" in n:
                assert_close(a, b, 1e-4, 1e-5, n)
    else:
        assert float((diff > 1e-4 + 1e-4 * lr.abs()).float().mean()) < 0.05


def test_linear1_forward_hook_sees_the_reference_global_feature(lib):
    """models/dg_ssm.py:41 (MultiHeadDGCNN) hangs extra regression heads on the INPUT of `linear1`, captured with a
    forward hook. Same hook on the reference dgcnn_opensrc.DGCNN and on the B200 twin, same weights and input: the
    captured (B, 2*emb_dims) global feature and the main output agree, and a loss on the captured feature back-propagates
    into the EdgeConv weights of the twin."""
    from types import SimpleNamespace
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    _, ref_opensrc, _ = reference_shim.load()
    from fissure_segmentation_b200.dgcnn_opensrc import DGCNN
    args = SimpleNamespace(k=20, emb_dims=256, dropout=0.0, static=True)
    torch.manual_seed(5)
    ref = ref_opensrc.DGCNN(args, 3, 10).to(DEV).eval()
    ours = DGCNN(args, 3, 10).to(DEV).eval()
    ours.load_state_dict(ref.state_dict())
    ours.precision = "fp32"
    seen = {}
    ref.linear1.register_forward_hook(lambda m, inp, out: seen.__setitem__("ref", inp[0]))
    ours.linear1.register_forward_hook(lambda m, inp, out: seen.__setitem__("ours", inp[0]))
    x, _ = synth.make_batch(3, 1024, seed=21, jitter=True)
    x = x.to(DEV)
    with torch.no_grad():
        out_ref = ref(x)
    out = ours(x)
    assert seen["ours"].shape == seen["ref"].shape == (3, 2 * args.emb_dims)
    assert_close(seen["ours"], seen["ref"], 1e-4, 1e-4, "hooked global feature")
    assert_close(out, out_ref, 1e-4, 1e-4, "main head")
    ours.zero_grad()
    seen["ours"].square().mean().backward()
    g = ours.conv1[0].weight.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0
