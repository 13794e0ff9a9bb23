"""Pins the oracle (oracle/dgcnn_oracle.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_golden.py). CPU only."""
import torch
import torch.nn.functional as F

from fissure_segmentation_b200 import synth
from oracle import dgcnn_oracle as O


def _checksum(p):
    return float(sum(v.double().abs().sum() for v in p.values() if v.dtype.is_floating_point))


def test_inputs_reproduce(golden):
    x, _ = synth.make_batch(2, 256, seed=3, jitter=True)
    assert abs(float(x.double().abs().sum()) - golden["knn_x_checksum"]) < 1e-9
    xl, _ = synth.make_batch(2, 256, seed=3, jitter=False)
    assert abs(float(xl.double().abs().sum()) - golden["knn_lattice_checksum"]) < 1e-9


def test_knn_matches_reference(golden):
    x, _ = synth.make_batch(2, 256, seed=3, jitter=True)
    for sl in (False, True):
        idx, d = O.knn(x, 8, self_loop=sl, return_dist=True)
        assert torch.equal(idx.to(torch.int32), golden[f"knn3d_idx_sl{int(sl)}"])
        assert torch.equal(d, golden[f"knn3d_dist_sl{int(sl)}"])
    assert torch.equal(O.knn_opensrc(x, 8).to(torch.int32), golden["knn3d_opensrc_idx"])
    gen = torch.Generator().manual_seed(17)
    feat = torch.randn(2, 64, 256, generator=gen)
    idx, d = O.knn(feat, 8, self_loop=True, return_dist=True)
    assert torch.equal(idx.to(torch.int32), golden["knnfeat_idx"])
    assert torch.equal(O.knn_opensrc(feat, 8).to(torch.int32), golden["knnfeat_opensrc_idx"])


def _edgeconv_shapes(widths):
    shapes, cin = [], 128
    for i, w in enumerate(widths):
        shapes.append((f"shared_mlp.{i}.layers.0.weight", (w, cin, 1, 1)))
        for nm in ("weight", "bias", "running_mean", "running_var"):
            shapes.append((f"shared_mlp.{i}.layers.1.{nm}", (w,)))
        shapes.append((f"shared_mlp.{i}.layers.1.num_batches_tracked", ()))
        cin = w
    return shapes


def test_edgeconv_layers_match_reference(golden):
    gen = torch.Generator().manual_seed(23)
    xin = torch.randn(2, 64, 256, generator=gen)
    for tag, widths in (("ec_single", [64]), ("ec_double", [64, 128])):
        g = golden[tag]
        p = O.make_params(_edgeconv_shapes(widths), 29)
        assert abs(_checksum(p) - g["param_checksum"]) < 1e-6
        p = {"ec." + k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
             for k, v in p.items()}
        xr = xin.clone().requires_grad_(True)
        stats = {}
        out = O.edgeconv(xr, p, "ec", len(widths), 8, g["graph"].long(), False, True, stats)
        assert torch.allclose(out, g["out"], rtol=1e-6, atol=1e-6)
        gen2 = torch.Generator().manual_seed(31)
        out.backward(torch.randn(out.shape, generator=gen2))
        assert torch.allclose(xr.grad, g["dx"], rtol=1e-5, atol=1e-6)
        for n, gr in g["grads"].items():
            assert torch.allclose(p["ec." + n].grad, gr, rtol=1e-5, atol=1e-6), n
        for n, v in g["running"].items():
            assert torch.allclose(stats["ec." + n], v, rtol=1e-6, atol=1e-7), n


def _run_seg(cfg, dynamic):
    x, y = synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)
    p = O.make_params(O.dgcnn_seg_param_shapes(cfg["in_features"], cfg["num_classes"]), cfg["param_seed"])
    return x, y, p


def test_dgcnn_seg_matches_reference(golden):
    for tag, cfg_key, dynamic in (("seg_small_dynamic", "config_small", True), ("seg_small_static", "config_small", False),
                                  ("seg_feat_static", "config_feat", False)):
        g = golden[tag]
        cfg = golden[cfg_key]
        x, y, p = _run_seg(cfg, dynamic)
        assert abs(float(x.double().abs().sum()) - g["x_checksum"]) < 1e-9
        assert abs(_checksum(p) - g["param_checksum"]) < 1e-6
        pr = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v)
              for k, v in p.items()}
        stats = {}
        logits = O.dgcnn_seg(pr, x, cfg["k"], dynamic=dynamic, training=True, stats_out=stats)
        assert torch.allclose(logits, g["logits"], rtol=1e-5, atol=1e-5), tag
        loss = F.cross_entropy(logits, y)
        assert abs(float(loss) - float(g["loss"])) < 1e-5
        loss.backward()
        for n, gr in g["grads"].items():
            assert torch.allclose(pr[n].grad, gr, rtol=1e-4, atol=1e-6), (tag, n)
        for n, v in g["grad_norms"].items():
            assert abs(float(pr[n].grad.double().norm()) - v) <= 1e-4 * max(v, 1e-6) + 1e-7, (tag, n)
        for n, v in g["running"].items():
            if "num_batches" in n:
                continue
            assert torch.allclose(stats[n], v, rtol=1e-5, atol=1e-6), (tag, n)
        with torch.no_grad():
            ev = O.dgcnn_seg(p | stats, x, cfg["k"], dynamic=dynamic, training=False)
        assert torch.allclose(ev, g["logits_eval"], rtol=1e-5, atol=1e-5), tag


# ---------------------------------------------------------------------------------------------------------------
# The measured configurations (tests/golden/make_golden_large.py): config A (B=2, N=2048, k=20) dynamic + static and
# config C (B=1, N=8192, k=40, 9 channels) static, produced by the unmodified reference at full size.
# ---------------------------------------------------------------------------------------------------------------
import os  # noqa: E402

import pytest  # noqa: E402

LARGE = os.path.join(os.path.dirname(__file__), "golden", "large_golden.pt")


@pytest.fixture(scope="module")
def large():
    return torch.load(LARGE, map_location="cpu", weights_only=False)


@pytest.mark.parametrize("tag,cfg_key,dynamic", [("A_dynamic", "config_A", True), ("A_static", "config_A", False),
                                                 ("C_static", "config_C", False)])
def test_oracle_matches_reference_at_measured_configs(large, tag, cfg_key, dynamic):
    g, cfg = large[tag], large[cfg_key]
    x, y = synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)
    p = O.make_params(O.dgcnn_seg_param_shapes(cfg["in_features"], cfg["num_classes"]), cfg["param_seed"])
    assert abs(float(x.double().abs().sum()) - g["x_checksum"]) <= 1e-9 * g["x_checksum"]
    assert abs(_checksum(p) - g["param_checksum"]) <= 1e-9 * g["param_checksum"]
    graphs, stats = [], {}
    logits = O.dgcnn_seg(p, x, cfg["k"], dynamic=dynamic, training=True, stats_out=stats, graphs_out=graphs)
    if dynamic:
        # the three graphs the reference's own knn() returned inside its forward, bit for bit
        for got, want in zip(graphs, g["graphs"]):
            assert torch.equal(got, want.long())
    else:
        assert torch.equal(graphs[0].sum(-1).to(torch.int32), g["static_graph_rowsum"])
    assert torch.allclose(logits, g["logits"], rtol=1e-6, atol=1e-6)
    assert abs(float(F.cross_entropy(logits, y)) - float(g["loss"])) < 1e-6
    for n, v in stats.items():
        assert torch.allclose(v, g["running"][n], rtol=1e-6, atol=1e-7), n
    with torch.no_grad():
        ev = O.dgcnn_seg({**p, **stats}, x, cfg["k"], dynamic=dynamic, training=False)
    assert torch.allclose(ev, g["logits_eval"], rtol=1e-5, atol=1e-5)


def test_oracle_predict_full_pointcloud_properties():
    """Restatement of models/point_seg_net.py:21-48: probabilities sum to one, every point is covered, and the result
    is the softmax of the accumulated per-run probabilities (checked with a forward that ignores its input)."""
    torch.manual_seed(0)
    const = torch.tensor([2.0, 0.0, -1.0, 0.5])

    def forward(xs):
        return const.view(1, 4, 1).expand(xs.shape[0], 4, xs.shape[-1]).clone()

    pc = torch.randn(1, 3, 700)
    prob = O.predict_full_pointcloud(forward, pc, 4, sample_points=256, n_runs_min=10)
    assert prob.shape == (1, 4, 700)
    assert torch.allclose(prob.sum(1), torch.ones(1, 700), atol=1e-6)
    assert bool((prob.argmax(1) == 0).all())          # every point was visited at least once


def test_staged_reference_copy_is_the_reference():
    """baseline/_ref (git-ignored, travels to the GPU box) holds byte-for-byte copies of the reference files."""
    import filecmp
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import stage_reference as S
    if not os.path.isdir(S.REFERENCE_ROOT):
        pytest.skip("reference tree not present on this machine")
    assert S.stage(verbose=False)
    for rel in S.FILES:
        assert filecmp.cmp(os.path.join(S.REFERENCE_ROOT, rel), os.path.join(S.DEST, rel), shallow=False), rel
