"""DGCNNSeg end to end on the CUDA path vs reference goldens / oracle."""
import os
import tempfile

import pytest
import torch
import torch.nn.functional as F

import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import synth
from oracle import dgcnn_oracle as O
from parity import assert_close, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _build(cfg, dynamic, precision="fp32"):
    x, y = synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)
    p = O.make_params(O.dgcnn_seg_param_shapes(cfg["in_features"], cfg["num_classes"]), cfg["param_seed"])
    m = fs.DGCNNSeg(k=cfg["k"], in_features=cfg["in_features"], num_classes=cfg["num_classes"], dynamic=dynamic).to(DEV)
    m.load_state_dict(p)
    m.precision = precision
    return m, x, y, p


@pytest.mark.parametrize("tag,cfg_key", [("seg_small_static", "config_small"), ("seg_feat_static", "config_feat")])
def test_static_graph_end_to_end_matches_reference(golden, lib, tag, cfg_key):
    """dynamic=False (the configuration the authors trained, bash_scripts/redo_dgcnn_seg.sh:6-8):
    logits, loss, gradients and BatchNorm running statistics within rtol 1e-4 of the reference."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g, cfg = golden[tag], golden[cfg_key]
    m, x, y, p = _build(cfg, dynamic=False)
    m.train()
    logits = m(x.to(DEV))
    assert logits.shape == g["logits"].shape and logits.dtype == torch.float32
    assert_close(logits, g["logits"], 1e-4, 1e-4, tag + " logits")
    loss = F.cross_entropy(logits, y.to(DEV))
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    grads = {n: q.grad for n, q in m.named_parameters()}
    # Gradients are discontinuous in the arg-max routing (max over k, global max-pool over N): ONE flipped
    # near-tie moves a weight gradient by ~1e-3 relative. The reference's own fp32 gradients deviate from
    # an fp64 run of itself by 1.7e-3 on the 9-channel fixture (and ours by 3e-6 there; on the xyz fixture
    # it is the other way round). End to end the check is therefore made against the fp64 oracle with a
    # flip-tolerant bound; the strict rtol-1e-4 gradient checks are the layer-wise ones in
    # test_gpu_edgeconv.py.
    p64 = {n: (v.double().requires_grad_(True) if v.dtype.is_floating_point and "running" not in n
               else (v.double() if v.dtype.is_floating_point else v)) for n, v in p.items()}
    ref64 = O.dgcnn_seg(p64, x.double(), cfg["k"], dynamic=False, training=True)
    F.cross_entropy(ref64, y).backward()
    worst_ours, worst_ref = 0.0, 0.0
    for n, q in grads.items():
        if float(p64[n].grad.norm()) < 1e-12:
            continue                                  # e.g. the bias in front of a BatchNorm: exactly zero
        e = rel_err(q, p64[n].grad)
        worst_ours = max(worst_ours, e)
        assert e < 5e-3, (n, e)
        cos = F.cosine_similarity(q.double().cpu().flatten(), p64[n].grad.flatten(), dim=0)
        assert float(cos) > 0.9999, (n, float(cos))
        if n in g["grads"]:
            worst_ref = max(worst_ref, rel_err(g["grads"][n], p64[n].grad))
    print("%s: worst gradient rel. error vs fp64 oracle: ours %.2e, reference fp32 %.2e" % (tag, worst_ours, worst_ref))
    for n, v in g["running"].items():
        if "num_batches" in n:
            assert int(m.state_dict()[n]) == int(v), n
        else:
            assert_close(m.state_dict()[n], v, 1e-4, 1e-5, n)
    m.eval()
    with torch.no_grad():
        ev = m(x.to(DEV))
    assert_close(ev, g["logits_eval"], 1e-4, 1e-4, tag + " eval logits")


def test_dynamic_graph_end_to_end_reported(golden, lib):
    """dynamic=True: one flipped feature-space neighbour moves the logits chaotically even for the
    reference against itself (SURVEY hard part 1), so this is a reported statistic with a loose bound;
    the strict check is the teacher-forced one below."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g, cfg = golden["seg_small_dynamic"], golden["config_small"]
    m, x, y, p = _build(cfg, dynamic=True)
    m.train()
    logits = m(x.to(DEV))
    diff = (logits.cpu() - g["logits"]).abs()
    frac = float((diff > 1e-4 + 1e-4 * g["logits"].abs()).float().mean())
    print("dynamic e2e: max |dlogit| %.3e, fraction outside rtol 1e-4: %.4f" % (float(diff.max()), frac))
    assert float(diff.max()) < 0.5

    # teacher-forced: feed the oracle's three graphs to both sides -> strict tolerance again
    graphs = []
    pr = {k: v.clone() for k, v in p.items()}
    ref = O.dgcnn_seg(pr, x, cfg["k"], dynamic=True, training=True, graphs_out=graphs)
    assert torch.allclose(ref, g["logits"], rtol=1e-5, atol=1e-5)
    m2, _, _, _ = _build(cfg, dynamic=True)
    m2.train()
    B, _, N = x.shape
    with torch.no_grad():
        from fissure_segmentation_b200 import ops
        xg = x.to(DEV)
        x_pm = ops.to_point_major(xg)
        kg = [ops.KnnGraph.from_reference(t.to(DEV)) for t in graphs]
        x1 = m2.ec1.forward_pm(x_pm, B, N, kg[0])
        x2 = m2.ec2.forward_pm(x1, B, N, kg[1])
        x3 = m2.ec3.forward_pm(x2, B, N, kg[2])
    pr2 = {k: v.clone() for k, v in p.items()}
    o1 = O.edgeconv(x, pr2, "ec1", 2, cfg["k"], graphs[0], True, True)
    o2 = O.edgeconv(o1, pr2, "ec2", 1, cfg["k"], graphs[1], False, True)
    o3 = O.edgeconv(o2, pr2, "ec3", 1, cfg["k"], graphs[2], False, True)
    for got, want, nm in ((x1, o1, "x1"), (x2, o2, "x2"), (x3, o3, "x3")):
        assert_close(got.view(B, N, -1).permute(0, 2, 1), want, 1e-4, 1e-4, nm)


def test_dynamic_graphs_match_oracle_graphs(golden, lib):
    """Layer-wise kNN parity inside the network: neighbour sets of the three dynamic graphs."""
    torch.backends.cuda.matmul.allow_tf32 = False
    from fissure_segmentation_b200 import ops
    from parity import compare_knn
    cfg = golden["config_small"]
    m, x, y, p = _build(cfg, dynamic=True)
    m.train()
    graphs = []
    O.dgcnn_seg({k: v.clone() for k, v in p.items()}, x, cfg["k"], dynamic=True, training=True, graphs_out=graphs)
    B, _, N = x.shape
    with torch.no_grad():
        x_pm = ops.to_point_major(x.to(DEV))
        g1 = m.ec1.build_graph(x_pm, B, N)
        assert torch.equal(g1.idx.cpu().long().sort(-1)[0], graphs[0].sort(-1)[0])
        x1 = m.ec1.forward_pm(x_pm, B, N, g1)
        g2 = m.ec2.build_graph(x1, B, N)
        rep = compare_knn(g2.idx, None, x1.view(B, N, -1).permute(0, 2, 1).cpu(), cfg["k"], True, O.knn_with_gap)
        print("layer-2 graph:", rep)
        assert rep["mismatch_non_tie_rows"] == 0, rep


def test_config_checkpoint_roundtrip_and_reference_state_dict(golden, lib):
    m = fs.DGCNNSeg(k=20, in_features=3, num_classes=4).to(DEV)
    assert list(m.state_dict().keys()) == golden["state_dict_keys"]
    assert m.config == golden["config"]
    clone = type(m)(**m.config)                       # train.py:505
    assert list(clone.state_dict().keys()) == golden["state_dict_keys"]
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "model.pth")
        m.save(path)
        ck = torch.load(path, map_location="cpu", weights_only=False)
        assert set(ck.keys()) == {"config", "model_state"}
        m2 = fs.DGCNNSeg.load(path, DEV).to(DEV)
    x, _ = synth.make_batch(2, 512, seed=2)
    m.eval(); m2.eval()
    with torch.no_grad():
        assert torch.equal(m(x.to(DEV)), m2(x.to(DEV)))


def test_autocast_gradscaler_step_like_model_trainer(lib):
    """model_trainer.py:154-195: autocast forward, GradScaler backward/step; bf16 tables inside."""
    m = fs.DGCNNSeg(k=20, in_features=3, num_classes=4).to(DEV)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)
    scaler = torch.amp.GradScaler("cuda")
    x, y = synth.make_batch(4, 1024, seed=3)
    m.train()
    losses = []
    for _ in range(3):
        with torch.autocast("cuda", dtype=torch.float16):
            out = m(x.to(DEV))
            loss = F.cross_entropy(out, y.to(DEV))
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        opt.zero_grad()
        losses.append(float(loss))
    assert all(l == l for l in losses)
    assert losses[-1] < losses[0]


def test_bf16_mode_no_worse_than_reference_amp(golden, lib):
    """bf16 mode end to end (static graph) vs the fp32 reference logits. Stated tolerance: max |dlogit| at
    most 2x that of the reference's own mixed-precision path (oracle under torch.autocast(bfloat16))."""
    g, cfg = golden["seg_small_static"], golden["config_small"]
    m, x, y, p = _build(cfg, dynamic=False, precision="bf16")
    m.train()
    logits = m(x.to(DEV))
    pc = {n: v.clone().to(DEV) for n, v in p.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        amp = O.dgcnn_seg(pc, x.to(DEV), cfg["k"], dynamic=False, training=True).float().cpu()
    err_ours = float((logits.cpu() - g["logits"]).abs().max())
    err_amp = float((amp - g["logits"]).abs().max())
    print("bf16 DGCNNSeg: max|dlogit| ours %.3e, reference AMP %.3e" % (err_ours, err_amp))
    assert err_ours <= 2 * err_amp + 1e-2


def test_predict_full_pointcloud_and_regression_net(lib):
    m = fs.DGCNNSeg(k=10, in_features=3, num_classes=4).to(DEV).eval()
    pc, _ = synth.make_batch(1, 3000, seed=4)
    with torch.no_grad():
        prob = m.predict_full_pointcloud(pc.to(DEV), sample_points=512, n_runs_min=10)
    assert prob.shape == (1, 4, 3000)
    assert torch.allclose(prob.sum(1), torch.ones(1, 3000, device=DEV), atol=1e-5)
    r = fs.DGCNNReg(k=10, in_features=3, num_classes=7).to(DEV).train()
    out = r(torch.randn(3, 3, 256, device=DEV))
    assert out.shape == (3, 7, 1)
    out.sum().backward()


def test_spatial_transformer_and_image_features_paths(lib):
    m = fs.DGCNNSeg(k=8, in_features=9, num_classes=4, spatial_transformer=True, image_feat_module=True).to(DEV).train()
    x, y = synth.make_batch(2, 300, seed=8, n_features=6)
    out = m(x.to(DEV))
    assert out.shape == (2, 4, 300)
    F.cross_entropy(out, y.to(DEV)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # all-zero cloud (thesis/utils.py:22-23 param_and_op_count) gives finite output
    m.eval()
    with torch.no_grad():
        assert torch.isfinite(m(torch.zeros(1, 9, 300, device=DEV))).all()
