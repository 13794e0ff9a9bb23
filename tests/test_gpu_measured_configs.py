"""Parity at the configurations bench.py measures (BASELINE.json configs A / B / C), against fixtures produced by the
UNMODIFIED reference at those sizes (tests/golden/make_golden_large.py) and against the oracle on the same
activations. Tolerances: fp32 static logits rtol 1e-4 / atol 1e-4; kNN neighbour sets bit-exact on every non-tie row;
teacher-forced layers rtol 1e-4; bf16 mode at most 2x the deviation of the reference's own bf16 autocast path."""
import os

import pytest
import torch
import torch.nn.functional as F

import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import ops, synth
from oracle import dgcnn_oracle as O
from parity import assert_close, compare_knn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LARGE = os.path.join(os.path.dirname(__file__), "golden", "large_golden.pt")


@pytest.fixture(scope="module")
def large():
    return torch.load(LARGE, map_location="cpu", weights_only=False)


def _build(cfg, dynamic, precision="fp32"):
    x, y = synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)
    p = O.make_params(O.dgcnn_seg_param_shapes(cfg["in_features"], cfg["num_classes"]), cfg["param_seed"])
    m = fs.DGCNNSeg(k=cfg["k"], in_features=cfg["in_features"], num_classes=cfg["num_classes"], dynamic=dynamic).to(DEV)
    m.load_state_dict(p)
    m.precision = precision
    return m, x, y, p


def _pm_to_bcn(t, B, N):
    return t.view(B, N, -1).permute(0, 2, 1)


@pytest.mark.parametrize("tag,cfg_key", [("A_static", "config_A"), ("C_static", "config_C")])
def test_static_end_to_end_at_measured_size(large, lib, tag, cfg_key, monkeypatch):
    """Config A (B=2, N=2048, k=20) and config C (B=1, N=8192, k=40, 9 input channels), static graph, fp32, against the
    reference's own run at that size.
    (1) Free-running: the coordinate graph must equal the oracle's on every non-tie row. On tie rows (the gap between
        the k-th and the (k+1)-th neighbour is below fp32 rounding: ~1 % of the rows at N=8192, k=40) either neighbour
        is a correct answer and the reference's own choice depends on its summation order, so the logits of the points
        that see such a row differ: reported, bounded by the tie-row count.
    (2) With the reference's graph teacher-forced: logits / loss / gradients / running statistics / eval logits strict."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g, cfg = large[tag], large[cfg_key]
    m, x, y, p = _build(cfg, dynamic=False)
    assert abs(float(x.double().abs().sum()) - g["x_checksum"]) <= 1e-9 * g["x_checksum"]
    m.train()
    logits = m(x.to(DEV))
    rep = compare_knn(m.knn_graph.to(torch.int32), None, x[:, :3], cfg["k"], False, O.knn_with_gap)
    same_as_ref = (m.knn_graph.cpu().sort(-1)[0] == g["static_graph"].long().sort(-1)[0]).all(-1)
    diff = (logits.detach().cpu() - g["logits"]).abs()
    frac = float((diff > 1e-4 + 1e-4 * g["logits"].abs()).float().mean())
    print("%s free-running: graph %s; rows with the reference's neighbour set %d of %d; logits outside rtol 1e-4: %.4f, max %.2e"
          % (tag, rep, int(same_as_ref.sum()), same_as_ref.numel(), frac, float(diff.max())))
    assert rep["mismatch_non_tie_rows"] == 0, rep
    # A tie row is broken by index, and the model numbers the points along the Morton curve (spatial_sort), the
    # reference in the caller's order: on tie rows the two may keep different (equally near) neighbours, and each such
    # row reaches the points around it through three EdgeConv layers. Measured: 1.9 % of the logits at 1.3 % tie rows
    # (config C), 0 at config A. Bound: 5x the tie-row fraction.
    assert frac <= 5.0 * rep["tie_rows"] / rep["rows"] + 1e-3, (frac, rep)

    # (2) teacher-forced: the model computes its static graph with ops.knn_coords; hand it the reference's graph
    ref_graph = g["static_graph"].to(torch.int32).to(DEV).contiguous()
    real_knn = ops.knn_coords

    def forced(xc, k, self_loop=False, diag_zero=True, return_dist=False):
        if (not self_loop) and k == cfg["k"] and not return_dist and tuple(xc.shape[::2]) == tuple(ref_graph.shape[:2]):
            return ref_graph
        return real_knn(xc, k, self_loop, diag_zero, return_dist)

    monkeypatch.setattr(ops, "knn_coords", forced)
    m, x, y, p = _build(cfg, dynamic=False)
    m.spatial_sort = False                       # the reference's graph is in the caller's point numbering
    m.train()
    logits = m(x.to(DEV))
    assert_close(logits, g["logits"], 1e-4, 1e-4, tag + " logits")
    loss = F.cross_entropy(logits, y.to(DEV))
    assert abs(float(loss) - float(g["loss"])) < 1e-4
    loss.backward()
    # Gradient norms against the reference's (CPU fp32) norms: 5e-3. One quantity is ill-conditioned at B = 2: the
    # gradient of the global feature is the per-cloud SUM of the BatchNorm input gradient of segmentation.0, and those
    # sums cancel exactly across the batch (sum_b = 0 in exact arithmetic), so d beta of global_feature is a difference
    # of rounding-level residues. The yardstick there is the reference module itself in PyTorch eager on this GPU:
    # at most 3x its own deviation from the CPU run.
    ref_dev = {}
    from oracle import reference_shim
    if reference_shim.available():
        ref_dgcnn, _, _ = reference_shim.load()
        rm = ref_dgcnn.DGCNNSeg(k=cfg["k"], in_features=cfg["in_features"], num_classes=cfg["num_classes"], dynamic=False).to(DEV)
        rm.load_state_dict(p)
        rm.train()
        F.cross_entropy(rm(x.to(DEV)), y.to(DEV)).backward()
        for n, q in rm.named_parameters():
            r = g["grad_norms"][n]
            if r >= 1e-9:
                ref_dev[n] = abs(float(q.grad.double().norm()) - r) / r
    devs = []
    for n, q in m.named_parameters():
        r = g["grad_norms"][n]
        if r < 1e-9:
            continue
        devs.append((abs(float(q.grad.double().norm()) - r) / r, n))
    devs.sort(reverse=True)
    print("%s: largest gradient-norm deviations from the reference (CPU): %s" % (tag, [(n, "%.2e" % e) for e, n in devs[:4]]))
    print("%s: the reference module on this GPU deviates by: %s"
          % (tag, sorted(((n, "%.2e" % e) for n, e in ref_dev.items()), key=lambda t: -float(t[1]))[:4]))
    for e, n in devs:
        assert e < max(5e-3, 3.0 * ref_dev.get(n, 0.0)), (n, e, ref_dev.get(n))
    for n, v in g["running"].items():
        if "num_batches" in n:
            assert int(m.state_dict()[n]) == int(v), n
        else:
            assert_close(m.state_dict()[n], v, 1e-4, 1e-5, n)
    m.eval()
    with torch.no_grad():
        ev = m(x.to(DEV))
    assert_close(ev, g["logits_eval"], 1e-4, 1e-4, tag + " eval logits")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_dynamic_graphs_at_config_A(large, lib, precision):
    """Config A dynamic (the CPU reference configuration). Layer by layer: the coordinate graph must equal the
    reference's recorded graph; the feature-space graphs come from the tcgen05 path (fs_knn_feat_tc) and are compared
    (a) with the oracle's kNN on the SAME activations - zero mismatches outside tie rows - and (b), in fp32 mode, with
    the graphs the reference built inside its own forward (flips reported); each layer's output is then checked
    with the reference's graph teacher-forced (rtol 1e-4)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g, cfg = large["A_dynamic"], large["config_A"]
    m, x, y, p = _build(cfg, dynamic=True, precision=precision)
    m.train()
    B, N, k = cfg["B"], cfg["N"], cfg["k"]
    ref_graphs = [t.long() for t in g["graphs"]]
    cdt = torch.float32 if precision == "fp32" else torch.bfloat16
    assert fs._lib.load().fs_knn_feat_tc_supported(B, N, 64, k, 1) == 1
    ops.knn_tc_report = {}
    try:
        with torch.no_grad():
            x_pm = ops.to_point_major(x.to(DEV))
            g1 = m.ec1.build_graph(x_pm, B, N)
            assert torch.equal(g1.idx.cpu().long().sort(-1)[0], ref_graphs[0].sort(-1)[0]), "coordinate graph differs"
            x1 = m.ec1.forward_pm(x_pm, B, N, g1, cdt)
            g2 = m.ec2.build_graph(x1, B, N)
            rep2 = compare_knn(g2.idx, None, _pm_to_bcn(x1.float(), B, N).cpu(), k, True, O.knn_with_gap)
            x2 = m.ec2.forward_pm(x1, B, N, g2, cdt)
            g3 = m.ec3.build_graph(x2, B, N)
            rep3 = compare_knn(g3.idx, None, _pm_to_bcn(x2.float(), B, N).cpu(), k, True, O.knn_with_gap)
        report = dict(ops.knn_tc_report)
    finally:
        ops.knn_tc_report = None
    print("config A dynamic (%s): layer-2 graph %s, layer-3 graph %s, tcgen05 kNN %s" % (precision, rep2, rep3, report))
    assert report.get("channels", []).count(64) == 2, "the feature-space graphs did not come from fs_knn_feat_tc"
    assert rep2["mismatch_non_tie_rows"] == 0 and rep3["mismatch_non_tie_rows"] == 0
    if precision == "fp32":
        flips2 = int((g2.idx.cpu().long().sort(-1)[0] != ref_graphs[1].sort(-1)[0]).any(-1).sum())
        flips3 = int((g3.idx.cpu().long().sort(-1)[0] != ref_graphs[2].sort(-1)[0]).any(-1).sum())
        print("rows whose neighbour set differs from the reference's in-network graph: layer 2 %d, layer 3 %d of %d"
              % (flips2, flips3, B * N))
        assert flips2 <= 0.002 * B * N            # the layer-2 input agrees to ~1e-6: only rounding-level near-ties flip
        # teacher-forced with the reference's graphs: strict per-layer parity at this size
        m2, _, _, _ = _build(cfg, dynamic=True)
        m2.train()
        pr = {n: v.clone() for n, v in p.items()}
        with torch.no_grad():
            kg = [ops.KnnGraph.from_reference(t.to(DEV)) for t in ref_graphs]
            a1 = m2.ec1.forward_pm(x_pm, B, N, kg[0])
            a2 = m2.ec2.forward_pm(a1, B, N, kg[1])
            a3 = m2.ec3.forward_pm(a2, B, N, kg[2])
        o1 = O.edgeconv(x, pr, "ec1", 2, k, ref_graphs[0], True, True)
        o2 = O.edgeconv(o1, pr, "ec2", 1, k, ref_graphs[1], False, True)
        o3 = O.edgeconv(o2, pr, "ec3", 1, k, ref_graphs[2], False, True)
        for got, want, nm in ((a1, o1, "x1"), (a2, o2, "x2"), (a3, o3, "x3")):
            assert_close(_pm_to_bcn(got, B, N), want, 1e-4, 1e-4, nm)
        # end to end, free-running: reported statistic (SURVEY hard part 1)
        logits = m(x.to(DEV)).detach().cpu()
        diff = (logits - g["logits"]).abs()
        print("config A dynamic e2e: max |dlogit| %.3e, fraction outside rtol 1e-4: %.4f"
              % (float(diff.max()), float((diff > 1e-4 + 1e-4 * g["logits"].abs()).float().mean())))
        assert float(diff.max()) < 0.5


def test_bf16_static_at_config_A(large, lib):
    """bf16 mode at config A (static graph): deviation from the fp32 reference logits at most 2x that of the
    reference's own bf16-autocast path (oracle under torch.autocast(bfloat16) on the GPU)."""
    g, cfg = large["A_static"], large["config_A"]
    m, x, y, p = _build(cfg, dynamic=False, precision="bf16")
    m.train()
    logits = m(x.to(DEV)).detach().cpu()
    pc = {n: v.clone().to(DEV) for n, v in p.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        amp = O.dgcnn_seg(pc, x.to(DEV), cfg["k"], dynamic=False, training=True).float().cpu()
    err_ours = float((logits - g["logits"]).abs().max())
    err_amp = float((amp - g["logits"]).abs().max())
    print("config A bf16: max|dlogit| ours %.3e, reference AMP %.3e" % (err_ours, err_amp))
    assert err_ours <= 2 * err_amp + 1e-2


def test_bench_shape_batch32_slice_against_oracle(lib):
    """The benchmarked configuration itself: B=32, N=2048, k=20, dynamic, bf16, train mode. The feature-space graphs of
    a 2-cloud slice are compared with the oracle's kNN on the same activations (kNN is per cloud), and the eval-mode
    logits of the slice (running statistics: clouds independent) with the oracle's eval forward in fp32."""
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, k = 32, 2048, 20
    x, y = synth.make_batch(B, N, seed=1234)
    p = O.make_params(O.dgcnn_seg_param_shapes(3, 4), 77)
    m = fs.DGCNNSeg(k=k, in_features=3, num_classes=4, dynamic=True).to(DEV)
    m.load_state_dict(p)
    m.precision = "bf16"
    m.train()
    ops.knn_tc_report = {}
    try:
        logits = m(x.to(DEV))
        F.cross_entropy(logits, y.to(DEV)).backward()
        assert all(q.grad is not None and torch.isfinite(q.grad).all() for q in m.parameters())
        with torch.no_grad():
            m.spatial_sort = False                      # layer-wise below in the caller's point order
            x_pm = ops.to_point_major(x.to(DEV))
            g1 = m.ec1.build_graph(x_pm, B, N)
            x1 = m.ec1.forward_pm(x_pm, B, N, g1, torch.bfloat16)
            g2 = m.ec2.build_graph(x1, B, N)
            x2 = m.ec2.forward_pm(x1, B, N, g2, torch.bfloat16)
            g3 = m.ec3.build_graph(x2, B, N)
        report = dict(ops.knn_tc_report)
    finally:
        ops.knn_tc_report = None
    sl = slice(0, 2)
    xs = x[sl]
    rep1 = compare_knn(g1.idx[sl], None, xs, k, True, O.knn_with_gap)
    rep2 = compare_knn(g2.idx[sl], None, _pm_to_bcn(x1.float(), B, N)[sl].cpu(), k, True, O.knn_with_gap)
    rep3 = compare_knn(g3.idx[sl], None, _pm_to_bcn(x2.float(), B, N)[sl].cpu(), k, True, O.knn_with_gap)
    print("bench shape, 2-cloud slice: graphs", rep1, rep2, rep3, "tcgen05 kNN", report)
    assert rep1["mismatch_non_tie_rows"] == 0 and rep2["mismatch_non_tie_rows"] == 0 and rep3["mismatch_non_tie_rows"] == 0
    assert report["channels"].count(64) == 4 and report["channels"].count(3) == 2      # every graph on the tcgen05 path
    assert report["redo_rows"] <= 0.01 * report["rows"]
    # eval mode, fp32: per-cloud independent -> the 2-cloud slice of a B=32 forward equals the oracle on 2 clouds
    # The bench clouds sit on the voxel lattice: exact distance ties are common and are broken by index, so the
    # comparison keeps the caller's point numbering on both sides (spatial_sort off; the sorted path is covered by
    # the kNN parity tests and by the free-running checks above).
    m.precision = "fp32"
    m.spatial_sort = False
    m.eval()
    with torch.no_grad():
        ev = m(x.to(DEV))[sl].cpu()
        # the graphs of the eval forward (same kernels, same numbering), handed to the oracle: on the lattice a tie row
        # may legitimately keep another equally near neighbour, which would otherwise spread through three layers
        e1 = m.ec1.build_graph(x_pm, B, N)
        y1 = m.ec1.forward_pm(x_pm, B, N, e1)
        e2 = m.ec2.build_graph(y1, B, N)
        y2 = m.ec2.forward_pm(y1, B, N, e2)
        e3 = m.ec3.build_graph(y2, B, N)
        gs = [t.idx[sl].cpu().long() for t in (e1, e2, e3)]
        sd = {n: v.detach().cpu().clone() for n, v in m.state_dict().items()}
        ref = O.dgcnn_seg(sd, xs, k, dynamic=True, training=False, fixed_graphs=gs)
    diff = (ev - ref).abs()
    frac = float((diff > 1e-4 + 1e-4 * ref.abs()).float().mean())
    print("bench shape eval slice: max |dlogit| %.3e, fraction outside rtol 1e-4: %.5f" % (float(diff.max()), frac))
    assert frac == 0.0 and float(diff.max()) < 1e-3


def test_tensor_core_knn_hostile_inputs(lib):
    """fs_knn_feat_tc on inputs that defeat the approximate selection: all rows equal (the all-zero cloud the reference
    forwards before training, thesis/utils.py:22-23), duplicated rows, a NaN and an Inf row, features quantised to
    1/64. Non-tie rows must have the oracle's neighbour set; rows handed to the exact kernel are counted."""
    B, N, C, k = 2, 2048, 64, 20
    gen = torch.Generator().manual_seed(5)
    base = F.leaky_relu(torch.randn(B, C, N, generator=gen), 0.2)
    cases = {}
    cases["all_equal"] = torch.zeros(B, C, N)
    cases["constant_rows"] = torch.full((B, C, N), 0.37)
    dup = base.clone()
    dup[:, :, 1::2] = dup[:, :, 0::2]                    # every point twice
    dup[:, :, :64] = dup[:, :, :1]                       # and one point 64 times (> k duplicates)
    cases["duplicated"] = dup
    cases["quantised"] = torch.round(base * 64) / 64
    nanrow = base.clone()
    nanrow[0, 3, 17] = float("nan")
    nanrow[1, 5, 99] = float("inf")
    cases["nan_inf_row"] = nanrow
    for name, feat in cases.items():
        pm = ops.to_point_major(feat.to(DEV)).contiguous()
        ops.knn_tc_report = {}
        try:
            idx = ops.knn_features(pm, B, N, k, self_loop=True)
            torch.cuda.synchronize()
            report = dict(ops.knn_tc_report)
        finally:
            ops.knn_tc_report = None
        assert report.get("calls") == 1, name
        ic = idx.cpu().long()
        assert int(ic.min()) >= 0 and int(ic.max()) < N, name
        if name == "nan_inf_row":
            # rows that do not involve the poisoned points must still be exact; compare on the finite sub-problem
            # through the exact kernel (same contract: fs_knn_feat), which shares the NaN semantics
            exact, _ = ops.knn_features(pm, B, N, k, self_loop=True, return_dist=True)
            same = (ic.sort(-1)[0] == exact.cpu().long().sort(-1)[0]).all(-1)
            finite_rows = torch.isfinite(feat).all(1)
            bad = int((~same & finite_rows).sum())
            print("hostile/%s: redo rows %d of %d, finite rows differing from the exact kernel: %d"
                  % (name, report["redo_rows"], report["rows"], bad))
            assert bad <= 4          # rows whose k-th neighbour distance ties with the distance to the Inf point
            continue
        rep = compare_knn(idx, None, feat, k, True, O.knn_with_gap)
        print("hostile/%s: %s, redo rows %d of %d" % (name, rep, report["redo_rows"], report["rows"]))
        assert rep["mismatch_non_tie_rows"] == 0, (name, rep)
        assert bool((ic.sort(-1)[0][..., 1:] != ic.sort(-1)[0][..., :-1]).all()), name + ": duplicate indices in a row"


def test_public_dense_helpers_match_oracle(lib):
    """pairwise_dist / create_neighbor_features (models/dgcnn.py:15-36) / get_graph_feature
    (models/dgcnn_opensrc.py:43-66): the dense tensors the reference API exposes, built on the CUDA kNN."""
    from fissure_segmentation_b200 import dgcnn_opensrc
    from fissure_segmentation_b200.knn import create_neighbor_features, pairwise_dist
    x, _ = synth.make_batch(2, 512, seed=31, n_features=6, jitter=True)
    xd = x.to(DEV)
    d = pairwise_dist(xd.transpose(2, 1))
    assert_close(d, O.pairwise_sqdist(x.transpose(2, 1)), 1e-5, 1e-5, "pairwise_dist")
    for coords_only in (True, False):
        e = create_neighbor_features(xd, 12, knn_only_over_coords=coords_only)
        assert torch.equal(e.cpu(), O.edge_features(x, 12, coords_only=coords_only)), "create_neighbor_features"
    graph = O.knn(x[:, :3], 12, self_loop=False)
    e = create_neighbor_features(xd, 12, fixed_knn_graph=graph.to(DEV))
    assert torch.equal(e.cpu(), O.edge_features(x, 12, idx=graph))
    gf = dgcnn_opensrc.get_graph_feature(xd, k=12)
    assert torch.equal(gf.cpu(), O.graph_feature(x, 12))
    gf = dgcnn_opensrc.get_graph_feature(xd, k=12, idx=graph.to(DEV))
    assert torch.equal(gf.cpu(), O.graph_feature(x, 12, idx=graph))


def test_predict_full_pointcloud_matches_oracle_loop(lib):
    """models/point_seg_net.py:21-48 with the permutations drawn from the same seeded CUDA generator: the batched
    implementation (one B=n forward per phase + one scatter kernel, CUDA graph) against the oracle's sequential loop
    over the oracle's eval forward. Static-graph model so that the comparison is strict (atol 2e-4 on probabilities)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    k, S, runs = 10, 512, 20
    p = O.make_params(O.dgcnn_seg_param_shapes(3, 4), 9, random_bn=True)
    m = fs.DGCNNSeg(k=k, in_features=3, num_classes=4, dynamic=False).to(DEV).eval()
    m.load_state_dict(p)
    m.precision = "fp32"
    # 1500 points, 16 subset runs of 512 (coverage ~99.6 %): the left-over phase is exercised as well
    pc, _ = synth.make_batch(1, 1500, seed=4, jitter=True)

    def cuda_randperm(n):
        return torch.randperm(n, device=DEV).cpu()

    def oracle_forward(xs):
        return O.dgcnn_seg(p, xs, k, dynamic=False, training=False)

    for graph_mode in (True, False):
        m.inference_cuda_graph = graph_mode
        torch.manual_seed(123)
        with torch.no_grad():
            prob = m.predict_full_pointcloud(pc.to(DEV), sample_points=S, n_runs_min=runs)
            if graph_mode:                                   # second call replays the captured graph
                torch.manual_seed(123)
                prob2 = m.predict_full_pointcloud(pc.to(DEV), sample_points=S, n_runs_min=runs)
                assert torch.allclose(prob, prob2, atol=1e-6)
        torch.manual_seed(123)
        want = O.predict_full_pointcloud(oracle_forward, pc, 4, sample_points=S, n_runs_min=runs, randperm=cuda_randperm)
        assert prob.shape == (1, 4, 1500)
        diff = float((prob.cpu() - want).abs().max())
        print("predict_full_pointcloud (cuda graph %s): max |dp| vs oracle loop %.2e" % (graph_mode, diff))
        assert diff < 2e-4
    # training mode falls back to the reference's sequential loop (batch statistics couple the clouds)
    m.train()
    torch.manual_seed(5)
    with torch.no_grad():
        pt = m.predict_full_pointcloud(pc.to(DEV), sample_points=S, n_runs_min=5)
    assert torch.allclose(pt.sum(1), torch.ones(1, 1500, device=DEV), atol=1e-5)


def test_auto_precision_follows_autocast(lib):
    """precision='auto' (the default the reference trainer gets): bf16 tables inside torch.autocast
    (model_trainer.py:75-76, 157), fp32 outside. The dtype is observed at the dense head's BatchNorm input."""
    m = fs.DGCNNSeg(k=8, in_features=3, num_classes=4).to(DEV).train()
    assert m.precision == "auto"
    x, _ = synth.make_batch(2, 256, seed=3)
    seen = []
    orig = ops.bn_act

    def spy(t, *a, **kw):
        seen.append(t.dtype)
        return orig(t, *a, **kw)

    ops.bn_act = spy
    try:
        with torch.autocast("cuda", dtype=torch.float16):
            out = m(x.to(DEV))
        inside = list(seen)
        seen.clear()
        out2 = m(x.to(DEV))
        outside = list(seen)
    finally:
        ops.bn_act = orig
    assert inside and all(d == torch.bfloat16 for d in inside), inside
    assert outside and all(d == torch.float32 for d in outside), outside
    assert out.dtype == torch.float32 and out2.dtype == torch.float32


def test_cuda_graph_replay_equals_eager_step(lib):
    """bench.py captures the whole training step in a CUDA graph: two replays on the same input must reproduce the
    eager step (same kernels, same arena memset). fp32 mode and a static graph, so that the only run-to-run variation
    is the summation order of the atomics (fp64 batch statistics, fp32 routed scatter); with dynamic graphs and bf16
    tiles a 1e-7 difference can flip a near-tie neighbour or a bf16 rounding and move single logits by 1e-2."""
    torch.manual_seed(0)
    B, N, k = 4, 2048, 20
    x, y = synth.make_batch(B, N, seed=77, jitter=True)
    xd, yd = x.to(DEV), y.to(DEV)
    m = fs.DGCNNSeg(k=k, in_features=3, num_classes=4, dynamic=False).to(DEV).train()
    m.precision = "fp32"
    for q in m.modules():
        if isinstance(q, torch.nn.modules.batchnorm._BatchNorm):
            q.momentum = 0.0                              # running statistics frozen: every step sees the same state
    out_static = torch.empty(B, 4, N, device=DEV)
    gw = torch.empty_like(m.ec2.shared_mlp[0].layers[0].weight)

    def step():
        for q in m.parameters():
            q.grad = None
        logits = m(xd)
        F.cross_entropy(logits, yd).backward()
        out_static.copy_(logits.detach())
        gw.copy_(m.ec2.shared_mlp[0].layers[0].weight.grad)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    eager_out, eager_gw = out_static.clone(), gw.clone()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        step()
    for _ in range(2):
        out_static.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert float((out_static - eager_out).abs().max()) < 1e-4
        assert float((gw - eager_gw).norm() / eager_gw.norm()) < 1e-4
