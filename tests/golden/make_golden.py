"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference) on seeded inputs.

Run in the build container only:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed here as small fixtures; the
oracle (oracle/dgcnn_oracle.py) is pinned against them by tests/test_oracle_golden.py and the CUDA path
is compared with both.
"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)

from fissure_segmentation_b200 import synth  # noqa: E402
from oracle import dgcnn_oracle as O  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "dgcnn_golden.pt")

SMALL = dict(B=2, N=256, k=8, in_features=3, num_classes=4, param_seed=11, data_seed=5)
FEAT = dict(B=2, N=256, k=8, in_features=9, num_classes=4, param_seed=13, data_seed=6)


def checksum(p):
    return float(sum(v.double().abs().sum() for v in p.values() if v.dtype.is_floating_point))


def run_seg(ref_dgcnn, cfg, dynamic):
    x, y = synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)
    p = O.make_params(O.dgcnn_seg_param_shapes(cfg["in_features"], cfg["num_classes"]), cfg["param_seed"])
    model = ref_dgcnn.DGCNNSeg(k=cfg["k"], in_features=cfg["in_features"], num_classes=cfg["num_classes"], dynamic=dynamic)
    model.load_state_dict(p)
    model.train()
    logits = model(x)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    grads = {n: q.grad.detach().clone() for n, q in model.named_parameters()}
    sd = model.state_dict()
    out = {
        "x_checksum": float(x.double().abs().sum()), "param_checksum": checksum(p),
        "logits": logits.detach().clone(), "loss": loss.detach().clone(),
        "grad_norms": {n: float(g.double().norm()) for n, g in grads.items()},
        "grads": {n: g for n, g in grads.items() if g.numel() <= 64 * 128},
        "running": {n: v.clone() for n, v in sd.items() if "running" in n or "num_batches" in n},
    }
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = model(x).clone()
    return out


def main():
    ref_dgcnn, ref_opensrc, ref_utils = reference_shim.load()
    gold = {"config_small": SMALL, "config_feat": FEAT}

    # --- kNN: utils.general_utils.knn on coordinates and on 64-d features; dgcnn_opensrc.knn
    x, _ = synth.make_batch(2, 256, seed=3, jitter=True)
    gold["knn_x_checksum"] = float(x.double().abs().sum())
    for self_loop in (False, True):
        idx, d = ref_utils.knn(x, 8, self_loop=self_loop, return_dist=True)
        gold[f"knn3d_idx_sl{int(self_loop)}"] = idx.to(torch.int32)
        gold[f"knn3d_dist_sl{int(self_loop)}"] = d
    gold["knn3d_opensrc_idx"] = ref_opensrc.knn(x, 8).to(torch.int32)
    xl, _ = synth.make_batch(2, 256, seed=3, jitter=False)
    gold["knn_lattice_checksum"] = float(xl.double().abs().sum())
    idx, d = ref_utils.knn(xl, 8, self_loop=False, return_dist=True)
    gold["knn3d_lattice_idx"] = idx.to(torch.int32)
    gold["knn3d_lattice_dist"] = d
    gen = torch.Generator().manual_seed(17)
    feat = torch.randn(2, 64, 256, generator=gen)
    idx, d = ref_utils.knn(feat, 8, self_loop=True, return_dist=True)
    gold["knnfeat_idx"] = idx.to(torch.int32)
    gold["knnfeat_dist"] = d
    gold["knnfeat_opensrc_idx"] = ref_opensrc.knn(feat, 8).to(torch.int32)

    # --- single EdgeConv layers with a teacher-forced graph: forward, gradients, running stats
    gen = torch.Generator().manual_seed(23)
    xin = torch.randn(2, 64, 256, generator=gen)
    graph = ref_utils.knn(xin, 8, self_loop=True)
    for tag, widths in (("ec_single", [64]), ("ec_double", [64, 128])):
        ec = ref_dgcnn.EdgeConv(64, widths, 8)
        shapes = []
        cin = 128
        for i, w in enumerate(widths):
            shapes.append((f"shared_mlp.{i}.layers.0.weight", (w, cin, 1, 1)))
            for nm in ("weight", "bias", "running_mean", "running_var"):
                shapes.append((f"shared_mlp.{i}.layers.1.{nm}", (w,)))
            shapes.append((f"shared_mlp.{i}.layers.1.num_batches_tracked", ()))
            cin = w
        p = O.make_params(shapes, 29)
        ec.load_state_dict(p)
        ec.train()
        xr = xin.clone().requires_grad_(True)
        out = ec(xr, graph)
        gen2 = torch.Generator().manual_seed(31)
        gout = torch.randn(out.shape, generator=gen2)
        out.backward(gout)
        gold[tag] = {
            "param_checksum": checksum(p), "x_checksum": float(xin.double().abs().sum()),
            "graph": graph.to(torch.int32), "out": out.detach().clone(), "dx": xr.grad.clone(),
            "grads": {n: q.grad.clone() for n, q in ec.named_parameters()},
            "running": {n: v.clone() for n, v in ec.state_dict().items() if "running" in n},
        }

    # --- DGCNNSeg end to end: dynamic and static, xyz only and xyz + 6 features
    gold["seg_small_dynamic"] = run_seg(ref_dgcnn, SMALL, True)
    gold["seg_small_static"] = run_seg(ref_dgcnn, SMALL, False)
    gold["seg_feat_static"] = run_seg(ref_dgcnn, FEAT, False)

    # --- initialisation stream: same seed => same initial weights as the reference
    torch.manual_seed(0)
    m = ref_dgcnn.DGCNNSeg(k=20, in_features=3, num_classes=4)
    gold["init_seed0_checksum"] = checksum({k: v for k, v in m.state_dict().items()})
    gold["init_seed0_ec2_weight"] = m.state_dict()["ec2.shared_mlp.0.layers.0.weight"].clone()
    gold["state_dict_keys"] = list(m.state_dict().keys())
    gold["config"] = dict(m.config)

    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
