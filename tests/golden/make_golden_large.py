"""Golden fixtures at the MEASURED configurations, produced by the UNMODIFIED reference (/root/reference):

  config A  DGCNNSeg(k=20, in_features=3, num_classes=4), B=2, N=2048, dynamic and static      (BASELINE configs[0])
  config C  DGCNNSeg(k=40, in_features=9, num_classes=4), B=1, N=8192, static                  (BASELINE configs[2] shape;
            static is what the authors trained, bash_scripts/redo_dgcnn_seg.sh:6-8)

Run in the build container only:  python tests/golden/make_golden_large.py
Stored per run: logits, loss, gradient norms, the small gradients, BatchNorm running statistics and - for the dynamic
run - the three kNN graphs the reference's own `knn` returned inside the network (recorded by wrapping
models.dgcnn.knn; int16, N <= 32767). Parameters and inputs are regenerated from seeds (checksums stored).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from fissure_segmentation_b200 import synth  # noqa: E402
from oracle import dgcnn_oracle as O  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.join(HERE, "large_golden.pt")

CONFIG_A = dict(B=2, N=2048, k=20, in_features=3, num_classes=4, param_seed=101, data_seed=103)
CONFIG_C = dict(B=1, N=8192, k=40, in_features=9, num_classes=4, param_seed=107, data_seed=109)


def make_inputs(cfg):
    return synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)


def make_params(cfg):
    return O.make_params(O.dgcnn_seg_param_shapes(cfg["in_features"], cfg["num_classes"]), cfg["param_seed"])


def checksum(p):
    return float(sum(v.double().abs().sum() for v in p.values() if v.dtype.is_floating_point))


def run(ref_dgcnn, cfg, dynamic, record_graphs=False):
    x, y = make_inputs(cfg)
    p = make_params(cfg)
    model = ref_dgcnn.DGCNNSeg(k=cfg["k"], in_features=cfg["in_features"], num_classes=cfg["num_classes"], dynamic=dynamic)
    model.load_state_dict(p)
    model.train()
    graphs = []
    orig_knn = ref_dgcnn.knn
    if record_graphs:
        def recording_knn(*a, **kw):
            out = orig_knn(*a, **kw)
            graphs.append(out.clone())
            return out
        ref_dgcnn.knn = recording_knn
    try:
        logits = model(x)
    finally:
        ref_dgcnn.knn = orig_knn
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    grads = {n: q.grad.detach().clone() for n, q in model.named_parameters()}
    out = {
        "x_checksum": float(x.double().abs().sum()), "param_checksum": checksum(p),
        "logits": logits.detach().clone(), "loss": loss.detach().clone(),
        "grad_norms": {n: float(g.double().norm()) for n, g in grads.items()},
        "grads": {n: g for n, g in grads.items() if g.numel() <= 64 * 128},
        "running": {n: v.clone() for n, v in model.state_dict().items() if "running" in n or "num_batches" in n},
    }
    if record_graphs:
        out["graphs"] = [g.to(torch.int16) for g in graphs]
    else:
        out["static_graph_rowsum"] = model.knn_graph.sum(-1).to(torch.int32)      # (B, N): pins the static graph cheaply
        out["static_graph"] = model.knn_graph.to(torch.int16)                     # the reference's graph, for teacher forcing
    model.eval()
    with torch.no_grad():
        out["logits_eval"] = model(x).clone()
    return out


def main():
    ref_dgcnn, _, _ = reference_shim.load()
    torch.manual_seed(0)
    gold = {"config_A": CONFIG_A, "config_C": CONFIG_C}
    gold["A_dynamic"] = run(ref_dgcnn, CONFIG_A, True, record_graphs=True)
    assert len(gold["A_dynamic"]["graphs"]) == 3
    gold["A_static"] = run(ref_dgcnn, CONFIG_A, False)
    gold["C_static"] = run(ref_dgcnn, CONFIG_C, False)
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
