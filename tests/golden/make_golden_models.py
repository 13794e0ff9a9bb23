"""Golden fixtures for the secondary model classes (DGCNNReg, DGCNNSeg with spatial transformer + image-feature
module, dgcnn_opensrc.DGCNN, folding_net.DGCNN_Cls_Encoder), produced by the UNMODIFIED reference (/root/reference).

Run in the build container only:  python tests/golden/make_golden_models.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import model_fixtures as MF  # noqa: E402
from oracle import reference_shim  # noqa: E402

OUT = os.path.join(HERE, "models_golden.pt")


def build_reference(cfg, ref_dgcnn, ref_opensrc):
    torch.manual_seed(cfg["seed"])
    if cfg["kind"] == "reg":
        m = ref_dgcnn.DGCNNReg(**cfg["kwargs"])
    elif cfg["kind"] == "seg":
        m = ref_dgcnn.DGCNNSeg(**cfg["kwargs"])
    elif cfg["kind"] == "cls_encoder":
        import models.folding_net as ref_folding          # /root/reference, through the same shims
        m = ref_folding.DGCNN_Cls_Encoder(**cfg["kwargs"])
    else:
        m = ref_opensrc.DGCNN(MF.opensrc_args(cfg), cfg["in_features"], cfg["output_channels"])
    return MF.perturb(m, cfg["seed"])


def main():
    ref_dgcnn, ref_opensrc, _ = reference_shim.load()
    gold = {}
    for tag, cfg in MF.CONFIGS.items():
        m = build_reference(cfg, ref_dgcnn, ref_opensrc)
        x = MF.inputs(cfg)
        entry = {"param_checksum": MF.checksum(m), "x_checksum": float(x.double().abs().sum()),
                 "state_dict_keys": list(m.state_dict().keys())}
        m.eval()
        with torch.no_grad():
            entry["out_eval"] = m(x).clone()
        m.train()
        out = m(x)
        (out * MF.cotangent(out.shape, cfg["seed"])).sum().backward()
        entry["out"] = out.detach().clone()
        entry["grad_norms"] = {n: float(p.grad.double().norm()) for n, p in m.named_parameters() if p.grad is not None}
        entry["grads"] = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None and p.numel() <= 8192}
        entry["running"] = {n: v.clone() for n, v in m.state_dict().items() if "running" in n}
        gold[tag] = entry
        print(tag, "out", tuple(out.shape), "params", sum(p.numel() for p in m.parameters()),
              "max|out|", float(out.abs().max()))
    torch.save(gold, OUT)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
