"""Secondary model classes of the hot path: DGCNNReg, DGCNNSeg with spatial transformer +
image-feature module, dgcnn_opensrc.DGCNN, against fixtures produced by the unmodified reference
(tests/golden/make_golden_models.py). Parameters are not stored: the same seed must give the reference's initial
weights (checked here on the CPU), followed by the shared deterministic perturbation."""
import os
import sys

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import model_fixtures as MF  # noqa: E402

import fissure_segmentation_b200 as fs  # noqa: E402
from fissure_segmentation_b200 import dgcnn_opensrc  # noqa: E402
from parity import rel_err  # noqa: E402

DEV = "cuda:0"


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(HERE, "golden", "models_golden.pt"), map_location="cpu", weights_only=False)


def build_ours(cfg):
    torch.manual_seed(cfg["seed"])
    if cfg["kind"] == "reg":
        m = fs.DGCNNReg(**cfg["kwargs"])
    elif cfg["kind"] == "seg":
        m = fs.DGCNNSeg(**cfg["kwargs"])
    elif cfg["kind"] == "cls_encoder":
        from fissure_segmentation_b200.folding_net import DGCNN_Cls_Encoder
        m = DGCNN_Cls_Encoder(**cfg["kwargs"])
    else:
        m = dgcnn_opensrc.DGCNN(MF.opensrc_args(cfg), cfg["in_features"], cfg["output_channels"])
    return MF.perturb(m, cfg["seed"])


@pytest.mark.parametrize("tag", list(MF.CONFIGS))
def test_same_seed_gives_the_reference_parameters(gold, tag):
    """No GPU needed: construction order, initialisers and state_dict keys are the reference's."""
    cfg, g = MF.CONFIGS[tag], gold[tag]
    m = build_ours(cfg)
    assert list(m.state_dict().keys()) == g["state_dict_keys"]
    assert abs(MF.checksum(m) - g["param_checksum"]) <= 1e-9 * g["param_checksum"]
    assert abs(float(MF.inputs(cfg).double().abs().sum()) - g["x_checksum"]) <= 1e-9 * g["x_checksum"]


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(MF.CONFIGS))
def test_forward_backward_match_the_reference(gold, tag):
    """fp32, TF32 off. Outputs rtol/atol 1e-3 (the 1024-wide heads amplify the 1e-5 layer-wise deviations), eval
    outputs likewise; gradients: every parameter-gradient norm within 2e-2 relative (+ a noise floor) and cosine >= 0.999 on the
    stored small gradients (arg-max routing is discontinuous: a flipped near-tie moves single entries, see
    DESIGN.md section 3); BatchNorm running statistics rtol 1e-3."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    cfg, g = MF.CONFIGS[tag], gold[tag]
    m = build_ours(cfg).to(DEV)
    if hasattr(m, "precision"):
        m.precision = "fp32"
    x = MF.inputs(cfg).to(DEV)
    m.eval()
    with torch.no_grad():
        out_eval = m(x).cpu()
    assert out_eval.shape == g["out_eval"].shape
    assert torch.allclose(out_eval, g["out_eval"], rtol=1e-3, atol=1e-3), float((out_eval - g["out_eval"]).abs().max())
    m.train()
    out = m(x)
    (out * MF.cotangent(out.shape, cfg["seed"]).to(DEV)).sum().backward()
    assert torch.allclose(out.detach().cpu(), g["out"], rtol=1e-3, atol=1e-3), float((out.detach().cpu() - g["out"]).abs().max())
    grads = {n: p.grad.detach().cpu() for n, p in m.named_parameters() if p.grad is not None}
    assert set(grads) == set(g["grad_norms"])
    # A Linear / Conv bias in front of a BatchNorm has an exactly zero gradient; what the reference stores for it is
    # fp32 rounding noise (e.g. spatial_transformer.mlp.0.bias: 1.5e-3 next to a median norm of ~20). Norms are
    # therefore compared with a floor of 1e-3 x the median gradient norm.
    floor = 1e-3 * float(torch.tensor(list(g["grad_norms"].values())).median())
    bad = {}
    for n, r in g["grad_norms"].items():
        mine = float(grads[n].double().norm())
        if abs(mine - r) > 2e-2 * r + floor:
            bad[n] = (mine, r)
    assert not bad, bad
    for n, ref in g["grads"].items():
        if float(ref.norm()) > 10 * floor:
            cos = float(torch.nn.functional.cosine_similarity(grads[n].flatten().double(), ref.flatten().double(), dim=0))
            assert cos >= 0.999, (n, cos, rel_err(grads[n], ref))
    sd = m.state_dict()
    for n, ref in g["running"].items():
        assert torch.allclose(sd[n].cpu(), ref, rtol=1e-3, atol=1e-5), n
