"""Shared between tests/golden/make_golden_models.py (runs the UNMODIFIED reference) and the GPU parity tests:
model configurations, deterministic inputs and a deterministic perturbation of the freshly initialised models, so
that no parameter tensor has to be stored in the fixture (the same seed gives the reference's initial weights:
tests/test_host_api.py checks the streams are identical)."""
import types
import zlib

import torch

CONFIGS = {
    # models/dgcnn.py:165-209 - regression net, static graph
    "reg_static": dict(kind="reg", seed=41, B=4, N=256, in_features=3, data_seed=51,
                       kwargs=dict(k=8, in_features=3, num_classes=5, dynamic=False)),
    # models/dgcnn.py:61-112, 246-343 - spatial transformer + image-feature module in front of DGCNNSeg
    "seg_st_imf_static": dict(kind="seg", seed=43, B=2, N=256, in_features=9, data_seed=53,
                              kwargs=dict(k=8, in_features=9, num_classes=4, spatial_transformer=True, dynamic=False,
                                          image_feat_module=True)),
    # models/dgcnn_opensrc.py:101-179 - the open-source DGCNN used by DG-SSM / affine DGCNN (dropout 0: deterministic)
    "opensrc_static": dict(kind="opensrc", seed=47, B=4, N=256, in_features=3, data_seed=57,
                           args=dict(k=8, emb_dims=128, dropout=0.0, static=True), output_channels=5),
    # models/folding_net.py:83-141 - PC-AE encoder (DGCNN_Cls_Encoder), static coordinate graph and dynamic graphs
    "cls_encoder_static": dict(kind="cls_encoder", seed=59, B=4, N=256, in_features=3, data_seed=61,
                               kwargs=dict(k=8, n_embedding=128, static=True)),
    "cls_encoder_dynamic": dict(kind="cls_encoder", seed=67, B=2, N=512, in_features=3, data_seed=71,
                                kwargs=dict(k=12, n_embedding=256, static=False)),
}


def opensrc_args(cfg):
    return types.SimpleNamespace(**cfg["args"])


def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)


def perturb(model, seed):
    """BatchNorm affine parameters of both signs and non-trivial running statistics; a non-zero last layer of the
    spatial transformer (its reference init is the identity map, models/dgcnn.py:276-279). Keyed by parameter NAME,
    so the result does not depend on module iteration order."""
    sd = model.state_dict()
    new = {}
    for name, v in sd.items():
        prefix, leaf = name.rsplit(".", 1) if "." in name else ("", name)
        is_bn = (prefix + ".running_mean") in sd
        g = _gen(name, seed)
        if is_bn and leaf == "weight":
            new[name] = torch.randn(v.shape, generator=g)
        elif is_bn and leaf == "bias":
            new[name] = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf == "running_mean":
            new[name] = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf == "running_var":
            new[name] = 0.5 + torch.rand(v.shape, generator=g)
        elif name == "spatial_transformer.transform.weight":
            new[name] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            new[name] = v
    model.load_state_dict(new)
    return model


def checksum(model):
    return float(sum(v.double().abs().sum() for v in model.state_dict().values() if v.dtype.is_floating_point))


def inputs(cfg):
    from fissure_segmentation_b200 import synth
    x, _ = synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)
    return x


def cotangent(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(1000 + seed))
