"""CPU-only checks of the host side: C-ABI surface, reference-shaped module interface, no fallback."""
import ctypes
import os
import tempfile

import pytest
import torch

import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import _lib, synth


def test_library_builds_and_exports_every_declared_symbol(lib):
    protos = _lib.parse_header()
    assert len(protos) >= 30
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(raw, name), name
    assert lib.fs_version() >= 100
    assert b"bad argument" in lib.fs_error_string(-1)
    assert b"unsupported" in lib.fs_error_string(-2)


def test_entry_points_reject_bad_arguments_without_launching(lib):
    # null pointers / bad sizes return FS_ERR_BAD_ARG before touching the device
    assert lib.fs_knn3d(0, None, None, 0, 0, 0, 1, 16, 4, 1, 1, None, None) == -1
    assert lib.fs_knnquery(0, None, 4, 4, None, None, None, None, 1, None, None) == -1
    assert lib.fs_nn_points(0, None, None, None, 1, 4, 4, None, None) == -1
    assert lib.fs_edgeconv_gather(0, None, None, 0, 128, None, 1, 4, 2, 64, None, None, None, None, None, None) == -1


def test_no_cpu_fallback():
    m = fs.DGCNNSeg(k=4, in_features=3, num_classes=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.randn(1, 3, 32))
    with pytest.raises(RuntimeError, match="CUDA"):
        fs.knn(torch.randn(1, 3, 32), 4)
    with pytest.raises(RuntimeError, match="CUDA"):
        fs.ChamferLoss()(torch.randn(1, 8, 3), torch.randn(1, 8, 3))


def test_state_dict_config_and_init_stream_match_reference(golden):
    torch.manual_seed(0)
    m = fs.DGCNNSeg(k=20, in_features=3, num_classes=4)
    sd = m.state_dict()
    assert list(sd.keys()) == golden["state_dict_keys"]
    assert m.config == golden["config"]
    # same construction order => same RNG stream => identical initial weights as the reference
    assert torch.equal(sd["ec2.shared_mlp.0.layers.0.weight"], golden["init_seed0_ec2_weight"])
    chk = float(sum(v.double().abs().sum() for v in sd.values() if v.dtype.is_floating_point))
    assert abs(chk - golden["init_seed0_checksum"]) < 1e-6
    assert sum(p.numel() for p in m.parameters()) == 631428
    assert m.num_classes == 4 and m.in_features == 3 and m.k == 20


def test_save_load_roundtrip_cpu():
    m = fs.DGCNNSeg(k=12, in_features=9, num_classes=3, dynamic=False)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "m.pth")
        m.save(path)
        m2 = fs.DGCNNSeg.load(path, "cpu")
    assert m2.config == m.config and not m2.dynamic
    for (k1, v1), (k2, v2) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2)


def test_optional_modules_have_reference_keys():
    m = fs.DGCNNSeg(k=8, in_features=9, num_classes=4, spatial_transformer=True, image_feat_module=True)
    keys = set(m.state_dict().keys())
    assert "spatial_transformer.ec.shared_mlp.1.layers.0.weight" in keys
    assert "spatial_transformer.transform.bias" in keys
    assert "image_feature_module.layers.1.layers.0.weight" in keys
    assert m.in_features == 15
    assert torch.equal(m.spatial_transformer.transform.bias.view(3, 3), torch.eye(3))
    with pytest.raises(ValueError):
        fs.DGCNNSeg(k=8, in_features=3, num_classes=4, image_feat_module=True)


def test_opensrc_dgcnn_keys():
    from types import SimpleNamespace
    from fissure_segmentation_b200.dgcnn_opensrc import DGCNN
    net = DGCNN(SimpleNamespace(k=20, emb_dims=1024, dropout=0.0, static=False), input_channels=3, output_channels=5)
    keys = list(net.state_dict().keys())
    assert "conv1.0.weight" in keys and "bn4.running_var" in keys and "linear3.bias" in keys
    assert sum(p.numel() for p in net.parameters()) == 1800581


def test_synthetic_clouds_are_deterministic_and_lung_shaped():
    a, la = synth.make_batch(3, 2048, seed=7)
    b, lb = synth.make_batch(3, 2048, seed=7)
    assert torch.equal(a, b) and torch.equal(la, lb)
    assert a.shape == (3, 3, 2048) and la.shape == (3, 2048)
    assert float(a.abs().max()) <= 1.25 and int(la.max()) == 3 and int(la.min()) == 0
    c, _ = synth.make_batch(2, 512, seed=7, n_features=6)
    assert c.shape == (2, 9, 512) and float(c[:, 3:].min()) > 0 and float(c[:, 3:].max()) <= 1.0
    lat, _ = synth.make_batch(1, 512, seed=1, augmentation=False)
    D, H, W = synth.SHAPE_DHW
    vox = (lat[0] * torch.tensor([W, H, D]).view(3, 1) + torch.tensor([W - 1, H - 1, D - 1]).view(3, 1)) / 2
    assert torch.allclose(vox, vox.round(), atol=1e-3)       # lattice variant sits on integer voxels


def test_pointops_cuda_module_importable():
    import pointops_cuda
    for name in ("furthestsampling_cuda", "knnquery_cuda", "grouping_forward_cuda", "grouping_backward_cuda",
                 "subtraction_forward_cuda", "subtraction_backward_cuda", "aggregation_forward_cuda",
                 "aggregation_backward_cuda", "interpolation_forward_cuda", "interpolation_backward_cuda"):
        assert callable(getattr(pointops_cuda, name))


def test_flat_data_parallel_single_process_semantics():
    """world size 1, CPU: gradients are moved bucket-wise into the flat buffer (no per-parameter accumulation),
    parameters that received no gradient count as zero, `.grad` views the flat buffer after the step."""
    from fissure_segmentation_b200.ddp import FlatDataParallel
    torch.manual_seed(3)
    used = torch.nn.Linear(5, 4)
    unused = torch.nn.Linear(5, 2)                       # never part of the loss: no gradient arrives
    net = torch.nn.ModuleDict({"used": used, "unused": unused})
    dp = FlatDataParallel(net, n_buckets=2)
    x = torch.randn(7, 5)
    for step in range(2):                                # second step: stale gradients must not leak
        dp.zero_grad()
        assert all(p.grad is None for p in dp.params)
        (used(x) ** 2).sum().backward()
        dp.finish_backward()
    ref = torch.nn.Linear(5, 4)
    ref.load_state_dict(used.state_dict())
    (ref(x) ** 2).sum().backward()
    got = {id(p): dp.flat_grad[o:o + n].view_as(p) for p, (o, n) in zip(dp.params, dp._slices)}
    assert torch.allclose(got[id(used.weight)], ref.weight.grad) and torch.allclose(got[id(used.bias)], ref.bias.grad)
    assert float(got[id(unused.weight)].abs().max()) == 0.0 and float(got[id(unused.bias)].abs().max()) == 0.0
    assert used.weight.grad.data_ptr() == got[id(used.weight)].data_ptr()      # .grad is the flat view
    # parameters are views of one flat buffer: an in-place update of the buffer moves the module's weights
    before = used.weight.detach().clone()
    dp.flat_param.add_(1.0)
    assert torch.allclose(used.weight.detach(), before + 1.0)


def test_statistics_arena_is_zero_filled_and_grows():
    """ops._ZeroArena: slices are disjoint, zero, 16-byte aligned; demand beyond the capacity falls back to a
    private buffer and enlarges the next step's arena (one memset per step instead of one per BatchNorm layer)."""
    from fissure_segmentation_b200 import ops
    ar = ops._ZeroArena()
    dev = torch.device("cpu")
    ar.begin_step(dev)                                   # first step: no capacity yet, every take() is private
    a = ar.take(5, dev)
    b = ar.take(8, dev)
    assert a.numel() == 6 and b.numel() == 8 and ar.buf is None
    a.fill_(1.0)
    ar.begin_step(dev)                                   # capacity follows the previous demand (6 + 8)
    assert ar.cap == 14 and ar.buf.numel() == 14 and float(ar.buf.abs().max()) == 0.0
    c = ar.take(5, dev)
    d = ar.take(8, dev)
    assert c.data_ptr() == ar.buf.data_ptr() and d.data_ptr() == ar.buf.data_ptr() + 6 * 8
    c.fill_(2.0)
    assert float(d.abs().max()) == 0.0                   # disjoint slices
    e = ar.take(4, dev)                                  # beyond the capacity: private zero buffer
    assert e.data_ptr() != ar.buf.data_ptr() and float(e.abs().max()) == 0.0
    ar.begin_step(dev)
    assert ar.cap == 18 and float(ar.buf.abs().max()) == 0.0


def test_tensor_core_knn_host_contract(lib):
    """Host-side entry points of the tcgen05 kNN that need no device: the supported shape range documented in
    include/fissure_b200.h and the workspace size (two fp16 operand tables of nboxes * 64 columns, two 128-slot
    survivor lists, two counts per row, norms and thresholds; every block 256-byte aligned)."""
    ok = lib.fs_knn_feat_tc_supported
    assert ok(32, 2048, 64, 20, 1) == 1 and ok(32, 2048, 64, 20, 0) == 1
    assert ok(2, 64, 64, 5, 1) == 1 and ok(2, 63, 64, 5, 1) == 0                    # N >= 64
    assert ok(2, 2048, 64, 40, 0) == 1 and ok(2, 8192, 64, 40, 0) == 1              # k = 40 static graphs (41 selected)
    assert ok(2, 2048, 64, 64, 1) == 1 and ok(2, 2048, 64, 64, 0) == 0              # k + !self_loop <= 64
    assert ok(2, 2048, 128, 20, 1) == 1 and ok(2, 2048, 256, 20, 1) == 1            # DGCNNReg / dgcnn_opensrc widths
    assert ok(2, 2048, 9, 20, 1) == 0 and ok(2, 2048, 3, 20, 1) == 0                # other widths: exact kernel
    assert ok(1, 32768, 64, 20, 1) == 1 and ok(1, 32769, 64, 20, 1) == 0
    ok3 = lib.fs_knn3d_tc_supported
    assert ok3(32, 2048, 20, 1) == 1 and ok3(8, 8192, 40, 0) == 1 and ok3(1, 63, 8, 1) == 0 and ok3(1, 2048, 64, 0) == 0
    ws = lib.fs_knn_feat_tc_workspace_bytes
    P = 32 * 2048
    need = 2 * P * 128 * 2 + 2 * P * 128 * 4 + P * 4 * 2 + 3 * P * 4 + P
    got = ws(32, 2048, 64, 20)
    assert need <= got <= need + 16 * 256 + 32 * 68 * 4                             # payload + alignment slack + per-cloud block
    assert ws(64, 2048, 64, 20) > got and ws(32, 2048, 64, 20) == got               # monotone, deterministic
    assert ws(1, 64, 64, 5) % 256 == 0
    assert lib.fs_knn_feat_tc_redo_offset(32, 2048, 64, 20) == got - P              # the flags are the last block
    need3 = 2 * P * 64 * 2 + 2 * P * 128 * 4 + P * 4 * 2 + 3 * P * 4 + P * 16 + P
    got3 = lib.fs_knn3d_tc_workspace_bytes(32, 2048, 20)
    assert need3 <= got3 <= need3 + 16 * 256 + 32 * 7 * 4


def test_built_library_contains_blackwell_instructions(lib):
    """SASS evidence (no GPU needed): the tensor-core kNN really issues tcgen05.mma with TMA-staged operands and
    TMEM loads/stores, and the shared-memory gather stages its indices with cp.async - the mnemonics listed in the
    B200 profiling recipe (tcgen05.mma -> UTCHMMA, cp.async.bulk.tensor -> UTMALDG, tcgen05.ld/st -> LDTM/STTM,
    cp.async -> LDGSTS). sm_100a only: no other architecture is embedded."""
    import shutil
    import subprocess
    from fissure_segmentation_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=600).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "LDGSTS", "SYNCS"):
        assert mnemonic in sass, mnemonic
    assert sass.count("UTCHMMA") >= 1
    archs = set(line.split("=")[1].strip() for line in sass.splitlines() if line.strip().startswith("arch ="))
    assert archs == {"sm_100a"}, archs
