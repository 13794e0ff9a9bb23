"""Out-of-bounds writes and races, checked without compute-sanitizer (closed on the GPU pool).

* Guard bands: every CUDA tensor that the package allocates from Python (outputs, workspaces, statistics arenas) while
  the test runs is carved out of a larger buffer with 4 KiB of 0xA5 on both sides; after a full training step, an eval
  forward and the stand-alone kNN / pooled-layer entry points the bands must be untouched. A kernel that writes before
  or past one of its output or scratch tensors fails here (reads are not covered).
* Races: the entry points whose results do not go through floating-point atomics must be bit-identical when repeated
  (kNN graphs on both tensor-core paths, the fused two-layer EdgeConv forward, eval logits, the pooled layer's
  backward). A shared-memory or TMEM hand-over race shows up as a run-to-run difference.
"""
import pytest
import torch
import torch.nn.functional as F

import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
PAD = 4096
FILL = 0xA5


class GuardedAllocations:
    """Context manager: torch.empty / zeros (+ _like) calls made from Python for CUDA tensors return views into guarded
    buffers. Library code that allocates in C++ (GEMM outputs, autograd) is unaffected."""

    def __init__(self):
        self.records = []
        self.orig = {}

    def _is_cuda(self, device):
        return device is not None and torch.device(device).type == "cuda"

    def _alloc(self, shape, dtype, device, zero):
        dtype = dtype or torch.get_default_dtype()
        numel = 1
        for s in shape:
            numel *= int(s)
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        body = (nbytes + 15) // 16 * 16
        base = self.orig["empty"](body + 2 * PAD, dtype=torch.uint8, device=device)
        base[:PAD] = FILL
        base[PAD + nbytes:] = FILL
        self.records.append((base, nbytes))
        view = base[PAD:PAD + nbytes].view(dtype).view(*shape) if numel else self.orig["empty"](*shape, dtype=dtype, device=device)
        if zero and numel:
            view.zero_()
        return view

    @staticmethod
    def _shape(args):
        if len(args) == 1 and isinstance(args[0], (tuple, list, torch.Size)):
            return tuple(args[0])
        return tuple(args)

    def __enter__(self):
        self.orig = {"empty": torch.empty, "zeros": torch.zeros, "empty_like": torch.empty_like,
                     "zeros_like": torch.zeros_like}

        def make(name, zero):
            def fn(*args, dtype=None, device=None, **kw):
                if not self._is_cuda(device) or kw:
                    return self.orig[name](*args, dtype=dtype, device=device, **kw)
                return self._alloc(self._shape(args), dtype, device, zero)
            return fn

        def make_like(name, zero):
            def fn(t, *args, **kw):
                if args or kw or not t.is_cuda or not t.is_contiguous():
                    return self.orig[name](t, *args, **kw)
                return self._alloc(tuple(t.shape), t.dtype, t.device, zero)
            return fn

        torch.empty, torch.zeros = make("empty", False), make("zeros", True)
        torch.empty_like, torch.zeros_like = make_like("empty_like", False), make_like("zeros_like", True)
        return self

    def __exit__(self, *exc):
        torch.empty, torch.zeros = self.orig["empty"], self.orig["zeros"]
        torch.empty_like, torch.zeros_like = self.orig["empty_like"], self.orig["zeros_like"]
        return False

    def check(self):
        torch.cuda.synchronize()
        bad = 0
        for base, nbytes in self.records:
            if not bool((base[:PAD] == FILL).all()) or not bool((base[PAD + nbytes:] == FILL).all()):
                bad += 1
        assert bad == 0, "%d of %d guarded allocations had their guard bands overwritten" % (bad, len(self.records))
        return len(self.records)


def _model(k, in_features, dynamic, precision, N, B, seed=3):
    torch.manual_seed(seed)
    m = fs.DGCNNSeg(k=k, in_features=in_features, num_classes=4, dynamic=dynamic).to(DEV)
    m.precision = precision
    x, y = synth.make_batch(B, N, seed=11, n_features=in_features - 3, jitter=True)
    return m, x.to(DEV), y.to(DEV)


@pytest.mark.parametrize("N,k,in_features,dynamic,precision", [(2048, 20, 3, True, "bf16"), (1000, 20, 3, True, "fp32"),
                                                                (1536, 40, 9, False, "bf16"), (300, 8, 3, True, "bf16")])
def test_guard_bands_survive_a_training_step_and_an_eval_forward(lib, N, k, in_features, dynamic, precision):
    ops._workspaces.clear()                   # cached workspaces would bypass the guarded allocator
    with GuardedAllocations() as guard:
        m, x, y = _model(k, in_features, dynamic, precision, N, 2)
        m.train()
        F.cross_entropy(m(x), y).backward()
        m.eval()
        with torch.no_grad():
            out = m(x)
        assert torch.isfinite(out).all()
        n = guard.check()
    ops._workspaces.clear()
    assert n > 50                              # the step really went through the guarded allocator


def test_guard_bands_knn_ragged_and_hostile(lib):
    ops._workspaces.clear()
    gen = torch.Generator().manual_seed(5)
    with GuardedAllocations() as guard:
        for N, k in ((64, 8), (1000, 20), (2049, 20), (4100, 40)):
            x = torch.randn(2, 3, N, generator=gen).to(DEV)
            ops.knn_coords(x, k, self_loop=True)
            f = torch.randn(2 * N, 64, generator=gen).to(DEV)
            ops.knn_features(f, 2, N, k, self_loop=True)
        z = torch.zeros(1, 3, 2048, device=DEV)          # every score ties: all rows take the exact re-do path
        ops.knn_coords(z, 20, self_loop=True)
        ops.knn_features(torch.zeros(2048, 64, device=DEV), 1, 2048, 20, self_loop=True)
        guard.check()
    ops._workspaces.clear()


def test_repeatable_results_where_no_float_atomics_are_involved(lib):
    """Bitwise repeatability = no hand-over race in the tcgen05 / shared-memory pipelines."""
    gen = torch.Generator().manual_seed(8)
    x = torch.randn(4, 3, 2048, generator=gen).to(DEV)
    f = torch.randn(4 * 2048, 64, generator=gen).to(DEV)
    first = None
    for _ in range(4):
        cur = (ops.knn_coords(x, 20, self_loop=True).clone(), ops.knn_features(f, 4, 2048, 20, self_loop=True).clone())
        if first is None:
            first = cur
        else:
            assert torch.equal(cur[0], first[0]) and torch.equal(cur[1], first[1])
    m, xb, _ = _model(20, 3, True, "bf16", 2048, 4)
    m.eval()
    with torch.no_grad():
        ref = m(xb).clone()
        for _ in range(3):
            assert torch.equal(m(xb), ref)


def test_pooled_layer_backward_is_deterministic(lib):
    B, N, K, C = 4, 1024, 192, 1024
    gen = torch.Generator().manual_seed(2)
    x = torch.relu(torch.randn(B * N, K, generator=gen)).to(DEV).bfloat16()
    w = (torch.randn(C, K, generator=gen) / K ** 0.5).to(DEV)
    gout = torch.randn(B, C, generator=gen).to(DEV)
    grads = []
    for _ in range(3):
        bn = torch.nn.BatchNorm1d(C).to(DEV)
        xg, wg = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
        ops.pool_linear_bn_act(xg, wg, bn, 0.2, B, N).float().backward(gout)
        grads.append((xg.grad.clone(), wg.grad.clone()))
    for gx, gw in grads[1:]:
        assert torch.equal(gx, grads[0][0])
        assert torch.equal(gw, grads[0][1])
