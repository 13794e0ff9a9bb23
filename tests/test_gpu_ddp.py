"""fs_adam_step_peers on one GPU: the 'peers' are two local gradient buffers, so the sum over ranks and the fused Adam update
can be checked against fs_adam_step on the explicit sum (the real multi-process run over symmetric memory is
tools/check_fused_tail.py under torchrun)."""
import pytest
import torch

from fissure_segmentation_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,offset,world", [(21376, 610052, 2), (1000, 0, 1), (70001, 13, 4)])
def test_adam_step_peers_equals_adam_on_the_sum(lib, n, offset, world):
    gen = torch.Generator().manual_seed(n)
    total = offset + n + 5
    bufs = [torch.randn(total, generator=gen).to(DEV) for _ in range(world)]
    peers = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=DEV)
    p = torch.randn(n, generator=gen).to(DEV)
    m = (0.1 * torch.randn(n, generator=gen)).to(DEV)
    v = (0.01 * torch.rand(n, generator=gen)).to(DEV)
    dyn = torch.tensor([3.0, 1e-3], device=DEV)
    hyper = (1e-3, 0.9, 0.999, 1e-8, 1e-5)
    gsum = bufs[0][offset:offset + n].clone()
    for b in bufs[1:]:
        gsum += b[offset:offset + n]                      # same order as the kernel: rank 0, 1, ...
    p_ref, m_ref, v_ref = p.clone(), m.clone(), v.clone()
    _lib.call("fs_adam_step", p_ref, p_ref, gsum, m_ref, v_ref, n, *hyper, 0, 1.0 / world, dyn)
    out = torch.zeros(n, device=DEV)
    _lib.call("fs_adam_step_peers", p, p, peers, world, offset, m, v, n, *hyper, 0, 1.0 / world, dyn, out)
    assert torch.equal(out, gsum)
    assert torch.equal(p, p_ref) and torch.equal(m, m_ref) and torch.equal(v, v_ref)
    for b in bufs:                                         # nothing is written to a peer
        assert torch.isfinite(b).all()


def test_multi_copy_moves_a_gradient_bucket(lib):
    import ctypes
    gen = torch.Generator().manual_seed(3)
    sizes = [4, 1024 * 192, 7, 256 * 1216, 64 * 6, 1, 100000] + [33] * 40          # more than 32 tensors: two launches
    srcs = [torch.randn(n, generator=gen).to(DEV) for n in sizes]
    flat = torch.zeros(sum(sizes) + 3, device=DEV)
    dsts, off = [], 1                                                                # odd offsets: no alignment assumed
    for n in sizes:
        dsts.append(flat[off:off + n]); off += n
    n = len(srcs)
    sp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
    dp = (ctypes.c_void_p * n)(*[t.data_ptr() for t in dsts])
    cn = (ctypes.c_longlong * n)(*sizes)
    _lib.call("fs_multi_copy_f32", flat, n, sp, dp, cn)
    assert torch.equal(flat[1:1 + sum(sizes)], torch.cat(srcs))
    assert float(flat[0]) == 0 and float(flat[1 + sum(sizes):].abs().sum()) == 0
