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
