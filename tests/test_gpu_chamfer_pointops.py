"""Chamfer loss and pointops operators vs their CPU restatements (both 'parity unpinned' upstream)."""
import numpy as np
import pytest
import torch

import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import ops, synth, pointops_cuda
from oracle import chamfer_oracle, pointops_oracle

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,N,M", [(2, 300, 300), (3, 257, 1000), (1, 2048, 2048)])
def test_chamfer_forward_backward(lib, B, N, M):
    gen = torch.Generator().manual_seed(N + M)
    x = torch.randn(B, N, 3, generator=gen)
    y = torch.randn(B, M, 3, generator=gen)
    xo, yo = x.clone().requires_grad_(True), y.clone().requires_grad_(True)
    ref = chamfer_oracle.chamfer_distance(xo, yo)
    ref.backward()
    xg, yg = x.to(DEV).requires_grad_(True), y.to(DEV).requires_grad_(True)
    loss = fs.chamfer_distance(xg, yg)[0]
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))       # scalar loss rtol 1e-5
    (loss * 3.0).backward()
    assert torch.allclose(xg.grad.cpu(), 3 * xo.grad, rtol=1e-4, atol=1e-7)
    assert torch.allclose(yg.grad.cpu(), 3 * yo.grad, rtol=1e-4, atol=1e-7)
    d, i = ops.nn_points(x.to(DEV), y.to(DEV))
    rd, ri = chamfer_oracle.nn_points(x, y)
    assert torch.equal(i.cpu().long(), ri)
    assert torch.allclose(d.cpu(), rd, rtol=1e-5, atol=1e-7)


def test_chamfer_loss_module_and_properties(lib):
    pred, target = synth.make_chamfer_pair(4, 2048, seed=3)
    loss_fn = fs.ChamferLoss()
    a = loss_fn(pred.to(DEV), target.to(DEV))
    b = loss_fn(pred.transpose(1, 2).to(DEV), target.transpose(1, 2).to(DEV))     # B x 3 x N convention
    assert torch.equal(a, b)
    assert abs(float(a) - float(chamfer_oracle.chamfer_loss(pred, target))) <= 1e-5 * float(a)
    # symmetric, zero on identical clouds, invariant to point order
    assert float(loss_fn(pred.to(DEV), pred.to(DEV))) == 0.0
    assert abs(float(loss_fn(target.to(DEV), pred.to(DEV))) - float(a)) <= 1e-6 * float(a)
    perm = torch.randperm(2048)
    assert abs(float(loss_fn(pred[:, perm].to(DEV), target.to(DEV))) - float(a)) <= 1e-5 * float(a)


def _segments(sizes):
    return torch.tensor(np.cumsum(sizes), dtype=torch.int32)


def test_knnquery_matches_restatement(lib):
    sizes, qsizes = [500, 37, 1200], [500, 37, 1200]
    gen = torch.Generator().manual_seed(1)
    xyz = torch.rand(sum(sizes), 3, generator=gen)
    off = _segments(sizes)
    for nsample in (8, 16, 40):
        idx = torch.zeros(sum(qsizes), nsample, dtype=torch.int32, device=DEV)
        d2 = torch.zeros(sum(qsizes), nsample, device=DEV)
        pointops_cuda.knnquery_cuda(sum(qsizes), nsample, xyz.to(DEV), xyz.to(DEV), off.to(DEV), off.to(DEV), idx, d2)
        ri, rd = pointops_oracle.knnquery(nsample, xyz.numpy(), xyz.numpy(), off.tolist(), off.tolist())
        got_i, got_d = idx.cpu().numpy(), d2.cpu().numpy()
        full = np.all(rd < 1e9, axis=1)
        assert np.array_equal(np.sort(got_i[full], 1), np.sort(ri[full], 1))
        assert np.allclose(got_d, rd, rtol=1e-5, atol=1e-7)
        assert np.all(got_i[:, 0] == np.arange(sum(qsizes)))             # self is the nearest
    # distinct query set (TransitionDown: new_xyz = sampled subset)
    q = xyz[:100].contiguous()
    noff = torch.tensor([100, 100, 100], dtype=torch.int32)              # all queries in segment 0
    idx = torch.zeros(100, 8, dtype=torch.int32, device=DEV)
    d2 = torch.zeros(100, 8, device=DEV)
    pointops_cuda.knnquery_cuda(100, 8, xyz.to(DEV), q.to(DEV), off.to(DEV), noff.to(DEV), idx, d2)
    ri, rd = pointops_oracle.knnquery(8, xyz.numpy(), q.numpy(), off.tolist(), noff.tolist())
    assert np.array_equal(np.sort(idx.cpu().numpy(), 1), np.sort(ri, 1))


def test_furthestsampling_matches_restatement(lib):
    sizes = [700, 64, 1500]
    new_sizes = [175, 16, 375]
    gen = torch.Generator().manual_seed(2)
    xyz = torch.rand(sum(sizes), 3, generator=gen)
    off, noff = _segments(sizes), _segments(new_sizes)
    idx = torch.zeros(sum(new_sizes), dtype=torch.int32, device=DEV)
    tmp = torch.full((sum(sizes),), 1e10, device=DEV)
    pointops_cuda.furthestsampling_cuda(3, max(sizes), xyz.to(DEV), off.to(DEV), noff.to(DEV), tmp, idx)
    ref = pointops_oracle.furthestsampling(xyz.numpy(), off.tolist(), noff.tolist())
    assert np.array_equal(idx.cpu().numpy(), ref)


@pytest.mark.parametrize("sizes,n_max", [([5000, 300], 5000), ([5000, 300], 0), ([9000], 9000), ([1024, 1023, 1025], 1025)])
def test_furthestsampling_variants_agree(lib, sizes, n_max):
    """Register-resident kernel (n_max <= 8192: 1, 2, 4 or 8 points per thread) and the global-memory kernel
    (n_max = 0 or larger segments) pick the same points as the restatement of the upstream operator."""
    new_sizes = [max(1, sz // 4) for sz in sizes]
    gen = torch.Generator().manual_seed(7)
    xyz = torch.rand(sum(sizes), 3, generator=gen)
    off, noff = _segments(sizes), _segments(new_sizes)
    idx = torch.zeros(sum(new_sizes), dtype=torch.int32, device=DEV)
    tmp = torch.full((sum(sizes),), 1e10, device=DEV)
    pointops_cuda.furthestsampling_cuda(len(sizes), n_max, xyz.to(DEV), off.to(DEV), noff.to(DEV), tmp, idx)
    ref = pointops_oracle.furthestsampling(xyz.numpy(), off.tolist(), noff.tolist())
    assert np.array_equal(idx.cpu().numpy(), ref)


def test_grouping_interpolation_subtraction_aggregation(lib):
    gen = torch.Generator().manual_seed(3)
    n, m, ns, c, wc = 200, 150, 8, 16, 4
    feat = torch.randn(n, c, generator=gen).to(DEV)
    idx = torch.randint(0, n, (m, ns), generator=gen, dtype=torch.int32).to(DEV)
    out = torch.zeros(m, ns, c, device=DEV)
    pointops_cuda.grouping_forward_cuda(m, ns, c, feat, idx, out)
    assert torch.equal(out, feat[idx.long()])
    go = torch.randn(m, ns, c, generator=gen).to(DEV)
    gi = torch.zeros(n, c, device=DEV)
    pointops_cuda.grouping_backward_cuda(m, ns, c, go, idx, gi)
    ref = torch.zeros(n, c, device=DEV).index_add_(0, idx.long().view(-1), go.view(-1, c))
    assert torch.allclose(gi, ref, atol=1e-5)

    w = torch.rand(m, 3, generator=gen).to(DEV)
    idx3 = idx[:, :3].contiguous()
    o = torch.zeros(m, c, device=DEV)
    pointops_cuda.interpolation_forward_cuda(m, c, 3, feat, idx3, w, o)
    assert torch.allclose(o, (feat[idx3.long()] * w.unsqueeze(-1)).sum(1), atol=1e-5)

    idxn = torch.randint(0, n, (n, ns), generator=gen, dtype=torch.int32).to(DEV)
    f2 = torch.randn(n, c, generator=gen).to(DEV)
    o = torch.zeros(n, ns, c, device=DEV)
    pointops_cuda.subtraction_forward_cuda(n, ns, c, feat, f2, idxn, o)
    assert torch.allclose(o, feat.unsqueeze(1) - f2[idxn.long()], atol=1e-6)

    pos = torch.randn(n, ns, c, generator=gen).to(DEV)
    wt = torch.randn(n, ns, wc, generator=gen).to(DEV)
    o = torch.zeros(n, c, device=DEV)
    pointops_cuda.aggregation_forward_cuda(n, ns, c, wc, feat, pos, wt, idxn, o)
    ref = ((feat[idxn.long()] + pos) * wt.repeat(1, 1, c // wc)).sum(1)
    assert torch.allclose(o, ref, atol=1e-4)
