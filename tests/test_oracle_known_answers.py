"""Known-answer tests for the two oracles whose third-party originals (pytorch3d, pointops_cuda) are not available:
hand-computed cases and the properties their documented behaviour implies (losses/chamfer_loss.py:9-20 with the
pytorch3d defaults quoted at train_pc_ae.py:85; models/pointtransformer/pointops.py:16-62). They do not replace a
pin against the originals - the oracles stay "parity unpinned" - but they fix what the restatements compute."""
import numpy as np
import torch

from oracle import chamfer_oracle as C
from oracle import pointops_oracle as P


def test_chamfer_hand_computed():
    # x = {(0,0,0), (1,0,0)}, y = {(0,0,0), (0,2,0), (3,0,0)}
    x = torch.tensor([[[0., 0., 0.], [1., 0., 0.]]])
    y = torch.tensor([[[0., 0., 0.], [0., 2., 0.], [3., 0., 0.]]])
    # x -> y: min d2 = 0, 1            -> mean 0.5
    # y -> x: min d2 = 0, 4, 4         -> mean 8/3
    assert abs(float(C.chamfer_distance(x, y)) - (0.5 + 8.0 / 3.0)) < 1e-6
    d, i = C.nn_points(x, y)
    assert d.tolist() == [[0.0, 1.0]] and i.tolist() == [[0, 0]]
    # batch mean: a second pair of identical clouds contributes 0
    x2 = torch.cat([x, torch.zeros(1, 2, 3)]); y2 = torch.cat([y, torch.zeros(1, 3, 3)])
    assert abs(float(C.chamfer_distance(x2, y2)) - 0.5 * (0.5 + 8.0 / 3.0)) < 1e-6


def test_chamfer_properties_and_gradient():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 40, 3, generator=g, requires_grad=True)
    y = torch.randn(3, 55, 3, generator=g)
    loss = C.chamfer_distance(x, y)
    assert float(loss) > 0 and abs(float(C.chamfer_distance(y, x.detach())) - float(loss)) < 1e-6      # symmetric
    assert float(C.chamfer_distance(y, y)) == 0.0
    t = torch.tensor([0.3, -1.0, 2.0])
    assert abs(float(C.chamfer_distance(x.detach() + t, y + t)) - float(loss)) < 1e-5                 # translation
    # gradient of the x -> y term: 2 (x_i - y_nn(i)) / (N B); the y -> x term adds 2 (x_i - y_j) / (M B) for every
    # y_j whose nearest point is x_i
    loss.backward()
    B, N, M = 3, 40, 55
    d2 = ((x.detach().unsqueeze(2) - y.unsqueeze(1)) ** 2).sum(-1)
    nn_xy = d2.argmin(2); nn_yx = d2.argmin(1)
    ref = 2 * (x.detach() - torch.gather(y, 1, nn_xy.unsqueeze(-1).expand(-1, -1, 3))) / (N * B)
    contrib = 2 * (torch.gather(x.detach(), 1, nn_yx.unsqueeze(-1).expand(-1, -1, 3)) - y) / (M * B)
    ref = ref.scatter_add(1, nn_yx.unsqueeze(-1).expand(-1, -1, 3), contrib)
    assert torch.allclose(x.grad, ref, atol=1e-6)
    # the module's transposition rule (B x 3 x N inputs)
    assert abs(float(C.chamfer_loss(x.detach().transpose(1, 2), y.transpose(1, 2))) - float(loss)) < 1e-6


def test_knnquery_hand_computed_segments():
    # two segments: {0,1,2} on the x axis and {3,4} on the y axis; queries are the points themselves
    xyz = np.array([[0, 0, 0], [1, 0, 0], [3, 0, 0], [0, 0, 0], [0, 5, 0]], dtype=np.float32)
    off = np.array([3, 5], dtype=np.int32)
    idx, d2 = P.knnquery(2, xyz, xyz, off, off)
    assert idx.tolist() == [[0, 1], [1, 0], [2, 1], [3, 4], [4, 3]]          # self first, indices global, per segment
    assert d2.tolist() == [[0, 1], [0, 1], [0, 4], [0, 25], [0, 25]]
    # nsample larger than a segment: padded with the segment's first index and 1e10 (callers never ask for it)
    idx, d2 = P.knnquery(3, xyz, xyz, off, off)
    assert idx[3].tolist() == [3, 4, 3] and d2[3, 2] == np.float32(1e10)


def test_furthestsampling_hand_computed():
    # one segment on a line: start at the first point, then always the point furthest from the chosen set
    xyz = np.array([[0, 0, 0], [1, 0, 0], [10, 0, 0], [4, 0, 0], [6, 0, 0]], dtype=np.float32)
    out = P.furthestsampling(xyz, np.array([5], dtype=np.int32), np.array([4], dtype=np.int32))
    # chosen: 0 -> 2 (d=100) -> then min-dist to {0,10}: pts 1:1, 3:16, 4:16 -> first max = 3 -> then 4 (d=4) vs 1 (1)
    assert out.tolist() == [0, 2, 3, 4]
    # two segments sample independently and return global indices
    xyz2 = np.concatenate([xyz, xyz + 100])
    out2 = P.furthestsampling(xyz2, np.array([5, 10], dtype=np.int32), np.array([2, 5], dtype=np.int32))
    assert out2.tolist() == [0, 2, 5, 7, 8]
