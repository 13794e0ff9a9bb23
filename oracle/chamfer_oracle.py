"""ORACLE — test infrastructure, not product code.

CPU restatement of the Chamfer loss the reference calls at losses/chamfer_loss.py:9-20.

PARITY UNPINNED: the arithmetic lives in the third-party dependency `pytorch3d`
(pytorch3d.loss.chamfer_distance -> pytorch3d.ops.knn_points), which the reference neither vendors nor
pins (no requirements file, lock file or submodule) and which is not installed in the build container.
This file restates pytorch3d's published default behaviour — squared L2 from direct differences,
nearest neighbour (K = 1) in both directions, mean over the points of each cloud, sum of the two
directions, mean over the batch — which is also what the reference's own comment at
train_pc_ae.py:85 says ("mean squared distance and returns d_cham(x,y)+d_cham(y,x)").
"""
import torch


def chamfer_distance(x, y):
    """x (B, N, 3), y (B, M, 3) -> scalar; differentiable w.r.t. both (gradient 2 (x_i - y_nn(i)) / (N B))."""
    diff = x.unsqueeze(2) - y.unsqueeze(1)              # (B, N, M, 3)
    d2 = (diff * diff).sum(-1)
    cham_x = d2.min(dim=2)[0].mean(dim=1)               # (B,)
    cham_y = d2.min(dim=1)[0].mean(dim=1)
    return (cham_x + cham_y).mean()


def nn_points(x, y):
    diff = x.unsqueeze(2) - y.unsqueeze(1)
    d2 = (diff * diff).sum(-1)
    return d2.min(dim=2)


def chamfer_loss(prediction, target):
    """losses/chamfer_loss.py:9-20 including the B x 3 x N -> B x N x 3 transposition rule."""
    if prediction.shape[1] == 3:
        prediction = prediction.transpose(1, 2)
    if target.shape[1] == 3:
        target = target.transpose(1, 2)
    assert prediction.shape[0] == target.shape[0] and prediction.shape[2] == target.shape[2]
    return chamfer_distance(prediction, target)
