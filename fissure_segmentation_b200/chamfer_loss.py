"""ChamferLoss with the reference's interface (losses/chamfer_loss.py:4-20); the bidirectional
nearest-neighbour reduction and its gradient run in the CUDA kernels of libfissure_b200.so instead of
pytorch3d.loss.chamfer_distance."""
from torch import nn

from . import ops


def chamfer_distance(x, y):
    """(loss, None) like pytorch3d.loss.chamfer_distance(x, y) with default arguments: squared L2,
    mean over points, both directions summed, mean over the batch. x (B, N, 3), y (B, M, 3)."""
    return ops.chamfer_distance(x, y), None


class ChamferLoss(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, prediction, target):
        # B x 3 x N inputs are transposed to B x N x 3 (losses/chamfer_loss.py:10-16)
        if prediction.shape[1] == 3:
            prediction = prediction.transpose(1, 2)
        if target.shape[1] == 3:
            target = target.transpose(1, 2)
        assert prediction.shape[0] == target.shape[0] and prediction.shape[2] == target.shape[2]
        loss, _ = chamfer_distance(prediction, target)
        return loss
