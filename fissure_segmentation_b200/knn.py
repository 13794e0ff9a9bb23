"""Reference-facing kNN helpers (utils/general_utils.py:43-53, 315-327; models/dgcnn.py:15-36)."""
import torch

from . import ops


def knn(x, k, self_loop=False, return_dist=False):
    """x (B, C, N) -> idx (B, N, k) int64 [, squared distances (B, N, k)], ascending distance.

    Same contract as utils/general_utils.py:315-327: distances in the expansion form of
    pairwise_dist with the diagonal forced to 0; with self_loop=False, k+1 neighbours are selected
    and the first column is dropped. The N x N distance matrix is never written to memory."""
    out = ops.knn_any(x, k, self_loop=self_loop, diag_zero=True, return_dist=return_dist)
    if return_dist:
        return out[0].long(), out[1]
    return out.long()


def pairwise_dist(x):
    """Dense squared distance matrix (B, N, N) of x (B, N, C), diagonal 0 (utils/general_utils.py:43-53).
    Provided for callers that really want the matrix; the kNN path does not use it."""
    sq = (x ** 2).sum(2, keepdim=True)
    dist = sq - 2.0 * torch.bmm(x, x.transpose(2, 1)) + sq.transpose(2, 1)
    n = dist.shape[1]
    dist[:, torch.arange(n), torch.arange(n)] = 0
    return dist


def create_neighbor_features(x, k, fixed_knn_graph=None, knn_only_over_coords=False):
    """Dense edge features [x_j - x_i, x_i] of shape (B, 2C, N, k) (models/dgcnn.py:15-36), for callers
    outside the fused EdgeConv that need the tensor itself. The graph comes from the CUDA kNN."""
    B, C, N = x.shape
    if fixed_knn_graph is None:
        idx = knn(x[:, :3] if knn_only_over_coords else x, k, self_loop=True)
    else:
        idx = fixed_knn_graph
    nbr = torch.take_along_dim(x, idx.reshape(B, 1, N * k), dim=-1).view(B, C, N, k)
    ctr = x.unsqueeze(-1).expand(B, C, N, k)
    return torch.cat([nbr - ctr, ctr], dim=1)
