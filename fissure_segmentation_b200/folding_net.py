"""B200-native twin of the PC-AE encoder `DGCNN_Cls_Encoder` (models/folding_net.py:83-141).

Same constructor (`k, n_embedding, static=False`), `config` capture, parameter names (bn1..bn5,
conv1..conv5) and output (B, 1, n_embedding) as the reference, so `DGCNNFoldingNet` can take this class
as its encoder. The reference calls the dense `get_graph_feature` + Conv2d four times; here the four
stages run on the fused EdgeConv kernels and the 512 -> n_embedding layer reduces the max over the points
without writing its activation.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .dgcnn import _compute_dtype
from .dgcnn_opensrc import edgeconv_stage
from .modelio import LoadableModel, store_config_args
from .ops import KnnGraph


class DGCNN_Cls_Encoder(LoadableModel):
    @store_config_args
    def __init__(self, k, n_embedding, static=False):
        super().__init__()
        self.static = static
        self.k = k
        self.n_embedding = n_embedding
        self.precision = "auto"

        self.bn1 = nn.BatchNorm2d(64)
        self.bn2 = nn.BatchNorm2d(64)
        self.bn3 = nn.BatchNorm2d(128)
        self.bn4 = nn.BatchNorm2d(256)
        self.bn5 = nn.BatchNorm1d(n_embedding)

        def stage(cin, cout, bn):
            return nn.Sequential(nn.Conv2d(cin * 2, cout, kernel_size=1, bias=False), bn,
                                 nn.LeakyReLU(negative_slope=0.2))

        self.conv1 = stage(3, 64, self.bn1)
        self.conv2 = stage(64, 64, self.bn2)
        self.conv3 = stage(64, 128, self.bn3)
        self.conv4 = stage(128, 256, self.bn4)
        self.conv5 = nn.Sequential(nn.Conv1d(512, n_embedding, kernel_size=1, bias=False), self.bn5,
                                   nn.LeakyReLU(negative_slope=0.2))

    def forward(self, x):
        """x (B, 3, N) -> (B, 1, n_embedding) (models/folding_net.py:113-141)."""
        B, _, N = x.shape
        cdt = _compute_dtype(self.precision)
        with torch.autocast("cuda", enabled=False):
            ops.begin_step(x.device, self)
            graph = None
            if self.static:
                # models/folding_net.py:114-115: dgcnn_opensrc.knn on the coordinates (self included, no diagonal fix)
                with torch.no_grad():
                    graph = KnnGraph(ops.knn_coords(x.detach()[:, :3], self.k, self_loop=True, diag_zero=False))
            x_pm = ops.to_point_major(x.float())
            x1 = edgeconv_stage(x_pm, B, N, self.k, self.conv1, graph, cdt)
            x2 = edgeconv_stage(x1, B, N, self.k, self.conv2, graph, cdt)
            x3 = edgeconv_stage(x2, B, N, self.k, self.conv3, graph, cdt)
            x4 = edgeconv_stage(x3, B, N, self.k, self.conv4, graph, cdt)
            feats = torch.cat((x1, x2, x3, x4), dim=1).to(cdt)                       # (B*N, 512)
            w5 = self.conv5[0].weight.view(self.n_embedding, 512)
            y = feats @ w5.to(cdt).t()
            E = self.n_embedding
            if E >= 64 and (E & (E - 1)) == 0:
                # conv5 + BN + LeakyReLU + max over the points: per-cloud max/min of the GEMM output is enough
                out = ops.pool_bn_act(y, self.bn5, 0.2, B, N)
            else:
                bn = self.bn5
                y = F.batch_norm(y.float(), bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training,
                                 bn.momentum, bn.eps)
                if bn.training:
                    bn.num_batches_tracked.add_(1)
                out = F.leaky_relu(y, 0.2).view(B, N, E).amax(dim=1)
            return out.float().unsqueeze(1)
