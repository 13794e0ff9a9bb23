"""B200-native twin of models/dgcnn_opensrc.py (the WangYueFt-style DGCNN the reference reuses in
models/dg_ssm.py:31-44 and whose knn / get_graph_feature models/folding_net.py:113-141 calls).

Same names and state_dict keys: `knn`, `get_graph_feature`, `DGCNN(args, input_channels,
output_channels)` with bn1..bn7, conv1..conv5, linear1..linear3. The four EdgeConv stages run on
the fused CUDA path; conv5 and the linear tail are small dense layers left to cuBLAS.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .dgcnn import _compute_dtype
from .ops import KnnGraph


def knn(x, k):
    """x (B, C, N) -> int64 (B, N, k) nearest neighbours, self included, no diagonal fix-up
    (models/dgcnn_opensrc.py:34-40)."""
    return ops.knn_any(x, k, self_loop=True, diag_zero=False).long()


def get_graph_feature(x, k=20, idx=None):
    """Edge features [x_j - x_i, x_i] as a dense (B, 2C, N, k) tensor (models/dgcnn_opensrc.py:43-66).
    Kept for API parity; the networks below never materialise it."""
    B, C, N = x.shape
    if idx is None:
        idx = knn(x, k=k)
    x_pm = x.transpose(2, 1).reshape(B * N, C)
    flat = (idx + torch.arange(B, device=x.device).view(-1, 1, 1) * N).reshape(-1)
    nbr = x_pm[flat].view(B, N, k, C)
    ctr = x_pm.view(B, N, 1, C).expand(B, N, k, C)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()


def edgeconv_stage(x_pm, B, N, k, conv, graph, cdt):
    """One [get_graph_feature -> Conv2d 1x1 -> BatchNorm2d -> LeakyReLU(0.2) -> max over k] stage
    (models/dgcnn_opensrc.py:143-157) on a point-major table, fused."""
    conv2d, bn = conv[0], conv[1]
    if graph is None:
        with torch.no_grad():
            xd = x_pm.detach()
            C = xd.shape[1]
            if C == 3:
                idx = ops.knn_coords(xd.view(B, N, C).permute(0, 2, 1), k, self_loop=True, diag_zero=False)
            else:
                idx = ops.knn_features(xd, B, N, k, self_loop=True, diag_zero=False)
            graph = KnnGraph(idx)
    C = x_pm.shape[1]
    w = conv2d.weight.view(conv2d.out_channels, 2 * C)
    w_cat = ops.edge_weight_table(w, C)
    table = ops.table_gemm(x_pm.float().contiguous(), w_cat.float())   # fp32 table in every mode (see dgcnn.EdgeConv.forward_pm)
    return ops.edgeconv_fused(table, bn.weight, bn.bias, graph, bn.running_mean, bn.running_var,
                              bn.num_batches_tracked, bn.training, eps=bn.eps, momentum=bn.momentum)


class DGCNN(nn.Module):
    def __init__(self, args, input_channels, output_channels=40):
        super().__init__()
        self.args = args
        self.k = args.k
        self.precision = "auto"

        self.bn1 = nn.BatchNorm2d(64)
        self.bn2 = nn.BatchNorm2d(64)
        self.bn3 = nn.BatchNorm2d(128)
        self.bn4 = nn.BatchNorm2d(256)
        self.bn5 = nn.BatchNorm1d(args.emb_dims)

        def stage(cin, cout, bn):
            return nn.Sequential(nn.Conv2d(cin * 2, cout, kernel_size=1, bias=False), bn,
                                 nn.LeakyReLU(negative_slope=0.2))

        self.conv1 = stage(input_channels, 64, self.bn1)
        self.conv2 = stage(64, 64, self.bn2)
        self.conv3 = stage(64, 128, self.bn3)
        self.conv4 = stage(128, 256, self.bn4)
        self.conv5 = nn.Sequential(nn.Conv1d(512, args.emb_dims, kernel_size=1, bias=False), self.bn5,
                                   nn.LeakyReLU(negative_slope=0.2))
        self.linear1 = nn.Linear(args.emb_dims * 2, 512, bias=False)
        self.bn6 = nn.BatchNorm1d(512)
        self.dp1 = nn.Dropout(p=args.dropout)
        self.linear2 = nn.Linear(512, 256)
        self.bn7 = nn.BatchNorm1d(256)
        self.dp2 = nn.Dropout(p=args.dropout)
        self.linear3 = nn.Linear(256, output_channels)

    def encode(self, x, cdt=None):
        """The four EdgeConv stages: x (B, C, N) -> point-major (B*N, 512)."""
        B, _, N = x.shape
        if cdt is None:
            cdt = _compute_dtype(self.precision)
        graph = None
        if self.args.static:
            with torch.no_grad():
                graph = KnnGraph(ops.knn_coords(x.detach(), self.k, self_loop=True, diag_zero=False))
        x_pm = ops.to_point_major(x.float())
        x1 = edgeconv_stage(x_pm, B, N, self.k, self.conv1, graph, cdt)
        x2 = edgeconv_stage(x1, B, N, self.k, self.conv2, graph, cdt)
        x3 = edgeconv_stage(x2, B, N, self.k, self.conv3, graph, cdt)
        x4 = edgeconv_stage(x3, B, N, self.k, self.conv4, graph, cdt)
        return torch.cat((x1, x2, x3, x4), dim=1)

    def forward(self, x):
        B, _, N = x.shape
        cdt = _compute_dtype(self.precision)          # the autocast flag is read before it is switched off below
        with torch.autocast("cuda", enabled=False):
            feats = self.encode(x, cdt).to(cdt)                                     # (B*N, 512)
            w5 = self.conv5[0].weight.view(self.conv5[0].out_channels, 512)
            e = feats @ w5.to(feats.dtype).t()
            bn = self.bn5
            e = F.batch_norm(e, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, bn.momentum, bn.eps)
            if bn.training:
                bn.num_batches_tracked.add_(1)
            e = F.leaky_relu(e, 0.2).view(B, N, -1).float()
            pooled = torch.cat((e.amax(dim=1), e.mean(dim=1)), 1)                   # (B, 2*emb)
            h = F.leaky_relu(self.bn6(self.linear1(pooled)), negative_slope=0.2)
            h = self.dp1(h)
            h = F.leaky_relu(self.bn7(self.linear2(h)), negative_slope=0.2)
            h = self.dp2(h)
            return self.linear3(h).unsqueeze(-1)

    def predict_full_pointcloud(self, pc, sample_points=1024, n_runs_min=50):
        acc = torch.zeros(pc.shape[0], self.linear3.out_features, 1, device=pc.device)
        for _ in range(n_runs_min):
            sub = torch.randperm(pc.shape[-1], device=pc.device)[:sample_points]
            acc += self(pc[..., sub])
        return acc / n_runs_min
