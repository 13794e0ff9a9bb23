"""B200-native DGCNN modules with the reference's interface (models/dgcnn.py).

Same constructors, `forward(x: B x C x N) -> B x classes x N`, `config`, `state_dict` keys and
checkpoint format as the reference, so `train.py` / `ModelTrainer` can use these classes in place of
`models.dgcnn`. Internally everything runs point-major ((B*N) x C tables) on the CUDA kernels of
libfissure_b200.so; the parameter containers are ordinary nn.Conv / nn.BatchNorm modules that are
created in the reference's order (identical initial weights for the same seed) but whose own
`forward` is never used on the EdgeConv path.
"""
import os

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init

from . import ops
from .modelio import PointSegmentationModelBase
from .ops import KnnGraph

FUSED_WIDTHS = (64, 128, 256)


def init_weights(m):
    """Xavier-normal weights, zero bias (utils/model_utils.py:11-15)."""
    if isinstance(m, (nn.modules.conv._ConvNd, nn.Linear, nn.modules.conv._ConvTransposeNd)):
        nn.init.xavier_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0.0)


def _compute_dtype(precision):
    if precision == "fp32":
        return torch.float32
    if precision == "bf16":
        return torch.bfloat16
    # 'auto': follow the trainer's autocast context (model_trainer.py:75-76,157); the kernels keep
    # half-precision tables in bf16 whatever the autocast dtype is
    return torch.bfloat16 if torch.is_autocast_enabled("cuda") else torch.float32


def _as_graph(g):
    if g is None or isinstance(g, KnnGraph):
        return g
    return KnnGraph.from_reference(g)


class ConvBlock(nn.Module):
    """conv -> [BatchNorm] -> [LeakyReLU]; parameter layout of models/dgcnn.py:282-315."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=True, dim=2,
                 negative_slope=1e-2, bn=True, activation=True):
        super().__init__()
        try:
            conv_cls, norm_cls = {1: (nn.Conv1d, nn.BatchNorm1d), 2: (nn.Conv2d, nn.BatchNorm2d),
                                  3: (nn.Conv3d, nn.BatchNorm3d)}[dim]
        except KeyError:
            raise ValueError(f'There is no Conv layer for dimensionality {dim}.')
        self.layers = nn.ModuleList([conv_cls(in_channels=in_channels, out_channels=out_channels,
                                              kernel_size=kernel_size, bias=not bn, stride=stride,
                                              padding=(kernel_size // 2) if padding else 0)])
        if bn:
            self.layers.append(norm_cls(num_features=out_channels))
        if activation:
            self.layers.append(nn.LeakyReLU(negative_slope=negative_slope))
        self.negative_slope = negative_slope

    def forward(self, x):
        for layer in self.layers:
            x = layer(x)
        return x

    # ---- point-major path -------------------------------------------------------------------
    @property
    def conv(self):
        return self.layers[0]

    @property
    def norm(self):
        return self.layers[1] if len(self.layers) > 1 and isinstance(self.layers[1], nn.modules.batchnorm._BatchNorm) else None

    @property
    def has_activation(self):
        return isinstance(self.layers[-1], nn.LeakyReLU)

    def weight_matrix(self):
        w = self.conv.weight
        return w.view(w.shape[0], w.shape[1])

    def norm_act_pm(self, y, rowbias=None, rows_per_cloud=1):
        """BatchNorm (over all rows) + LeakyReLU on a point-major (rows, C) table; `rowbias` (B, C) is an
        optional per-cloud bias added to y first (fused into the kernels)."""
        bn = self.norm
        C = y.shape[1]
        if bn is not None and y.is_cuda and C >= 64 and (C & (C - 1)) == 0 and y.stride(1) == 1:
            slope = self.negative_slope if self.has_activation else 1.0
            return ops.bn_act(y, bn, slope, rowbias, rows_per_cloud)
        if rowbias is not None:
            y = (y.view(-1, rows_per_cloud, C) + rowbias.to(y.dtype).unsqueeze(1)).view(-1, C)
        if bn is not None:
            y = F.batch_norm(y, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training, bn.momentum, bn.eps)
            if bn.training:
                bn.num_batches_tracked.add_(1)
        if self.has_activation:
            y = F.leaky_relu(y, self.negative_slope)
        return y

    def forward_pool_pm(self, x, B, N):
        """Conv + BatchNorm + LeakyReLU followed by the max over the N points of each cloud -> (B, C_out).
        LeakyReLU(BN(.)) is monotone per channel, so only per-cloud max/min of the GEMM output is needed."""
        bn = self.norm
        C = self.conv.out_channels
        fused = bn is not None and self.conv.bias is None and x.is_cuda and C >= 64 and (C & (C - 1)) == 0
        slope = self.negative_slope if self.has_activation else 1.0
        if fused and ops.pool_linear_supported(x, self.weight_matrix()):
            # GEMM + pooled BatchNorm as ONE autograd node: the backward needs neither y nor the dense gradient
            return ops.pool_linear_bn_act(x, self.weight_matrix(), bn, slope, B, N)
        y = x @ self.weight_matrix().to(x.dtype).t()
        if fused:
            return ops.pool_bn_act(y, bn, slope, B, N)
        if self.conv.bias is not None:
            y = y + self.conv.bias.to(y.dtype)
        return self.norm_act_pm(y).view(B, N, -1).amax(dim=1)

    def forward_pm(self, x):
        """x (rows, C_in) -> (rows, C_out): 1x1 conv as a GEMM, then BatchNorm + LeakyReLU."""
        y = ops.linear_pm(x, self.weight_matrix(), self.conv.bias)       # bias added by the GEMM, its gradient by a GEMM too
        return self.norm_act_pm(y)


class SharedFullyConnected(ConvBlock):
    """1x1 conv shared over points (and edges); models/dgcnn.py:318-323."""

    def __init__(self, in_features, out_features, dim=2, last_layer=False):
        super().__init__(in_features, out_features, dim=dim, kernel_size=1, padding=False, bn=not last_layer,
                         activation=not last_layer, negative_slope=0.2)


class EdgeConv(nn.Module):
    """Dynamic-graph edge convolution (models/dgcnn.py:212-243), fused.

    forward(x: B x C x N, fixed_knn_graph=None) keeps the reference signature; `forward_pm` is the
    point-major entry the networks use. The first shared-MLP layer is evaluated with the identity
    W [x_j - x_i ; x_i] = W1 x_j + (W2 - W1) x_i (one per-point GEMM, no edge tensor); a single-layer
    EdgeConv then needs only a gather/max pass over the k neighbour rows.
    """

    def __init__(self, in_features, out_features_list, k, first_layer=False):
        super().__init__()
        self.k = k
        self.first_layer = first_layer
        self.precision = "auto"
        widths = [in_features * 2] + list(out_features_list)
        self.shared_mlp = nn.ModuleList()
        for i in range(len(out_features_list)):
            self.shared_mlp.append(SharedFullyConnected(widths[i], widths[i + 1]))

    def forward(self, x, fixed_knn_graph=None):
        B, C, N = x.shape
        cdt = _compute_dtype(self.precision)         # read the trainer's autocast flag BEFORE it is switched off below
        with torch.autocast("cuda", enabled=False):
            out = self.forward_pm(ops.to_point_major(x.float()), B, N, _as_graph(fixed_knn_graph), cdt)
            return out.view(B, N, -1).permute(0, 2, 1).float()

    def build_graph(self, x_pm, B, N):
        """kNN graph of this layer's input, self loop included (models/dgcnn.py:24-27)."""
        C = x_pm.shape[1]
        if self.first_layer or C == 3:
            coords = x_pm.view(B, N, C).permute(0, 2, 1)[:, :3]
            idx = ops.knn_coords(coords, self.k, self_loop=True)
        else:
            idx = ops.knn_features(x_pm, B, N, self.k, self_loop=True)
        return KnnGraph(idx)

    def forward_pm(self, x_pm, B, N, graph=None, cdt=torch.float32):
        if graph is None:
            with torch.no_grad():
                graph = self.build_graph(x_pm.detach(), B, N)
        first = self.shared_mlp[0]
        Cp = first.conv.out_channels
        if Cp not in FUSED_WIDTHS:
            raise NotImplementedError(f"fused EdgeConv supports widths {FUSED_WIDTHS}, got {Cp}")
        C = x_pm.shape[1]
        if (len(self.shared_mlp) > 1 and C == 3 and Cp in (64, 128) and not x_pm.requires_grad
                and first.norm is not None and first.has_activation and abs(first.negative_slope - 0.2) < 1e-12):
            last = self.shared_mlp[-1]
            if (len(self.shared_mlp) == 2 and cdt == torch.bfloat16 and last.norm is not None and last.has_activation
                    and abs(last.negative_slope - 0.2) < 1e-12 and last.norm.training == first.norm.training
                    and ops.edge2_supported(graph.k, Cp, last.conv.out_channels)):
                # both layers in one tcgen05 kernel each way: no edge tensor is written (csrc/edge2.cu); bf16 operands,
                # like the cuBLAS product of the materialised path in this precision mode
                return ops.edgeconv2_fused(x_pm.float().contiguous(), first, last, graph)
            # first layer on raw coordinates: batch statistics from the moments of the 6-D edge vectors, H written
            # once, weight gradient accumulated straight from dH (csrc/edge3.cu)
            h = ops.edge_first3(x_pm.float().contiguous(), first.conv.weight, first.norm, graph, cdt)
            return self._finish_multilayer(h, graph)
        w = first.weight_matrix()
        w_cat = ops.edge_weight_table(w, C)                                    # [W1 ; W2 - W1]  (2Cp, C)
        # The per-point table is STORED in fp32 in every precision mode: y = a_j + b_i = W1 (x_j - x_i) + W2 x_i
        # cancels the common part of a_j and -b_i, so rounding a to 16 bits would wipe out the local
        # differences the layer is about (measured: cosine 0.998 on gradients with bf16 tables). In the 'bf16' mode the
        # GEMM that forms it may round its OPERANDS to TF32 (fp32 accumulation and output); 'fp32' mode does not.
        table = ops.table_gemm(x_pm.float().contiguous(), w_cat.float(), tf32=cdt == torch.bfloat16)
        bn = first.norm
        if len(self.shared_mlp) == 1:
            return ops.edgeconv_fused(table, bn.weight, bn.bias, graph, bn.running_mean, bn.running_var,
                                      bn.num_batches_tracked, bn.training, eps=bn.eps, momentum=bn.momentum)
        # two (or more) layers: layer 1 pre-activations as an edge tensor (compute dtype), middle layers in
        # torch, last layer's BatchNorm + LeakyReLU + max over k in one reduction kernel
        h = first.norm_act_pm(ops.edge_build(table, graph, cdt))
        return self._finish_multilayer(h, graph)

    def _finish_multilayer(self, h, graph):
        """Layers 2.. of a multi-layer EdgeConv on the hidden edge tensor h (P*k, C1), then max over k."""
        for layer in self.shared_mlp[1:-1]:
            h = layer.forward_pm(h)
        last = self.shared_mlp[-1]
        if last.conv.out_channels not in FUSED_WIDTHS:
            raise NotImplementedError(f"fused EdgeConv supports widths {FUSED_WIDTHS}")
        z = h @ last.weight_matrix().to(h.dtype).t()
        bn = last.norm
        out = ops.edge_reduce(z, bn.weight, bn.bias, graph.k, bn.running_mean, bn.running_var,
                              bn.num_batches_tracked, bn.training, eps=bn.eps, momentum=bn.momentum)
        return out.float()


class SpatialTransformer(nn.Module):
    """Learned 3x3 alignment of the coordinates (models/dgcnn.py:246-279)."""

    def __init__(self, k):
        super().__init__()
        self.in_features = 3
        self.ec = EdgeConv(self.in_features, [64, 128], k)
        self.shared_fc = SharedFullyConnected(128, 1024, dim=1)
        self.mlp = nn.Sequential(
            nn.Linear(1024, 512), nn.BatchNorm1d(512), nn.LeakyReLU(negative_slope=0.2),
            nn.Linear(512, 256), nn.BatchNorm1d(256), nn.LeakyReLU(negative_slope=0.2),
        )
        self.transform = nn.Linear(256, self.in_features * self.in_features)

    def forward(self, x, fixed_knn_graph=None):
        B, _, N = x.shape
        coords = x[:, :self.in_features].float()
        cdt = _compute_dtype(self.ec.precision)
        with torch.autocast("cuda", enabled=False):
            feat = self.ec.forward_pm(ops.to_point_major(coords), B, N, _as_graph(fixed_knn_graph), cdt)
            feat = self.shared_fc.forward_pool_pm(feat.to(cdt), B, N).float()   # conv + BN + LReLU + max over points
            mat = self.transform(self.mlp(feat)).view(B, self.in_features, self.in_features)
            moved = torch.bmm(coords.transpose(2, 1), mat).transpose(2, 1)
        return torch.cat([moved.to(x.dtype), x[:, self.in_features:]], dim=1)

    def init_weights(self):
        self.apply(init_weights)
        init.constant_(self.transform.weight, 0)
        init.eye_(self.transform.bias.view(self.in_features, self.in_features))


class ImageFeatures(nn.Module):
    """Per-point MLP on the non-coordinate channels (models/dgcnn.py:326-343)."""

    def __init__(self, in_channels=6, out_channels=(6, 12), kernel_size=1):
        super().__init__()
        self.layers = nn.ModuleList()
        widths = [in_channels, *out_channels[:-1]]
        for i, o in zip(widths, out_channels):
            self.layers.append(ConvBlock(in_channels=i, out_channels=o, kernel_size=kernel_size, dim=1))

    def forward(self, x):
        feat = x[:, 3:]
        for layer in self.layers:
            feat = layer(feat)
        return torch.cat([x[:, :3], feat], dim=1)


class DGCNNBase(PointSegmentationModelBase):
    """Shared front end: static graph, image-feature module, spatial transformer (models/dgcnn.py:61-112)."""

    def __init__(self, k, in_features, num_classes, spatial_transformer=False, dynamic=True, image_feat_module=False):
        super().__init__(in_features, num_classes, k=k, spatial_transformer=spatial_transformer, dynamic=dynamic,
                         image_feat_module=image_feat_module)
        self.k = k
        self.dynamic = dynamic
        self._graph = None
        self._perm = None
        self.precision = "auto"   # 'auto' | 'fp32' | 'bf16'
        # Sort every cloud along the Morton curve of its coordinates before the EdgeConvs: consecutive rows are
        # then spatial neighbours and the k-row gathers of a CTA's point range overlap (L1/L2 hits). Every
        # operator is permutation-equivariant; per-point outputs are un-permuted before they are returned.
        self.spatial_sort = os.environ.get("FS_SPATIAL_SORT", "1") != "0"

        if image_feat_module:
            if in_features < 4:
                raise ValueError('Number of In-Features for DGCNN too low if you want to use the image feature '
                                 'module! Need at 3, as the first 3 are assumed to be the point coordinates.')
            self.image_feature_module = ImageFeatures(in_channels=in_features - 3, out_channels=(6, 12))
            self.in_features = 3 + 12
        else:
            self.image_feature_module = None
            self.in_features = in_features

        self.spatial_transformer = SpatialTransformer(k) if spatial_transformer else None
        self.output_activation = nn.Identity()

    @property
    def knn_graph(self):
        """Static graph of the last forward as the reference exposes it: int64 (B, N, k) or None,
        in the caller's point numbering."""
        if self._graph is None:
            return None
        idx = self._graph.idx.long()
        if self._perm is None:
            return idx
        B, N, k = idx.shape
        orig = torch.gather(self._perm, 1, idx.reshape(B, N * k)).view(B, N, k)      # neighbour ids -> original
        out = torch.empty_like(orig)
        out.scatter_(1, self._perm.unsqueeze(-1).expand(B, N, k), orig)              # rows -> original
        return out

    def _unsort_points(self, y):
        """(B, C, N) in sorted order -> the caller's order."""
        if self._perm is None:
            return y
        out = torch.empty_like(y)
        return out.scatter(2, self._perm.unsqueeze(1).expand_as(y), y)

    def forward(self, x):
        self._perm = None
        ops.begin_step(x.device, self)      # one zero-filled arena (per model and stream) for all statistics buffers of this step
        if self.spatial_sort and x.is_cuda and x.shape[1] >= 3:
            with torch.no_grad():
                self._perm = ops.spatial_order(x.detach())
            x = torch.gather(x, 2, self._perm.unsqueeze(1).expand_as(x))
        if not self.dynamic:
            with torch.no_grad():
                self._graph = KnnGraph(ops.knn_coords(x.detach(), self.k, self_loop=False))
        if self.image_feature_module is not None:
            x = self.image_feature_module(x)
        if self.spatial_transformer is not None:
            self.spatial_transformer.ec.precision = self.precision
            x = self.spatial_transformer(x)
        return x

    def init_weights(self):
        self.apply(init_weights)
        if self.spatial_transformer is not None:
            self.spatial_transformer.init_weights()


class DGCNNSeg(DGCNNBase):
    """Point segmentation DGCNN (models/dgcnn.py:115-162): three EdgeConvs, 1024-d global feature,
    four shared FC layers; forward(x: B x C x N) -> logits B x num_classes x N."""

    def __init__(self, k, in_features, num_classes, spatial_transformer=False, dynamic=True, image_feat_module=False):
        super().__init__(k, in_features, num_classes, spatial_transformer, dynamic, image_feat_module)
        self.ec1 = EdgeConv(self.in_features, [64, 64], self.k, first_layer=True)
        self.ec2 = EdgeConv(64, [64], self.k)
        self.ec3 = EdgeConv(64, [64], self.k)
        self.global_feature = nn.Sequential(SharedFullyConnected(3 * 64, 1024, dim=1), nn.AdaptiveMaxPool1d(1))
        self.segmentation = nn.Sequential(
            SharedFullyConnected(3 * 64 + 1024, 256, dim=1),
            SharedFullyConnected(256, 256, dim=1),
            SharedFullyConnected(256, 128, dim=1),
            SharedFullyConnected(128, self.num_classes, dim=1, last_layer=True),
        )
        self.init_weights()

    def forward(self, x):
        x = super().forward(x)
        B, _, N = x.shape
        cdt = _compute_dtype(self.precision)
        with torch.autocast("cuda", enabled=False):
            x_pm = ops.to_point_major(x.float())
            g = self._graph if not self.dynamic else None
            x1 = self.ec1.forward_pm(x_pm, B, N, g, cdt)
            x2 = self.ec2.forward_pm(x1, B, N, g, cdt)
            x3 = self.ec3.forward_pm(x2, B, N, g, cdt)
            feats = ops.cat_cast([x1, x2, x3], cdt)                              # (B*N, 192), one pass

            glob = self.global_feature[0].forward_pool_pm(feats, B, N)           # (B, 1024), activation never written

            # segmentation[0] on [feats | broadcast global]: split the weight instead of materialising
            # the 1216-wide concat (models/dgcnn.py:159): local GEMM + one per-cloud bias row
            seg0 = self.segmentation[0]
            w0 = seg0.weight_matrix()
            local = ops.linear_pm(feats, w0[:, :feats.shape[1]])
            per_cloud = glob @ w0[:, feats.shape[1]:].to(cdt).t()                # (B, 256)
            h = seg0.norm_act_pm(local, rowbias=per_cloud, rows_per_cloud=N)
            h = self.segmentation[1].forward_pm(h)
            h = self.segmentation[2].forward_pm(h)
            last = self.segmentation[3]
            if (last.norm is None and not last.has_activation and h.is_contiguous()
                    and ops.final_linear_supported(h, last.weight_matrix())):
                # classes-wide 1x1 conv + bias written straight as (B, classes, N) fp32 in the caller's point order
                return ops.final_linear_out(h, last.weight_matrix(), last.conv.bias, self._perm, B, N)
            logits = last.forward_pm(h)                                          # (B*N, classes)
            return ops.logits_out(logits, self._perm, B, N)          # (B, classes, N) fp32 in the caller's point order


class DGCNNReg(DGCNNBase):
    """Regression DGCNN (models/dgcnn.py:165-209): four EdgeConvs (64, 64, 128, 256), global 1024."""

    def __init__(self, k, in_features, num_classes, spatial_transformer=False, dynamic=True, image_feat_module=False):
        super().__init__(k, in_features, num_classes, spatial_transformer, dynamic, image_feat_module)
        self.ec1 = EdgeConv(self.in_features, [64], self.k, first_layer=True)
        self.ec2 = EdgeConv(64, [64], self.k)
        self.ec3 = EdgeConv(64, [128], self.k)
        self.ec4 = EdgeConv(128, [256], self.k)
        self.global_feature = nn.Sequential(SharedFullyConnected(2 * 64 + 128 + 256, 1024, dim=1),
                                            nn.AdaptiveMaxPool1d(1))
        self.regression = nn.Sequential(
            SharedFullyConnected(1024, 512, dim=1),
            SharedFullyConnected(512, 256, dim=1),
            SharedFullyConnected(256, self.num_classes, dim=1, last_layer=True),
        )
        self.init_weights()

    def forward(self, x):
        x = super().forward(x)
        B, _, N = x.shape
        cdt = _compute_dtype(self.precision)
        with torch.autocast("cuda", enabled=False):
            x_pm = ops.to_point_major(x.float())
            g = self._graph if not self.dynamic else None
            x1 = self.ec1.forward_pm(x_pm, B, N, g, cdt)
            x2 = self.ec2.forward_pm(x1, B, N, g, cdt)
            x3 = self.ec3.forward_pm(x2, B, N, g, cdt)
            x4 = self.ec4.forward_pm(x3, B, N, g, cdt)
            feats = ops.cat_cast([x1, x2, x3, x4], cdt)
            glob = self.global_feature[0].forward_pool_pm(feats, B, N)                   # (B, 1024)
            h = glob
            for layer in self.regression:
                h = layer.forward_pm(h)
            return h.float().unsqueeze(-1)                                               # (B, out, 1)

    def predict_full_pointcloud(self, pc, sample_points=1024, n_runs_min=50):
        acc = torch.zeros(pc.shape[0], self.num_classes, 1, device=pc.device)
        for _ in range(n_runs_min):
            sub = torch.randperm(pc.shape[-1], device=pc.device)[:sample_points]
            acc += self(pc[..., sub])
        return acc / n_runs_min
