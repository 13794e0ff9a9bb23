"""ORACLE — test infrastructure, not product code.

CPU (PyTorch, fp32 or fp64) restatement of the reference's DGCNN EdgeConv path, written functionally
over a `state_dict`-shaped parameter dictionary. Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this package; the product path
(fissure_segmentation_b200) never does.

Every function cites the reference lines it restates (paths relative to the reference tree).
Pinned against the reference itself: tests/golden/make_golden.py imports the real reference modules in
the build container and writes fixtures; tests/test_oracle_golden.py checks this file against them.
"""
import torch
import torch.nn.functional as F

LEAKY = 0.2
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------
# kNN (utils/general_utils.py:43-53, 315-327; models/dgcnn_opensrc.py:34-40)
# ----------------------------------------------------------------------------------------------
def pairwise_sqdist(x_bnc):
    """utils/general_utils.py:43-53: |x|^2 - 2 x x^T + |x|^2^T with the diagonal forced to 0."""
    sq = (x_bnc ** 2).sum(2, keepdim=True)
    gram = torch.bmm(x_bnc, x_bnc.transpose(2, 1))
    dist = sq - 2.0 * gram + sq.transpose(2, 1)
    n = dist.shape[1]
    ar = torch.arange(n)
    dist[:, ar, ar] = 0
    return dist


def knn(x_bcn, k, self_loop=False, return_dist=False):
    """utils/general_utils.py:315-327: topk smallest of k (+1 when the self match is dropped)."""
    extra = 0 if self_loop else 1
    dist = pairwise_sqdist(x_bcn.transpose(2, 1))
    top_d, idx = dist.topk(k=k + extra, dim=-1, largest=False)
    top_d, idx = top_d[..., extra:], idx[..., extra:]
    return (idx, top_d) if return_dist else idx


def knn_with_gap(x_bcn, k, self_loop=False):
    """Oracle graph plus the material for tie classification: sorted distances of the k
    kept neighbours, the distance of the first rejected candidate and the scale of the rounding error."""
    extra = 0 if self_loop else 1
    xt = x_bcn.transpose(2, 1)
    dist = pairwise_sqdist(xt)
    kk = min(k + extra + 1, dist.shape[-1])
    top_d, idx = dist.topk(k=kk, dim=-1, largest=False)
    kept_d, kept_i = top_d[..., extra:k + extra], idx[..., extra:k + extra]
    next_d = top_d[..., k + extra] if kk > k + extra else torch.full_like(top_d[..., 0], float("inf"))
    sq = (xt ** 2).sum(2)
    return kept_i, kept_d, next_d, sq


def knn_opensrc(x_bcn, k):
    """models/dgcnn_opensrc.py:34-40: largest k of -|xi|^2 + 2 xi.xj - |xj|^2, self included."""
    inner = -2 * torch.matmul(x_bcn.transpose(2, 1), x_bcn)
    sq = torch.sum(x_bcn ** 2, dim=1, keepdim=True)
    neg = -sq - inner - sq.transpose(2, 1)
    return neg.topk(k=k, dim=-1)[1]


# ----------------------------------------------------------------------------------------------
# EdgeConv (models/dgcnn.py:15-36, 212-243, 282-323)
# ----------------------------------------------------------------------------------------------
def edge_features(x_bcn, k, idx=None, coords_only=False):
    """models/dgcnn.py:15-36: gather neighbours, stack [x_j - x_i, x_i] -> (B, 2C, N, k)."""
    B, C, N = x_bcn.shape
    if idx is None:
        idx = knn(x_bcn[:, :3] if coords_only else x_bcn, k, self_loop=True)
    flat = idx.reshape(B, 1, N * k)
    nbr = torch.take_along_dim(x_bcn, flat, dim=-1).view(B, C, N, k)
    ctr = x_bcn.unsqueeze(-1).repeat(1, 1, 1, k)
    return torch.cat([nbr - ctr, ctr], dim=1)


def _bn(x, p, prefix, training, stats_out):
    """torch BatchNorm with momentum 0.1, eps 1e-5 (models/dgcnn.py:306-307); running statistics are
    updated functionally into stats_out so the caller can compare them."""
    rm = p[prefix + ".running_mean"].clone()
    rv = p[prefix + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, p[prefix + ".weight"], p[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)
    if stats_out is not None and training:
        stats_out[prefix + ".running_mean"] = rm
        stats_out[prefix + ".running_var"] = rv
    return y


def shared_fc(x, p, prefix, dim, training, last_layer=False, stats_out=None):
    """SharedFullyConnected (models/dgcnn.py:318-323): 1x1 conv (bias only when last) [+ BN + LeakyReLU]."""
    w = p[prefix + ".layers.0.weight"]
    b = p.get(prefix + ".layers.0.bias")
    x = F.conv2d(x, w, b) if dim == 2 else F.conv1d(x, w, b)
    if not last_layer:
        x = _bn(x, p, prefix + ".layers.1", training, stats_out)
        x = F.leaky_relu(x, LEAKY)
    return x


def edgeconv(x_bcn, p, prefix, n_layers, k, idx=None, first_layer=False, training=True, stats_out=None):
    """EdgeConv.forward (models/dgcnn.py:226-243): edge features -> shared MLP -> max over k."""
    e = edge_features(x_bcn, k, idx, coords_only=first_layer)
    for i in range(n_layers):
        e = shared_fc(e, p, f"{prefix}.shared_mlp.{i}", 2, training, stats_out=stats_out)
    return e.max(dim=-1)[0]


def dgcnn_seg(p, x, k, dynamic=True, training=True, stats_out=None, graphs_out=None, fixed_graphs=None):
    """DGCNNSeg.forward (models/dgcnn.py:141-162) without spatial transformer / image features.
    fixed_graphs: optional list of three (B, N, k) index tensors to teacher-force the graphs."""
    graph = None
    if not dynamic:
        graph = knn(x[:, :3], k, self_loop=False)                      # models/dgcnn.py:96
    g = [graph] * 3 if fixed_graphs is None else list(fixed_graphs)
    if graphs_out is not None and dynamic and fixed_graphs is None:
        g1 = knn(x[:, :3], k, self_loop=True)
        x1 = edgeconv(x, p, "ec1", 2, k, g1, True, training, stats_out)
        g2 = knn(x1, k, self_loop=True)
        x2 = edgeconv(x1, p, "ec2", 1, k, g2, False, training, stats_out)
        g3 = knn(x2, k, self_loop=True)
        x3 = edgeconv(x2, p, "ec3", 1, k, g3, False, training, stats_out)
        graphs_out.extend([g1, g2, g3])
    else:
        x1 = edgeconv(x, p, "ec1", 2, k, g[0], True, training, stats_out)
        x2 = edgeconv(x1, p, "ec2", 1, k, g[1], False, training, stats_out)
        x3 = edgeconv(x2, p, "ec3", 1, k, g[2], False, training, stats_out)
        if graphs_out is not None:
            graphs_out.extend(g)
    multi = torch.cat([x1, x2, x3], dim=1)
    glob = shared_fc(multi, p, "global_feature.0", 1, training, stats_out=stats_out)
    glob = F.adaptive_max_pool1d(glob, 1)
    h = torch.cat([multi, glob.repeat(1, 1, multi.shape[-1])], dim=1)
    h = shared_fc(h, p, "segmentation.0", 1, training, stats_out=stats_out)
    h = shared_fc(h, p, "segmentation.1", 1, training, stats_out=stats_out)
    h = shared_fc(h, p, "segmentation.2", 1, training, stats_out=stats_out)
    return shared_fc(h, p, "segmentation.3", 1, training, last_layer=True)


def dgcnn_seg_param_shapes(in_features, num_classes):
    """state_dict layout of DGCNNSeg, in the reference's registration order."""
    shapes = []

    def fc(prefix, cin, cout, dim, last=False):
        shapes.append((prefix + ".layers.0.weight", (cout, cin, 1, 1) if dim == 2 else (cout, cin, 1)))
        if last:
            shapes.append((prefix + ".layers.0.bias", (cout,)))
        else:
            for nm in ("weight", "bias", "running_mean", "running_var"):
                shapes.append((prefix + ".layers.1." + nm, (cout,)))
            shapes.append((prefix + ".layers.1.num_batches_tracked", ()))

    fc("ec1.shared_mlp.0", 2 * in_features, 64, 2)
    fc("ec1.shared_mlp.1", 64, 64, 2)
    fc("ec2.shared_mlp.0", 128, 64, 2)
    fc("ec3.shared_mlp.0", 128, 64, 2)
    fc("global_feature.0", 192, 1024, 1)
    fc("segmentation.0", 1216, 256, 1)
    fc("segmentation.1", 256, 256, 1)
    fc("segmentation.2", 256, 128, 1)
    fc("segmentation.3", 128, num_classes, 1, last=True)
    return shapes


def make_params(shapes, seed, dtype=torch.float32, random_bn=True):
    """Deterministic parameters for parity tests: Xavier-like conv weights, BatchNorm affine with random
    sign gamma (so both the max and the min branch of the fused kernel are exercised), non-trivial
    running statistics."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    for name, shape in shapes:
        if name.endswith("num_batches_tracked"):
            p[name] = torch.zeros((), dtype=torch.long)
        elif name.endswith("layers.0.weight"):
            fan = shape[0] + shape[1]
            p[name] = (torch.randn(shape, generator=gen) * (2.0 / fan) ** 0.5).to(dtype)
        elif name.endswith("layers.0.bias"):
            p[name] = (0.1 * torch.randn(shape, generator=gen)).to(dtype)
        elif name.endswith("layers.1.weight"):
            g = 0.5 + torch.rand(shape, generator=gen)
            if random_bn:
                g = g * torch.where(torch.rand(shape, generator=gen) < 0.3, -1.0, 1.0)
            p[name] = g.to(dtype)
        elif name.endswith("layers.1.bias"):
            p[name] = (0.2 * torch.randn(shape, generator=gen) if random_bn else torch.zeros(shape)).to(dtype)
        elif name.endswith("running_mean"):
            p[name] = (0.1 * torch.randn(shape, generator=gen)).to(dtype)
        elif name.endswith("running_var"):
            p[name] = (0.5 + torch.rand(shape, generator=gen)).to(dtype)
        else:
            raise KeyError(name)
    return p


# ----------------------------------------------------------------------------------------------
# dgcnn_opensrc.DGCNN (models/dgcnn_opensrc.py:43-66, 136-171)
# ----------------------------------------------------------------------------------------------
def graph_feature(x_bcn, k, idx=None):
    """models/dgcnn_opensrc.py:43-66."""
    B, C, N = x_bcn.shape
    if idx is None:
        idx = knn_opensrc(x_bcn, k)
    flat = (idx + torch.arange(B).view(-1, 1, 1) * N).view(-1)
    xt = x_bcn.transpose(2, 1).contiguous()
    nbr = xt.view(B * N, -1)[flat, :].view(B, N, k, C)
    ctr = xt.view(B, N, 1, C).repeat(1, 1, k, 1)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2).contiguous()


def opensrc_stage(x_bcn, p, conv, bn, k, idx, training):
    """One stage of models/dgcnn_opensrc.py:143-157: graph feature -> conv/BN/LeakyReLU -> max."""
    e = graph_feature(x_bcn, k, idx)
    e = F.conv2d(e, p[conv + ".0.weight"])
    rm, rv = p[bn + ".running_mean"].clone(), p[bn + ".running_var"].clone()
    e = F.batch_norm(e, rm, rv, p[bn + ".weight"], p[bn + ".bias"], training, BN_MOMENTUM, BN_EPS)
    return F.leaky_relu(e, LEAKY).max(dim=-1)[0]


# ----------------------------------------------------------------------------------------------
# PointSegmentationModelBase.predict_full_pointcloud (models/point_seg_net.py:21-48)
# ----------------------------------------------------------------------------------------------
def predict_full_pointcloud(forward, pc, num_classes, sample_points=1024, n_runs_min=50, randperm=torch.randperm):
    """Restatement of models/point_seg_net.py:21-48 over a forward callable `forward(x: 1 x C x n) -> logits`.
    `randperm(n)` is injectable so that a test can draw the permutations from the same (device) generator stream
    as the implementation under test; draws happen in the reference's order. Includes the reference's quirk at :41-43
    of indexing the point cloud with positions *into* `other_pts` rather than with `other_pts[...]`."""
    n_leftover = n_runs_min // 5                                             # :24
    n_initial = n_runs_min - n_leftover                                      # :25
    acc = torch.zeros(pc.shape[0], num_classes, *pc.shape[2:])               # :26
    for _ in range(n_initial):                                               # :27-29
        perm = randperm(pc.shape[-1])[:sample_points]
        acc[..., perm] += torch.softmax(forward(pc[..., perm]), dim=1)
    left_out = torch.nonzero(acc.sum(1) == 0)[..., 1]                        # :32
    if left_out.shape[0] > 0:
        other = torch.nonzero(acc.sum(1))[..., 1]                            # :35
        point_mix = sample_points // 2
        fill_out = sample_points - point_mix
        perm = randperm(n_leftover * point_mix) % len(left_out)              # :38
        for r in range(n_leftover):
            lo = left_out[perm[r * point_mix:(r + 1) * point_mix]]
            oth = randperm(len(other))[:fill_out]                            # :41 (positions, not other[...])
            pts = torch.cat((lo, oth), dim=0)
            acc[..., pts] += torch.softmax(forward(pc[..., pts]), dim=1)     # :43
    return torch.softmax(acc, dim=1)                                         # :48
