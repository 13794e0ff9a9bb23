import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import ops, synth
from oracle import dgcnn_oracle as O
DEV = 'cuda:0'
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
def bcn(t, B, N): return t.view(B, N, -1).permute(0, 2, 1)
def err(a, b, name):
    a, b = a.float().cpu(), b.float().cpu()
    d = (a - b).abs()
    print("   %-10s max|d| %.3e  frac>1e-4 %.4f" % (name, float(d.max()), float((d > 1e-4 + 1e-4 * b.abs()).float().mean())), flush=True)

# ---- issue 1: sorted vs unsorted at N=8192, k=40, static
B, N, k, cin = 1, 8192, 40, 9
x, y = synth.make_batch(B, N, seed=109, n_features=cin - 3, jitter=True)
p = O.make_params(O.dgcnn_seg_param_shapes(cin, 4), 107)
m = fs.DGCNNSeg(k=k, in_features=cin, num_classes=4, dynamic=False).to(DEV); m.load_state_dict(p); m.precision = "fp32"; m.train()
xd = x.to(DEV)
with torch.no_grad():
    perm = ops.spatial_order(xd)
    xs = torch.gather(xd, 2, perm.unsqueeze(1).expand_as(xd)).contiguous()
    outs = {}
    for tag, xx in (("unsorted", xd), ("sorted", xs)):
        g = ops.KnnGraph(ops.knn_coords(xx, k, self_loop=False))
        xpm = ops.to_point_major(xx)
        x1 = m.ec1.forward_pm(xpm, B, N, g)
        x2 = m.ec2.forward_pm(x1, B, N, g)
        x3 = m.ec3.forward_pm(x2, B, N, g)
        feats = torch.cat([x1, x2, x3], 1)
        glob = m.global_feature[0].forward_pool_pm(feats, B, N)
        outs[tag] = (x1, x2, x3, glob, g)
    inv = torch.empty_like(perm); inv.scatter_(1, perm, torch.arange(N, device=DEV).unsqueeze(0))
    for i, nm in enumerate(("x1", "x2", "x3")):
        a = outs["sorted"][i].view(B, N, -1)
        a_un = torch.gather(a, 1, inv.unsqueeze(-1).expand_as(a))     # sorted row r corresponds to original perm[r]
        err(a_un.reshape(B * N, -1), outs["unsorted"][i], nm + " s/u")
    err(outs["sorted"][3], outs["unsorted"][3], "glob s/u")
    # full model: sort on/off
    m.spatial_sort = True; l1 = m(xd)
    m.spatial_sort = False; l0 = m(xd)
    err(l1, l0, "logits s/u")

# ---- issue 2: eval after a bf16 training step
B, N, k = 32, 2048, 20
x, y = synth.make_batch(B, N, seed=1234)
p = O.make_params(O.dgcnn_seg_param_shapes(3, 4), 77)
for train_prec in ("fp32", "bf16"):
    m = fs.DGCNNSeg(k=k, in_features=3, num_classes=4, dynamic=True).to(DEV); m.load_state_dict(p); m.precision = train_prec; m.train()
    logits = m(x.to(DEV)); F.cross_entropy(logits, y.to(DEV)).backward()
    m.precision = "fp32"; m.eval()
    sd = {n: v.detach().cpu().clone() for n, v in m.state_dict().items()}
    with torch.no_grad():
        ev = m(x.to(DEV))[:2].cpu()
        xpm = ops.to_point_major(x.to(DEV))
        g1 = m.ec1.build_graph(xpm, B, N); x1 = m.ec1.forward_pm(xpm, B, N, g1)
        g2 = m.ec2.build_graph(x1, B, N); x2 = m.ec2.forward_pm(x1, B, N, g2)
        g3 = m.ec3.build_graph(x2, B, N); x3 = m.ec3.forward_pm(x2, B, N, g3)
    gs = [g.idx[:2].cpu().long() for g in (g1, g2, g3)]
    xs = x[:2]
    o1 = O.edgeconv(xs, sd, "ec1", 2, k, gs[0], True, False)
    o2 = O.edgeconv(o1, sd, "ec2", 1, k, gs[1], False, False)
    o3 = O.edgeconv(o2, sd, "ec3", 1, k, gs[2], False, False)
    print("== eval after a", train_prec, "training step")
    err(bcn(x1, B, N)[:2], o1, "x1"); err(bcn(x2, B, N)[:2], o2, "x2"); err(bcn(x3, B, N)[:2], o3, "x3")
    ref = O.dgcnn_seg(sd, xs, k, dynamic=True, training=False, fixed_graphs=gs)
    err(ev, ref, "logits")
    for n in ("ec1.shared_mlp.1.layers.1.running_var", "ec1.shared_mlp.1.layers.1.running_mean", "segmentation.0.layers.1.running_var"):
        print("   ", n, float(sd[n].min()), float(sd[n].max()))
