"""Layer-wise comparison against the oracle with OUR graphs teacher-forced (diagnostics)."""
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
    print("   %-8s max|d| %.3e  rel-fro %.3e  frac>1e-4 %.4f" % (name, float(d.max()), float((a - b).norm() / b.norm()), float((d > 1e-4 + 1e-4 * b.abs()).float().mean())))

def run(tag, B, N, k, cin, dynamic, training, sl=None, seed=7, sort=True):
    print("==", tag)
    x, y = synth.make_batch(B, N, seed=seed, n_features=cin - 3, jitter=True)
    p = O.make_params(O.dgcnn_seg_param_shapes(cin, 4), 77)
    m = fs.DGCNNSeg(k=k, in_features=cin, num_classes=4, dynamic=dynamic).to(DEV)
    m.load_state_dict(p); m.precision = "fp32"; m.spatial_sort = sort
    m.train(training)
    sl = sl or slice(0, B)
    with torch.no_grad():
        xd = x.to(DEV)
        x_pm = ops.to_point_major(xd)
        if dynamic:
            g1 = m.ec1.build_graph(x_pm, B, N)
        else:
            g1 = ops.KnnGraph(ops.knn_coords(xd, k, self_loop=False))
        x1 = m.ec1.forward_pm(x_pm, B, N, g1)
        g2 = m.ec2.build_graph(x1, B, N) if dynamic else g1
        x2 = m.ec2.forward_pm(x1, B, N, g2)
        g3 = m.ec3.build_graph(x2, B, N) if dynamic else g1
        x3 = m.ec3.forward_pm(x2, B, N, g3)
        logits = m(xd)
    pr = {n: v.clone() for n, v in p.items()}
    if training:
        # batch statistics couple the clouds: the oracle must see the whole batch
        xs, gs = x, [g.idx.cpu().long() for g in (g1, g2, g3)]
        sl = slice(0, B)
    else:
        xs, gs = x[sl], [g.idx[sl].cpu().long() for g in (g1, g2, g3)]
    o1 = O.edgeconv(xs, pr, "ec1", 2, k, gs[0], True, training)
    o2 = O.edgeconv(o1, pr, "ec2", 1, k, gs[1], False, training)
    o3 = O.edgeconv(o2, pr, "ec3", 1, k, gs[2], False, training)
    err(bcn(x1, B, N)[sl], o1, "x1"); err(bcn(x2, B, N)[sl], o2, "x2"); err(bcn(x3, B, N)[sl], o3, "x3")
    # layer-wise with the ORACLE's previous activation as input (isolates each layer)
    with torch.no_grad():
        if not training or True:
            a2 = m.ec2.forward_pm(ops.to_point_major(o1.to(DEV)).contiguous(), xs.shape[0], N, ops.KnnGraph(gs[1].to(DEV).int().contiguous()))
            err(bcn(a2, xs.shape[0], N), o2, "ec2|o1")
    ref = O.dgcnn_seg(pr, xs, k, dynamic=dynamic, training=training, fixed_graphs=gs)
    err(logits[sl], ref, "logits")

run("config C static train B=1 N=8192 k=40 cin=9", 1, 8192, 40, 9, False, True)
run("config C static train, no spatial sort", 1, 8192, 40, 9, False, True, sort=False)
run("cin=9 static train B=1 N=2048 k=40", 1, 2048, 40, 9, False, True)
run("cin=9 static train B=1 N=8192 k=20", 1, 8192, 20, 9, False, True)
run("cin=3 static train B=1 N=8192 k=40", 1, 8192, 40, 3, False, True)
run("B=32 eval dynamic slice", 32, 2048, 20, 3, True, False, sl=slice(0, 2))
run("B=8 eval dynamic slice", 8, 2048, 20, 3, True, False, sl=slice(0, 2))
run("B=2 eval dynamic", 2, 2048, 20, 3, True, False)
run("B=32 eval static slice", 32, 2048, 20, 3, False, False, sl=slice(0, 2))
