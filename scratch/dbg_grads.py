import sys, torch, torch.nn.functional as F
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import synth
from oracle import dgcnn_oracle as O
torch.backends.cuda.matmul.allow_tf32 = False
g = torch.load('tests/golden/dgcnn_golden.pt', weights_only=False)
for tag, ck in (("seg_small_static", "config_small"), ("seg_feat_static", "config_feat")):
    cfg = g[ck]
    x, y = synth.make_batch(cfg["B"], cfg["N"], seed=cfg["data_seed"], n_features=cfg["in_features"] - 3, jitter=True)
    p = O.make_params(O.dgcnn_seg_param_shapes(cfg["in_features"], cfg["num_classes"]), cfg["param_seed"])
    m = fs.DGCNNSeg(k=cfg["k"], in_features=cfg["in_features"], num_classes=cfg["num_classes"], dynamic=False).cuda()
    m.load_state_dict(p); m.train()
    pr = {n: (v.clone().double().requires_grad_(True) if v.dtype.is_floating_point and "running" not in n else (v.double() if v.dtype.is_floating_point else v)) for n, v in p.items()}
    ref64 = O.dgcnn_seg(pr, x.double(), cfg["k"], dynamic=False, training=True)
    F.cross_entropy(ref64, y).backward()
    logits = m(x.cuda()); F.cross_entropy(logits, y.cuda()).backward()
    print(tag, "logit err vs fp64", float((logits.cpu().double()-ref64).abs().max()), "golden(fp32 ref) vs fp64", float((g[tag]["logits"].double()-ref64).abs().max()))
    for n, q in m.named_parameters():
        r64 = pr[n].grad
        mine = float((q.grad.cpu().double() - r64).norm() / r64.norm())
        gold = g[tag]["grads"].get(n)
        gd = float((gold.double() - r64).norm() / r64.norm()) if gold is not None else float('nan')
        print("  %-40s mine-vs-fp64 %.2e   ref32-vs-fp64 %.2e" % (n, mine, gd))
