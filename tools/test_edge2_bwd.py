"""Fused two-layer EdgeConv (forward + backward) against the materialised paths: bf16 (old) and fp32."""
import sys, torch
sys.path.insert(0, '.')
import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import ops, synth
dev = 'cuda'
torch.backends.cuda.matmul.allow_tf32 = False

def run(B, N, k, mode, training=True, seed=0):
    torch.manual_seed(seed)
    ec = fs.EdgeConv(3, [64, 64], k, first_layer=True).to(dev)
    for l in ec.shared_mlp:
        torch.nn.init.normal_(l.layers[1].weight, 0.0, 1.0)       # both signs of gamma
        torch.nn.init.normal_(l.layers[1].bias, 0.0, 0.2)
    ec.train(training)
    x, _ = synth.make_batch(B, N, seed=3, jitter=True)
    x = x.to(dev)
    xpm = ops.to_point_major(x).contiguous()
    graph = ops.KnnGraph(ops.knn_coords(x, k, self_loop=True))
    ops.USE_FUSED_EDGE2 = mode == "fused"
    cdt = torch.float32 if mode == "fp32" else torch.bfloat16
    out = ec.forward_pm(xpm, B, N, graph, cdt)
    gen = torch.Generator(device=dev).manual_seed(5)
    go = torch.randn(out.shape, device=dev, generator=gen)
    (out * go).sum().backward()
    grads = {n: p.grad.detach().clone() for n, p in ec.named_parameters() if p.grad is not None}
    stats = {n: v.detach().clone() for n, v in ec.state_dict().items() if "running" in n}
    return out.detach(), grads, stats

def rel(a, b): return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))

for (B, N, k) in ((2, 256, 8), (4, 2048, 20), (1, 300, 20), (2, 1024, 40)):
    for training in (True, False):
        o32, g32, s32 = run(B, N, k, "fp32", training)
        o16, g16, s16 = run(B, N, k, "bf16", training)
        ofu, gfu, sfu = run(B, N, k, "fused", training)
        print("B=%d N=%d k=%d train=%s" % (B, N, k, training))
        print("   out        : fused vs fp32 %.2e | old-bf16 vs fp32 %.2e | fused vs old-bf16 %.2e" % (rel(ofu, o32), rel(o16, o32), rel(ofu, o16)))
        for n in g32:
            print("   %-34s: fused vs fp32 %.2e | old-bf16 vs fp32 %.2e" % (n, rel(gfu[n], g32[n]), rel(g16[n], g32[n])))
        for n in s32:
            print("   %-34s: fused vs fp32 %.2e | old-bf16 vs fp32 %.2e" % (n, rel(sfu[n], s32[n]), rel(s16[n], s32[n])))

# timing at the bench shape
B, N, k = 32, 2048, 20
for mode in ("bf16", "fused"):
    torch.manual_seed(0)
    ec = fs.EdgeConv(3, [64, 64], k, first_layer=True).to(dev).train()
    x, _ = synth.make_batch(B, N, seed=3)
    x = x.to(dev); xpm = ops.to_point_major(x).contiguous()
    graph = ops.KnnGraph(ops.knn_coords(x, k, self_loop=True)); graph.reverse()
    ops.USE_FUSED_EDGE2 = mode == "fused"
    def step():
        for p in ec.parameters(): p.grad = None
        out = ec.forward_pm(xpm, B, N, graph, torch.bfloat16)
        out.sum().backward()
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print("ec1 fwd+bwd [%s]: %.1f us (eager, includes launch gaps)" % (mode, e0.elapsed_time(e1) * 100))
