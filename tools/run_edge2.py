"""ec1 (two-layer EdgeConv on coordinates) forward + backward at the bench shape, fused tcgen05 path."""
import sys, torch
sys.path.insert(0, '.')
import fissure_segmentation_b200 as fs
from fissure_segmentation_b200 import ops, synth
dev = 'cuda'
B, N, k = 32, 2048, 20
torch.manual_seed(0)
ec = fs.EdgeConv(3, [64, 64], k, first_layer=True).to(dev).train()
x, _ = synth.make_batch(B, N, seed=3)
x = x.to(dev)
perm = ops.spatial_order(x)
x = torch.gather(x, 2, perm.unsqueeze(1).expand_as(x)).contiguous()
xpm = ops.to_point_major(x).contiguous()
graph = ops.KnnGraph(ops.knn_coords(x, k, self_loop=True)); graph.reverse()
mode = sys.argv[1] if len(sys.argv) > 1 else "fused"
ops.USE_FUSED_EDGE2 = mode == "fused"
for _ in range(3):
    for p in ec.parameters(): p.grad = None
    out = ec.forward_pm(xpm, B, N, graph, torch.bfloat16)
    out.sum().backward()
torch.cuda.synchronize()
print("done")
