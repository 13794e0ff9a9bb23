"""Three launches of the train-mode EdgeConv gather at the bench shape (target of the ncu --set full capture)."""
import sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth, _lib
B, N, k, Cp = 32, 2048, 20, 64
dev = 'cuda'
x, _ = synth.make_batch(B, N, seed=5)
x = x.to(dev)
perm = ops.spatial_order(x)
x = torch.gather(x, 2, perm.unsqueeze(1).expand_as(x))
xpm = x.transpose(1, 2).reshape(B * N, 3).contiguous()
W = torch.randn(3, 64, device=dev)
feat = torch.nn.functional.leaky_relu(torch.sin(xpm @ W * 3) + 0.3, 0.2).contiguous()
idx = ops.knn_features(feat, B, N, k, self_loop=True)
graph = ops.KnnGraph(idx)
rev_ptr, rev_src = graph.reverse()
table = torch.randn(B * N, 2 * Cp, device=dev)
gamma = torch.randn(Cp, device=dev)
P = B * N
sel = torch.empty(P, Cp, device=dev); arg = torch.empty(P, Cp, dtype=torch.uint8, device=dev)
sy = torch.empty(P, Cp, device=dev); stats = ops._stats_buffer(Cp, dev)
for _ in range(3):
    _lib.call("fs_edgeconv_gather", table, table, 0, table.stride(0), idx, B, N, k, Cp, gamma, rev_ptr, sel, arg, sy, stats)
torch.cuda.synchronize()
print('done')
