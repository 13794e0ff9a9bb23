import os, sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth, _lib
B, N, k = 32, 2048, 20
x, _ = synth.make_batch(B, N, seed=5)
x = x.cuda()
perm = ops.spatial_order(x)
x = torch.gather(x, 2, perm.unsqueeze(1).expand_as(x)).contiguous()
xpm = x.transpose(1, 2).reshape(B * N, 3)
W = torch.randn(3, 64, device='cuda')
feat = torch.nn.functional.leaky_relu(torch.sin(xpm @ W * 3) + 0.3, 0.2).contiguous()
for _ in range(2):
    ops.knn_coords(x, k, self_loop=True)
    ops.knn_features(feat, B, N, k, self_loop=True)
torch.cuda.synchronize()
