"""Four feature-space kNN builds at the bench shape (target of the ncu captures of the tcgen05 path)."""
import sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth
B, N, k = 32, 2048, 20
x, _ = synth.make_batch(B, N, seed=5)
x = x.cuda().transpose(1, 2).reshape(B * N, 3)
W = torch.randn(3, 64, device='cuda')
feat = torch.nn.functional.leaky_relu(torch.sin(x @ W * 3) + 0.3, 0.2).contiguous()
for _ in range(4):
    ops.knn_features(feat, B, N, k, self_loop=True)
torch.cuda.synchronize()
print("done")
