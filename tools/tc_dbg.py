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
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
for dbg in (0, 1, 8, 9, 2, 3, 4, 6, 7, 15):
    os.environ["FS_TC_DBG"] = str(dbg)
    a = t(lambda: ops.knn_coords(x, k, self_loop=True))
    b = t(lambda: ops.knn_features(feat, B, N, k, self_loop=True))
    print("dbg=%2d  knn_coords %7.1f us   knn_features %7.1f us" % (dbg, a, b), flush=True)
