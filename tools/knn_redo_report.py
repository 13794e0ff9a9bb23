import sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth
dev = 'cuda'
for (B, N, k, sl) in ((1, 8192, 40, False), (1, 8192, 40, True), (32, 2048, 20, True), (32, 2048, 40, False), (4, 8192, 20, False)):
    x, y = synth.make_batch(B, N, seed=1234, n_features=0)
    x = x.to(dev)
    perm = ops.spatial_order(x)
    xs = torch.gather(x, 2, perm.unsqueeze(1).expand_as(x))
    for name, xx in (("sorted", xs), ("unsorted", x)):
        ops.knn_tc_report = {}
        ops.knn_coords(xx, k, self_loop=sl)
        torch.cuda.synchronize()
        print(B, N, k, sl, name, ops.knn_tc_report)
    ops.knn_tc_report = None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): ops.knn_coords(xs, k, self_loop=sl)
    e0.record()
    for _ in range(10): ops.knn_coords(xs, k, self_loop=sl)
    e1.record(); torch.cuda.synchronize()
    print("   time per call us", e0.elapsed_time(e1) * 100)
