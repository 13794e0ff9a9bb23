"""Entry-point timings of the kNN builds at the bench shapes (CUDA events, L2 flushed between runs)."""
import sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth
dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=15, label=''):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print('%-54s min %8.1f us  med %8.1f us' % (label, ts[0], ts[len(ts) // 2]), flush=True)

for (B, N, k) in ((32, 2048, 20), (32, 2048, 40), (8, 8192, 40)):
    x, _ = synth.make_batch(B, N, seed=5)
    x = x.to(dev)
    perm = ops.spatial_order(x)
    x = torch.gather(x, 2, perm.unsqueeze(1).expand_as(x)).contiguous()
    xpm = x.transpose(1, 2).reshape(B * N, 3).contiguous()
    W = torch.randn(3, 64, device=dev)
    feat = torch.nn.functional.leaky_relu(torch.sin(xpm @ W * 3) + 0.3, 0.2).contiguous()
    for tc in (True, False):
        ops.USE_TENSOR_CORE_KNN = ops.USE_TENSOR_CORE_KNN3D = tc
        tag = 'tcgen05' if tc else 'SIMT'
        timeit(lambda: ops.knn_coords(x, k, self_loop=True), label='knn_coords   B=%d N=%d k=%d [%s]' % (B, N, k, tag))
        if tc or N <= 2048:
            timeit(lambda: ops.knn_features(feat, B, N, k, self_loop=True), label='knn_features B=%d N=%d k=%d C=64 [%s]' % (B, N, k, tag))
    ops.USE_TENSOR_CORE_KNN = ops.USE_TENSOR_CORE_KNN3D = True
    ops.knn_tc_report = {}
    ops.knn_coords(x, k, self_loop=True); ops.knn_features(feat, B, N, k, self_loop=True)
    print('   redo report', ops.knn_tc_report)
    ops.knn_tc_report = None
