"""Micro-benchmarks of single entry points at the bench shape (B=32, N=2048, k=20, Cp=64)."""
import os, sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth, _lib
B, N, k, Cp = 32, 2048, 20, 64
dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, n=20, label=''):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print('%-40s min %8.1f us  med %8.1f us' % (label, ts[0], ts[len(ts) // 2]), flush=True)
    return ts[len(ts) // 2]

x, _ = synth.make_batch(B, N, seed=5)
x = x.to(dev)
perm = ops.spatial_order(x)
x = torch.gather(x, 2, perm.unsqueeze(1).expand_as(x))
xpm = x.transpose(1, 2).reshape(B * N, 3).contiguous()
W = torch.randn(3, 64, device=dev)
feat = torch.nn.functional.leaky_relu(torch.sin(xpm @ W * 3) + 0.3, 0.2).contiguous()

timeit(lambda: ops.knn_coords(x, k, self_loop=True), label='knn_coords (3-D)')
timeit(lambda: ops.knn_features(feat, B, N, k, self_loop=True), label='knn_features (tc path)')
_lib.time_calls.clear()

idx = ops.knn_features(feat, B, N, k, self_loop=True)
graph = ops.KnnGraph(idx)
rev_ptr, rev_src = graph.reverse()
table = torch.randn(B * N, 2 * Cp, device=dev)
gamma = torch.randn(Cp, device=dev)
P = B * N
def gather(mode):
    os.environ['FS_GATHER'] = mode
    sel = torch.empty(P, Cp, device=dev); arg = torch.empty(P, Cp, dtype=torch.uint8, device=dev)
    sy = torch.empty(P, Cp, device=dev); stats = ops._stats_buffer(Cp, dev)
    def run():
        stats.zero_()
        _lib.call("fs_edgeconv_gather", table, table, 0, table.stride(0), idx, B, N, k, Cp, gamma, rev_ptr, sel, arg, sy, stats)
    t = timeit(run, label='edgeconv_gather train [%s]' % mode)
    return sel, arg, sy, stats, t
a = gather('global'); b = gather('smem')
print('sel equal', torch.equal(a[0], b[0]), 'arg equal', torch.equal(a[1], b[1]), 'sy maxdiff', float((a[2] - b[2]).abs().max()),
      'stats rel', float(((a[3][:2 * Cp] - b[3][:2 * Cp]).abs() / (a[3][:2 * Cp].abs() + 1e-9)).max()))
alg = P * (2 * Cp * 4 + 4 * k + 4 * Cp + Cp + 4 * Cp)
for m, r in (('global', a), ('smem', b)):
    print('%s: %.1f GB/s algorithmic (%.3f of 6556)' % (m, alg / r[4] / 1e3, alg / r[4] / 1e3 / 6556.2))
# idx from the 3-D graph (spatially local neighbours)
idx3 = ops.knn_coords(x, k, self_loop=True)
idx_save = idx; idx = idx3
rev_ptr, rev_src = ops.KnnGraph(idx3).reverse()
gather('global'); gather('smem')
