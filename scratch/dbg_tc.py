import sys, time, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, _lib
lib = _lib.load()
torch.manual_seed(0)
def al(v, a=256): return (v + a - 1) // a * a
for (B, N, k, sl) in [(2, 2048, 20, True), (2, 2048, 20, False), (3, 1000, 16, True), (32, 2048, 20, True)]:
    feat = torch.randn(B * N, 64, device='cuda') + 0.5 * torch.randn(1, 64, device='cuda')
    ops.USE_TENSOR_CORE_KNN = False
    i0, d0 = ops.knn_features(feat, B, N, k, self_loop=sl, return_dist=True)
    torch.cuda.synchronize()
    ops.USE_TENSOR_CORE_KNN = True
    P = B * N
    nbytes = lib.fs_knn_feat_tc_workspace_bytes(B, N, 64, k)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
    idx = torch.empty(B, N, k, dtype=torch.int32, device='cuda'); dist = torch.empty(B, N, k, device='cuda')
    _lib.call("fs_knn_feat_tc", feat, feat, 64, B, N, 64, k, int(sl), 1, idx, dist, ws, nbytes)
    torch.cuda.synchronize()
    off = 2 * al(P * 512) + al(P * 128) + 2 * al(P * 4) + al(B * 256) + al(B * 4)
    redo = ws[off:off + P]
    same = (idx == i0).all(dim=-1)
    print((B, N, k, sl), "rows", P, "identical rows", int(same.sum()), "redo rows", int(redo.sum()),
          "max dist diff", float((dist - d0).abs().max()))
    if B == 32:
        for name, flag in (("exact simt", False), ("tcgen05", True)):
            ops.USE_TENSOR_CORE_KNN = flag
            for _ in range(3): ops.knn_features(feat, B, N, k, self_loop=sl)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10): ops.knn_features(feat, B, N, k, self_loop=sl)
            torch.cuda.synchronize(); print(name, (time.perf_counter() - t0) / 10 * 1e3, "ms")
