import sys, time, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, _lib, synth
lib = _lib.load()
torch.manual_seed(0)
def al(v, a=256): return (v + a - 1) // a * a
def feats(B, N, kind):
    if kind == "randn":
        return torch.randn(B * N, 64, device='cuda') + 0.5 * torch.randn(1, 64, device='cuda')
    # smooth low-dimensional features of lung-shaped clouds (like EdgeConv activations)
    x, _ = synth.make_batch(B, N, seed=5)
    x = x.cuda().transpose(1, 2).reshape(B * N, 3)
    W = torch.randn(3, 64, device='cuda')
    return torch.nn.functional.leaky_relu(torch.sin(x @ W * 3) + 0.3, 0.2).contiguous()
for kind in ("randn", "smooth"):
  for (B, N, k, sl) in [(2, 2048, 20, True), (2, 2048, 20, False), (3, 1000, 16, True), (1, 8192, 20, True), (32, 2048, 20, True)]:
    feat = feats(B, N, kind)
    ops.USE_TENSOR_CORE_KNN = False
    i0 = ops.knn_features(feat, B, N, k, self_loop=sl)
    torch.cuda.synchronize()
    P = B * N
    nbytes = lib.fs_knn_feat_tc_workspace_bytes(B, N, 64, k)
    ws = torch.zeros(nbytes, dtype=torch.uint8, device='cuda')
    idx = torch.full((B, N, k), -1, dtype=torch.int32, device='cuda')
    _lib.call("fs_knn_feat_tc", feat, feat, 64, B, N, 64, k, int(sl), 1, idx, None, ws, nbytes)
    torch.cuda.synchronize()
    off_n = 2 * al(P * 512) + 2 * al(P * 512)
    cnt = ws[off_n:off_n + 4 * P].view(torch.int32)
    off_redo = off_n + 3 * al(P * 4) + al(B * 256) + al(B * 8)
    redo = ws[off_redo:off_redo + P]
    same = (idx.sort(-1)[0] == i0.sort(-1)[0]).all(dim=-1)
    print(kind, (B, N, k, sl), "rows", P, "identical sets", int(same.sum()), "redo", int(redo.sum()),
          "survivors mean %.1f max %d" % (float(cnt.float().mean()), int(cnt.max())))
    if B == 32:
        for name, flag in (("exact simt", False), ("tcgen05", True)):
            ops.USE_TENSOR_CORE_KNN = flag
            for _ in range(3): ops.knn_features(feat, B, N, k, self_loop=sl)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(10): ops.knn_features(feat, B, N, k, self_loop=sl)
            torch.cuda.synchronize(); print("  ", name, (time.perf_counter() - t0) / 10 * 1e3, "ms")
