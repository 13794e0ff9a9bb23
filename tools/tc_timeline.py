import ctypes, sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth, _lib
lib = _lib.load()
B, N, k = 32, 2048, 20
x, _ = synth.make_batch(B, N, seed=5)
x = x.cuda()
perm = ops.spatial_order(x)
x = torch.gather(x, 2, perm.unsqueeze(1).expand_as(x)).contiguous()
xpm = x.transpose(1, 2).reshape(B * N, 3)
W = torch.randn(3, 64, device='cuda')
feat = torch.nn.functional.leaky_relu(torch.sin(xpm @ W * 3) + 0.3, 0.2).contiguous()
for _ in range(2): ops.knn_features(feat, B, N, k, self_loop=True)
ncta = 16 * B
buf = torch.zeros(ncta * 16 * 8, dtype=torch.int64, device='cuda')
lib.fs_tc_set_timeline.argtypes = [ctypes.c_void_p]
lib.fs_tc_set_timeline(buf.data_ptr())
ops.knn_features(feat, B, N, k, self_loop=True)
torch.cuda.synchronize()
lib.fs_tc_set_timeline(None)
t = buf.view(ncta, 16, 8).cpu().double()
# per CTA: warp 0 (epilogue) phases relative to its start; clock64 is per-SM so only differences within a CTA are meaningful
ep = t[:, 0, :]
d = ep - ep[:, :1]
names = ["start", "setup done", "A loaded", "sweep1 done", "between done", "sweep2 done", "exit"]
for i, n in enumerate(names):
    print("%-14s mean %9.0f cyc   min %9.0f   max %9.0f" % (n, d[:, i].mean(), d[:, i].min(), d[:, i].max()))
mma = t[:, 9, :]
print("MMA warp: setup->exit mean %.0f" % (mma[:, 6] - mma[:, 0]).mean())
for w in (0, 3, 4, 7):
    e = t[:, w, :]
    print("warp %d: sweep1 %.0f  between %.0f  sweep2 %.0f" % (w, (e[:, 3] - e[:, 2]).mean(), (e[:, 4] - e[:, 3]).mean(), (e[:, 5] - e[:, 4]).mean()))
