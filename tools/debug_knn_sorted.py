import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from fissure_segmentation_b200 import ops, synth, _lib
from oracle import dgcnn_oracle as O
from parity import compare_knn
lib = _lib.load()
for (B, N, k, sl) in ((1, 8192, 40, False), (2, 8192, 40, True), (2, 2048, 40, False)):
    x, _ = synth.make_batch(B, N, seed=7, n_features=0, jitter=True)
    xd = x.cuda()
    perm = ops.spatial_order(xd)
    xs = torch.gather(xd, 2, perm.unsqueeze(1).expand_as(xd)).contiguous()
    idx = ops.knn_coords(xs, k, self_loop=sl)
    nbytes = lib.fs_knn3d_tc_workspace_bytes(B, N, k)
    ws = ops._workspace(nbytes, xs.device, "knn")
    off = lib.fs_knn_feat_tc_redo_offset(B, N, 3, k)
    redo = ws[off:off + B * N].clone().view(B, N).bool()
    exact, _ = ops.knn_coords(xs, k, self_loop=sl, return_dist=True)
    same = (idx.sort(-1)[0] == exact.sort(-1)[0]).all(-1)
    rep = compare_knn(idx, None, xs.cpu(), k, sl, O.knn_with_gap)
    print("B=%d N=%d k=%d sl=%s: %s | rows differing from SIMT exact: %d, of which redo rows: %d; redo rows total %d"
          % (B, N, k, sl, rep, int((~same).sum()), int((~same & redo).sum()), int(redo.sum())))
    bad = (~same).nonzero()[:3]
    for b, r in bad.tolist():
        print("   row", b, r, "redo", bool(redo[b, r]), "tc", idx[b, r].sort()[0][:12].tolist(), "exact", exact[b, r].sort()[0][:12].tolist())
