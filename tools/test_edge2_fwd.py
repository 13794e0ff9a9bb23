"""Forward of the fused two-layer EdgeConv (fs_edge2_fwd) against the materialised path (fs_edge3_hidden + torch)."""
import sys, torch
sys.path.insert(0, '.')
from fissure_segmentation_b200 import ops, synth, _lib
lib = _lib.load()
dev = 'cuda'
torch.manual_seed(0)
for (B, N, k) in ((2, 256, 8), (4, 2048, 20), (1, 300, 20), (2, 1024, 40), (32, 2048, 20)):
    x, _ = synth.make_batch(B, N, seed=3, jitter=True)
    x = x.to(dev)
    xpm = ops.to_point_major(x).contiguous()
    idx = ops.knn_coords(x, k, self_loop=True)
    P = B * N
    w1 = torch.randn(64, 6, device=dev) * 0.5
    g1 = torch.randn(64, device=dev); b1 = 0.1 * torch.randn(64, device=dev)
    w2 = torch.randn(64, 64, device=dev) * 0.2
    g2 = torch.randn(64, device=dev)
    coef1 = torch.empty(256, device=dev)
    mom = torch.zeros(lib.fs_edge3_moment_doubles(), dtype=torch.float64, device=dev)
    _lib.call("fs_edge3_bn_coef", xpm, xpm, xpm.stride(0), idx, B, N, k, w1, 64, g1, b1, 1e-5, 0.1, mom, coef1, None, None, None)
    h = torch.empty(P * k, 64, dtype=torch.bfloat16, device=dev)
    _lib.call("fs_edge3_hidden", xpm, xpm, xpm.stride(0), idx, B, N, k, w1, 64, coef1, h, 1)
    sgn = torch.where(g2 >= 0, 1.0, -1.0)
    w2b = (w2 * sgn[:, None]).bfloat16().float()
    z = (h.float() @ w2b.t()).view(P, k, 64)              # flipped pre-activation
    best, barg = z.max(dim=1)
    sel_ref = best * sgn
    zz = z * sgn
    s1_ref = zz.double().sum((0, 1)); s2_ref = (zz.double() ** 2).sum((0, 1))
    gram_ref = h.float().t().double() @ h.float().double()
    hsum_ref = h.float().double().sum(0)
    sel = torch.empty(P, 64, device=dev); arg = torch.empty(P, 64, dtype=torch.uint8, device=dev)
    stats = torch.zeros(lib.fs_stats_buffer_doubles(64), dtype=torch.float64, device=dev)
    gram = torch.zeros(64, 64, device=dev); hsum = torch.zeros(64, device=dev)
    assert lib.fs_edge2_supported(k, 64, 64)
    _lib.call("fs_edge2_fwd", xpm, xpm, xpm.stride(0), idx, B, N, k, w1, coef1, w2, 64, g2, sel, arg, stats, gram, hsum)
    torch.cuda.synchronize()
    e_sel = float((sel - sel_ref).abs().max() / sel_ref.abs().max())
    arg_ok = float((arg.long() == barg).float().mean())
    # where arg differs the values must tie
    zsel = torch.gather(z, 1, arg.long().unsqueeze(1)).squeeze(1)
    tie = float((zsel - best).abs().max())
    e_s1 = float(((stats[:64] - s1_ref).abs() / (s1_ref.abs() + 1e-3 * s2_ref.sqrt())).max())
    e_s2 = float(((stats[64:128] - s2_ref).abs() / s2_ref).max())
    e_g = float((gram.double() - gram_ref).abs().max() / gram_ref.abs().max())
    e_h = float((hsum.double() - hsum_ref).abs().max() / hsum_ref.abs().max())
    print("B=%d N=%d k=%d: sel rel %.2e | arg equal %.4f (value gap at differing slots %.2e) | sum z %.2e sum z2 %.2e | gram %.2e hsum %.2e"
          % (B, N, k, e_sel, arg_ok, tie, e_s1, e_s2, e_g, e_h), flush=True)
    # eval variant
    sel2 = torch.empty_like(sel); arg2 = torch.empty_like(arg)
    _lib.call("fs_edge2_fwd", xpm, xpm, xpm.stride(0), idx, B, N, k, w1, coef1, w2, 64, g2, sel2, arg2, None, None, None)
    torch.cuda.synchronize()
    print("   eval variant equal:", bool(torch.equal(sel, sel2)), bool(torch.equal(arg, arg2)))
    if B == 32:
        for fn, name in ((lambda: _lib.call("fs_edge2_fwd", xpm, xpm, xpm.stride(0), idx, B, N, k, w1, coef1, w2, 64, g2, sel, arg, stats, gram, hsum), "train"),
                         (lambda: _lib.call("fs_edge2_fwd", xpm, xpm, xpm.stride(0), idx, B, N, k, w1, coef1, w2, 64, g2, sel, arg, None, None, None), "eval")):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            print("   fs_edge2_fwd [%s] %.1f us" % (name, e0.elapsed_time(e1) * 100))
