"""Shared helpers for the GPU parity tests: tie-aware kNN comparison (SURVEY 8c protocol)."""
import torch

EPS32 = 1.1920929e-07


def compare_knn(idx_gpu, dist_gpu, x_bcn_cpu, k, self_loop, oracle_knn_with_gap):
    """Compare a CUDA kNN result with the oracle row by row.

    A row is a *tie row* when the oracle's gap between the last kept and the first rejected neighbour
    is within 16 eps (|x_i|^2 + max_j |x_j|^2) — there the reference's own fp32 answer is arbitrary. Non-tie
    rows must have exactly the oracle's index SET; sorted distances must agree to rtol 1e-5 (abs 1e-6
    of the squared-norm scale). Returns a report dict.
    """
    ref_i, ref_d, next_d, sq = oracle_knn_with_gap(x_bcn_cpu, k, self_loop)
    gi = idx_gpu.cpu().long()
    B, N, _ = gi.shape
    scale = 16 * EPS32 * (sq.unsqueeze(-1) + sq.max(dim=1, keepdim=True)[0].unsqueeze(-1))     # (B, N, 1) upper bound
    gap = (next_d - ref_d[..., -1]).unsqueeze(-1)
    # also ties inside the kept list do not matter for sets; only the boundary does
    tie_row = (gap <= scale).squeeze(-1)
    if not self_loop:
        # the reference drops column 0 of a k+1 search (general_utils.py:317-322); when the two smallest
        # distances are within rounding of each other (duplicate keypoints) WHICH point is dropped is arbitrary
        _, d01, _, _ = oracle_knn_with_gap(x_bcn_cpu, 2, True)
        tie_row |= ((d01[..., 1] - d01[..., 0]).unsqueeze(-1) <= scale).squeeze(-1)
    same_set = (gi.sort(dim=-1)[0] == ref_i.sort(dim=-1)[0]).all(dim=-1)
    report = {
        "rows": B * N,
        "tie_rows": int(tie_row.sum()),
        "mismatch_rows": int((~same_set).sum()),
        "mismatch_non_tie_rows": int((~same_set & ~tie_row).sum()),
    }
    if dist_gpu is not None:
        gd = dist_gpu.cpu()
        tol = 1e-5 * ref_d.abs() + 8 * EPS32 * (sq.unsqueeze(-1) * 2 + 1e-6)
        report["dist_bad"] = int(((gd.sort(dim=-1)[0] - ref_d.sort(dim=-1)[0]).abs() > tol + scale).sum())
    return report


def rel_err(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def assert_close(a, b, rtol, atol, what=""):
    a, b = a.float().cpu(), b.float().cpu()
    bad = (a - b).abs() > atol + rtol * b.abs()
    assert not bad.any(), "%s: %d / %d elements out of tolerance, max abs diff %.3e" % (
        what, int(bad.sum()), bad.numel(), float((a - b).abs().max()))
