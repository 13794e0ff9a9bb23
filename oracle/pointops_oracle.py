"""ORACLE — test infrastructure, not product code.

CPU restatement of the two live `pointops_cuda` operators behind
models/pointtransformer/pointops.py:16-62.

PARITY UNPINNED: `pointops_cuda` is the CUDA extension of POSTECH-CVLab/point-transformer
(lib/pointops), named at pointops.py:1-4, not vendored, not pinned and CUDA-only. The restatement
follows the docstrings and call sites in pointops.py: knnquery = per-segment brute-force kNN of new_xyz
in xyz with ascending squared distances and global indices (sqrt is applied by the caller, :60);
furthestsampling = per-segment farthest point sampling that starts at the segment's first point with
all running distances initialised to 1e10 (:32).
"""
import numpy as np


def knnquery(nsample, xyz, new_xyz, offset, new_offset):
    """xyz (n,3), new_xyz (m,3), cumulative offsets -> idx (m,nsample) int32, dist2 (m,nsample) f32."""
    xyz = np.asarray(xyz, dtype=np.float32)
    new_xyz = np.asarray(new_xyz, dtype=np.float32)
    m = new_xyz.shape[0]
    idx = np.zeros((m, nsample), dtype=np.int32)
    dist2 = np.zeros((m, nsample), dtype=np.float32)
    s0 = q0 = 0
    for s1, q1 in zip(offset, new_offset):
        ref = xyz[s0:s1]
        qry = new_xyz[q0:q1]
        d = ((qry[:, None, :] - ref[None, :, :]) ** 2).sum(-1, dtype=np.float32)
        order = np.argsort(d, axis=1, kind="stable")[:, :nsample]
        got = order.shape[1]
        idx[q0:q1, :got] = order + s0
        dist2[q0:q1, :got] = np.take_along_axis(d, order, axis=1)
        if got < nsample:
            idx[q0:q1, got:] = s0
            dist2[q0:q1, got:] = 1e10
        s0, q0 = s1, q1
    return idx, dist2


def furthestsampling(xyz, offset, new_offset):
    xyz = np.asarray(xyz, dtype=np.float32)
    out = np.zeros(int(new_offset[-1]), dtype=np.int32)
    s0 = q0 = 0
    for s1, q1 in zip(offset, new_offset):
        pts = xyz[s0:s1]
        tmp = np.full(pts.shape[0], 1e10, dtype=np.float32)
        last = 0
        out[q0] = s0
        for r in range(1, q1 - q0):
            diff = pts - pts[last]
            d = (diff[:, 2] * diff[:, 2] + (diff[:, 1] * diff[:, 1] + diff[:, 0] * diff[:, 0])).astype(np.float32)
            tmp = np.minimum(tmp, d)
            last = int(np.argmax(tmp))
            out[q0 + r] = s0 + last
        s0, q0 = s1, q1
    return out
