"""Seeded synthetic "lung-keypoint-shaped" point clouds.

A case is built in a 300 x 250 x 320 voxel box (1 mm spacing): two ellipsoidal lungs, three smooth
fissure sheets (two in the right lung, one in the left) sampled as height fields with sigma = 3 voxel
noise (data_processing/keypoint_extraction.py:33-50) and mixed 50/50 with thinned interior lattice
points (Foerstner-like, spacing 5 voxels, keypoint_extraction.py:182). Integer voxel coordinates map
to grid coordinates with kpts_to_grid(align_corners=False) (utils/general_utils.py:105-130), so xyz
lies on a lattice in [-1, 1]. Labels: 0 background, 1..3 nearest sheet within 2 voxels. Each training
sample is a random N-subset of a case (data.py:451) with the train-time augmentation of
augmentations.py:52-75 (rotation <= 0.1 pi about a random axis, translation +-0.1, scale in [0.9, 1]).
Everything is generated on the CPU from a torch.Generator, so GPU and CPU arms see identical tensors.
"""
import math

import torch

SHAPE_DHW = (320, 250, 300)   # D, H, W (z, y, x)
MAX_KPTS = 20000


def _sheet_height(u, v, coeff):
    """Smooth height field z(u, v) in voxels; u, v in [0, 1]."""
    a0, a1, a2, a3, a4, a5 = coeff
    return a0 + a1 * u + a2 * v + a3 * torch.sin(math.pi * u) + a4 * torch.cos(math.pi * v) + a5 * u * v


def make_case(gen, n_points=MAX_KPTS):
    """One synthetic case: integer voxel keypoints (n, 3) in xyz order and labels (n,)."""
    D, H, W = SHAPE_DHW
    # lungs: ellipsoids (centre, radii) in voxel xyz
    lungs = [((95.0, 125.0, 160.0), (62.0, 88.0, 118.0)),    # right lung (image left)
             ((205.0, 125.0, 160.0), (58.0, 85.0, 112.0))]   # left lung
    # three sheets: (lung, coefficients of the height field in voxels)
    jitter = lambda s: (torch.rand(6, generator=gen) * 2 - 1) * s  # noqa: E731
    sheets = [
        (0, torch.tensor([150.0, 55.0, -25.0, 14.0, 8.0, 10.0]) + jitter(4.0)),     # right oblique
        (0, torch.tensor([200.0, -8.0, 12.0, 6.0, -5.0, 4.0]) + jitter(3.0)),       # right horizontal
        (1, torch.tensor([145.0, 60.0, -30.0, 12.0, 10.0, -8.0]) + jitter(4.0)),    # left oblique
    ]
    n_sheet = n_points // 2
    n_interior = n_points - n_sheet

    pts, labels = [], []
    per_sheet = [n_sheet // 3 + (1 if i < n_sheet % 3 else 0) for i in range(3)]
    for si, ((lung, coeff), cnt) in enumerate(zip(sheets, per_sheet)):
        (cx, cy, cz), (rx, ry, rz) = lungs[lung]
        got = 0
        chunks = []
        while got < cnt:
            m = (cnt - got) * 2 + 64
            u = torch.rand(m, generator=gen)
            v = torch.rand(m, generator=gen)
            x = cx + (u * 2 - 1) * rx
            y = cy + (v * 2 - 1) * ry
            z = _sheet_height(u, v, coeff)
            inside = ((x - cx) / rx) ** 2 + ((y - cy) / ry) ** 2 + ((z - cz) / rz) ** 2 <= 1.0
            p = torch.stack([x, y, z], dim=1)[inside]
            chunks.append(p)
            got += p.shape[0]
        p = torch.cat(chunks)[:cnt]
        noisy = (p + torch.randn(p.shape, generator=gen) * 3.0).long().float()   # get_noisy_keypoints
        # label = sheet id if still within 2 voxels of the sheet, else background
        u = ((noisy[:, 0] - cx) / rx + 1) / 2
        v = ((noisy[:, 1] - cy) / ry + 1) / 2
        dz = (noisy[:, 2] - _sheet_height(u, v, coeff)).abs()
        pts.append(noisy)
        labels.append(torch.where(dz <= 2.0, torch.full_like(dz, si + 1), torch.zeros_like(dz)).long())

    # interior keypoints: lattice with 5-voxel spacing inside the lungs, randomly thinned
    gx = torch.arange(2, W, 5, dtype=torch.float32)
    gy = torch.arange(2, H, 5, dtype=torch.float32)
    gz = torch.arange(2, D, 5, dtype=torch.float32)
    grid = torch.stack(torch.meshgrid(gx, gy, gz, indexing="ij"), dim=-1).reshape(-1, 3)
    inside = torch.zeros(grid.shape[0], dtype=torch.bool)
    for (cx, cy, cz), (rx, ry, rz) in lungs:
        inside |= ((grid[:, 0] - cx) / rx) ** 2 + ((grid[:, 1] - cy) / ry) ** 2 + ((grid[:, 2] - cz) / rz) ** 2 <= 1.0
    grid = grid[inside]
    keep = torch.randperm(grid.shape[0], generator=gen)[:n_interior]
    interior = grid[keep]
    if interior.shape[0] < n_interior:   # not enough lattice sites: top up with jittered copies
        extra = interior[torch.randint(0, interior.shape[0], (n_interior - interior.shape[0],), generator=gen)]
        interior = torch.cat([interior, (extra + torch.randint(-2, 3, extra.shape, generator=gen)).float()])
    pts.append(interior)
    labels.append(torch.zeros(interior.shape[0], dtype=torch.long))

    kp = torch.cat(pts)
    lab = torch.cat(labels)
    lim = torch.tensor([W - 1, H - 1, D - 1], dtype=torch.float32)
    kp = torch.minimum(torch.clamp(kp, min=0), lim)
    return kp, lab


def voxels_to_grid(kp_xyz):
    """kpts_to_grid(align_corners=False): v -> (2 v - (S - 1)) / S per axis (xyz order)."""
    D, H, W = SHAPE_DHW
    size = torch.tensor([W, H, D], dtype=torch.float32)
    return (2 * kp_xyz - (size - 1)) / size


def _rotation_matrices(rotvec):
    """Rodrigues formula for a batch of axis-angle vectors (B, 3) -> (B, 3, 3)."""
    theta = rotvec.norm(dim=1, keepdim=True).clamp_min(1e-12)
    kx, ky, kz = (rotvec / theta).unbind(1)
    zero = torch.zeros_like(kx)
    K = torch.stack([zero, -kz, ky, kz, zero, -kx, -ky, kx, zero], dim=1).view(-1, 3, 3)
    th = theta.view(-1, 1, 1)
    eye = torch.eye(3).expand_as(K)
    return eye + torch.sin(th) * K + (1 - torch.cos(th)) * (K @ K)


def augment(points_b3n, gen, rotation_amount=0.1, translation_amount=0.1, scale_amount=0.1):
    B = points_b3n.shape[0]
    axis = torch.rand(B, 3, generator=gen) * 2 - 1
    axis = axis / axis.norm(dim=1, keepdim=True)
    R = _rotation_matrices(axis * math.pi * rotation_amount)
    t = (torch.rand(B, 3, generator=gen) * 2 - 1) * translation_amount
    s = 1 - torch.rand(B, 1, generator=gen) * scale_amount
    return (R @ points_b3n) * s.view(B, 1, 1) + t.view(B, 3, 1)


def make_batch(batch, n_points, seed=1234, n_features=0, jitter=False, augmentation=True, n_cases=4):
    """Returns x (batch, 3 + n_features, n_points) float32 and labels (batch, n_points) int64.

    jitter=True adds U(-0.5, 0.5) voxel noise before the grid mapping: the continuous variant used for
    the bit-exact kNN claim; the default lattice variant is the tie-heavy, realistic one."""
    gen = torch.Generator().manual_seed(seed)
    cases = [make_case(gen) for _ in range(min(n_cases, batch))]
    xs, ys = [], []
    for b in range(batch):
        kp, lab = cases[b % len(cases)]
        sub = torch.randperm(kp.shape[0], generator=gen)[:n_points]
        p = kp[sub]
        if jitter:
            p = p + (torch.rand(p.shape, generator=gen) - 0.5)
        xs.append(voxels_to_grid(p).t())
        ys.append(lab[sub])
    x = torch.stack(xs)
    y = torch.stack(ys)
    if augmentation:
        x = augment(x, gen)
    if n_features > 0:
        # MIND-like descriptors in (0, 1]: smooth functions of position plus noise, exp(-d / var)
        freq = torch.rand(n_features, 3, generator=gen) * 4 + 1
        phase = torch.rand(n_features, 1, generator=gen) * 2 * math.pi
        resp = torch.sin(torch.einsum("fc,bcn->bfn", freq, x) + phase) ** 2
        resp = resp + 0.1 * torch.rand(resp.shape, generator=gen)
        x = torch.cat([x, torch.exp(-resp)], dim=1)
    return x.contiguous().float(), y.contiguous()


def make_chamfer_pair(batch, n_points, seed=1234, sigma=0.02):
    """Prediction / target clouds (B, N, 3) for the Chamfer configuration: target = prediction + N(0, sigma)."""
    x, _ = make_batch(batch, n_points, seed=seed, jitter=True)
    gen = torch.Generator().manual_seed(seed + 7)
    pred = x.transpose(1, 2).contiguous()
    target = pred + sigma * torch.randn(pred.shape, generator=gen)
    return pred, target.contiguous()
