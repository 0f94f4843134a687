"""Restatement of the JSD occupancy metric -- TEST INFRASTRUCTURE ONLY.

Follows reference utils/metrics/jsd.py: unit_cube_grid_point_cloud (:10-20), the brute-force voting of
entropy_of_occupancy_grid (:42-84, distances as (p-g).pow(2).sum(-1) in f32, argmin = first minimum),
_entropy (:95-107, including its in-place ``p += eps``) and _jensen_shannon_divergence (:110-121).
Pinned by tests/golden/jsd_cpu.npz (the reference's own compute_jsd / grid_counters).
"""
import numpy as np
import torch


def grid_points(resolution, clip_sphere=True):
    spacing = 1.0 / float(resolution - 1)
    steps = torch.arange(resolution)
    grid = torch.stack(torch.meshgrid(steps, steps, steps, indexing="ij"), dim=-1) * spacing - 0.5
    grid = grid.reshape(-1, 3)
    if clip_sphere:
        grid = grid[torch.norm(grid, dim=1) <= 0.5]
    return grid.numpy().astype(np.float32), spacing


def vote(pcs, resolution=28, in_sphere=True):
    """-> (grid_counters (Ng,), clouds_touching (Ng,)) as int64."""
    grid, _ = grid_points(resolution, in_sphere)
    pcs = np.asarray(pcs, np.float32)
    B, Np, _ = pcs.shape
    counters = np.zeros(len(grid), np.int64)
    touching = np.zeros(len(grid), np.int64)
    for b in range(B):
        idx = np.empty(Np, np.int64)
        for j in range(0, Np, 256):
            d = pcs[b, j:j + 256, None, :] - grid[None, :, :]          # f32
            d = d * d
            dist = (d[..., 0] + d[..., 1]) + d[..., 2]
            idx[j:j + 256] = dist.argmin(axis=1)
        np.add.at(counters, idx, 1)
        touching[np.unique(idx)] += 1
    return counters, touching


def _entropy(p, base2=True, eps=np.float32(1e-8)):
    p += eps                                                           # in place, like the reference
    log_p = np.log2(p) if base2 else np.log(p)
    return (-p * log_p).sum(dtype=np.float32)


def jsd_from_counts(P, Q):
    P = np.asarray(P, np.float32); Q = np.asarray(Q, np.float32)
    P_ = P / P.sum(dtype=np.float32); Q_ = Q / Q.sum(dtype=np.float32)
    e1 = _entropy(P_); e2 = _entropy(Q_)
    e_sum = _entropy((P_ + Q_) / np.float32(2.0))
    return float(e_sum - ((e1 + e2) / np.float32(2.0)))


def compute_jsd(pcs_gen, pcs_ref, resolution=28):
    return jsd_from_counts(vote(pcs_gen, resolution)[0], vote(pcs_ref, resolution)[0])
