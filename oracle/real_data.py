"""Restatement of the real-data side of the path -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

    KITTIOdometry.preprocess / transform      reference datasets/kitti.py:54-78        (numpy, per scan)
    preprocess_reals                          reference evaluate_synthesis.py:49-57    (torch, per batch)
    cache + time-series subsampling           reference evaluate_synthesis.py:69-110

The numpy half is written with explicit float32 operations in the order numpy executes them
(``np.linalg.norm(xyz, ord=2, axis=2)`` = ``sqrt(add.reduce(x*x, axis=2))`` = sqrt((x^2+y^2)+z^2));
the torch half issues the reference's ATen ops in the reference's order on whatever device the inputs
live on (on the GPU it is the same-device reference, SURVEY.md trap T2). Pinned against the
reference's own classes by tests/golden/real_data.npz (oracle/gen_golden.py).
"""
import numpy as np
import torch


def dataset_preprocess(scan, min_depth=0.9, max_depth=120.0):
    """reference datasets/kitti.py:54-67 on one (Hs,Ws,>=3) float32 scan -> xyz (Hs,Ws,3), depth, mask."""
    f32 = np.float32
    xyz = np.array(scan[..., :3], dtype=np.float32, copy=True)
    x, y, z = xyz[..., 0], xyz[..., 1], xyz[..., 2]
    depth = np.sqrt((x * x + y * y) + z * z, dtype=np.float32)
    mask = (depth > f32(0.0)) & (depth > f32(min_depth)) & (depth < f32(max_depth))
    depth = (depth - f32(min_depth)) / f32(max_depth - min_depth)
    xyz = xyz / f32(max_depth)
    depth[~mask] = 0
    xyz[~mask] = 0
    return xyz, depth, mask


def nearest_index(out_size, in_size):
    """torch 'nearest' source indices (ATen nearest_neighbor_compute_source_index): f32 scale."""
    scale = np.float32(in_size) / np.float32(out_size)
    src = np.floor(np.arange(out_size, dtype=np.float32) * scale).astype(np.int64)
    return np.minimum(src, in_size - 1)


def dataset_item(scan, shape, min_depth=0.9, max_depth=120.0):
    """preprocess + transform (to_tensor, nearest resize; flip=False) -> dict of torch CPU tensors
    xyz (3,H,W) f32, depth (1,H,W) f32, mask (1,H,W) bool -- what __getitem__ returns."""
    xyz, depth, mask = dataset_preprocess(scan, min_depth, max_depth)
    ri = nearest_index(shape[0], scan.shape[0])
    ci = nearest_index(shape[1], scan.shape[1])
    pick = lambda a: a[ri][:, ci]
    return {"xyz": torch.from_numpy(np.ascontiguousarray(pick(xyz).transpose(2, 0, 1))),
            "depth": torch.from_numpy(np.ascontiguousarray(pick(depth)))[None],
            "mask": torch.from_numpy(np.ascontiguousarray(pick(mask)))[None]}


def invert_depth(norm_depth, min_depth, max_depth):
    """reference utils/lidar.py:31-36 with normalize/denormalize_minmax (:23-29) inlined."""
    depth = norm_depth * (max_depth - min_depth) + min_depth
    disp = 1 / depth
    return (disp - 1 / max_depth) / (1 / min_depth - 1 / max_depth)


def preprocess_reals(raw_batch, min_depth=0.9, max_depth=120.0, drop_const=-1, device=None):
    """reference evaluate_synthesis.py:49-57 -> (inv (B,1,H,W), mask (B,1,H,W) f32, points (B,N,3) contiguous)."""
    xyz = raw_batch["xyz"].to(device)
    points = xyz.flatten(2).transpose(1, 2)
    depth = raw_batch["depth"].to(device)
    mask = raw_batch["mask"].to(device).float()
    inv = invert_depth(depth, min_depth, max_depth)
    inv = inv * 2.0 - 1.0                                   # sigmoid_to_tanh, utils/__init__.py:70-73
    inv = mask * inv + (1 - mask) * drop_const
    return inv, mask, points.contiguous()


def subsample_time_series(t, num_test):
    """reference evaluate_synthesis.py:102-110."""
    if num_test == -1:
        return t
    skip = len(t) // num_test
    limit = skip * num_test + 1
    return t[skip:limit:skip]


def synthetic_scans(n, seed=0, hs=64, ws=2048, channels=4, dropout=0.25):
    """Random organised scans in the (64,2048,4) layout process_kitti.py writes: smooth ranges of
    2..110 m on an HDL-64E-like grid plus returns below min_depth, beyond max_depth and empty pixels
    (all zeros), so that every branch of the mask is exercised."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(2.0, -24.8, hs))[:, None]
    azim = np.linspace(np.pi, -np.pi, ws, endpoint=False)[None, :]
    out = np.zeros((n, hs, ws, channels), np.float32)
    for i in range(n):
        base = 20 + 15 * np.sin(3 * azim + rng.uniform(0, 6)) + 10 * np.cos(5 * elev * 7 + rng.uniform(0, 6))
        r = np.abs(base + rng.standard_normal((hs, ws)) * 8.0) + 0.2
        r[rng.random((hs, ws)) < 0.02] *= 8.0                # some beyond max_depth
        r[rng.random((hs, ws)) < 0.02] *= 0.02               # some below min_depth
        out[i, ..., 0] = r * np.cos(elev) * np.cos(azim)
        out[i, ..., 1] = r * np.cos(elev) * np.sin(azim)
        out[i, ..., 2] = r * np.sin(elev)
        if channels > 3:
            out[i, ..., 3] = rng.random((hs, ws))
        out[i][rng.random((hs, ws)) < dropout] = 0           # empty pixels
    return out
