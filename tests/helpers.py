"""Seeded synthetic inputs shared by the tests (and mirrored by bench.py)."""
import numpy as np
import torch


def lidar_like_clouds(n_clouds, n_points, seed, dropped=0.3, near=0.15):
    """(n,P,3) f32 clouds shaped like projected range images: points on rings around the sensor at
    normalised range in (0,1], a fraction exactly at the origin (dropped pixels) and a fraction
    inside the FPS exclusion radius (|p|^2 <= 1e-3)."""
    rng = np.random.default_rng(seed)
    az = rng.uniform(-np.pi, np.pi, (n_clouds, n_points))
    el = np.deg2rad(rng.uniform(-24.8, 2.0, (n_clouds, n_points)))
    r = np.exp(rng.uniform(np.log(0.035), np.log(0.7), (n_clouds, n_points)))
    r = np.where(rng.uniform(size=r.shape) < near, rng.uniform(0.008, 0.03, r.shape), r)
    r = np.where(rng.uniform(size=r.shape) < dropped, 0.0, r)
    xyz = np.stack([r * np.cos(el) * np.cos(az), r * np.cos(el) * np.sin(az), r * np.sin(el)], -1)
    return xyz.astype(np.float32)


def sampled_clouds(n_clouds, n_points, seed):
    """FPS-like clouds: no dropped points, ranges spread over the scene."""
    return lidar_like_clouds(n_clouds, n_points, seed, dropped=0.0, near=0.0)


def head_inputs(B, C, H, W, seed, device="cpu", calibrated=True):
    """Backbone-like outputs: depth in tanh space, confidence logits, U1/U2 for the fixed noise."""
    g = torch.Generator().manual_seed(seed)
    depth = torch.tanh(torch.randn(B, 1, H, W, generator=g) * (1.2 if calibrated else 0.115) - (0.3 if calibrated else 0.04))
    conf = torch.randn(B, C, H, W, generator=g) * 2.0
    u1 = torch.rand(1, 1, H, W, generator=g)
    u2 = torch.rand(1, 1, H, W, generator=g)
    return depth.to(device), conf.to(device), u1.to(device), u2.to(device)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
