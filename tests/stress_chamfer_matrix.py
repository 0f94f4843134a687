"""Randomised parity sweep of the matrix front end's three search kernels against the CUDA-rounding oracle (not collected
by pytest: run by hand on a B200, `python tests/stress_chamfer_matrix.py [rounds] [seed]`). Shapes straddle every dispatch
boundary (256 / 2048 / 32768 points, chunk and row-group multiples), clouds mix LiDAR-like points, dropped (0,0,0) points,
duplicates, tight far-away clusters, lines and planes, tiny clouds."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import lidar_like_clouds  # noqa: E402
from oracle import native  # noqa: E402
from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix  # noqa: E402

ULP = 2.0 ** -23


def make(rng, n, p):
    kind = rng.integers(0, 7)
    if kind == 0:
        a = lidar_like_clouds(n, p, int(rng.integers(1 << 30)), dropped=float(rng.uniform(0, 0.9)))
    elif kind == 1:
        a = lidar_like_clouds(n, p, int(rng.integers(1 << 30)), dropped=0.0, near=0.0)
    elif kind == 2:                                   # tight cluster far from the origin (search rounding window)
        a = (rng.uniform(0, 0.05, (n, p, 3)) + rng.uniform(-1, 1, (n, 1, 3))).astype(np.float32)
    elif kind == 3:                                   # many exact duplicates
        base = rng.uniform(-1, 1, (n, max(1, p // 37), 3)).astype(np.float32)
        a = base[:, rng.integers(0, base.shape[1], p)]
    elif kind == 4:                                   # a line / a plane (degenerate boxes)
        a = np.zeros((n, p, 3), np.float32)
        a[..., : int(rng.integers(1, 3))] = rng.uniform(-1, 1, (n, p, int(a[..., :1].shape[-1])))[..., :1]
    elif kind == 5:                                   # integer lattice: exact ties everywhere
        a = rng.integers(-3, 4, (n, p, 3)).astype(np.float32) * 0.125
    else:                                             # mostly zeros
        a = lidar_like_clouds(n, p, int(rng.integers(1 << 30)), dropped=0.97)
    if rng.uniform() < 0.2:
        a[rng.integers(0, n)] = 0.0
    return np.ascontiguousarray(a, np.float32)


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    sizes = [1, 2, 31, 32, 33, 63, 64, 65, 255, 256, 257, 300, 511, 512, 513, 1000, 1023, 1024, 1025, 2016, 2047, 2048, 2049,
             2111, 3000, 4095, 4096, 4097, 8000]
    big = [16384, 20000, 32767, 32768]
    bad = 0
    for it in range(rounds):
        pa = int(rng.choice(big if it % 15 == 14 else sizes)); pb = int(rng.choice(big if it % 15 == 14 else sizes))
        na, nb = (2, 2) if max(pa, pb) > 8000 else (int(rng.integers(1, 5)), int(rng.integers(1, 5)))
        a, b = make(rng, na, pa), make(rng, nb, pb)
        O = native.pairwise_cd(a, b, rounding="cuda")
        for merge in (None, True, False):
            M = chamfer_matrix(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), merge_origin=merge).cpu().numpy()
            ok = np.all(np.abs(M.astype(np.float64) - O) <= ULP * np.abs(O))
            if not ok:
                bad += 1
                print("MISMATCH", it, pa, pb, na, nb, merge, np.abs(M - O).max(), flush=True)
        if pa == pb:
            S = chamfer_matrix(torch.from_numpy(a).cuda()).cpu().numpy()
            Os = native.pairwise_cd(a, None, rounding="cuda")
            if not (np.array_equal(S, S.T) and np.all(np.abs(S.astype(np.float64) - Os) <= ULP * np.abs(Os))):
                bad += 1
                print("MISMATCH symmetric", it, pa, na, flush=True)
        print("round", it, pa, pb, na, nb, "ok" if not bad else "FAILED so far", flush=True)
    print("stress done: %d mismatches" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
