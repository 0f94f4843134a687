"""Golden vectors from the reference's OWN CUDA kernels, run on a B200 -- TEST INFRASTRUCTURE.

    gpurun -- python oracle/gen_golden_gpu.py          (writes gpurun_out/golden_gpu/*.npz)

Uses the reference extensions compiled from /root/reference into oracle/_ref (oracle/build_ref.py):
``fps.furthest_point_sampling`` / ``fps.gather_points`` (reference has no CPU FPS) and
``cd.forward_cuda``. Inputs are regenerated from seeds by tests/helpers.py, so only the reference's
outputs (indices, distances) are stored; the files are then copied into tests/golden/.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import lidar_like_clouds, sampled_clouds  # noqa: E402
from oracle import refload  # noqa: E402

FPS_CASES = [  # (clouds, points, samples, seed, dropped, near)
    (2, 32768, 2048, 101, 0.3, 0.15), (2, 32768, 512, 102, 0.0, 0.0), (3, 5000, 512, 103, 0.3, 0.15),
    (2, 700, 64, 104, 0.2, 0.1), (2, 513, 100, 105, 0.2, 0.1), (2, 31, 8, 106, 0.2, 0.1), (1, 40000, 96, 107, 0.1, 0.05)]
CD_CASES = [(3, 2048, 2048, 201), (2, 1000, 777, 202), (2, 4096, 4096, 203)]  # (b, n, m, seed)


def main():
    out = os.path.join(ROOT, "gpurun_out", "golden_gpu")
    os.makedirs(out, exist_ok=True)
    fps = refload.load("dustyref_fps")
    cd = refload.load("dustyref_cd")
    assert fps is not None and cd is not None, "build oracle/_ref first (python oracle/build_ref.py)"
    res = {"gpu": np.array(torch.cuda.get_device_name(0)), "torch": np.array(torch.__version__)}
    for i, (b, n, m, seed, dropped, near) in enumerate(FPS_CASES):
        x = torch.from_numpy(lidar_like_clouds(b, n, seed, dropped=dropped, near=near)).cuda()
        idx = fps.furthest_point_sampling(x, m)
        g = fps.gather_points(x.transpose(1, 2).contiguous(), idx)
        torch.cuda.synchronize()
        res[f"fps{i}_case"] = np.array([b, n, m, seed, int(dropped * 1000), int(near * 1000)])
        res[f"fps{i}_idx"] = idx.cpu().numpy()
        res[f"fps{i}_gather_sum"] = g.double().sum((1, 2)).cpu().numpy()
    # degenerate clouds (ties, dropped seed, nothing eligible) spelled out explicitly
    deg = np.zeros((3, 1500, 3), np.float32)
    deg[0, [5, 77, 517, 1101, 300, 1401]] = [0.3, 0.1, 0.05]
    deg[0, 300] = [0.5, -0.2, 0.01]
    deg[1] = (np.random.default_rng(3).standard_normal((1500, 3)) * 0.005).astype(np.float32)
    half = np.random.default_rng(4).uniform(0.05, 0.5, (750, 3)).astype(np.float32)
    deg[2] = np.concatenate([half, half * np.array([1, -1, 1], np.float32)])
    res["fps_deg_input"] = deg
    res["fps_deg_idx"] = fps.furthest_point_sampling(torch.from_numpy(deg).cuda(), 200).cpu().numpy()
    for i, (b, n, m, seed) in enumerate(CD_CASES):
        a = torch.from_numpy(sampled_clouds(b, n, seed)).cuda()
        c = torch.from_numpy(lidar_like_clouds(b, m, seed + 1)).cuda()
        d1 = torch.zeros(b, n, device="cuda"); d2 = torch.zeros(b, m, device="cuda")
        i1 = torch.zeros(b, n, dtype=torch.int32, device="cuda"); i2 = torch.zeros(b, m, dtype=torch.int32, device="cuda")
        cd.forward_cuda(a, c, d1, d2, i1, i2)
        torch.cuda.synchronize()
        res[f"cd{i}_case"] = np.array([b, n, m, seed])
        res[f"cd{i}_dist1"] = d1.cpu().numpy(); res[f"cd{i}_dist2"] = d2.cpu().numpy()
        res[f"cd{i}_idx1"] = i1.cpu().numpy(); res[f"cd{i}_idx2"] = i2.cpu().numpy()
    np.savez_compressed(os.path.join(out, "gpu_reference_kernels.npz"), **res)
    print("wrote", os.path.join(out, "gpu_reference_kernels.npz"))


if __name__ == "__main__":
    main()
