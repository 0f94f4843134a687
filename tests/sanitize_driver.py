"""Small end-to-end pass over every kernel, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tests/sanitize_driver.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import head_inputs, lidar_like_clouds  # noqa: E402

from dusty_gan_b200 import pipeline  # noqa: E402
from dusty_gan_b200.models.dusty import DUSty1, DUSty2  # noqa: E402
from dusty_gan_b200.utils.lidar import LiDAR, synthetic_hdl64e_angles  # noqa: E402
from dusty_gan_b200.utils.metrics.cov_mmd_1nna import compute_cov_mmd_1nna  # noqa: E402
from dusty_gan_b200.utils.metrics.distance import chamfer_distance  # noqa: E402
from dusty_gan_b200.utils.metrics.jsd import compute_jsd  # noqa: E402
from dusty_gan_b200.utils.sampling.fps import downsample_point_clouds  # noqa: E402

H, W = 16, 128
lidar = LiDAR(H, W, 0.9, 120.0, angle=synthetic_hdl64e_angles()).cuda()
for kind in (1, 2):
    head = (DUSty1 if kind == 1 else DUSty2)(torch.nn.Identity(), tau=1.0).cuda().eval()
    depth, conf, u1, u2 = head_inputs(3, kind, H, W, 7, "cuda")
    gate = head.gumbel if kind == 1 else head.gumbel_pixel
    gate.fixed_noise = gate._logistic_from_uniform(u1, u2)
    out = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0, compact=True)
    pts = downsample_point_clouds(out["points"], 64)
x = torch.from_numpy(lidar_like_clouds(3, 3000, 5)).cuda()
for algo in ("single", "multi", "flat"):
    os.environ["DUSTY_FPS_ALGO"] = algo          # read once per process: only the first takes effect
    downsample_point_clouds(x, 40)
a = torch.from_numpy(lidar_like_clouds(2, 700, 1)).cuda().requires_grad_(True)
b = torch.from_numpy(lidar_like_clouds(2, 2500, 2)).cuda()
d1, d2 = chamfer_distance(a, b)
(d1.sum() + d2.sum()).backward()
gen = downsample_point_clouds(torch.from_numpy(lidar_like_clouds(5, 2000, 3)).cuda(), 300)
ref = downsample_point_clouds(torch.from_numpy(lidar_like_clouds(4, 2000, 4)).cuda(), 300)
print(compute_cov_mmd_1nna(gen, ref, 512, ("cd",), verbose=False))
print(compute_jsd(gen / 2, ref / 2))
# real-data side and the merged-origin Chamfer matrix
from dusty_gan_b200.datasets import preprocess_scans  # noqa: E402
from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix  # noqa: E402
scans = (torch.randn(2, 8, 96, 4, device="cuda") * 20) * (torch.rand(2, 8, 96, 1, device="cuda") > 0.3)
preprocess_scans(scans, (8, 24), want=("xyz", "depth", "mask", "inv", "points"))
preprocess_scans(scans[..., :3].contiguous(), (8, 96))
u = torch.from_numpy(lidar_like_clouds(3, 4200, 6, dropped=0.5)).cuda()       # > 1024 kept points: sorted + pruned path
u[1] = 0
print(chamfer_matrix(u, merge_origin=True))
print(chamfer_matrix(u, u[:2, :600].contiguous(), merge_origin=True))
torch.cuda.synchronize()
print("sanitize driver done")
