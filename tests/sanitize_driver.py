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
    for mode in ("segment", "image"):          # both ordered-compaction kernels
        os.environ["DUSTY_HEAD_COMPACT"] = mode
        out = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0, compact=True)
    os.environ.pop("DUSTY_HEAD_COMPACT")
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
# round 2: batch front end on sorted clouds (perm / inv maps, unsort), its backward (own + scatter kernels), an
# adversarial cluster that drives every row through the guard's whole-tile re-evaluation, the fused epilogue with
# unequal point counts (three launches into one key set), the k-NN vote, the symmetric assembly of row shards
ua = u.clone().requires_grad_(True)
e1, e2 = chamfer_distance(ua, torch.flip(u, dims=[0]))
(e1.sum() + e2.sum()).backward()
cube = (torch.rand(2, 2300, 3, device="cuda") * 0.05 + torch.tensor([0.7, 0.5, 0.1], device="cuda"))
print(chamfer_matrix(cube, merge_origin=False))
chamfer_distance(cube, torch.flip(cube, dims=[0]))
print(compute_cov_mmd_1nna(gen, ref[:, :200].contiguous(), 512, ("cd",), verbose=False))
from dusty_gan_b200.utils.metrics import cov_mmd_1nna as M  # noqa: E402
mats = M.pairwise_matrices(gen, ref)
print(M._compute_nna(*mats, 3))
from dusty_gan_b200 import sharding  # noqa: E402
blocks = torch.stack([chamfer_matrix(gen, None, rows=(r, 5, 2), compact_rows=True, out=torch.zeros(3, 5, device="cuda")) for r in range(2)])
print(sharding.assemble_symmetric(blocks, 5))
many = torch.from_numpy(lidar_like_clouds(150, 1500, 8)).cuda()          # > 148 clouds: the on-chip-distance FPS variant
downsample_point_clouds(many, 24)
print(chamfer_matrix(u, merge_origin=True))
print(chamfer_matrix(u, u[:2, :600].contiguous(), merge_origin=True))
# round 2, second session: k-d ordered clouds (prep_sort_kernel<true>), the resident-pair kernel (dynamic task counter, TMA loads,
# guard from shared memory on the adversarial cluster, ragged point counts with a merged origin) and the two-level walk kernel
# in both front ends (matrix above 2048 points, batch on un-sampled clouds)
print(chamfer_matrix(cube[:, :2000].contiguous()))
print(chamfer_matrix(u[:, :1999].contiguous(), u[:2, :777].contiguous()))
print(chamfer_matrix(cube, u[:, :2100].contiguous()))
torch.cuda.synchronize()
print("sanitize driver done")
