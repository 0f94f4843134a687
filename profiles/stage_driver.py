"""Tiny driver for ncu: the stage kernels at their bench sizes. run_ncu.sh skips the warm-up launches and captures:
head+projection (batch 256) without and with compaction, scan preprocess (256 scans, full width), FPS pruned
(148 clouds), FPS throughput variant (888 clouds), sorted/pruned Chamfer matrix (24 un-sampled clouds) with its
prep_sort, and the sorted batch front end on 8 un-sampled pairs."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dusty_gan_b200 import pipeline  # noqa: E402
from dusty_gan_b200.datasets import preprocess_scans  # noqa: E402
from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix  # noqa: E402
from dusty_gan_b200.utils.sampling.fps import downsample_point_clouds  # noqa: E402

dev = torch.device("cuda:0")
lidar = bench.make_lidar(dev)
head = bench.make_head(1, dev)
depth, conf = bench.backbone_like(888, 1, 11, dev)
d256, c256 = depth[:256].contiguous(), conf[:256].contiguous()
scans = torch.randn(256, 64, 2048, 4, device=dev) * 20


def head256():
    return pipeline.maskout_and_project(head, {"depth": d256, "confidence": c256}, lidar, tol=0.0)


def head256c():
    return pipeline.maskout_and_project(head, {"depth": d256, "confidence": c256}, lidar, tol=0.0, compact=True)


# ---- warm-up ----
head256(); head256(); head256c()
pts = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0)["points"]
downsample_point_clouds(pts, 2048)
preprocess_scans(scans, (64, 2048), 0.9, 120.0, -1)
torch.cuda.synchronize()
print("WARMUP_DONE", flush=True)
# ---- captured ----
head256()
head256c()
preprocess_scans(scans, (64, 2048), 0.9, 120.0, -1)
downsample_point_clouds(pts[:148].contiguous(), 2048)
downsample_point_clouds(pts, 2048)
chamfer_matrix(pts[:24].contiguous())
from dusty_gan_b200.utils.metrics.distance import chamfer_distance  # noqa: E402
chamfer_distance(pts[:8].contiguous(), pts[8:16].contiguous())
torch.cuda.synchronize()
