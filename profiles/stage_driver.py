"""Tiny driver for ncu: a few launches of the head+projection kernel (batch 256) and the FPS kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dusty_gan_b200 import pipeline  # noqa: E402
from dusty_gan_b200.utils.sampling.fps import downsample_point_clouds  # noqa: E402

dev = torch.device("cuda:0")
lidar = bench.make_lidar(dev)
head = bench.make_head(1, dev)
depth, conf = bench.backbone_like(256, 1, 11, dev)
for _ in range(3):
    out = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0)
pts = out["points"][:148].contiguous()
for _ in range(2):
    downsample_point_clouds(pts, 2048)
torch.cuda.synchronize()
