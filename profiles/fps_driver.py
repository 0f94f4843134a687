"""ncu driver: the FPS throughput kernel on 888 clouds of 32 768 points (bench shape); DUSTY_FPS_ALGO=l2 selects
the round-1 layout (six clouds per SM, distances in L2) for comparison."""
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
depth, conf = bench.backbone_like(888, 1, 12, dev)
pts = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0)["points"]
downsample_point_clouds(pts, 2048)
torch.cuda.synchronize()
downsample_point_clouds(pts, 2048)
torch.cuda.synchronize()
