import os, sys, statistics, torch
sys.path.insert(0, os.getcwd())
import bench
from dusty_gan_b200 import pipeline
from dusty_gan_b200.utils.sampling.fps import downsample_point_clouds
dev = torch.device("cuda:0")
lidar = bench.make_lidar(dev); head = bench.make_head(1, dev)
nclouds = int(os.environ.get("CLOUDS", "148"))
depth, conf = bench.backbone_like(nclouds, 1, 12, dev)
pts = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0)["points"]
for m in (2048, 512, 64):
    ms = bench.time_events(lambda: downsample_point_clouds(pts, m), 3, 1)
    t = statistics.median(ms)
    print(os.environ.get("DUSTY_FPS_ALGO", "auto"), "clouds", nclouds, "samples", m, t, "ms", round(nclouds / t * 1e3), "clouds/s")
