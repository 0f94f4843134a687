"""A/B of rows-per-thread (DUSTY_CHAMFER_R) for the batch front end on small batches; not a test."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from helpers import sampled_clouds
from dusty_gan_b200.utils.metrics.distance import chamfer_distance
for B, n in ((1, 2048), (8, 2048), (32, 2048), (128, 2048), (512, 2048), (32, 512), (8, 32768)):
    a = torch.from_numpy(sampled_clouds(B, n, 1)).cuda(); b = torch.from_numpy(sampled_clouds(B, n, 2)).cuda()
    line = f"B={B:4d} n={n:6d}:"
    for r in ("8", "4", "2", "1", ""):
        if r: os.environ["DUSTY_CHAMFER_R"] = r
        else: os.environ.pop("DUSTY_CHAMFER_R", None)
        ms = statistics.median(bench.time_events(lambda: chamfer_distance(a, b), 5, 2))
        line += f"  R={r or 'auto'}: {ms * 1e3:8.1f} us"
    print(line)
