"""Chamfer matrix throughput by points per cloud (the training-time validation uses 512); not a test."""
import os, statistics, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench
from helpers import sampled_clouds
from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix
for P, N in ((128, 2000), (256, 2000), (512, 2000), (1024, 1000), (2048, 500), (4096, 300)):
    a = torch.from_numpy(sampled_clouds(N, P, 1)).cuda()
    ms = statistics.median(bench.time_events(lambda: chamfer_matrix(a), 3, 1))
    entries = N * (N + 1) / 2
    tf = entries * 12.0 * P * P / (ms * 1e-3) / 1e12
    print(f"P={P:5d} N={N:5d}: {ms:9.2f} ms  {entries / (ms * 1e-3) / 1e6:8.2f} M executed entries/s  {tf:6.2f} TFLOP/s executed ({tf / 74.45 * 100:5.1f} %)")
