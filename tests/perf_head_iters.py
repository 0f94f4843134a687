"""A/B of the head kernel's groups-per-thread (DUSTY_HEAD_ITERS) on configs[1]; not a test."""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from dusty_gan_b200 import pipeline  # noqa: E402

dev = torch.device("cuda:0")
lidar = bench.make_lidar(dev)
for kind in (1, 2):
    head = bench.make_head(kind, dev)
    for batch in (256, 250, 32):
        depth, conf = bench.backbone_like(batch, kind, 11, dev)
        bufs = {"mask": torch.empty_like(conf), "depth": torch.empty_like(depth),
                "points": torch.empty(batch, bench.H * bench.W, 3, device=dev)}
        for iters in ("4", "2", "1", ""):
            if iters:
                os.environ["DUSTY_HEAD_ITERS"] = iters
            else:
                os.environ.pop("DUSTY_HEAD_ITERS", None)

            def run():
                pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0, buffers=bufs)
            run(); torch.cuda.synchronize()
            side = torch.cuda.Stream()
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(20):
                        run()
            ms = statistics.median(bench.time_events(graph.replay, 7, 3)) / 20
            gb = batch * bench.H * bench.W * (28 if kind == 1 else 36) / 1e9
            print(f"dusty{kind} batch {batch:4d} iters {iters or 'auto':>4}: {ms * 1e3:7.2f} us  {gb / (ms * 1e-3):7.1f} GB/s")
