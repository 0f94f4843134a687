"""A/B of the ordered-compaction paths of the head kernel (env DUSTY_HEAD_COMPACT=precount|image|segment); not a test."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device("cuda:0")
hbm, _, _ = bench.peaks()
for kind in (1, 2):
    print(os.environ.get("DUSTY_HEAD_COMPACT", "default"), "dusty", kind, json.dumps(bench.r4(bench.bench_head(dev, hbm, kind, True))))
