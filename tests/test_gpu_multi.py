"""Multi-rank runs of the sharded evaluation: every rank must return exactly the scores and matrices of the
unsharded call. With two or more GPUs the ranks talk NCCL, one GPU each (the production set-up); on a
single-GPU box the same code runs as 2 and 3 ranks that share cuda:0 over gloo (the all-gathers are then staged
through the host, sharding._all_gather_flat), so the row deal, the fused epilogue keys, their reduction and the
matrix assembly are exercised by the driver's one-GPU test run as well."""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["DUSTY_ROOT"]); sys.path.insert(0, os.path.join(os.environ["DUSTY_ROOT"], "tests"))
from helpers import sampled_clouds, lidar_like_clouds
rank = int(os.environ["RANK"]); backend = os.environ["DUSTY_BACKEND"]
dev = int(os.environ["LOCAL_RANK"]) if backend == "nccl" else 0
torch.cuda.set_device(dev)
from dusty_gan_b200.utils.metrics import cov_mmd_1nna as M
cases = [(sampled_clouds(37, 512, 1), sampled_clouds(41, 512, 2)),            # one stacked launch
         (sampled_clouds(9, 300, 3), sampled_clouds(14, 640, 4)),              # unequal point counts: three launches
         (lidar_like_clouds(5, 5000, 5, dropped=0.4), lidar_like_clouds(6, 5000, 6, dropped=0.4))]   # merged + sorted + pruned
cases = [(torch.from_numpy(g).cuda(), torch.from_numpy(r).cuda()) for g, r in cases]
single = [(M.compute_cov_mmd_1nna(g, r, 512, ("cd",), verbose=False), [m.clone() for m in M.pairwise_matrices(g, r)]) for g, r in cases]
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
else:
    dist.init_process_group("gloo")
ok = True
for (g, r), (scores1, mats1) in zip(cases, single):
    sharded = M.compute_cov_mmd_1nna(g, r, 512, ("cd",), verbose=False)
    M.FUSED_EPILOGUE = False
    unfused = M.compute_cov_mmd_1nna(g, r, 512, ("cd",), verbose=False)
    M.FUSED_EPILOGUE = True
    mats2 = M.pairwise_matrices(g, r)
    ok = ok and sharded == scores1 and unfused == scores1 and all(torch.equal(a, b) for a, b in zip(mats1, mats2))
from dusty_gan_b200 import sharding          # sharded upload: 1/G of the clouds per rank over PCIe, one all-gather
host = cases[2][0].cpu().pin_memory()
ok = ok and torch.equal(sharding.upload_sharded(host, torch.device("cuda", dev)), cases[2][0])
flags = [None] * dist.get_world_size(); dist.all_gather_object(flags, bool(ok))
if rank == 0: print("RESULT", json.dumps({"ok": all(flags), "world": dist.get_world_size(), "backend": backend}))
dist.destroy_process_group()
'''


def _run(tmp_path, world, backend, port):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DUSTY_ROOT=ROOT, DUSTY_BACKEND=backend)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    assert '"ok": true' in line and f'"world": {world}' in line, line


def test_sharded_scores_equal_unsharded_over_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the single-GPU variant below covers the same code over gloo)")
    _run(tmp_path, 2, "nccl", 29541)


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_scores_equal_unsharded_ranks_sharing_one_gpu(tmp_path, world):
    _run(tmp_path, world, "gloo", 29550 + world)
