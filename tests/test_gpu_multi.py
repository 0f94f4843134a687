"""Two-rank NCCL run of the sharded evaluation (skipped on a single-GPU box): every rank must return
exactly the scores and matrices of the unsharded call."""
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
from helpers import sampled_clouds
rank = int(os.environ["RANK"]); torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
from dusty_gan_b200.utils.metrics.cov_mmd_1nna import compute_cov_mmd_1nna, pairwise_matrices
gen = torch.from_numpy(sampled_clouds(37, 512, 1)).cuda(); ref = torch.from_numpy(sampled_clouds(41, 512, 2)).cuda()
single = compute_cov_mmd_1nna(gen, ref, 512, ("cd",), verbose=False)
mats1 = [m.clone() for m in pairwise_matrices(gen, ref)]
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
sharded = compute_cov_mmd_1nna(gen, ref, 512, ("cd",), verbose=False)
mats2 = pairwise_matrices(gen, ref)
ok = sharded == single and all(torch.equal(a, b) for a, b in zip(mats1, mats2))
flags = [None] * dist.get_world_size(); dist.all_gather_object(flags, bool(ok))
if rank == 0: print("RESULT", json.dumps({"ok": all(flags), "world": dist.get_world_size()}))
dist.destroy_process_group()
'''


def test_sharded_scores_equal_unsharded_over_nccl(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, DUSTY_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")][-1]
    assert '"ok": true' in line and '"world": 2' in line
