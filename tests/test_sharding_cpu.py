"""Host-side logic of the row-sharded Chamfer matrix with world_size 2 and 3 over gloo on CPU.
The per-rank compute is stood in by the oracle (tests may use it); what is under test is the deal of
rows, the compact block layout, the single all-gather and the re-assembly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n, P, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dusty_gan_b200 import sharding
        from helpers import sampled_clouds
        from oracle import native
        clouds = sampled_clouds(n, P, 5)
        assert sharding.world() == (rank, world)
        begin, end, stride = sharding.owned_rows(n, rank, world)
        cap = sharding.rows_per_rank(n, world)
        mine = torch.zeros(cap, n)
        for r, i in enumerate(range(begin, end, stride)):      # what the kernel writes with COMPACT_ROWS
            row = native.pairwise_cd(clouds, None, rows=(i, i + 1))[i]
            mine[r, i:] = torch.from_numpy(row[i:])
        gathered = sharding.all_gather_blocks(mine)
        M = sharding.symmetrize_upper(sharding.assemble_upper(gathered, n, world))
        full = native.pairwise_cd(clouds, None)
        ok = bool(np.array_equal(M.numpy(), full))
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        if rank == 0:
            ret["ok"] = all(flags)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 9), (3, 10), (2, 1)])
def test_row_sharded_matrix_equals_unsharded(world, n):
    port = 29600 + world * 10 + n
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, n, 64, ret), nprocs=world, join=True)
        assert ret.get("ok") is True


def test_row_deal_balances_the_triangle():
    from dusty_gan_b200 import sharding
    n, G = 2000, 8
    work = []
    for r in range(G):
        b, e, s = sharding.owned_rows(n, r, G)
        work.append(sum(n - i for i in range(b, e, s)))
    assert max(work) / min(work) < 1.01
    assert sharding.rows_per_rank(n, G) == 250 and sharding.rows_per_rank(2001, G) == 251
    blocks = torch.arange(8 * 251 * 3, dtype=torch.float32).reshape(8, 251, 3)
    full = sharding.assemble_upper(blocks, 2001, 8)
    assert full.shape == (2001, 3) and torch.equal(full[9], blocks[1, 1])
