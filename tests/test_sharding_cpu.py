"""Host-side logic of the row-sharded Chamfer evaluation with world_size 2 and 3 over gloo on CPU.
The per-rank compute is stood in by the oracle (tests may use it); what is under test is the deal of
rows, the compact block layout, the single all-gather of each path (row blocks for the matrix, packed
per-cloud minima for the scores) and the re-assembly / reduction rules the CUDA kernels implement."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _keys_from_rows(M, rows, nr):
    """What nn_kernel's epilogue leaves for the entries (i, j >= i) of the given rows of the stacked matrix."""
    n = M.shape[0]
    keys = np.full(3 * n, np.iinfo(np.uint64).max, np.uint64)
    bits = M.astype(np.float32).view(np.uint32).astype(np.uint64)
    for i in rows:
        for j in range(i + 1, n):
            vb = bits[i, j] << np.uint64(32)
            keys[j] = min(keys[j], vb | np.uint64(i))
            keys[i] = min(keys[i], vb | np.uint64(j))
            if i < nr <= j:
                keys[n + j] = min(keys[n + j], vb | np.uint64(i))
                keys[2 * n + i] = min(keys[2 * n + i], vb | np.uint64(j))
    return keys


def _scores_from_keys(gathered, nr, ng):
    """keys_kernel + final_kernel of csrc/metrics.cu in numpy."""
    n = nr + ng
    k = gathered.reshape(-1, 3, n).min(axis=0)
    nn_is_ref = (k[0] & np.uint64(0xffffffff)).astype(np.int64) < nr
    val = lambda a: (a >> np.uint64(32)).astype(np.uint32).view(np.float32)
    label = np.arange(n) < nr
    return {"mmd": float(np.float32(val(k[2][:nr]).astype(np.float64).mean())),
            "mmd-sample": float(np.float32(val(k[1][nr:]).astype(np.float64).mean())),
            "cov": len(np.unique(k[1][nr:] & np.uint64(0xffffffff))) / nr,
            "tp": float((nn_is_ref & label).sum()), "fp": float((nn_is_ref & ~label).sum()),
            "fn": float((~nn_is_ref & label).sum()), "tn": float((~nn_is_ref & ~label).sum())}


def _worker(rank, world, port, n, P, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dusty_gan_b200 import sharding
        from helpers import sampled_clouds
        from oracle import metrics as om, native
        clouds = sampled_clouds(n, P, 5)
        assert sharding.world() == (rank, world)
        begin, end, stride = sharding.owned_rows(n, rank, world)
        cap = sharding.rows_per_rank(n, world)
        full = native.pairwise_cd(clouds, None)
        # ---- matrix path: compact row blocks, one all-gather, every entry read from the shard that owns it ----
        mine = torch.zeros(cap, n)
        for r, i in enumerate(range(begin, end, stride)):      # what the kernel writes with COMPACT_ROWS
            row = native.pairwise_cd(clouds, None, rows=(i, i + 1))[i]
            mine[r, i:] = torch.from_numpy(row[i:])
        gathered = sharding.all_gather_blocks(mine).numpy()
        M = np.empty((n, n), np.float32)
        for i in range(n):
            for j in range(n):                                  # symmetric_kernel of csrc/metrics.cu
                g, r = sharding.shard_of_row(min(i, j), world)
                M[i, j] = gathered[g, r, max(i, j)]
        ok = bool(np.array_equal(M, full))
        # ---- score path: packed per-cloud minima, one all-gather of 24 n bytes per rank ----
        nr = n // 2
        if nr >= 1 and n - nr >= 1:
            keys = _keys_from_rows(full, range(begin, end, stride), nr)
            g = sharding.all_gather_keys(torch.from_numpy(keys.view(np.int64))).numpy().view(np.uint64)
            assert g.shape == (world, 3 * n)
            got = _scores_from_keys(g, nr, n - nr)
            want = om.scores_from_matrices(full[:nr, :nr], full[:nr, nr:], full[nr:, nr:])
            for key in ("mmd", "mmd-sample", "cov"):
                ok = ok and got[key] == want[key + "-cd"]
            for key in ("tp", "fp", "fn", "tn"):
                ok = ok and got[key] == want["1-nn-" + key + "-cd"]
        # ---- sharded upload: every rank ends up with the full tensor, rows in order, for n not divisible by the world ----
        host = torch.from_numpy(clouds)
        ok = ok and torch.equal(sharding.upload_sharded(host, torch.device("cpu")), host)
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        if rank == 0:
            ret["ok"] = all(flags)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 9), (3, 10), (2, 1)])
def test_row_sharded_matrix_and_scores_equal_unsharded(world, n):
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
    assert sharding.shard_of_row(9, 8) == (1, 1) and sharding.shard_of_row(2000, 8) == (0, 250)
    assert sharding.world() == (0, 1)


def test_packed_keys_order_by_value_then_index():
    from dusty_gan_b200 import sharding
    bits = lambda v: int(np.float32(v).view(np.uint32))
    assert sharding.pack_key(bits(0.25), 7) < sharding.pack_key(bits(0.5), 0)
    assert sharding.pack_key(bits(0.25), 3) < sharding.pack_key(bits(0.25), 7)
    assert sharding.pack_key(bits(0.0), 5) < sharding.pack_key(bits(1e-30), 0)
