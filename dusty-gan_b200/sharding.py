"""Row sharding of the Chamfer matrix across the GPUs of one NVSwitch box.

The reference is single-GPU on this path; the matrix entries are independent, so
rank r computes rows r, r+G, r+2G, ... of the symmetric stacked matrix (a cyclic deal balances the
triangular work: row i has n-i entries). Two ways to combine the shards, one collective each:

* scores (``compute_cov_mmd_1nna``): the matrix kernel's epilogue reduces every entry into per-cloud packed
  (value, index) minima (``dusty_chamfer_matrix_fused``); the ranks all-gather those vectors -- 24 bytes per
  stacked cloud and rank, 48 KB at 1000 vs 1000 -- and no rank ever holds an n x n tensor;
* the matrix itself (``pairwise_matrices`` / ``_pairwise_distance``): compact (rows_owned, n) blocks, one
  all-gather, assembled into the full symmetric matrix by one kernel (``dusty_symmetric_from_shards``).

Entry values do not depend on G: each entry is produced by one CTA with a fixed reduction order.
"""
import torch
import torch.distributed as dist

from . import _lib


def world(group=None):
    """(rank, world size) of ``group`` (default process group when None); (0, 1) without torch.distributed."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def owned_rows(n, rank, world_size):
    """Rows of an n-row matrix owned by ``rank`` under the cyclic deal: (begin, end, stride)."""
    return rank, n, world_size


def rows_per_rank(n, world_size):
    """Row capacity of every rank's block (the all-gather needs equal shapes)."""
    return (n + world_size - 1) // world_size


def shard_of_row(i, world_size):
    """(rank, row inside that rank's compact block) holding global row ``i`` under the cyclic deal."""
    return i % world_size, i // world_size


def pack_key(value_f32_bits, index):
    """The epilogue's packed minimum: float bits of a (non-negative) matrix entry << 32 | stacked index, so
    that an unsigned 64-bit min orders by value, then by index (csrc/chamfer.cu, nn_kernel epilogue)."""
    return (int(value_f32_bits) << 32) | int(index)


def assemble_symmetric(blocks, n):
    """(G, cap, n) gathered compact blocks -> full symmetric (n, n) matrix (one kernel; entries below the
    diagonal are read from their mirror, bit-equal in the reference)."""
    G, cap, width = blocks.shape
    assert width == n and blocks.is_contiguous()
    out = torch.empty(n, n, device=blocks.device, dtype=torch.float32)
    if n == 0:
        return out
    lib = _lib.load()
    with torch.cuda.device(blocks.device):
        _lib.check(lib.dusty_symmetric_from_shards(_lib.ptr(blocks), G, cap, n, _lib.ptr(out), out.stride(0),
                                                   _lib.stream_of(blocks)), "dusty_symmetric_from_shards")
    return out


def _all_gather_flat(t, group):
    """all_gather_into_tensor of a contiguous tensor. NCCL gathers device buffers directly over NVLink; gloo
    (CPU tests, and ranks that share one GPU on a single-GPU box) has no device all-gather, so device tensors
    are staged through the host there."""
    _, G = world(group)
    out = torch.empty(G * t.numel(), device=t.device, dtype=t.dtype)
    if t.is_cuda and dist.get_backend(group) == "gloo":
        host = torch.empty(G * t.numel(), dtype=t.dtype)
        dist.all_gather_into_tensor(host, t.reshape(-1).cpu(), group=group)
        out.copy_(host)
    else:
        dist.all_gather_into_tensor(out, t.reshape(-1), group=group)
    return out


def all_gather_blocks(mine, group=None):
    """The single collective of the matrix path: (cap, n) per rank -> (G, cap, n) on every rank."""
    _, G = world(group)
    return _all_gather_flat(mine.contiguous(), group).view(G, mine.shape[0], mine.shape[1])


def all_gather_keys(keys, group=None):
    """The single collective of the score path: (3 n,) int64 packed minima per rank -> (G, 3 n)."""
    _, G = world(group)
    if G == 1:
        return keys.view(1, -1)
    return _all_gather_flat(keys.contiguous(), group).view(G, keys.numel())


def upload_sharded(host, device, group=None):
    """Host clouds (n, ...) -> the same tensor on ``device`` of every rank, each byte crossing PCIe once: rank r uploads
    rows [r cap, (r+1) cap) and the ranks all-gather the blocks over NVLink (the optional pre-step:
    what a data-parallel generation would hand to the evaluation). Every rank must pass the same host tensor. With one
    rank it is a plain (non-blocking, if ``host`` is pinned) copy."""
    _, G = world(group)
    if G == 1:
        return host.to(device, non_blocking=True)
    rank, _ = world(group)
    n = host.shape[0]
    cap = rows_per_rank(n, G)
    lo, hi = min(rank * cap, n), min((rank + 1) * cap, n)
    mine = torch.zeros((cap,) + tuple(host.shape[1:]), device=device, dtype=host.dtype)
    if hi > lo:
        mine[: hi - lo].copy_(host[lo:hi], non_blocking=True)
    return _all_gather_flat(mine, group).view((G * cap,) + tuple(host.shape[1:]))[:n]


def symmetric_chamfer_matrix(clouds, group=None):
    """Full symmetric (n,n) Chamfer matrix of ``clouds`` (n,P,3); every rank returns the same tensor."""
    from .utils.metrics.cov_mmd_1nna import chamfer_matrix
    rank, G = world(group)
    n = clouds.size(0)
    if G == 1:
        return chamfer_matrix(clouds)
    cap = rows_per_rank(n, G)
    mine = torch.zeros(cap, n, device=clouds.device, dtype=torch.float32)
    chamfer_matrix(clouds, None, rows=owned_rows(n, rank, G), compact_rows=True, out=mine)
    return assemble_symmetric(all_gather_blocks(mine, group), n)
