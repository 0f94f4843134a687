"""Row sharding of the Chamfer matrix across the GPUs of one NVSwitch box.

The reference is single-GPU on this path; the matrix entries are independent, so
rank r computes rows r, r+G, r+2G, ... of the symmetric stacked matrix (a cyclic deal balances the
triangular work: row i has n-i entries) into a compact (rows_owned, n) block, and ONE all-gather of
those blocks gives every rank the whole upper triangle. No other collective is on the path. Entry
values do not depend on G: each entry is produced by one CTA with a fixed reduction order.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def owned_rows(n, rank, world_size):
    """Rows of an n-row matrix owned by ``rank`` under the cyclic deal: (begin, end, stride)."""
    return rank, n, world_size


def rows_per_rank(n, world_size):
    """Row capacity of every rank's block (the all-gather needs equal shapes)."""
    return (n + world_size - 1) // world_size


def assemble_upper(blocks, n, world_size):
    """(G, cap, n) gathered compact blocks -> (n, n) matrix whose upper triangle (incl. diagonal) is valid."""
    cap = blocks.shape[1]
    # block g row r is global row g + r*G: interleave
    full = blocks.permute(1, 0, 2).reshape(cap * world_size, blocks.shape[2])
    return full[:n]


def symmetrize_upper(U):
    """Mirror the strict upper triangle into the lower one (entries are bit-equal in the reference,
    SURVEY.md S8)."""
    upper = torch.triu(U)
    return upper + torch.triu(U, diagonal=1).t()


def all_gather_blocks(mine, group=None):
    """The single collective on the path: (cap, n) per rank -> (G, cap, n) on every rank."""
    _, G = world()
    flat = torch.empty(G * mine.shape[0], mine.shape[1], device=mine.device, dtype=mine.dtype)
    dist.all_gather_into_tensor(flat, mine.contiguous(), group=group)
    return flat.view(G, mine.shape[0], mine.shape[1])


def symmetric_chamfer_matrix(clouds, group=None):
    """Full symmetric (n,n) Chamfer matrix of ``clouds`` (n,P,3); every rank returns the same tensor."""
    from .utils.metrics.cov_mmd_1nna import chamfer_matrix
    rank, G = world()
    n = clouds.size(0)
    if G == 1:
        return chamfer_matrix(clouds)
    cap = rows_per_rank(n, G)
    mine = torch.zeros(cap, n, device=clouds.device, dtype=torch.float32)
    chamfer_matrix(clouds, None, rows=owned_rows(n, rank, G), compact_rows=True, out=mine)
    gathered = all_gather_blocks(mine, group)
    return symmetrize_upper(assemble_upper(gathered, n, G))
