"""Jensen-Shannon divergence between occupancy grids on libdustyb200 (mirror of reference
utils/metrics/jsd.py; SURVEY.md "next" row 8f-2, called at evaluate_synthesis.py:174-177)."""
import warnings

import numpy as np
import torch

from ... import _lib


def unit_cube_grid_point_cloud(resolution, clip_sphere, device):
    """The reference's grid, built with the same torch ops (jsd.py:10-20)."""
    spacing = 1.0 / float(resolution - 1)
    steps = torch.arange(resolution, device=device)
    grids = torch.meshgrid(steps, steps, steps, indexing="ij")
    grid = torch.stack(grids, dim=-1) * spacing - 0.5
    if clip_sphere:
        grid = grid.reshape(-1, 3)
        grid = grid[torch.norm(grid, dim=1) <= 0.5]
    return grid, spacing


_grid_cache = {}


def _grid_tables(resolution, in_sphere, device):
    key = (resolution, bool(in_sphere), str(device))
    if key not in _grid_cache:
        full, spacing = unit_cube_grid_point_cloud(resolution, False, device)
        full = full.reshape(-1, 3)
        keep = (torch.norm(full, dim=1) <= 0.5) if in_sphere else torch.ones(len(full), dtype=torch.bool, device=device)
        cell_to_idx = torch.full((resolution ** 3,), -1, dtype=torch.int32, device=device)
        cell_to_idx[keep] = torch.arange(int(keep.sum()), dtype=torch.int32, device=device)
        axis = (torch.arange(resolution, device=device) * spacing - 0.5).float().contiguous()
        _grid_cache[key] = (full[keep].float().contiguous(), cell_to_idx.contiguous(), axis, spacing)
    return _grid_cache[key]


def _vote(pcs, resolution, in_sphere):
    _lib.require_cuda(pcs, "pcs")
    if pcs.dim() != 3 or pcs.size(2) != 3:
        raise ValueError(f"expected (B,N,3), got {tuple(pcs.shape)}")
    x = pcs.contiguous()
    grid, cell_to_idx, axis, spacing = _grid_tables(resolution, in_sphere, x.device)
    ng = grid.shape[0]
    counters = torch.empty(ng, dtype=torch.int32, device=x.device)
    touching = torch.empty(ng, dtype=torch.int32, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.dusty_jsd_vote(_lib.ptr(x), x.shape[0], x.shape[1], resolution, ng, _lib.ptr(grid),
                                      _lib.ptr(cell_to_idx), _lib.ptr(axis), np.float32(spacing), _lib.ptr(counters),
                                      _lib.ptr(touching), _lib.stream_of(x)), "dusty_jsd_vote")
    return counters, touching


def _entropy(p, base=None, dim=-1, eps=1e-8):
    p += eps
    if base is None:
        log_p = torch.log(p)
    elif base == 2:
        log_p = torch.log2(p)
    elif base == 10:
        log_p = torch.log10(p)
    else:
        raise NotImplementedError
    return (-p * log_p).sum(dim=dim)


def entropy_of_occupancy_grid(pcs, resolution, in_sphere=False, batch_size=128, verbose=True):
    """(acc_entropy, grid_counters) like the reference (jsd.py:23-92); ``batch_size`` and ``verbose``
    only shaped its Python loops and have no effect here."""
    epsilon = 1e-3
    bound = 0.5 + epsilon
    if abs(pcs.max()) > bound or abs(pcs.min()) > bound:
        warnings.warn("Point-clouds are not in unit cube.")
    if in_sphere and torch.norm(pcs, p=2, dim=2).max() > bound:
        warnings.warn("Point-clouds are not in unit sphere.")
    counters, touching = _vote(pcs, resolution, in_sphere)
    grid_counters = counters.float()
    bern = touching.float()
    p = bern[bern > 0] / float(len(pcs))
    acc_entropy = _entropy(torch.cat([p, 1 - p])) / len(grid_counters)
    return acc_entropy, grid_counters


def _jensen_shannon_divergence(P, Q):
    assert (P >= 0).all() and (Q >= 0).all(), "Negative values."
    assert len(P) == len(Q), "Non equal size."
    P_ = P / P.sum()
    Q_ = Q / Q.sum()
    e1 = _entropy(P_, base=2)
    e2 = _entropy(Q_, base=2)
    e_sum = _entropy((P_ + Q_) / 2.0, base=2)
    return e_sum - ((e1 + e2) / 2.0)


@torch.no_grad()
def compute_jsd(pcs_gen, pcs_ref, resolution=28, batchsize=128, verbose=True):
    """JSD of the two sets' occupancy histograms; two voting launches + one reduction launch, one
    4-byte read-back (the reference: three nested Python loops over 128-sized blocks)."""
    gen_counts, _ = _vote(pcs_gen, resolution, True)
    ref_counts, _ = _vote(pcs_ref, resolution, True)
    out = torch.empty(1, device=gen_counts.device, dtype=torch.float32)
    lib = _lib.load()
    with torch.cuda.device(out.device):
        _lib.check(lib.dusty_jsd_from_counts(_lib.ptr(gen_counts), _lib.ptr(ref_counts), gen_counts.numel(), _lib.ptr(out),
                                             _lib.stream_of(out)), "dusty_jsd_from_counts")
    return out.item()
