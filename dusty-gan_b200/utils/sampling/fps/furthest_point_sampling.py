"""Farthest-point sampling on libdustyb200 (mirror of reference
utils/sampling/fps/furthest_point_sampling.py:19-93)."""
from typing import Any

import torch

from .... import _lib


def _fps(xyz, npoint, want_points):
    _lib.require_cuda(xyz, "xyz")
    if xyz.dim() != 3 or xyz.size(2) != 3:
        raise ValueError("expected (B,N,3), but got {}".format(tuple(xyz.shape)))
    x = xyz.contiguous()
    B, N, _ = x.shape
    idx = torch.zeros(B, npoint, device=x.device, dtype=torch.int32)
    pts = torch.empty(B, npoint, 3, device=x.device, dtype=torch.float32) if want_points else None
    if B == 0 or npoint == 0:
        return idx, pts
    lib = _lib.load()
    nbytes = lib.dusty_fps_workspace_bytes(B, N, npoint)
    ws = _lib.workspace(nbytes, x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.dusty_fps(_lib.ptr(x), B, N, npoint, _lib.ptr(idx), _lib.ptr(pts), _lib.ptr(ws), nbytes,
                                 _lib.stream_of(x)), "dusty_fps")
    return idx, pts


class FurthestPointSampling(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        # type: (Any, torch.Tensor, int) -> torch.Tensor
        """xyz (B,N,3) -> (B,npoint) int32 indices of the iteratively farthest points."""
        out, _ = _fps(xyz, npoint, False)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return ()


furthest_point_sampling = FurthestPointSampling.apply


class GatherOperation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        # type: (Any, torch.Tensor, torch.Tensor) -> torch.Tensor
        """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint)."""
        _lib.require_cuda(features, "features")
        _lib.require_cuda(idx, "idx", torch.int32)
        if not features.is_contiguous() or not idx.is_contiguous():
            raise RuntimeError("features and idx must be contiguous")     # reference fps/...cpp:28-29
        B, Cc, N = features.shape
        M = idx.size(1)
        out = torch.zeros(B, Cc, M, device=features.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(features.device):
            _lib.check(lib.dusty_gather_points(_lib.ptr(features), _lib.ptr(idx), B, Cc, N, M, _lib.ptr(out),
                                               _lib.stream_of(features)), "dusty_gather_points")
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        g = grad_out.contiguous()
        B, Cc, M = g.shape
        grad = torch.empty(B, Cc, ctx.N, device=g.device, dtype=torch.float32)
        lib = _lib.load()
        with torch.cuda.device(g.device):
            _lib.check(lib.dusty_gather_points_grad(_lib.ptr(g), _lib.ptr(idx), B, Cc, ctx.N, M, _lib.ptr(grad),
                                                    _lib.stream_of(g)), "dusty_gather_points_grad")
        return grad, None


gather_operation = GatherOperation.apply


def downsample_point_clouds(xyz, k):
    """(B,N,3) -> (B,k,3): FPS indices and the gather in one kernel launch (the reference makes two
    full copies of the cloud and launches two kernels, fps/furthest_point_sampling.py:84-93). When a
    gradient is required the gather goes through ``gather_operation`` so that it flows back to ``xyz``."""
    assert xyz.ndim == 3, "expected 3-dim, but got {}-dim tensor".format(xyz.ndim)
    assert xyz.size(2) == 3, "expected (B,N,3), but got {}".format(xyz.shape)
    assert xyz.is_cuda
    if torch.is_grad_enabled() and xyz.requires_grad:
        # the reference's composition (fps/furthest_point_sampling.py:88-92): non-differentiable indices, then
        # GatherOperation, whose backward scatters the gradient to the sampled points
        inds = furthest_point_sampling(xyz, k)
        return gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2)
    return _fps(xyz, k, True)[1]
