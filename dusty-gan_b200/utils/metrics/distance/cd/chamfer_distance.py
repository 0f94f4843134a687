"""Chamfer distance on libdustyb200.

Public names and call signatures are those of the reference module
(utils/metrics/distance/cd/chamfer_distance.py:16-69): ``ChamferDistanceFunction``, ``ChamferDistance`` and the
callable ``chamfer_distance(xyz1, xyz2) -> (dist1, dist2)`` with a backward pass. The body is two C-ABI calls.
"""
import torch

from ..... import _lib


def _nn_both_ways(a, b):
    """a (B,N,3), b (B,M,3) contiguous CUDA f32 -> squared NN distances and arg-mins in both directions."""
    B, N, M = a.size(0), a.size(1), b.size(1)
    opts = dict(device=a.device)
    d_ab = torch.empty(B, N, dtype=torch.float32, **opts)
    d_ba = torch.empty(B, M, dtype=torch.float32, **opts)
    i_ab = torch.empty(B, N, dtype=torch.int32, **opts)
    i_ba = torch.empty(B, M, dtype=torch.int32, **opts)
    lib = _lib.load()
    ws_bytes = lib.dusty_chamfer_forward_workspace_bytes(B, N, M)
    ws = _lib.workspace(ws_bytes, a.device)
    with torch.cuda.device(a.device):      # the reference launches on whatever device is current
        rc = lib.dusty_chamfer_forward(_lib.ptr(a), _lib.ptr(b), B, N, M, _lib.ptr(d_ab), _lib.ptr(d_ba), _lib.ptr(i_ab),
                                       _lib.ptr(i_ba), _lib.ptr(ws), ws_bytes, _lib.stream_of(a))
    _lib.check(rc, "dusty_chamfer_forward")
    return d_ab, d_ba, i_ab, i_ba


def _scatter_grads(a, b, g_ab, g_ba, i_ab, i_ba):
    """Gradients of sum(g_ab * d_ab) + sum(g_ba * d_ba) with respect to both clouds (reference .cu:148-190)."""
    ga, gb = torch.empty_like(a), torch.empty_like(b)
    lib = _lib.load()
    with torch.cuda.device(a.device):
        rc = lib.dusty_chamfer_backward(_lib.ptr(a), _lib.ptr(b), a.size(0), a.size(1), b.size(1), _lib.ptr(g_ab), _lib.ptr(g_ba),
                                        _lib.ptr(i_ab), _lib.ptr(i_ba), _lib.ptr(ga), _lib.ptr(gb), None, 0, _lib.stream_of(a))
    _lib.check(rc, "dusty_chamfer_backward")
    return ga, gb


class ChamferDistanceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        for t, name in ((xyz1, "xyz1"), (xyz2, "xyz2")):
            _lib.require_cuda(t, name)
        ok = xyz1.dim() == 3 and xyz2.dim() == 3 and xyz1.size(2) == 3 and xyz2.size(2) == 3 and xyz1.size(0) == xyz2.size(0)
        if not ok:
            raise ValueError(f"expected (B,N,3) and (B,M,3), got {tuple(xyz1.shape)} and {tuple(xyz2.shape)}")
        a, b = xyz1.contiguous(), xyz2.contiguous()
        d_ab, d_ba, i_ab, i_ba = _nn_both_ways(a, b)
        ctx.save_for_backward(a, b, i_ab, i_ba)
        return d_ab, d_ba

    @staticmethod
    def backward(ctx, grad_ab, grad_ba):
        a, b, i_ab, i_ba = ctx.saved_tensors
        return _scatter_grads(a, b, grad_ab.contiguous(), grad_ba.contiguous(), i_ab, i_ba)


class ChamferDistance(torch.nn.Module):
    def forward(self, xyz1, xyz2):
        return ChamferDistanceFunction.apply(xyz1, xyz2)


chamfer_distance = ChamferDistanceFunction.apply
