"""LiDAR projection model on libdustyb200 (mirror of reference utils/lidar.py:11-68, 111-130).

``inv_to_xyz`` -- the only method on the generate-and-evaluate path -- is one CUDA kernel. The
remaining small helpers keep the reference's names and formulas (they are set-up code, not hot).
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .. import _lib


class _InvToXYZ(torch.autograd.Function):
    """``Coordinate.inv_to_xyz`` with the gradient the reference's op chain has (utils/lidar.py:38-47,61-68):
    forward is the CUDA kernel; backward differentiates d = valid * ((1/disp - min) / range * range + min) / max,
    disp = inv * (1/min - 1/max) + 1/max, times the (constant) direction of every pixel. The backward pass is
    off the evaluate path (demo.py's inversion loss uses it) and is plain torch arithmetic."""

    @staticmethod
    def forward(ctx, inv, coord, tol):
        B = inv.shape[0]
        out = torch.empty(B, 3, coord.H, coord.W, device=inv.device, dtype=torch.float32)
        p = coord._params(B, tol, 0)
        lib = _lib.load()
        with torch.cuda.device(inv.device):
            _lib.check(lib.dusty_inv_to_xyz(C.byref(p), _lib.ptr(inv), _lib.ptr(coord.trig_table(inv.device)),
                                            _lib.ptr(out), _lib.stream_of(inv)), "dusty_inv_to_xyz")
        ctx.coord, ctx.tol = coord, tol
        ctx.save_for_backward(inv)
        return out

    @staticmethod
    def backward(ctx, grad_xyz):
        (inv,) = ctx.saved_tensors
        c = ctx.coord
        t = c.trig_table(inv.device)
        valid = ((inv - c.drop_const).abs() > ctx.tol).float()
        disp = inv * (1 / c.min_depth - 1 / c.max_depth) + 1 / c.max_depth
        ddepth = -(1 / c.min_depth - 1 / c.max_depth) / (disp * disp) / c.max_depth * valid      # d(range)/d(inv)
        direction = torch.stack([t[0] * t[2], t[0] * t[3], t[1]])[None]                          # (1,3,H,W)
        return (grad_xyz * direction).sum(dim=1, keepdim=True) * ddepth, None, None


class Coordinate(nn.Module):
    def __init__(self, min_depth, max_depth, shape, drop_const=0) -> None:
        super().__init__()
        self.min_depth = min_depth
        self.max_depth = max_depth
        self.H, self.W = shape
        self.drop_const = drop_const
        self.register_buffer("angle", self.init_coordmap(self.H, self.W))
        self._trig = {}

    def init_coordmap(self, H, W):
        raise NotImplementedError

    # -- value-range maps (reference utils/lidar.py:23-47) --
    @staticmethod
    def normalize_minmax(tensor, vmin: float, vmax: float):
        return (tensor - vmin) / (vmax - vmin)

    @staticmethod
    def denormalize_minmax(tensor, vmin: float, vmax: float):
        return tensor * (vmax - vmin) + vmin

    def invert_depth(self, norm_depth):
        depth = self.denormalize_minmax(norm_depth, self.min_depth, self.max_depth)
        disp = 1 / depth
        return self.normalize_minmax(disp, 1 / self.max_depth, 1 / self.min_depth)

    def revert_depth(self, norm_disp, norm=True):
        disp = self.denormalize_minmax(norm_disp, 1 / self.max_depth, 1 / self.min_depth)
        depth = 1 / disp
        return self.normalize_minmax(depth, self.min_depth, self.max_depth) if norm else depth

    def xyz_to_pol(self, xyz):
        return torch.norm(xyz, p=2, dim=1, keepdim=True)

    # -- the hot method --
    def trig_table(self, device):
        """(4,H,W) f32: cos(elev), sin(elev), cos(azim), sin(azim), evaluated by torch on ``device``
        exactly as pol_to_xyz evaluates them on every call (reference utils/lidar.py:51-52)."""
        key = (str(device), self.angle._version, self.angle.data_ptr())
        tab = self._trig.get(key)
        if tab is None:
            ang = self.angle.to(device)
            cos, sin = torch.cos(ang), torch.sin(ang)
            tab = torch.stack([cos[0, 0], sin[0, 0], cos[0, 1], sin[0, 1]]).contiguous()
            self._trig = {key: tab}
        return tab

    def _params(self, B, tol, layout):
        from ..models.dusty import _projection_fields
        if self.drop_const != 0:
            raise NotImplementedError("the kernel implements the LiDAR drop value 0 used by every caller "
                                      "(reference utils/lidar.py:12,121-125)")
        p = _lib.HeadParams()
        p.b, p.h, p.w, p.conf_channels = B, self.H, self.W, 1
        p.points_layout = layout
        _projection_fields(p, self, tol)
        return p

    def inv_to_xyz(self, inv_depth, tol=1e-8):
        """Normalised inverse depth (B,1,H,W) in [0,1] -> xyz (B,3,H,W); dropped pixels map to the origin.
        Differentiable with respect to ``inv_depth`` like the reference's op chain."""
        _lib.require_cuda(inv_depth, "inv_depth")
        if inv_depth.dim() != 4 or inv_depth.shape[1:] != (1, self.H, self.W):
            raise ValueError(f"expected (B,1,{self.H},{self.W}), got {tuple(inv_depth.shape)}")
        self._params(1, tol, 0)       # argument checks before the autograd node is built
        return _InvToXYZ.apply(inv_depth.contiguous(), self, tol)

    def pol_to_xyz(self, polar):
        """Range image (B,1,H,W) -> xyz (B,3,H,W) (reference utils/lidar.py:49-56); set-up helper."""
        assert polar.dim() == 4
        t = self.trig_table(polar.device)
        return torch.cat((polar * t[0] * t[2], polar * t[0] * t[3], polar * t[1]), dim=1)


class LiDAR(Coordinate):
    """Angle grid from ``angles.pt`` (2,64,2048: elevation, azimuth), bilinearly resized to (H,W)."""

    def __init__(self, num_ring, num_points, min_depth, max_depth, angle_file=None, angle=None):
        if angle is None:
            assert os.path.exists(angle_file), angle_file
            angle = torch.load(angle_file)
        self.angle_file = angle_file
        self._raw_angle = angle
        super().__init__(min_depth=min_depth, max_depth=max_depth, shape=(num_ring, num_points))
        del self._raw_angle

    def init_coordmap(self, H, W):
        angle = self._raw_angle[None].float()
        return F.interpolate(angle, size=(H, W), mode="bilinear")


def synthetic_hdl64e_angles(rings=64, columns=2048):
    """An ``angles.pt`` stand-in with the format process_kitti.py writes (reference
    process_kitti.py:109-111,163-170,222): channel 0 elevation (+2 deg ... -24.8 deg over the rings),
    channel 1 azimuth (pi ... -pi over the columns). Used by tests and the benchmark only."""
    elev = torch.linspace(np.deg2rad(2.0), np.deg2rad(-24.8), rings, dtype=torch.float64)
    azim = torch.linspace(np.pi, -np.pi, columns + 1, dtype=torch.float64)[:-1]
    grid = torch.stack([elev[:, None].expand(rings, columns), azim[None, :].expand(rings, columns)])
    return grid.float().contiguous()
