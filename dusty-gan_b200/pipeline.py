"""Fused generate -> points entry: what evaluate_synthesis.py does between ``G.backbone(latent)`` and
``compute_cov_mmd_1nna`` (reference evaluate_synthesis.py:59-64,151-163; trainers/dcgan_amp.py:344-349).

    backbone output {depth, confidence}
      -> maskout + tanh_to_sigmoid + clamp + inv_to_xyz + flatten/transpose      (one kernel)
      -> downsample_point_clouds                                                  (one kernel)
"""
import torch

from .models.dusty import _head_call
from .utils.sampling.fps import downsample_point_clouds


@torch.no_grad()
def maskout_and_project(head, output, lidar, tol=0.0, threshold=0.5, compact=False, buffers=None):
    """``head`` is a DUSty1/DUSty2 module, ``output`` the backbone's dict. Returns the dict updated like
    ``maskout`` plus ``points`` (B,H*W,3) -- the contiguous layout FPS consumes -- and, with
    ``compact=True``, ``valid_count`` (B,), ``valid_index`` (B,H*W) and ``valid_points`` (B,H*W,3):
    the valid pixels of each image in ascending pixel order (what ``points[i][valid[i]]`` selects)."""
    out, points, count, index, compacted = _head_call(head, output, threshold, lidar=lidar, tol=tol,
                                                      points_layout=1, compact=compact, buffers=buffers)
    out["points"] = points
    if compact:
        out["valid_count"], out["valid_index"], out["valid_points"] = count, index, compacted
    return out


@torch.no_grad()
def generate_points(head, output, lidar, num_points, tol=0.0, threshold=0.5):
    """range images -> FPS-sampled clouds (B,num_points,3): project_2d_to_3d of the reference."""
    out = maskout_and_project(head, output, lidar, tol=tol, threshold=threshold)
    return downsample_point_clouds(out["points"], num_points), out
