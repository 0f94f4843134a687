"""Real-data side of the path: raw scans in, GPU preprocessing per batch (see datasets/kitti.py)."""
from . import kitti
from .kitti import KITTIOdometry, preprocess_scans

__all__ = ["define_dataset", "KITTIOdometry", "preprocess_scans", "kitti"]


def define_dataset(cfg, phase: str = "train", modality=["depth"]):
    """Factory with the reference's signature (datasets/__init__.py:4). Only the KITTI odometry scans are on
    this path; any other ``cfg.name`` raises like the reference does for unknown names."""
    if cfg.name != "kitti_odometry":
        raise NotImplementedError(cfg.name)
    augment = bool(cfg.flip) and phase == "train"
    return KITTIOdometry(cfg.root, phase, shape=cfg.shape, min_depth=cfg.min_depth, max_depth=cfg.max_depth, flip=augment,
                         modality=modality)
