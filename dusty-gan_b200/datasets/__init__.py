"""Real-data side of the path (reference datasets/__init__.py:4-27, datasets/kitti.py)."""
from . import kitti
from .kitti import KITTIOdometry, preprocess_scans


def define_dataset(cfg, phase: str = "train", modality=["depth"]):
    """Same factory as the reference (datasets/__init__.py:4-27) for the dataset this path reads."""
    if cfg.name == "kitti_odometry":
        return kitti.KITTIOdometry(root=cfg.root, split=phase, shape=cfg.shape, min_depth=cfg.min_depth,
                                   max_depth=cfg.max_depth, flip=cfg.flip and phase == "train", modality=modality)
    raise NotImplementedError(cfg.name)
