"""Golden vectors for the real-data side from the reference's OWN code -- TEST INFRASTRUCTURE.

    python oracle/gen_golden_real.py          (build container only: needs /root/reference)

Runs, unmodified and on CPU: ``datasets.kitti.KITTIOdometry.preprocess`` + ``.transform``
(datasets/kitti.py:54-78, through torchvision's TF.to_tensor / TF.resize(NEAREST)),
``utils.lidar.LiDAR.invert_depth`` (utils/lidar.py:31-36) and ``utils.sigmoid_to_tanh``
(utils/__init__.py:70-73). ``preprocess_reals`` itself is a closure inside evaluate_synthesis.py's
``__main__`` (:49-57) and cannot be imported; its three glue lines (flatten/transpose, the mask
blend) are repeated here verbatim around those reference calls.
Output: tests/golden/real_data.npz.
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from gen_golden import OUT, hdl64e_angles, import_reference, np_  # noqa: E402
from oracle import real_data  # noqa: E402


def main():
    torch.set_grad_enabled(False)
    _, ref_lidar, _, _, ref_utils, _ = import_reference()
    from datasets.kitti import KITTIOdometry
    scans = real_data.synthetic_scans(3, seed=5, hs=16, ws=256, channels=4)
    # razor edges of the mask: exactly min_depth / max_depth, just inside, the origin, a NaN return
    f = np.float32
    scans[0, 0, 0, :3] = (f(0.9), 0, 0)
    scans[0, 0, 4, :3] = (np.nextafter(f(0.9), f(1)), 0, 0)
    scans[0, 0, 8, :3] = (0, f(120.0), 0)
    scans[0, 0, 12, :3] = (0, np.nextafter(f(120.0), f(0)), 0)
    scans[0, 0, 16, :3] = (0, 0, 0)
    scans[0, 0, 20, :3] = (np.nan, 1, 1)
    scans[0, 0, 24, :3] = (f(1e-30), f(1e-30), 0)            # squares underflow: depth 0 -> masked
    out = {"scans": scans}
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "angles.pt")
        torch.save(hdl64e_angles(), path)
        for H, W in ((16, 64), (12, 96), (16, 256)):
            ds = KITTIOdometry.__new__(KITTIOdometry)        # __init__ only globs the (absent) data root
            ds.min_depth, ds.max_depth, ds.shape, ds.flip, ds.modality = 0.9, 120.0, (H, W), False, ("depth",)
            items = []
            for s in scans:
                points = s.astype(np.float32)                # datasets/kitti.py:82-85
                item = ds.transform(ds.preprocess({"xyz": points[..., :3]}))
                items.append(item)
            raw_batch = {k: torch.stack([it[k] for it in items]) for k in items[0]}
            lidar = ref_lidar.LiDAR(num_ring=H, num_points=W, min_depth=0.9, max_depth=120.0, angle_file=path)
            # preprocess_reals, evaluate_synthesis.py:49-57 (device = cpu, drop_const = -1)
            xyz = raw_batch["xyz"]
            points = xyz.flatten(2).transpose(1, 2)
            depth = raw_batch["depth"]
            mask = raw_batch["mask"].float()
            inv = lidar.invert_depth(depth)
            inv = ref_utils.sigmoid_to_tanh(inv)
            inv = mask * inv + (1 - mask) * -1
            tag = f"_{H}x{W}"
            out.update({"xyz" + tag: np_(xyz), "depth" + tag: np_(depth), "mask" + tag: np_(raw_batch["mask"]),
                        "inv" + tag: np_(inv), "points" + tag: np_(points.contiguous())})
    np.savez_compressed(os.path.join(OUT, "real_data.npz"), **out)
    print("wrote real_data.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
