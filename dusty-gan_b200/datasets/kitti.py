"""KITTI odometry scans on libdustyb200 (mirror of reference datasets/kitti.py:21-98).

The reference preprocesses every scan in numpy on DataLoader workers (norm, mask, min-max map, unit
scaling, nearest resize: datasets/kitti.py:54-78) and finishes on the device with ~12 element-wise
kernels (``preprocess_reals``, evaluate_synthesis.py:49-57). Here the dataset hands out the RAW
``(64,2048,4)`` scan that process_kitti.py stored and one kernel does both halves per batch
(``preprocess_scans``); ``KITTIOdometry.collate`` + ``preprocess_scans`` reproduce the dict the
reference's ``__getitem__`` + default collate yield, bit for bit.
"""
import ctypes as C
import os.path as osp
from glob import glob

import numpy as np
import torch

from .. import _lib

CONFIG = {
    "split": {
        "train": [0, 1, 2, 3, 4, 5, 6, 7, 9, 10],
        "val": [8],
        "test": [11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21],
        "custom": [16],
    },
}


def _scan_params(scans, shape, min_depth, max_depth, drop_const):
    f32 = np.float32
    B, Hs, Ws, Cc = scans.shape
    p = _lib.ScanParams()
    p.b, p.hs, p.ws, p.channels = B, Hs, Ws, Cc
    p.h, p.w = int(shape[0]), int(shape[1])
    p.scale_h = f32(Hs) / f32(p.h)           # ATen nearest: compute_scales_value<float>(input/output)
    p.scale_w = f32(Ws) / f32(p.w)
    p.min_depth = f32(min_depth)
    p.max_depth = f32(max_depth)
    p.range = f32(max_depth - min_depth)
    p.disp_lo = f32(1 / max_depth)
    p.inv_disp_range = f32(1.0 / (1 / min_depth - 1 / max_depth))   # ATen CUDA: x / scalar = x * f32(1/double(scalar))
    p.drop_const = f32(drop_const)
    return p


@torch.no_grad()
def preprocess_scans(scans, shape, min_depth=0.9, max_depth=120.0, drop_const=-1, want=("inv", "mask", "points"),
                     buffers=None):
    """scans (B,Hs,Ws,C>=3) f32 CUDA, xyz in metres in channels 0..2 -> dict with the requested keys:

        "xyz" (B,3,H,W), "depth" (B,1,H,W), "mask" (B,1,H,W) f32   -- the dataset's outputs after collate
        "inv" (B,1,H,W), "points" (B,H*W,3)                        -- preprocess_reals' outputs

    One launch of ``dusty_scan_preprocess``; "mask" and "inv" are always produced. ``buffers`` may hold
    preallocated contiguous f32 outputs under the same keys (steady-state callers reuse them)."""
    _lib.require_cuda(scans, "scans")
    if scans.dim() != 4 or scans.shape[-1] < 3:
        raise ValueError(f"expected scans (B,Hs,Ws,C>=3), got {tuple(scans.shape)}")
    unknown = set(want) - {"xyz", "depth", "mask", "inv", "points"}
    if unknown:
        raise KeyError(f"unknown outputs {sorted(unknown)}")
    s = scans.contiguous()
    B = s.shape[0]
    H, W = int(shape[0]), int(shape[1])
    buffers = buffers or {}

    def new(key, *sh):
        t = buffers.get(key)
        if t is None:
            return torch.empty(*sh, device=s.device, dtype=torch.float32)
        if tuple(t.shape) != sh or t.dtype != torch.float32 or t.device != s.device or not t.is_contiguous():
            raise ValueError(f"preallocated buffer {key!r} does not match the expected shape/dtype/device")
        return t

    out = {"mask": new("mask", B, 1, H, W), "inv": new("inv", B, 1, H, W)}
    if "depth" in want:
        out["depth"] = new("depth", B, 1, H, W)
    if "points" in want:
        out["points"] = new("points", B, H * W, 3)
    if "xyz" in want:
        out["xyz"] = new("xyz", B, 3, H, W)
    p = _scan_params(s, shape, min_depth, max_depth, drop_const)
    lib = _lib.load()
    with torch.cuda.device(s.device):
        _lib.check(lib.dusty_scan_preprocess(C.byref(p), _lib.ptr(s), _lib.ptr(out.get("depth")), _lib.ptr(out["mask"]),
                                             _lib.ptr(out["inv"]), _lib.ptr(out.get("points")), _lib.ptr(out.get("xyz")),
                                             _lib.stream_of(s)), "dusty_scan_preprocess")
    return out


class KITTIOdometry(torch.utils.data.Dataset):
    """Same constructor and file layout as the reference (datasets/kitti.py:21-52); items are raw scans."""

    def __init__(self, root, split, shape=(64, 256), min_depth=0.9, max_depth=120.0, flip=False, config=CONFIG,
                 modality=("depth")):
        super().__init__()
        self.root = osp.join(root, "sequences")
        self.split = split
        self.config = config
        self.subsets = np.asarray(self.config["split"][split])
        self.shape = tuple(shape)
        self.min_depth = min_depth
        self.max_depth = max_depth
        if flip:
            raise NotImplementedError("random horizontal flip is a training-time augmentation "
                                      "(reference datasets/__init__.py:12); the evaluate path runs with flip=False")
        self.flip = flip
        assert "depth" in modality, '"depth" is required'
        self.modality = modality
        self.datalist = None
        self.load_datalist()

    def load_datalist(self):
        datalist = []
        for subset in self.subsets:
            subset_dir = osp.join(self.root, str(subset).zfill(2))
            datalist += list(sorted(glob(osp.join(subset_dir, "velodyne/*"))))
        self.datalist = datalist

    def __getitem__(self, index):
        """The raw (64,2048,4) scan (reference :81-82); everything after it runs on the GPU per batch."""
        return {"scan": torch.from_numpy(np.load(self.datalist[index]).astype(np.float32))}

    def preprocess_batch(self, raw_batch, device="cuda"):
        """{"scan": (B,Hs,Ws,4)} from the default collate -> {"xyz","depth","mask"} on ``device``: the batch
        the reference's DataLoader yields (mask as bool, like the reference's collated numpy mask)."""
        scans = raw_batch["scan"].to(device, non_blocking=True)
        out = preprocess_scans(scans, self.shape, self.min_depth, self.max_depth, want=("xyz", "depth", "mask"))
        return {"xyz": out["xyz"], "depth": out["depth"], "mask": out["mask"] > 0}

    def __len__(self):
        return len(self.datalist)

    def __repr__(self) -> str:
        head = "Dataset " + self.__class__.__name__
        body = ["Number of datapoints: {}".format(self.__len__())]
        body.append("Root location: {}".format(self.root))
        lines = [head] + ["    " + line for line in body]
        return "\n".join(lines)
