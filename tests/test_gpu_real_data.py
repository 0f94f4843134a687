"""Real-data side: raw scans -> range images + clouds -> FPS -> cache, against the
reference's dataset class (golden, CPU) and against the oracle run on the same GPU."""
import numpy as np
import pytest
import torch

from oracle import native
from oracle import real_data as rd

pytestmark = pytest.mark.gpu


def bits(t):
    return t.contiguous().view(torch.int32)


def assert_bit_equal(a, b, what):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = bits(a) == bits(b)
    assert bool(same.all()), f"{what}: {(~same).sum().item()} of {same.numel()} differ"


def oracle_batch(scans, shape, device):
    items = [rd.dataset_item(s, shape) for s in scans]
    raw = {k: torch.stack([it[k] for it in items]) for k in items[0]}
    inv, mask, points = rd.preprocess_reals(raw, device=device)
    return raw, inv, mask, points


@pytest.mark.parametrize("shape", [(16, 64), (12, 96), (16, 256)])
def test_against_reference_dataset_golden(golden, shape):
    """Dataset half (numpy, IEEE): bit-equal to the reference's KITTIOdometry outputs whatever the
    device. Device half: the golden ran ATen's CPU kernels (true division), the kernel follows ATen's
    CUDA kernels (reciprocal multiply): equal to 2 ulp of the [-1,1] image (SURVEY.md trap T2)."""
    from dusty_gan_b200.datasets import preprocess_scans
    g = golden("real_data.npz")
    tag = f"_{shape[0]}x{shape[1]}"
    out = preprocess_scans(torch.from_numpy(g["scans"]).cuda(), shape, 0.9, 120.0, -1,
                           want=("xyz", "depth", "mask", "inv", "points"))
    assert np.array_equal(out["xyz"].cpu().numpy().view(np.int32), g["xyz" + tag].view(np.int32))
    assert np.array_equal(out["depth"].cpu().numpy().view(np.int32), g["depth" + tag].view(np.int32))
    assert np.array_equal(out["mask"].cpu().numpy() > 0, g["mask" + tag])
    assert np.array_equal(out["points"].cpu().numpy().view(np.int32), g["points" + tag].view(np.int32))
    inv = out["inv"].cpu().numpy()
    assert np.all(inv[~g["mask" + tag]] == -1)
    assert np.max(np.abs(inv - g["inv" + tag])) <= 2 * 2.0 ** -23


@pytest.mark.parametrize("channels,shape,src", [(4, (64, 512), (64, 2048)), (4, (64, 2048), (64, 2048)),
                                                (3, (16, 64), (16, 256)), (5, (10, 36), (16, 100)),
                                                (4, (32, 128), (16, 64))])
def test_bit_exact_on_device(channels, shape, src):
    from dusty_gan_b200.datasets import preprocess_scans
    scans = rd.synthetic_scans(3, seed=11, hs=src[0], ws=src[1], channels=channels)
    raw, inv, mask, points = oracle_batch(scans, shape, "cuda")
    out = preprocess_scans(torch.from_numpy(scans).cuda(), shape, 0.9, 120.0, -1,
                           want=("xyz", "depth", "mask", "inv", "points"))
    assert_bit_equal(out["xyz"], raw["xyz"].cuda(), "xyz")
    assert_bit_equal(out["depth"], raw["depth"].cuda(), "depth")
    assert_bit_equal(out["mask"], mask, "mask")
    assert_bit_equal(out["inv"], inv, "inv")
    assert_bit_equal(out["points"], points, "points")
    # optional outputs may be omitted; the mandatory ones do not change
    slim = preprocess_scans(torch.from_numpy(scans).cuda(), shape, 0.9, 120.0, -1, want=("inv",))
    assert sorted(slim) == ["inv", "mask"]
    assert_bit_equal(slim["inv"], inv, "inv (slim)")


def test_unaligned_scans_and_other_limits():
    """A (…,4) view that is not 16-byte aligned takes the scalar-load path; other depth limits/drop value."""
    from dusty_gan_b200.datasets import preprocess_scans
    scans = rd.synthetic_scans(2, seed=3, hs=8, ws=64, channels=4)
    flat = torch.zeros(scans.size + 1, device="cuda")
    flat[1:] = torch.from_numpy(scans).cuda().flatten()
    view = flat[1:].view(2, 8, 64, 4)
    assert view.data_ptr() % 16 != 0
    a = preprocess_scans(view, (8, 32), 2.0, 50.0, -2, want=("inv", "points", "depth"))
    b = preprocess_scans(torch.from_numpy(scans).cuda(), (8, 32), 2.0, 50.0, -2, want=("inv", "points", "depth"))
    for k in a:
        assert_bit_equal(a[k], b[k], k)
    items = [rd.dataset_item(s, (8, 32), 2.0, 50.0) for s in scans]
    raw = {k: torch.stack([it[k] for it in items]) for k in items[0]}
    inv, mask, points = rd.preprocess_reals(raw, 2.0, 50.0, -2, device="cuda")
    assert_bit_equal(a["inv"], inv, "inv")
    assert_bit_equal(a["points"], points, "points")


def test_cache_and_subsampling_match_the_reference_loop():
    """evaluate_synthesis.py:76-110: per-batch preprocess -> FPS -> cat -> [skip:limit:skip]."""
    from dusty_gan_b200 import pipeline
    from dusty_gan_b200.utils.lidar import LiDAR, synthetic_hdl64e_angles
    H, W, P = 64, 512, 256
    lidar = LiDAR(H, W, 0.9, 120.0, angle=synthetic_hdl64e_angles()).cuda()
    scans = rd.synthetic_scans(7, seed=21)
    batches = [torch.from_numpy(scans[:3]), {"scan": torch.from_numpy(scans[3:6])}, torch.from_numpy(scans[6:])]
    cache = pipeline.build_real_cache(batches, lidar, P)
    assert cache["2d"].shape == (7, 1, H, W) and cache["3d"].shape == (7, P, 3)
    _, inv, _, points = oracle_batch(scans, (H, W), "cuda")
    assert_bit_equal(cache["2d"], inv, "2d")
    sub, _ = native.downsample_point_clouds(points.cpu().numpy(), P)
    assert np.array_equal(cache["3d"].cpu().numpy(), sub)
    for num_test in (-1, 2, 3, 7):
        got = pipeline.subsample_time_series(cache["3d"], num_test)
        want = rd.subsample_time_series(cache["3d"], num_test)
        assert torch.equal(got, want)
    # the reference's slice starts at `skip`, so num_test == len drops the first sample (kept as is)
    assert len(pipeline.subsample_time_series(cache["3d"], 7)) == 6 and len(pipeline.subsample_time_series(cache["3d"], 3)) == 3


def test_dataset_mirror_yields_the_reference_batch(tmp_path):
    from dusty_gan_b200.datasets import define_dataset
    import types
    scans = rd.synthetic_scans(3, seed=8, hs=16, ws=128)
    d = tmp_path / "sequences" / "08" / "velodyne"
    d.mkdir(parents=True)
    for i, s in enumerate(scans):
        np.save(d / f"{i:06d}.npy", s)
    cfg = types.SimpleNamespace(name="kitti_odometry", root=str(tmp_path), shape=[16, 32], min_depth=0.9,
                                max_depth=120.0, flip=False)
    ds = define_dataset(cfg, phase="val")
    assert len(ds) == 3 and "Number of datapoints: 3" in repr(ds)
    loader = torch.utils.data.DataLoader(ds, batch_size=3, shuffle=False)
    batch = ds.preprocess_batch(next(iter(loader)))
    items = [rd.dataset_item(s, (16, 32)) for s in scans]
    for k in ("xyz", "depth", "mask"):
        want = torch.stack([it[k] for it in items]).cuda()
        assert batch[k].dtype == want.dtype and torch.equal(batch[k], want), k
    with pytest.raises(NotImplementedError):
        define_dataset(types.SimpleNamespace(**{**vars(cfg), "flip": True}), phase="train")


def test_argument_errors():
    from dusty_gan_b200.datasets import preprocess_scans
    with pytest.raises(RuntimeError, match="no CPU path"):
        preprocess_scans(torch.zeros(1, 4, 8, 4), (4, 8))
    with pytest.raises(ValueError):
        preprocess_scans(torch.zeros(1, 4, 8, 2, device="cuda"), (4, 8))
    with pytest.raises(RuntimeError, match="multiple of 4"):
        preprocess_scans(torch.zeros(1, 4, 8, 4, device="cuda"), (4, 6))
    with pytest.raises(KeyError):
        preprocess_scans(torch.zeros(1, 4, 8, 4, device="cuda"), (4, 8), want=("normals",))
    out = preprocess_scans(torch.zeros(0, 4, 8, 4, device="cuda"), (4, 8))
    assert out["inv"].shape == (0, 1, 4, 8)
