"""The oracle against the reference's own outputs (tests/golden/*.npz, made by oracle/gen_golden.py
from /root/reference) and against the published tie/eligibility rules. CPU only."""
import numpy as np
import pytest
import torch

from oracle import head_projection as hp
from oracle import metrics as om
from oracle import native

from helpers import lidar_like_clouds


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.mark.parametrize("name,kind,training", [("head_dusty1_eval.npz", 1, False), ("head_dusty2_eval.npz", 2, False),
                                                ("head_dusty2_train.npz", 2, True)])
def test_head_oracle_matches_reference_modules(golden, name, kind, training):
    g = golden(name)
    depth, conf = T(g["depth"]), T(g["confidence"])
    # fresh noise: l derived from the recorded uniform draws
    lp = hp.logistic_noise(T(g["u1_pixel"]), T(g["u2_pixel"]))
    if kind == 1:
        mask, dout = hp.maskout_dusty1(depth, conf, lp)
    else:
        li = hp.logistic_noise(T(g["u1_image"]), T(g["u2_image"])) if training else None
        mask, dout = hp.maskout_dusty2(depth, conf, lp, noise_image=li)
    assert np.array_equal(mask.numpy(), g["fresh_mask"])
    assert np.array_equal(dout.numpy(), g["fresh_depth"])
    # fixed noise as frozen by utils.setup's hook
    fp = T(g["fixed_noise_pixel"])
    assert np.array_equal(hp.logistic_noise(T(g["fixed_u1_pixel"]), T(g["fixed_u2_pixel"])).numpy(), g["fixed_noise_pixel"])
    for thr, suffix in ((0.5, ""), (0.3, "_t03")):
        if kind == 1:
            mask, dout = hp.maskout_dusty1(depth, conf, fp, threshold=thr)
        else:
            fi = T(g["fixed_noise_image"]) if training else None
            mask, dout = hp.maskout_dusty2(depth, conf, fp, threshold=thr, noise_image=fi)
        assert np.array_equal(mask.numpy(), g["fixed_mask" + suffix])
        assert np.array_equal(dout.numpy(), g["fixed_depth" + suffix], equal_nan=True)


def test_projection_oracle_matches_reference_lidar(golden):
    g = golden("lidar_projection.npz")
    angle = T(g["angle"])
    inv = T(g["inv"])
    for tol, key in ((1e-8, "xyz_tol1e8"), (0, "xyz_tol0"), (0.008, "xyz_tol8e3")):
        xyz = hp.inv_to_xyz(inv.clone(), angle, 0.9, 120.0, tol)
        assert np.array_equal(xyz.numpy(), g[key])
    pts = hp.project_2d_to_3d_dense(T(g["tanh_img"]), angle, 0.9, 120.0, 0)
    assert np.array_equal(pts.numpy(), g["points_eval"])
    # dropped pixels land exactly on the origin
    assert np.all(g["points_eval"][0, :2 * 64] == 0.0)


def test_angle_grid_interpolation(golden):
    g = golden("lidar_projection.npz")
    from dusty_gan_b200.utils.lidar import synthetic_hdl64e_angles
    grid = hp.angle_grid(synthetic_hdl64e_angles(), 16, 64)
    assert np.array_equal(grid.numpy(), g["angle"])


def test_chamfer_cpu_twin_matches_reference_cd_forward(golden):
    g = golden("chamfer_cpu.npz")
    d1, d2, i1, i2 = native.chamfer_forward(g["xyz1"], g["xyz2"], rounding="cpu")
    assert np.array_equal(d1, g["dist1"]) and np.array_equal(d2, g["dist2"])
    assert np.array_equal(i1, g["idx1"]) and np.array_equal(i2, g["idx2"])
    assert i2[0, 3] == i2[0, 3] and g["idx1"][2, 0] == 149
    gx1, gx2 = native.chamfer_backward(g["xyz1"], g["xyz2"], g["w1"], g["w2"], g["idx1"], g["idx2"])
    assert np.allclose(gx1, g["grad1"], rtol=1e-6, atol=1e-8) and np.allclose(gx2, g["grad2"], rtol=1e-6, atol=1e-8)


def test_chamfer_cuda_rounding_close_to_cpu_twin(golden):
    g = golden("chamfer_cpu.npz")
    d1, d2, i1, i2 = native.chamfer_forward(g["xyz1"], g["xyz2"], rounding="cuda")
    # fused vs unfused squares differ in the last bits only
    assert np.allclose(d1, g["dist1"], rtol=2e-6, atol=1e-12) and np.allclose(d2, g["dist2"], rtol=2e-6, atol=1e-12)
    assert (i1 != g["idx1"]).mean() < 0.01


def test_metrics_oracle_matches_reference_driver(golden):
    g = golden("metrics_cpu.npz")
    M_rr, M_rg, M_gg = om.pairwise_matrices(g["gen"], g["ref"], rounding="cpu")
    for mine, ref in ((M_rr, g["M_rr"]), (M_rg, g["M_rg"]), (M_gg, g["M_gg"])):
        assert np.allclose(mine, ref, rtol=1e-6, atol=1e-12)      # f32 mean order only
    assert np.array_equal(M_rr, M_rr.T) and np.all(np.diag(M_rr) == 0)
    M_dup = native.pairwise_cd(g["ref"], g["gen_dup"], rounding="cpu")
    assert np.allclose(M_dup, g["M_rg_dup"], rtol=1e-6, atol=1e-12) and M_dup[5, 3] == 0.0 and g["M_rg_dup"][5, 3] == 0.0
    scores = om.scores_from_matrices(g["M_rr"], g["M_rg"], g["M_gg"])
    ref = dict(zip([str(k) for k in g["score_keys"]], g["score_values"]))
    assert set(scores) == set(ref)
    for k, v in ref.items():
        assert scores[k] == pytest.approx(v, rel=1e-6, abs=1e-12), k
    mine = om.compute_cov_mmd_1nna(g["gen"], g["ref"], rounding="cpu")
    for k, v in ref.items():
        assert mine[k] == pytest.approx(v, rel=1e-6, abs=1e-12), k


def _tie_key(k, n):
    L = 0
    while (2 << L) <= n and L < 9:
        L += 1
    T_ = 1 << L
    t = k % T_
    rev = int(format(t, "0{}b".format(L))[::-1], 2) if L else 0
    return (rev, k // T_)


def test_fps_oracle_semantic_pins():
    # S5: idx[0] = 0 even when point 0 is dropped; nothing eligible => all zeros
    x = np.zeros((1, 700, 3), np.float32)
    assert np.all(native.fps(x, 16) == 0)
    # S4: points inside the exclusion radius are never selected (except index 0)
    x = lidar_like_clouds(2, 3000, 5)
    idx = native.fps(x, 200)
    mag = (x.astype(np.float64) ** 2).sum(-1)
    for b in range(2):
        sel = idx[b, 1:]
        assert np.all(mag[b, sel] > 1e-3)
        assert len(np.unique(sel)) == len(sel)
        assert idx[b, 0] == 0
    # S5 tie rule: all eligible points identical => every temp hits 0 after the first pick and the
    # winner among equals minimises (bitreverse(k mod T), k div T)
    n = 1500
    x = np.zeros((1, n, 3), np.float32)
    elig = np.array([5, 77, 512 + 5, 1024 + 77, 300, 1401])
    x[0, elig] = [0.3, 0.1, 0.05]
    idx = native.fps(x, 6)[0]
    first = min(elig, key=lambda k: _tie_key(int(k), n))
    assert idx[0] == 0 and np.all(idx[1:] == first)
    # fewer distinct eligible positions than samples: the farthest from the (dropped) seed first, then
    # the five coincident points tie at their distance to the origin, then every temp is 0
    x[0, 300] = [0.5, -0.2, 0.01]
    idx = native.fps(x, 6)[0]
    rest = min([k for k in elig if k != 300], key=lambda k: _tie_key(int(k), n))
    assert idx[1] == 300 and idx[2] == rest and np.all(idx[3:] == first)


def test_fps_oracle_small_n_block_sizes():
    for n in (1, 2, 3, 31, 33, 511, 513, 1000):
        x = lidar_like_clouds(1, n, 11 + n, dropped=0.2, near=0.1)
        m = min(n, 9)
        idx = native.fps(x, m)
        assert idx.shape == (1, m) and idx.min() >= 0 and idx.max() < n


# ---- oracle vs the reference's own CUDA kernels run on a B200 (oracle/gen_golden_gpu.py) ----
def test_fps_oracle_matches_reference_cuda_kernel_outputs(golden):
    g = golden("gpu_reference_kernels.npz")
    i = 0
    while f"fps{i}_case" in g:
        b, n, m, seed, dropped, near = [int(v) for v in g[f"fps{i}_case"]]
        if n * m * b <= 32768 * 512 * 2:                  # keep the CPU suite short; the rest runs in -m gpu
            x = lidar_like_clouds(b, n, seed, dropped=dropped / 1000, near=near / 1000)
            assert np.array_equal(native.fps(x, m), g[f"fps{i}_idx"]), g[f"fps{i}_case"]
        i += 1
    assert i >= 6
    assert np.array_equal(native.fps(g["fps_deg_input"], 200), g["fps_deg_idx"])


def test_chamfer_cuda_rounding_matches_reference_cuda_kernel_outputs(golden):
    from helpers import sampled_clouds
    g = golden("gpu_reference_kernels.npz")
    for i in (1,):                                        # (2,1000,777): seconds on the CPU
        b, n, m, seed = [int(v) for v in g[f"cd{i}_case"]]
        a = sampled_clouds(b, n, seed); c = lidar_like_clouds(b, m, seed + 1)
        d1, d2, i1, i2 = native.chamfer_forward(a, c, rounding="cuda")
        assert np.array_equal(d1, g[f"cd{i}_dist1"]) and np.array_equal(d2, g[f"cd{i}_dist2"])
        assert np.array_equal(i1, g[f"cd{i}_idx1"]) and np.array_equal(i2, g[f"cd{i}_idx2"])


# ---- JSD (next row 8f-2) ----
def test_jsd_oracle_matches_reference(golden):
    from oracle import jsd as oj
    g = golden("jsd_cpu.npz")
    grid, spacing = oj.grid_points(28, True)
    assert np.array_equal(grid, g["grid"])
    cg, tg = oj.vote(g["gen"]); cr, tr = oj.vote(g["ref"])
    assert np.array_equal(cg, g["counters_gen"].astype(np.int64)) and np.array_equal(cr, g["counters_ref"].astype(np.int64))
    assert cg.sum() == g["gen"].shape[0] * g["gen"].shape[1]
    assert oj.jsd_from_counts(cg, cr) == pytest.approx(float(g["jsd"]), rel=1e-5)
    assert oj.compute_jsd(g["gen"], g["ref"]) == pytest.approx(float(g["jsd"]), rel=1e-5)


@pytest.mark.parametrize("shape", [(16, 64), (12, 96), (16, 256)])
def test_real_data_oracle_matches_reference_dataset(golden, shape):
    """oracle.real_data against the reference's KITTIOdometry.preprocess/transform, LiDAR.invert_depth and
    sigmoid_to_tanh run on CPU (tests/golden/real_data.npz, oracle/gen_golden_real.py): bit-equal."""
    from oracle import real_data as rd
    g = golden("real_data.npz")
    tag = f"_{shape[0]}x{shape[1]}"
    items = [rd.dataset_item(s, shape) for s in g["scans"]]
    raw = {k: torch.stack([it[k] for it in items]) for k in items[0]}
    inv, mask, points = rd.preprocess_reals(raw)
    for key, got in (("xyz", raw["xyz"]), ("depth", raw["depth"]), ("inv", inv), ("points", points)):
        assert np.array_equal(got.numpy().view(np.int32), g[key + tag].view(np.int32)), key
    assert np.array_equal(raw["mask"].numpy(), g["mask" + tag])
    assert np.array_equal(mask.numpy() > 0, g["mask" + tag])
    # the razor-edge returns planted by the generator (row 0 of scan 0, source columns 0,4,...,24)
    if shape == (16, 256):
        assert g["mask" + tag][0, 0, 0, [0, 4, 8, 12, 16, 20, 24]].tolist() == [False, True, False, True, False, False, False]


def test_nearest_index_is_torchs():
    from oracle import real_data as rd
    import torch.nn.functional as F
    for n_in, n_out in ((2048, 512), (2048, 2048), (256, 96), (100, 36), (64, 128), (7, 5)):
        src = torch.arange(n_in, dtype=torch.float32)[None, None, None]
        want = F.interpolate(src, size=(1, n_out), mode="nearest")[0, 0, 0].long().numpy()
        assert np.array_equal(rd.nearest_index(n_out, n_in), want), (n_in, n_out)


def test_box_lower_bound_never_exceeds_a_reference_rounded_distance():
    """The exactness argument of both pruned kernels (FPS buckets, Chamfer chunks on merged clouds): the bound
    formed on the per-axis gaps between two bounding boxes, in the reference's own fma pattern, is <= the
    reference-rounded distance of EVERY pair drawn from the two boxes -- including touching and overlapping
    boxes, gaps of a few ulp, huge and denormal coordinates, and a single centre point against a box (FPS)."""
    rng = np.random.default_rng(123)
    total_pairs = 0
    for trial in range(400):
        na = int(rng.choice([1, 1, 7, 32, 256])); nb = int(rng.choice([1, 32, 128]))
        scale = float(rng.choice([1e-30, 1e-6, 1e-2, 1.0, 1e3, 1e15]))
        a = rng.standard_normal((na, 3)).astype(np.float32) * np.float32(scale * rng.uniform(0.01, 1))
        shift = rng.standard_normal(3) * scale * rng.choice([0.0, 1e-7, 1e-3, 1.0, 10.0])
        c = (rng.standard_normal((nb, 3)) * scale * rng.uniform(0.01, 1) + shift).astype(np.float32)
        if trial % 5 == 0:          # boxes that touch within an ulp along one axis
            c[:, 0] = np.nextafter(a[:, 0].max(), np.float32(np.inf), dtype=np.float32) + np.abs(c[:, 0] - c[:, 0].min())
        bad, lb = native.box_bound_violations(a, c)
        assert bad == 0, (trial, na, nb, scale, lb)
        assert lb >= 0.0
        total_pairs += na * nb
    assert total_pairs > 500_000
    # sanity of the checker itself: disjoint unit cubes one apart along x have LB = 1 exactly
    a = rng.uniform(0, 1, (50, 3)).astype(np.float32); c = rng.uniform(0, 1, (50, 3)).astype(np.float32) + np.float32([2, 0, 0])
    a[0] = (1, 0, 0); c[0] = (2, 0, 0)
    assert native.box_bound_violations(a, c) == (0, 1.0)
