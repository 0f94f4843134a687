"""SURVEY.md section 8(a) semantic pins S1-S9, one test each, kernel against oracle on the GPU."""
import numpy as np
import pytest
import torch

from helpers import head_inputs, lidar_like_clouds, sampled_clouds
from oracle import head_projection as hp
from oracle import metrics as om
from oracle import native

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _lidar(H, W):
    from dusty_gan_b200.utils.lidar import LiDAR, synthetic_hdl64e_angles
    return LiDAR(H, W, 0.9, 120.0, angle=synthetic_hdl64e_angles()).cuda()


def test_S1_literal_sigmoid_strict_threshold_and_broadcast_noise():
    """x = logit + l with |x| below 1.2e-7 still goes through 1/(1+exp(-x)) > 0.5, not `x > 0`;
    the fixed (1,1,H,W) noise is shared by the whole batch."""
    from dusty_gan_b200.models.dusty import GumbelSigmoid
    g = GumbelSigmoid(tau=1.0).cuda()
    H, W = 8, 64
    noise = torch.zeros(1, 1, H, W, device="cuda")
    g.fixed_noise = noise
    tiny = torch.tensor([0.0, 1e-8, 5.9e-8, 6.1e-8, 1.2e-7, 2.4e-7, -1e-8, -1.2e-7], device="cuda")
    logits = tiny.repeat(H * W // 8).view(1, 1, H, W).repeat(3, 1, 1, 1).contiguous()
    out = g(logits)
    ref = hp.gumbel_sigmoid(logits, noise)
    assert torch.equal(out, ref)
    assert not torch.equal(out, (logits > 0).float())            # the algebraic shortcut is NOT what runs
    assert torch.equal(out[0], out[1]) and torch.equal(out[0], out[2])


def test_S2_validity_is_rederived_from_the_value():
    """A kept pixel whose generated depth is exactly -1 is dropped downstream; a dropped pixel is 0."""
    from dusty_gan_b200.models.dusty import DUSty1
    from dusty_gan_b200 import pipeline
    H, W = 8, 64
    lidar = _lidar(H, W)
    head = DUSty1(torch.nn.Identity(), tau=1.0).cuda().eval()
    head.gumbel.fixed_noise = torch.zeros(1, 1, H, W, device="cuda")
    depth = torch.full((1, 1, H, W), 0.25, device="cuda")
    conf = torch.full((1, 1, H, W), 5.0, device="cuda")          # everything kept ...
    depth[0, 0, 0, :4] = -1.0                                     # ... but these four are -1 by themselves
    conf[0, 0, 1, :4] = -5.0                                      # and these four are dropped by the mask
    out = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0, compact=True)
    pts = out["points"].view(H, W, 3)
    assert torch.all(out["mask"][0, 0, 0, :4] == 1) and torch.all(pts[0, :4] == 0)
    assert torch.all(out["mask"][0, 0, 1, :4] == 0) and torch.all(pts[1, :4] == 0)
    assert int(out["valid_count"][0]) == H * W - 8
    assert torch.all(pts[2].abs().sum(-1) > 0)


def test_S3_distance_rounding_pattern():
    """d = fma(dz,dz,fma(dx,dx,dy*dy)) on differences: bit-equal to the restatement of the reference
    kernel's SASS for (practically) every point, not merely close."""
    from test_gpu_chamfer import run_forward
    a = sampled_clouds(2, 1500, 801); b = sampled_clouds(2, 1700, 802)
    d1, d2, _, _ = run_forward(a, b)
    o1, o2, _, _ = native.chamfer_forward(a, b, rounding="cuda")
    u1, _, _, _ = native.chamfer_forward(a, b, rounding="cpu")   # unfused squares differ in the last bit for many points
    assert (d1 == o1).mean() > 0.999 and (d2 == o2).mean() > 0.999
    assert (o1 != u1).mean() > 0.05


def test_S4_fps_eligibility_threshold_in_double():
    from test_gpu_fps import fps_gpu
    r_in, r_out = np.float32(np.sqrt(1e-3) * 0.9999), np.float32(np.sqrt(1e-3) * 1.0001)
    x = np.zeros((1, 600, 3), np.float32)
    x[0, 1:300, 0] = r_in                        # mag <= 1e-3: never takes part
    x[0, 300:, 0] = r_out
    x[0, 300:, 1] = np.linspace(0.0, 0.3, 300, dtype=np.float32)
    idx = fps_gpu(x, 50)
    assert np.array_equal(idx, native.fps(x, 50))
    assert idx[0, 0] == 0 and np.all(idx[0, 1:] >= 300)


def test_S5_fps_seed_and_tie_rule():
    from test_gpu_fps import fps_gpu
    n = 2048
    x = np.zeros((1, n, 3), np.float32)          # index 0 dropped: still the seed
    ks = np.array([3, 515, 1027, 1539, 7, 519, 64, 1088])
    x[0, ks] = [0.2, 0.1, -0.05]                 # eight coincident eligible points: pure tie
    idx = fps_gpu(x, 4)
    assert np.array_equal(idx, native.fps(x, 4))
    key = lambda k: (int(format(k % 512, "09b")[::-1], 2), k // 512)
    assert idx[0, 0] == 0 and np.all(idx[0, 1:] == min(ks, key=key))


def test_S6_chamfer_ties_take_the_lowest_index():
    from test_gpu_chamfer import run_forward
    b = sampled_clouds(1, 500, 811)
    b = np.concatenate([b, b, b], axis=1)         # every candidate three times, 500 apart (other chunks, other tiles)
    a = sampled_clouds(1, 300, 812)
    d1, _, i1, _ = run_forward(a, b)
    o1, _, j1, _ = native.chamfer_forward(a, b, rounding="cuda")
    same = d1 == o1
    assert same.mean() > 0.999 and np.array_equal(i1[same], j1[same]) and np.all(i1 < 500)


def test_S7_means_divide_by_the_full_point_count():
    """Un-sampled clouds: dropped pixels are origin points and count in both means (config 5's shape)."""
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix, compute_cd
    a = lidar_like_clouds(3, 4096, 821, dropped=0.5); b = lidar_like_clouds(2, 4096, 822, dropped=0.4)
    M = chamfer_matrix(cuda(a), cuda(b)).cpu().numpy()
    O = native.pairwise_cd(a, b, rounding="cuda")
    assert np.abs(M - O).max() <= 1e-5 * O.max()
    row = compute_cd(cuda(a[[1]]).expand(2, -1, -1), cuda(b)).cpu().numpy()
    assert np.allclose(row, O[1], rtol=1e-5)


def test_S8_symmetric_matrices_are_exactly_symmetric_with_zero_diagonal():
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix
    a = cuda(sampled_clouds(9, 700, 831))
    S = chamfer_matrix(a)                          # upper triangle computed, mirrored
    R = chamfer_matrix(a, a.clone())               # same data through the rectangular path: all 81 entries
    assert torch.equal(S, S.t()) and torch.all(S.diagonal() == 0)
    assert torch.equal(S, R)                       # (a-b)^2 == (b-a)^2 bit for bit


def test_S9_scores_on_degenerate_inputs():
    """Literal random-init regime (trap T1): every FPS pick is index 0, clouds collapse to one point;
    scores are compared, not indices."""
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import compute_cov_mmd_1nna
    from dusty_gan_b200.utils.sampling.fps import downsample_point_clouds
    rng = np.random.default_rng(5)
    raw = (rng.standard_normal((10, 4096, 3)) * 0.004).astype(np.float32)
    sub = downsample_point_clouds(cuda(raw), 64)
    assert torch.equal(sub, cuda(raw)[:, :1].expand(-1, 64, -1))
    gen, ref = sub[:5].contiguous(), sub[5:].contiguous()
    s = compute_cov_mmd_1nna(gen, ref, 512, ("cd",), verbose=False)
    e = om.compute_cov_mmd_1nna(gen.cpu().numpy(), ref.cpu().numpy())
    for k in ("mmd-cd", "mmd-sample-cd", "cov-cd", "1-nn-accuracy-cd"):
        assert s[k] == pytest.approx(e[k], rel=1e-5, abs=1e-12), k
