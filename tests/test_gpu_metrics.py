"""MMD / COV / 1-NNA driver and the end-to-end generate-and-evaluate slice on the GPU."""
import numpy as np
import pytest
import torch

from helpers import head_inputs, sampled_clouds
from oracle import head_projection as hp
from oracle import metrics as om
from oracle import native

pytestmark = pytest.mark.gpu

KEYS = ["mmd-cd", "mmd-sample-cd", "cov-cd"] + ["1-nn-%s-cd" % k for k in
                                               ("tp", "fp", "fn", "tn", "precision", "recall", "accuracy_t", "accuracy_f", "accuracy")]


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_scores_match_reference_golden(golden):
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import compute_cov_mmd_1nna
    g = golden("metrics_cpu.npz")
    scores = compute_cov_mmd_1nna(cuda(g["gen"]), cuda(g["ref"]), 512, ("cd",), verbose=False)
    ref = dict(zip([str(k) for k in g["score_keys"]], g["score_values"]))
    assert sorted(scores) == sorted(ref) == sorted(KEYS)
    for k, v in ref.items():
        assert isinstance(scores[k], float)
        assert scores[k] == pytest.approx(v, rel=1e-5, abs=1e-12), k


def test_scores_match_oracle_on_unequal_sets():
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import compute_cov_mmd_1nna
    ref = sampled_clouds(40, 512, 61); gen = sampled_clouds(33, 512, 62) * np.float32(1.05)
    scores = compute_cov_mmd_1nna(cuda(gen), cuda(ref), 512, ("cd",), verbose=False)
    expect = om.compute_cov_mmd_1nna(gen, ref)
    for k in KEYS:
        assert scores[k] == pytest.approx(expect[k], rel=1e-5, abs=1e-12), k
    # unequal point counts take the three-launch path
    gen2 = sampled_clouds(12, 300, 63)
    s2 = compute_cov_mmd_1nna(cuda(gen2), cuda(ref[:15]), 512, ("cd",), verbose=False)
    e2 = om.compute_cov_mmd_1nna(gen2, ref[:15])
    for k in KEYS:
        assert s2[k] == pytest.approx(e2[k], rel=1e-5, abs=1e-12), k


def test_finalize_kernel_against_numpy():
    from dusty_gan_b200.utils.metrics import cov_mmd_1nna as m
    rng = np.random.default_rng(5)
    nr, ng = 57, 41
    A = rng.uniform(0.1, 1.0, (nr + ng, nr + ng)).astype(np.float32)
    A = np.minimum(A, A.T)
    np.fill_diagonal(A, 0)
    Mrr, Mrg, Mgg = A[:nr, :nr], A[:nr, nr:], A[nr:, nr:]
    cm, nna = m._finalize_device(cuda(Mrr), cuda(Mrg), cuda(Mgg))
    e = om.scores_from_matrices(Mrr, Mrg, Mgg)
    assert cm["mmd"] == pytest.approx(e["mmd-cd"], rel=1e-6) and cm["mmd-sample"] == pytest.approx(e["mmd-sample-cd"], rel=1e-6)
    assert cm["cov"] == e["cov-cd"]
    for k in ("tp", "fp", "fn", "tn", "accuracy"):
        assert nna[k] == pytest.approx(e["1-nn-%s-cd" % k], rel=1e-7), k
    # the torch-op mirrors of the reference helpers agree with the kernels
    t = m._compute_cov_mmd(cuda(Mrg))
    assert t["cov"] == cm["cov"] and t["mmd"] == pytest.approx(cm["mmd"], rel=1e-6)
    assert m._compute_nna(cuda(Mrr), cuda(Mrg), cuda(Mgg), 1)["tp"] == nna["tp"]
    # k > 1 (and the monotone sqrt option) on the device kernel against the reference's topk formulation
    label = torch.cat([torch.ones(nr), torch.zeros(ng)]).cuda()
    M = torch.cat([torch.cat((cuda(Mrr), cuda(Mrg)), 1), torch.cat((cuda(Mrg).t(), cuda(Mgg)), 1)], 0)
    for k, sqrt in ((3, False), (4, True), (7, False), (64, False)):
        Mk = (M.abs().sqrt() if sqrt else M) + torch.diag(float("inf") * torch.ones_like(label))
        _, idx = Mk.topk(k=k, dim=0, largest=False)
        count = sum(label.index_select(0, idx[i]) for i in range(k))
        pred = (count / k >= 0.5).float()
        got = m._compute_nna(cuda(Mrr), cuda(Mrg), cuda(Mgg), k, sqrt=sqrt)
        assert got["tp"] == (pred * label).sum().item() and got["fp"] == (pred * (1 - label)).sum().item()
        assert got["fn"] == ((1 - pred) * label).sum().item() and got["tn"] == ((1 - pred) * (1 - label)).sum().item()


def test_pairwise_distance_signature():
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import _pairwise_distance, compute_cd
    a = cuda(sampled_clouds(6, 256, 71)); b = cuda(sampled_clouds(4, 256, 72))
    d = _pairwise_distance(a, b, 512, ("cd",), False)
    assert set(d) == {"cd"} and d["cd"].shape == (6, 4)
    row = compute_cd(a[[2]].expand(4, -1, -1), b)           # the reference's inner-loop call
    assert torch.allclose(d["cd"][2], row, rtol=1e-6, atol=0)
    s = _pairwise_distance(a, a, 512, ("cd",), False)["cd"]
    assert torch.equal(s, s.t())
    with pytest.raises(NotImplementedError):
        _pairwise_distance(a, b, 512, ("cd", "emd"), False)


def test_generate_and_evaluate_slice():
    """Config 1 in miniature: head -> projection -> FPS -> matrices -> scores, GPU against the oracle
    chain fed the same tensors stage by stage."""
    from dusty_gan_b200.models.dusty import DUSty1
    from dusty_gan_b200.utils.lidar import LiDAR, synthetic_hdl64e_angles
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import compute_cov_mmd_1nna
    from dusty_gan_b200 import pipeline
    H, W, P, N = 64, 512, 256, 6
    lidar = LiDAR(H, W, 0.9, 120.0, angle=synthetic_hdl64e_angles()).cuda()
    head = DUSty1(torch.nn.Identity(), tau=1.0).cuda().eval()
    sets = {}
    for name, seed in (("gen", 1), ("ref", 2)):
        depth, conf, u1, u2 = head_inputs(N, 1, H, W, seed, "cuda")
        if head.gumbel.fixed_noise is None:
            head.gumbel.fixed_noise = head.gumbel._logistic_from_uniform(u1, u2)
        pts, out = pipeline.generate_points(head, {"depth": depth, "confidence": conf}, lidar, P, tol=0.0)
        _, dout = hp.maskout_dusty1(depth, conf, head.gumbel.fixed_noise)
        dense = hp.project_2d_to_3d_dense(dout, lidar.angle, 0.9, 120.0, 0.0)
        assert torch.equal(out["points"], dense)
        sub, _ = native.downsample_point_clouds(dense.cpu().numpy(), P)
        assert np.array_equal(pts.cpu().numpy(), sub)
        sets[name] = pts
    scores = compute_cov_mmd_1nna(sets["gen"], sets["ref"], 512, ("cd",), verbose=False)
    expect = om.compute_cov_mmd_1nna(sets["gen"].cpu().numpy(), sets["ref"].cpu().numpy())
    for k in KEYS:
        assert scores[k] == pytest.approx(expect[k], rel=1e-5, abs=1e-12), k


def test_config0_64_vs_64_at_full_size():
    """BASELINE configs[0] at its own size: 64 vs 64 range images of 64x512 through the DUSty-I head,
    projection and FPS to 2048 points, then every entry of M_rr / M_rg / M_gg and every score against the
    oracle (the C restatement in the reference CUDA kernel's rounding, rows spread over host threads --
    ctypes releases the GIL)."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    from dusty_gan_b200.models.dusty import DUSty1
    from dusty_gan_b200.utils.lidar import LiDAR, synthetic_hdl64e_angles
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import compute_cov_mmd_1nna, pairwise_matrices
    from dusty_gan_b200 import pipeline
    H, W, P, N = 64, 512, 2048, 64
    lidar = LiDAR(H, W, 0.9, 120.0, angle=synthetic_hdl64e_angles()).cuda()
    head = DUSty1(torch.nn.Identity(), tau=1.0).cuda().eval()
    sets = {}
    for name, seed in (("gen", 11), ("ref", 12)):
        depth, conf, u1, u2 = head_inputs(N, 1, H, W, seed, "cuda")
        if head.gumbel.fixed_noise is None:
            head.gumbel.fixed_noise = head.gumbel._logistic_from_uniform(u1, u2)
        sets[name] = pipeline.generate_points(head, {"depth": depth, "confidence": conf}, lidar, P, tol=0.0)[0]
    gen, ref = sets["gen"], sets["ref"]
    scores = compute_cov_mmd_1nna(gen, ref, 512, ("cd",), verbose=False)
    Mrr, Mrg, Mgg = [m.cpu().numpy() for m in pairwise_matrices(gen, ref)]
    stacked = np.concatenate([ref.cpu().numpy(), gen.cpu().numpy()])
    workers = min(32, os.cpu_count() or 1)
    rows = [(r, min(r + 4, 2 * N)) for r in range(0, 2 * N, 4)]
    with ThreadPoolExecutor(workers) as pool:
        parts = list(pool.map(lambda rr: native.pairwise_cd(stacked, None, rows=rr, rounding="cuda"), rows))
    U = np.zeros((2 * N, 2 * N), np.float32)
    for (r0, r1), part in zip(rows, parts):
        U[r0:r1] = part[r0:r1]
    O = np.triu(U) + np.triu(U, 1).T
    for got, want in ((Mrr, O[:N, :N]), (Mrg, O[:N, N:]), (Mgg, O[N:, N:])):
        assert np.abs(got - want).max() <= 1e-5 * want.max()
        off = want > 0
        assert (np.abs(got - want)[off] / want[off]).max() <= 1e-5
    expect = om.scores_from_matrices(O[:N, :N], O[:N, N:], O[N:, N:])
    for k in KEYS:
        assert scores[k] == pytest.approx(expect[k], rel=1e-5, abs=1e-12), k


def test_fused_epilogue_scores_are_bit_identical_to_the_matrix_path():
    """compute_cov_mmd_1nna reduces every entry inside the matrix kernel's epilogue (packed 64-bit minima)
    instead of storing (Nr+Ng)^2 floats; the scores must be exactly those of matrices + finaliser kernels, for
    equal point counts (one stacked launch) and unequal ones (three launches), including exact ties."""
    from dusty_gan_b200.utils.metrics import cov_mmd_1nna as m
    cases = [(sampled_clouds(37, 512, 81), sampled_clouds(41, 512, 82)),
             (sampled_clouds(9, 300, 83), sampled_clouds(14, 640, 84)),
             (sampled_clouds(1, 64, 85), sampled_clouds(1, 64, 86))]
    dup = sampled_clouds(12, 256, 87)
    cases.append((dup[[0, 0, 1, 2, 2, 3]], dup[[0, 4, 4, 5, 2, 2, 6]]))          # duplicated clouds: exact ties, zeros
    for gen, ref in cases:
        g, r = cuda(gen), cuda(ref)
        m.FUSED_EPILOGUE = True
        try:
            fused = m.compute_cov_mmd_1nna(g, r, 512, ("cd",), verbose=False)
            m.FUSED_EPILOGUE = False
            plain = m.compute_cov_mmd_1nna(g, r, 512, ("cd",), verbose=False)
        finally:
            m.FUSED_EPILOGUE = True
        assert fused == plain
        expect = om.compute_cov_mmd_1nna(gen, ref)
        for k in KEYS:
            assert fused[k] == pytest.approx(expect[k], rel=1e-6, abs=1e-12), k


def test_fused_evaluation_never_allocates_an_n_by_n_tensor():
    """configs[3]'s cloud count (5000 vs 5000) at 64 points per cloud: the stacked matrix would be 10000^2 x 4 B
    = 400 MB (the reference adds a 400 MB diagonal temporary, cov_mmd_1nna.py:82); the fused path may only
    allocate O(N) scratch on top of the stacked copy of the clouds."""
    from dusty_gan_b200.utils.metrics import cov_mmd_1nna as m
    N, P = 5000, 64
    gen = cuda(sampled_clouds(N, P, 91)); ref = cuda(sampled_clouds(N, P, 92))
    m.compute_cov_mmd_1nna(gen[:8], ref[:8], 512, ("cd",), verbose=False)          # warm-up: library, context
    torch.cuda.synchronize(); torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    scores = m.compute_cov_mmd_1nna(gen, ref, 512, ("cd",), verbose=False)
    torch.cuda.synchronize()
    extra = torch.cuda.max_memory_allocated() - base
    clouds_bytes = 2 * N * P * 3 * 4
    scan_bytes = 2 * N * P * 16
    box_bytes = scan_bytes // 16 + 2 * N * 1024      # a box per 32 points + 32 block boxes per cloud
    assert extra <= clouds_bytes + scan_bytes + box_bytes + (8 << 20), extra        # stacked copy + scan-format copy + boxes + 8 MB; no 400 MB matrix
    m.FUSED_EPILOGUE = False
    try:
        plain = m.compute_cov_mmd_1nna(gen, ref, 512, ("cd",), verbose=False)
    finally:
        m.FUSED_EPILOGUE = True
    assert scores == plain
