"""The Chamfer kernels return the reference CUDA kernel's distances and arg-mins bit for bit, also where the
|b|^2 - 2 a.b search cannot tell candidates apart (csrc/chamfer.cu: guard). Adversarial inputs: compact
clusters far from the origin, coincident clouds with a large offset, coordinates of 10..100, exact ties on a
lattice, duplicated candidates -- through the batch front end (dist + idx) and the matrix front end."""
import numpy as np
import pytest
import torch

from helpers import lidar_like_clouds, sampled_clouds
from oracle import native

pytestmark = pytest.mark.gpu
ULP = 2.0 ** -23


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def batch_exact(a, b):
    from test_gpu_chamfer import run_forward
    d1, d2, i1, i2 = run_forward(a, b)
    o1, o2, j1, j2 = native.chamfer_forward(a, b, rounding="cuda")
    for d, o, i, j, name in ((d1, o1, i1, j1, "1->2"), (d2, o2, i2, j2, "2->1")):
        assert np.array_equal(d, o), f"{name}: {(d != o).sum()} of {d.size} distances differ, max rel {np.abs(d - o).max() / max(o.max(), 1e-30):.2e}"
        assert np.array_equal(i, j), f"{name}: {(i != j).sum()} arg-mins differ"


def matrix_exact(a, b=None, **kw):
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix
    M = chamfer_matrix(cuda(a), None if b is None else cuda(b), **kw).cpu().numpy()
    O = native.pairwise_cd(a, b, rounding="cuda")
    assert np.all(np.abs(M.astype(np.float64) - O) <= ULP * np.abs(O)), np.abs(M - O).max()
    return M


def cube(n_clouds, n_points, seed, side=0.05, centre=(0.7, 0.5, 0.1)):
    rng = np.random.default_rng(seed)
    return (rng.uniform(0, side, (n_clouds, n_points, 3)) + np.asarray(centre)).astype(np.float32)


@pytest.mark.parametrize("points", [2048, 16384])
def test_compact_cluster_far_from_the_origin(points):
    # |p| ~ 0.87, extent 0.05: nearest-neighbour d ~ 1e-5 ... 1e-6 is of the order of the search's rounding
    # error 2^-23 (|a|^2 + |b|^2) ~ 2e-7, so a search-only kernel picks wrong neighbours here
    a, b = cube(2, points, 1), cube(2, points, 2)
    batch_exact(a, b)
    matrix_exact(cube(3, points, 3), cube(2, points, 4), merge_origin=False)
    if points > 4096:
        matrix_exact(cube(3, points, 3), cube(2, points, 4))          # default path: merged + sorted + pruned


def test_coincident_clouds_with_a_large_offset():
    base = sampled_clouds(2, 2048, 11) + np.float32([50.0, -30.0, 10.0])
    jitter = (np.random.default_rng(12).standard_normal(base.shape) * 1e-4).astype(np.float32)
    batch_exact(base, base.copy())                        # d = 0 everywhere, arg-min = lowest index of a duplicate
    batch_exact(base, base + jitter)
    M = matrix_exact(np.concatenate([base, base + jitter]))
    assert np.all(np.diag(M) == 0)


@pytest.mark.parametrize("scale", [10.0, 100.0])
def test_coordinates_in_metres(scale):
    a = lidar_like_clouds(2, 4096, 21) * np.float32(scale)
    b = lidar_like_clouds(2, 4096, 22) * np.float32(scale)
    batch_exact(a, b)
    matrix_exact(sampled_clouds(4, 2048, 23) * np.float32(scale), sampled_clouds(3, 2048, 24) * np.float32(scale))


def test_exact_ties_on_a_lattice_take_the_lowest_index():
    # candidates on a binary lattice, queries at cell centres: eight candidates at bit-equal distance each
    g = np.arange(12, dtype=np.float32) * 0.125
    lattice = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(1, -1, 3)          # 1728 points
    centres = np.stack(np.meshgrid(g[:-1], g[:-1], g[:-1], indexing="ij"), -1).reshape(1, -1, 3) + np.float32(0.0625)
    perm = np.random.default_rng(5).permutation(lattice.shape[1])
    batch_exact(centres, lattice[:, perm])
    batch_exact(centres + np.float32([3.0, 5.0, -2.0]), lattice[:, perm] + np.float32([3.0, 5.0, -2.0]))


def test_duplicated_candidates_and_single_point_clouds():
    a = sampled_clouds(3, 700, 31)
    b = np.repeat(sampled_clouds(3, 350, 32), 2, axis=1)            # every candidate twice, 2k and 2k+1
    batch_exact(a, b)
    batch_exact(a, np.repeat(a[:, :1], 64, axis=1))                 # one point, 64 copies
    matrix_exact(a, b)


def test_guard_is_rare_on_lidar_clouds_and_results_do_not_depend_on_it(monkeypatch):
    """The bench workload: sampled LiDAR clouds. Matrix entries bit-equal between the dense and the merged
    instantiations (different chunk / row layouts, hence different near-tie sets for the guard)."""
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix
    a = cuda(lidar_like_clouds(6, 5000, 41, dropped=0.4))
    plain = chamfer_matrix(a, merge_origin=False)
    merged = chamfer_matrix(a, merge_origin=True)
    assert bool((torch.abs(plain - merged) <= ULP * plain).all())
