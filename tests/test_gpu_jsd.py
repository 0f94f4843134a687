"""JSD occupancy voting and divergence against the reference's golden
outputs and the brute-force oracle."""
import numpy as np
import pytest
import torch

from helpers import lidar_like_clouds
from oracle import jsd as oj

pytestmark = pytest.mark.gpu


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_votes_and_divergence_match_reference_golden(golden):
    from dusty_gan_b200.utils.metrics import jsd as J
    g = golden("jsd_cpu.npz")
    grid, _ = J.unit_cube_grid_point_cloud(28, True, "cuda")
    assert np.array_equal(grid.cpu().numpy(), g["grid"])
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ent_g, cnt_g = J.entropy_of_occupancy_grid(cuda(g["gen"]), 28, True)
        ent_r, cnt_r = J.entropy_of_occupancy_grid(cuda(g["ref"]), 28, True)
    assert np.array_equal(cnt_g.cpu().numpy(), g["counters_gen"]) and np.array_equal(cnt_r.cpu().numpy(), g["counters_ref"])
    assert float(ent_g) == pytest.approx(float(g["entropy_gen"]), rel=1e-5)
    assert float(ent_r) == pytest.approx(float(g["entropy_ref"]), rel=1e-5)
    assert J.compute_jsd(cuda(g["gen"]), cuda(g["ref"])) == pytest.approx(float(g["jsd"]), rel=1e-5)
    # the torch-op mirror of the divergence agrees with the kernel
    t = J._jensen_shannon_divergence(cnt_g.clone(), cnt_r.clone()).item()
    assert t == pytest.approx(float(g["jsd"]), rel=1e-5)


@pytest.mark.parametrize("shape,seed", [((12, 2048), 1), ((3, 5000), 2), ((40, 100), 3)])
def test_votes_match_bruteforce_oracle(shape, seed):
    from dusty_gan_b200.utils.metrics import jsd as J
    x = lidar_like_clouds(shape[0], shape[1], 900 + seed, dropped=0.2, near=0.2) / np.float32(2.0)
    x[0, :5] = [[0.7, 0.0, 0.0], [0.0, -0.9, 0.1], [0.4, 0.4, 0.4], [np.float32(1 / 54), 0.0, 0.0], [0.0, 0.0, 0.0]]
    counters, touching = J._vote(cuda(x), 28, True)
    oc, ot = oj.vote(x)
    assert np.array_equal(counters.cpu().numpy().astype(np.int64), oc)
    assert np.array_equal(touching.cpu().numpy().astype(np.int64), ot)


def test_other_resolutions_and_empty_input():
    from dusty_gan_b200.utils.metrics import jsd as J
    x = lidar_like_clouds(4, 600, 950) / np.float32(2.0)
    for res in (8, 16, 33):
        c, _ = J._vote(cuda(x), res, True)
        assert np.array_equal(c.cpu().numpy().astype(np.int64), oj.vote(x, res)[0])
    c, t = J._vote(torch.zeros(0, 10, 3, device="cuda"), 28, True)
    assert int(c.sum()) == 0 and int(t.sum()) == 0
