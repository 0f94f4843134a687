"""Shape sweeps around every dispatch boundary of the kernels (threads per CTA, rows per thread, tile and
chunk edges, merged origins), each against the oracle on seeded inputs."""
import numpy as np
import pytest
import torch

from helpers import lidar_like_clouds, rel_err, sampled_clouds
from oracle import native

pytestmark = pytest.mark.gpu
REL_TOL = 1e-5


def cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("pa,pb", [(1, 1), (1, 40), (31, 33), (32, 32), (127, 128), (128, 129), (129, 64), (255, 256), (256, 257),
                                   (511, 512), (513, 200), (1023, 1024), (1024, 1025), (1025, 2047), (2049, 100), (4097, 4096)])
def test_matrix_point_count_boundaries(pa, pb):
    """The dense matrix front end picks (threads, rows per thread) from the larger cloud size: 32/64/128/256
    threads at 128/256/512/1024 points; clouds above 4096 points take the merged-origin instantiation."""
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix
    a = lidar_like_clouds(3, pa, 1000 + pa, dropped=0.2); b = lidar_like_clouds(2, pb, 2000 + pb, dropped=0.2)
    M = chamfer_matrix(cuda(a), cuda(b)).cpu().numpy()
    O = native.pairwise_cd(a, b, rounding="cuda")
    assert np.all(np.abs(M - O) <= 2.0 ** -23 * np.abs(O)), (M, O)          # per entry, one ulp
    S = chamfer_matrix(cuda(a)).cpu().numpy()
    Os = native.pairwise_cd(a, None, rounding="cuda")
    assert np.array_equal(S, S.T) and np.all(np.diag(S) == 0)
    assert np.all(np.abs(S - Os) <= 2.0 ** -23 * np.abs(Os))


@pytest.mark.parametrize("b,n,m", [(1, 1, 1), (1, 33, 7), (2, 256, 257), (3, 513, 100), (5, 1024, 1025), (70, 64, 64),
                                   (2, 2049, 2048), (1, 4100, 9000), (40, 300, 300)])
def test_batch_forward_shapes(b, n, m):
    """Batch front end: rows per thread follow the cloud size and the number of CTAs; every distance bit-equal
    to the reference kernel's restatement."""
    from dusty_gan_b200.utils.metrics.distance import chamfer_distance
    a = sampled_clouds(b, n, 31 + n); c = sampled_clouds(b, m, 47 + m)
    d1, d2 = chamfer_distance(cuda(a), cuda(c))
    o1, o2, _, _ = native.chamfer_forward(a, c, rounding="cuda")
    assert np.array_equal(d1.cpu().numpy(), o1) and np.array_equal(d2.cpu().numpy(), o2)


@pytest.mark.parametrize("n,m,clouds", [(1, 1, 2), (5, 9, 3), (127, 64, 2), (128, 128, 3), (129, 40, 2), (4097, 300, 2), (20000, 700, 150),
                                        (32768, 33, 2), (33000, 50, 2)])
def test_fps_shapes(n, m, clouds):
    """Single-cloud-per-SM, six-clouds-per-SM (clouds > 148) and flat (n > 32768) variants; m may exceed the
    number of eligible (or of all) points, in which case the tie rule decides the repeats."""
    from dusty_gan_b200.utils.sampling.fps import furthest_point_sampling
    x = lidar_like_clouds(clouds, n, 77 + n)
    idx = furthest_point_sampling(cuda(x), m).cpu().numpy()
    want = native.fps(x[: min(clouds, 6)], m)
    assert np.array_equal(idx[: min(clouds, 6)], want)
    if clouds > 6:          # the remaining clouds against themselves through the single-cloud variant
        again = furthest_point_sampling(cuda(x[6:40]), m).cpu().numpy()
        assert np.array_equal(idx[6:40], again)


@pytest.mark.parametrize("hs,ws,c,h,w", [(1, 4, 3, 1, 4), (64, 2048, 4, 64, 256), (64, 2048, 4, 32, 1024), (7, 50, 6, 9, 44), (3, 1000, 4, 3, 128)])
def test_scan_preprocess_shapes(hs, ws, c, h, w):
    from dusty_gan_b200.datasets import preprocess_scans
    from oracle import real_data as rd
    scans = rd.synthetic_scans(2, seed=hs + ws, hs=hs, ws=ws, channels=c)
    out = preprocess_scans(cuda(scans), (h, w), want=("xyz", "depth", "mask", "inv", "points"))
    items = [rd.dataset_item(s, (h, w)) for s in scans]
    raw = {k: torch.stack([it[k] for it in items]) for k in items[0]}
    inv, mask, points = rd.preprocess_reals(raw, device="cuda")
    assert torch.equal(out["xyz"], raw["xyz"].cuda()) and torch.equal(out["depth"], raw["depth"].cuda())
    assert torch.equal(out["mask"], mask) and torch.equal(out["inv"], inv) and torch.equal(out["points"], points)
