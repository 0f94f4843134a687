"""A bounded slice of the randomised parity sweeps (tests/stress_chamfer_matrix.py, tests/stress_chamfer_batch.py) for the
regular GPU run: random shapes around every dispatch boundary of the Chamfer front ends, random cloud kinds (LiDAR-like with
dropped points, far-away clusters, duplicates, lines, lattices with exact ties, mostly zeros), every entry / distance / arg-min
against the CUDA-rounding oracle."""
import numpy as np
import pytest
import torch

from oracle import native
from stress_chamfer_matrix import ULP, make
from test_gpu_chamfer import run_forward

pytestmark = pytest.mark.gpu

MATRIX_SIZES = [1, 31, 33, 64, 255, 256, 257, 511, 513, 1000, 1024, 2016, 2047, 2048, 2049, 2111, 4096, 4097, 8000]
BATCH_SIZES = [1, 100, 2048, 4096, 4097, 8191, 8193, 16384, 20000, 32767, 32768]


@pytest.mark.parametrize("seed", range(6))
def test_matrix_front_end_random_shapes(seed):
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix
    rng = np.random.default_rng(1000 + seed)
    for _ in range(6):
        pa, pb = int(rng.choice(MATRIX_SIZES)), int(rng.choice(MATRIX_SIZES))
        a, b = make(rng, int(rng.integers(1, 4)), pa), make(rng, int(rng.integers(1, 4)), pb)
        O = native.pairwise_cd(a, b, rounding="cuda")
        for merge in (None, True, False):
            M = chamfer_matrix(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), merge_origin=merge).cpu().numpy()
            assert np.all(np.abs(M.astype(np.float64) - O) <= ULP * np.abs(O)), (pa, pb, merge, np.abs(M - O).max())


@pytest.mark.parametrize("seed", range(4))
def test_batch_front_end_random_shapes(seed):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(4):
        n, m = int(rng.choice(BATCH_SIZES)), int(rng.choice(BATCH_SIZES))
        b = int(rng.integers(1, 3))
        x, y = make(rng, b, n), make(rng, b, m)
        d1, d2, i1, i2 = run_forward(x, y)
        o1, o2, j1, j2 = native.chamfer_forward(x, y, rounding="cuda")
        assert np.array_equal(d1, o1) and np.array_equal(d2, o2), (n, m)
        assert np.array_equal(i1, j1) and np.array_equal(i2, j2), (n, m)
