"""Randomised parity sweep of the batch front end (chamfer_distance: per-point dist AND arg-min, both directions) against the
CUDA-rounding oracle; not collected by pytest: `python tests/stress_chamfer_batch.py [rounds] [seed]` on a B200. Sizes
straddle the sorted path's threshold (4096) and capacity (32768); lattice and duplicate clouds make every tie rule fire."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle import native  # noqa: E402
from stress_chamfer_matrix import make  # noqa: E402
from test_gpu_chamfer import run_forward  # noqa: E402


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    sizes = [1, 100, 2048, 4096, 4097, 5000, 8191, 8192, 8193, 12000, 16384, 20000, 32767, 32768, 32769]
    bad = 0
    for it in range(rounds):
        n, m = int(rng.choice(sizes)), int(rng.choice(sizes))
        b = int(rng.integers(1, 4))
        x, y = make(rng, b, n), make(rng, b, m)
        d1, d2, i1, i2 = run_forward(x, y)
        o1, o2, j1, j2 = native.chamfer_forward(x, y, rounding="cuda")
        ok = np.array_equal(d1, o1) and np.array_equal(d2, o2) and np.array_equal(i1, j1) and np.array_equal(i2, j2)
        if not ok:
            bad += 1
            print("MISMATCH", it, b, n, m, (d1 != o1).sum(), (d2 != o2).sum(), (i1 != j1).sum(), (i2 != j2).sum(), flush=True)
        print("round", it, b, n, m, "ok" if ok else "FAILED", flush=True)
    print("stress done: %d mismatches" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
