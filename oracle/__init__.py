"""CPU oracle for the DUSty-GAN hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package, and only as the checker. The product (``dusty-gan_b200/``) never does.

Parity status (DESIGN.md "Oracle"):
  head + projection   ``oracle.head_projection`` -- the reference's element-wise ATen chains restated
                      op by op; pinned against the reference's own modules (tests/golden/head_*.npz,
                      made by oracle/gen_golden.py from /root/reference).
  Chamfer / metrics   ``oracle.native`` (C) + ``oracle.metrics`` -- pinned against the reference's own
                      cd.forward and compute_cov_mmd_1nna outputs (tests/golden/chamfer_*.npz,
                      metrics_*.npz) and against its CUDA kernel run on a B200 (tests/golden/gpu_*.npz).
  FPS                 ``oracle.native.fps`` -- the reference has no CPU FPS; pinned against the
                      reference's own CUDA kernel run on a B200 through oracle/_ref
                      (tests/golden/gpu_fps_*.npz, made by oracle/gen_golden_gpu.py).
"""
