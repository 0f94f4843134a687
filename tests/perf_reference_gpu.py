"""Informative only: the reference's OWN CUDA path on this B200, from oracle/_ref,
timed next to ours on the same inputs. Not a test and not part of bench.py.

    gpurun -- python tests/perf_reference_gpu.py       (prints one JSON line, writes gpurun_out/ref_gpu.json)
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from oracle import refload  # noqa: E402


def timed(fn, iters=5, warmup=2):
    return statistics.median(bench.time_events(fn, iters, warmup)) * 1e-3


def main():
    dev = torch.device("cuda:0")
    fps = refload.load("dustyref_fps"); cd = refload.load("dustyref_cd")
    from dusty_gan_b200 import pipeline
    from dusty_gan_b200.utils.sampling.fps import downsample_point_clouds
    from dusty_gan_b200.utils.metrics.cov_mmd_1nna import chamfer_matrix
    lidar = bench.make_lidar(dev); head = bench.make_head(1, dev)
    depth, conf = bench.backbone_like(148, 1, 12, dev)
    pts = pipeline.maskout_and_project(head, {"depth": depth, "confidence": conf}, lidar, tol=0.0)["points"]
    out = {}
    # FPS: reference wrapper = 2 copies + fps kernel + gather kernel (fps/furthest_point_sampling.py:84-93)

    def ref_downsample():
        xyz = pts.contiguous(); src = xyz.transpose(1, 2).contiguous()
        return fps.gather_points(src, fps.furthest_point_sampling(xyz, 2048)).transpose(1, 2)
    out["fps_ref_clouds_per_s"] = 148 / timed(ref_downsample)
    out["fps_ours_clouds_per_s"] = 148 / timed(lambda: downsample_point_clouds(pts, 2048))
    assert torch.equal(ref_downsample().contiguous(), downsample_point_clouds(pts, 2048))
    # Chamfer matrix: the reference's Python loop (cov_mmd_1nna.py:24-51) on 64 x 512 entries
    clouds = downsample_point_clouds(pts, 2048)
    big = clouds.repeat(4, 1, 1)[:512].contiguous()

    def ref_rows(rows=64):
        M = torch.zeros(rows, 512, device=dev)
        for i in range(rows):
            b1 = clouds[[i]].expand(512, -1, -1).contiguous()
            d1 = torch.zeros(512, 2048, device=dev); d2 = torch.zeros(512, 2048, device=dev)
            i1 = torch.zeros(512, 2048, dtype=torch.int, device=dev); i2 = torch.zeros(512, 2048, dtype=torch.int, device=dev)
            cd.forward_cuda(b1, big, d1, d2, i1, i2)
            M[i] = d1.mean(1) + d2.mean(1)
        return M
    t_ref = timed(ref_rows, 2, 1)
    t_ours = timed(lambda: chamfer_matrix(clouds[:64].contiguous(), big), 3, 1)
    out["chamfer_ref_entries_per_s"] = 64 * 512 / t_ref
    out["chamfer_ours_entries_per_s_same_shape"] = 64 * 512 / t_ours
    err = (ref_rows() - chamfer_matrix(clouds[:64].contiguous(), big)).abs().max().item()
    out["chamfer_max_abs_diff"] = err
    # batch front end (per-point dist + idx, forward and backward) as evaluate_reconstruction.py:124-131 and
    # demo.py:510-515 use it: 32 pairs of sampled clouds, and 8 pairs of un-sampled 32768-point clouds
    from dusty_gan_b200.utils.metrics.distance import chamfer_distance
    for tag, a, b in (("2048", clouds[:32].contiguous(), clouds[32:64].contiguous()),
                      ("32768", pts[:8].contiguous(), pts[8:16].contiguous())):
        B, n, _ = a.shape

        def ref_fwd():
            d1 = torch.zeros(B, n, device=dev); d2 = torch.zeros(B, n, device=dev)
            i1 = torch.zeros(B, n, dtype=torch.int, device=dev); i2 = torch.zeros(B, n, dtype=torch.int, device=dev)
            cd.forward_cuda(a, b, d1, d2, i1, i2)
            return d1, d2, i1, i2
        d1, d2, i1, i2 = ref_fwd()
        g = torch.ones(B, n, device=dev)

        def ref_bwd():
            g1 = torch.zeros_like(a); g2 = torch.zeros_like(b)
            cd.backward_cuda(a, b, g1, g2, g, g, i1, i2)
            return g1, g2
        out[f"batch{tag}_ref_fwd_pairs_per_s"] = B / timed(ref_fwd)
        out[f"batch{tag}_ours_fwd_pairs_per_s"] = B / timed(lambda: chamfer_distance(a, b))
        o1, o2 = chamfer_distance(a, b)
        out[f"batch{tag}_fwd_bit_equal_frac"] = float(((o1 == d1).float().mean() + (o2 == d2).float().mean()) / 2)
        out[f"batch{tag}_ref_bwd_pairs_per_s"] = B / timed(ref_bwd)
        ar = a.clone().requires_grad_(True); br = b.clone().requires_grad_(True)
        with torch.enable_grad():
            q1, q2 = chamfer_distance(ar, br)
            loss = q1.sum() + q2.sum()
        out[f"batch{tag}_ours_bwd_pairs_per_s"] = B / timed(lambda: torch.autograd.grad(loss, (ar, br), retain_graph=True))
        # like for like with ref_bwd: the C-ABI call alone (the line above also pays autograd's graph walk)
        from dusty_gan_b200.utils.metrics.distance.cd.chamfer_distance import _scatter_grads
        out[f"batch{tag}_ours_bwd_abi_pairs_per_s"] = B / timed(lambda: _scatter_grads(a, b, g, g, i1, i2))
        r1, r2 = ref_bwd(); m1, m2 = _scatter_grads(a, b, g, g, i1, i2)
        out[f"batch{tag}_bwd_max_abs_diff"] = float(max((r1 - m1).abs().max(), (r2 - m2).abs().max()))
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ref_gpu.json"), "w") as fh:
        json.dump(out, fh)


if __name__ == "__main__":
    main()
