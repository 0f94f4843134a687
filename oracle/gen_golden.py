"""Generate golden vectors by running the reference's OWN Python/C++ code on CPU -- TEST INFRASTRUCTURE.

Run in the build container (needs /root/reference and oracle/_ref built by oracle/build_ref.py):

    python oracle/gen_golden.py

Imports models/dusty.py, utils/lidar.py, utils/metrics/cov_mmd_1nna.py and
utils/metrics/distance/cd/chamfer_distance.py from /root/reference unmodified. Off-path imports that
are not installed here (matplotlib, kornia, omegaconf) and the EMD extension (does not compile on
torch >= 2) are replaced by empty stub modules -- none is touched by the functions called. The
reference's JIT ``load(name="cd")`` is redirected to the already compiled oracle/_ref/dustyref_cd
(same sources, same flags). Outputs: tests/golden/*.npz (a few hundred KB in total).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("DUSTY_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)


def import_reference():
    for name in ("matplotlib", "matplotlib.cm", "matplotlib.colors", "kornia", "omegaconf"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["omegaconf"].OmegaConf = object
    for name in ("utils.metrics.distance.emd", "utils.metrics.distance.emd.earth_mover_distance"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["utils.metrics.distance.emd.earth_mover_distance"].earth_mover_distance = None
    sys.modules["utils.metrics.distance.emd.earth_mover_distance"].__all__ = ["earth_mover_distance"]
    from oracle import refload
    import torch.utils.cpp_extension as ext
    real_load = ext.load

    def load(name, sources, **kw):
        if name == "cd":
            return refload.load("dustyref_cd")
        return real_load(name=name, sources=sources, **kw)

    ext.load = load
    sys.path.insert(0, REF)
    import models.dusty as ref_dusty
    import utils.lidar as ref_lidar
    import utils.metrics.cov_mmd_1nna as ref_metrics
    # `from .cd.chamfer_distance import *` re-exports the extension handle `cd`, shadowing the sub-package
    ref_cd = sys.modules["utils.metrics.distance.cd.chamfer_distance"]
    import utils as ref_utils
    import utils.metrics.jsd as ref_jsd
    return ref_dusty, ref_lidar, ref_metrics, ref_cd, ref_utils, ref_jsd


def hdl64e_angles():
    elev = torch.linspace(np.deg2rad(2.0), np.deg2rad(-24.8), 64, dtype=torch.float64)
    azim = torch.linspace(np.pi, -np.pi, 2049, dtype=torch.float64)[:-1]
    return torch.stack([elev[:, None].expand(64, 2048), azim[None, :].expand(64, 2048)]).float().contiguous()


def np_(t):
    return t.detach().cpu().numpy()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_grad_enabled(False)
    torch.set_num_threads(1)
    ref_dusty, ref_lidar, ref_metrics, ref_cd, ref_utils, ref_jsd = import_reference()
    B, H, W = 3, 16, 64
    g = torch.Generator().manual_seed(1234)

    # ---------------- point-drop heads ----------------
    def head_case(kind, training):
        depth = torch.tanh(torch.randn(B, 1, H, W, generator=g))
        C = 1 if kind == 1 else 2
        conf = torch.randn(B, C, H, W, generator=g) * 2.0
        # a few razor-edge logits: exactly zero, +-tiny, +-huge
        conf.view(-1)[:8] = torch.tensor([0.0, 1e-8, -1e-8, 1e-4, -1e-4, 40.0, -40.0, 90.0])
        depth.view(-1)[:4] = torch.tensor([-1.0, 1.0, 0.0, -0.99999994])
        cls = ref_dusty.DUSty1 if kind == 1 else ref_dusty.DUSty2
        m = cls(backbone=torch.nn.Identity(), tau=1.0, drop_const=-1)
        m.train(training)
        out = {}
        # (a) fresh noise: replay the module's RNG draws to record U1/U2
        seed = 77 + kind + (10 if training else 0)
        torch.manual_seed(seed)
        res = m.maskout({"depth": depth.clone(), "confidence": conf.clone()})
        torch.manual_seed(seed)
        u1p = torch.rand(B, 1, H, W); u2p = torch.rand_like(u1p)
        out.update(depth=np_(depth), confidence=np_(conf), u1_pixel=np_(u1p), u2_pixel=np_(u2p),
                   fresh_mask=np_(res["mask"]), fresh_depth=np_(res["depth"]))
        if kind == 2 and training:
            u1i = torch.rand(B, 1, 1, 1); u2i = torch.rand_like(u1i)
            out.update(u1_image=np_(u1i), u2_image=np_(u2i))
        # (b) fixed noise, as utils.setup's pre-hook freezes it (reference utils/__init__.py:141-149)
        gates = [m.gumbel] if kind == 1 else [m.gumbel_pixel, m.gumbel_image]
        torch.manual_seed(seed + 1)
        for gate, name in zip(gates, ("pixel", "image")):
            logits = conf[:, :1]
            u_seed_state = torch.get_rng_state()
            gate.fixed_noise = gate.logistic_noise(logits)[[0]]
            torch.set_rng_state(u_seed_state)
            shape = (B, 1, H, W) if gate.pixelwise else (B, 1, 1, 1)
            fu1 = torch.rand(*shape); fu2 = torch.rand_like(fu1)
            out["fixed_noise_" + name] = np_(gate.fixed_noise)
            out["fixed_u1_" + name] = np_(fu1[[0]]); out["fixed_u2_" + name] = np_(fu2[[0]])
        res = m.maskout({"depth": depth.clone(), "confidence": conf.clone()})
        out.update(fixed_mask=np_(res["mask"]), fixed_depth=np_(res["depth"]))
        # thresholds other than 0.5 exercise the literal (hard - soft) + soft
        res = m.maskout({"depth": depth.clone(), "confidence": conf.clone()}, threshold=0.3)
        out.update(fixed_mask_t03=np_(res["mask"]), fixed_depth_t03=np_(res["depth"]))
        np.savez_compressed(os.path.join(OUT, f"head_dusty{kind}_{'train' if training else 'eval'}.npz"), **out)

    head_case(1, False)
    head_case(2, False)
    head_case(2, True)

    # ---------------- projection ----------------
    angles = hdl64e_angles()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "angles.pt")
        torch.save(angles, path)
        lidar = ref_lidar.LiDAR(num_ring=H, num_points=W, min_depth=0.9, max_depth=120.0, angle_file=path)
    inv = torch.rand(B, 1, H, W, generator=g)
    inv.view(-1)[:6] = torch.tensor([0.0, 1.0, 1e-9, 5e-9, 2e-8, 0.5])
    inv[1, 0, 3] = 0.0
    out = dict(angles_small=np_(angles[:, :, ::32].contiguous()), angle=np_(lidar.angle), inv=np_(inv),
               xyz_tol1e8=np_(lidar.inv_to_xyz(inv.clone(), tol=1e-8)), xyz_tol0=np_(lidar.inv_to_xyz(inv.clone(), tol=0)),
               xyz_tol8e3=np_(lidar.inv_to_xyz(inv.clone(), tol=0.008)))
    tanh_img = torch.tanh(torch.randn(B, 1, H, W, generator=g) * 1.5)
    tanh_img[0, 0, :2] = -1.0
    inv2 = ref_utils.tanh_to_sigmoid(tanh_img).clamp_(0, 1)
    pts = lidar.inv_to_xyz(inv2, 0).flatten(2).transpose(1, 2)
    out.update(tanh_img=np_(tanh_img), points_eval=np_(pts.contiguous()))
    np.savez_compressed(os.path.join(OUT, "lidar_projection.npz"), **out)

    # ---------------- Chamfer (the reference's CPU twin through its own autograd wrapper) ----------------
    a = torch.randn(3, 200, 3, generator=g) * 0.3
    b = torch.randn(3, 150, 3, generator=g) * 0.3
    b[0, 10] = b[0, 3]                      # duplicate candidates: lowest index must win
    a[1, :5] = 0.0; b[1, :7] = 0.0          # origin points on both sides (dropped pixels)
    a[2, 0] = b[2, 149]                     # an exact hit
    a.requires_grad_(True); b.requires_grad_(True)
    with torch.enable_grad():
        d1, d2 = ref_cd.chamfer_distance(a, b)
        w1 = torch.rand(3, 200, generator=g); w2 = torch.rand(3, 150, generator=g)
        ((d1 * w1).sum() + (d2 * w2).sum()).backward()
    cd = ref_cd.cd
    i1 = torch.zeros(3, 200, dtype=torch.int); i2 = torch.zeros(3, 150, dtype=torch.int)
    cd.forward(a.detach(), b.detach(), torch.zeros(3, 200), torch.zeros(3, 150), i1, i2)
    np.savez_compressed(os.path.join(OUT, "chamfer_cpu.npz"), xyz1=np_(a), xyz2=np_(b), dist1=np_(d1), dist2=np_(d2),
                        idx1=np_(i1), idx2=np_(i2), w1=np_(w1), w2=np_(w2), grad1=np_(a.grad), grad2=np_(b.grad))

    # ---------------- MMD / COV / 1-NNA driver ----------------
    ref = torch.randn(12, 128, 3, generator=g) * 0.25
    gen = torch.randn(10, 128, 3, generator=g) * 0.25 + 0.02
    M_rr = ref_metrics._pairwise_distance(ref, ref, 5, ("cd",), False)["cd"]
    M_rg = ref_metrics._pairwise_distance(ref, gen, 5, ("cd",), False)["cd"]
    M_gg = ref_metrics._pairwise_distance(gen, gen, 5, ("cd",), False)["cd"]
    scores = ref_metrics.compute_cov_mmd_1nna(gen, ref, 5, ("cd",), False)
    keys = sorted(scores)
    # a generated cloud identical to a reference cloud: a zero off-diagonal entry. Only the matrix is
    # recorded for it -- the 1-NN vote then has exact ties between rows of different labels and
    # torch.topk's choice among equals is implementation-defined.
    gen_dup = gen.clone()
    gen_dup[3] = ref[5]
    M_rg_dup = ref_metrics._pairwise_distance(ref, gen_dup, 5, ("cd",), False)["cd"]
    np.savez_compressed(os.path.join(OUT, "metrics_cpu.npz"), ref=np_(ref), gen=np_(gen), M_rr=np_(M_rr), M_rg=np_(M_rg),
                        M_gg=np_(M_gg), gen_dup=np_(gen_dup), M_rg_dup=np_(M_rg_dup), score_keys=np.array(keys),
                        score_values=np.array([scores[k] for k in keys], np.float64))
    # ---------------- JSD (next row 8f-2): the reference's own voting and divergence ----------------
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import lidar_like_clouds
    gen_j = torch.from_numpy(lidar_like_clouds(6, 700, 41, dropped=0.1, near=0.1)) / 2.0       # evaluate_synthesis.py:174-177 halves
    ref_j = torch.from_numpy(lidar_like_clouds(5, 700, 42, dropped=0.1, near=0.1)) / 2.0
    gen_j[0, :4] = torch.tensor([[0.5, 0.0, 0.0], [-0.5, 0.0, 0.0], [0.2887, 0.2887, 0.2887], [0.0185185, 0.0185185, 0.0]])
    ref_j[0, :3] = torch.tensor([[0.6, 0.1, 0.0], [0.0, 0.0, 0.52], [-0.3, 0.41, 0.1]])           # outside the sphere
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ent_g, cnt_g = ref_jsd.entropy_of_occupancy_grid(gen_j, 28, True, 128, False)
        ent_r, cnt_r = ref_jsd.entropy_of_occupancy_grid(ref_j, 28, True, 128, False)
        jsd = ref_jsd.compute_jsd(gen_j, ref_j, verbose=False)
        grid, spacing = ref_jsd.unit_cube_grid_point_cloud(28, True, "cpu")
    np.savez_compressed(os.path.join(OUT, "jsd_cpu.npz"), gen=np_(gen_j), ref=np_(ref_j), counters_gen=np_(cnt_g),
                        counters_ref=np_(cnt_r), entropy_gen=np.float64(ent_g.item()), entropy_ref=np.float64(ent_r.item()),
                        jsd=np.float64(jsd), grid=np_(grid))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
