"""numpy restatement of the MMD / COV / 1-NNA driver -- TEST INFRASTRUCTURE ONLY.

Follows reference utils/metrics/cov_mmd_1nna.py: _pairwise_distance (:24-51) through oracle.native,
_compute_cov_mmd (:54-65), _compute_nna with k=1 (:68-106), compute_cov_mmd_1nna (:109-139).
"""
import numpy as np

from . import native


def compute_cov_mmd(M_rg):
    N_ref, N_gen = M_rg.shape
    min_idx_gen = M_rg.argmin(axis=0)
    mmd_gen = M_rg.min(axis=0)
    mmd_ref = M_rg.min(axis=1)
    return {"mmd": float(mmd_ref.astype(np.float32).mean(dtype=np.float64).astype(np.float32)),
            "mmd-sample": float(mmd_gen.astype(np.float32).mean(dtype=np.float64).astype(np.float32)),
            "cov": float(len(np.unique(min_idx_gen))) / float(N_ref)}


def compute_nna(M_rr, M_rg, M_gg):
    N_ref, N_gen = M_rg.shape
    label = np.concatenate([np.ones(N_ref, np.float32), np.zeros(N_gen, np.float32)])
    M = np.block([[M_rr, M_rg], [M_rg.T, M_gg]]).astype(np.float32)
    M = M + np.diag(np.full(N_ref + N_gen, np.inf, np.float32))
    idx = M.argmin(axis=0)                      # topk(k=1, dim=0, largest=False)
    pred = label[idx]                           # count/k >= 0.5 with k = 1
    s = {"tp": float((pred * label).sum()), "fp": float((pred * (1 - label)).sum()),
         "fn": float(((1 - pred) * label).sum()), "tn": float(((1 - pred) * (1 - label)).sum())}
    s.update({"precision": s["tp"] / (s["tp"] + s["fp"] + 1e-10),
              "recall": s["tp"] / (s["tp"] + s["fn"] + 1e-10),
              "accuracy_t": s["tp"] / (s["tp"] + s["fn"] + 1e-10),
              "accuracy_f": s["tn"] / (s["tn"] + s["fp"] + 1e-10),
              "accuracy": float((label == pred).astype(np.float32).mean(dtype=np.float32))})
    return s


def pairwise_matrices(pcs_gen, pcs_ref, rounding="cuda"):
    M_rr = native.pairwise_cd(pcs_ref, None, rounding=rounding)
    M_rg = native.pairwise_cd(pcs_ref, pcs_gen, rounding=rounding)
    M_gg = native.pairwise_cd(pcs_gen, None, rounding=rounding)
    return M_rr, M_rg, M_gg


def scores_from_matrices(M_rr, M_rg, M_gg):
    results = {}
    for k, v in compute_cov_mmd(M_rg).items():
        results["{}-{}".format(k, "cd")] = v
    for k, v in compute_nna(M_rr, M_rg, M_gg).items():
        results["1-nn-{}-{}".format(k, "cd")] = v
    return results


def compute_cov_mmd_1nna(pcs_gen, pcs_ref, rounding="cuda"):
    return scores_from_matrices(*pairwise_matrices(pcs_gen, pcs_ref, rounding))
