import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    meta = json.loads(str(z["meta"])) if "meta" in z.files else {}
    return z, meta


def t(a, dtype=None):
    out = torch.from_numpy(np.array(a))
    return out.to(dtype) if dtype is not None else out


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): error relative to the tensor's scale."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def sets_equal_rows(idx_a, idx_b):
    a = torch.sort(torch.as_tensor(idx_a).reshape(-1, idx_a.shape[-1]).long(), dim=-1).values
    b = torch.sort(torch.as_tensor(idx_b).reshape(-1, idx_b.shape[-1]).long(), dim=-1).values
    return (a == b).all(-1)


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 (aggregate relative error; complements rel_err's worst element)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def selection_report(idx_cuda, ref_pre, ref_idx, k, tie_tol):
    """Compare per-row selected SETS with the oracle's.  Returns (same [N] bool, flip_rate, worst_gap) where
    worst_gap is, over the rows that differ, the largest relative distance between a value one side selected and
    the other did not and the row's k-th largest pre-activation: a flip is legitimate only when the disputed values tie with
    the k-th value within GEMM rounding (`tie_tol`)."""
    n = ref_pre.shape[-1]
    pre = ref_pre.reshape(-1, n)
    same = sets_equal_rows(idx_cuda, ref_idx.reshape(-1, k))
    bad = (~same).nonzero().flatten()
    worst = 0.0
    for r in bad.tolist():
        a = set(torch.as_tensor(idx_cuda)[r].tolist())
        b = set(ref_idx.reshape(-1, k)[r].tolist())
        kth = float(torch.sort(pre[r], descending=True).values[k - 1])
        for j in a ^ b:
            worst = max(worst, abs(float(pre[r, j]) - kth) / max(abs(kth), 1e-3))
    flip_rate = float((~same).float().mean())
    assert worst <= tie_tol, f"a row selects a value {worst:.2e} (relative) away from its k-th largest: not a tie"
    return same, flip_rate, worst
