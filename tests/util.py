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
