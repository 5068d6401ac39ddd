"""Fused optimisers and gradient clipping over the C ABI (train_sae.py:374-381,449-450).

FusedAdam / FusedRAdam are torch.optim.Optimizer subclasses with the SAME state layout as torch.optim.Adam /
RAdam (`step`, `exp_avg`, `exp_avg_sq`), so `optimizer.state_dict()` checkpoints (train_sae.py:232-248) round-trip
between this build and the reference, and torch's LR schedulers attach unchanged.  `clip_grad_norm_` mirrors
torch.nn.utils.clip_grad_norm_ (total L2 norm, coef = clamp(max_norm/(norm+1e-6), max=1), always multiplies).
"""
from __future__ import annotations

import torch

from . import ops


def _collect(params):
    ps = [p for p in params if p.grad is not None]
    for p in ps:
        if not p.is_cuda:
            raise RuntimeError("freud_b200 optimisers run on CUDA parameters only (no CPU fallback)")
    return ps


def clip_grad_norm_(parameters, max_norm: float) -> torch.Tensor:
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    ps = _collect(list(parameters))
    if not ps:
        return torch.tensor(0.0)
    total = None
    sumsq = None
    for i in range(0, len(ps), ops._lib.MAX_TENSORS):
        chunk = ps[i:i + ops._lib.MAX_TENSORS]
        tl = ops.make_tensor_list([p.data for p in chunk], [p.grad for p in chunk])
        s = ops.grad_sumsq(tl, chunk[0].device)
        sumsq = s if sumsq is None else sumsq + s
    for i in range(0, len(ps), ops._lib.MAX_TENSORS):
        chunk = ps[i:i + ops._lib.MAX_TENSORS]
        tl = ops.make_tensor_list([p.data for p in chunk], [p.grad for p in chunk])
        total = ops.clip_grads(tl, sumsq, float(max_norm))
    return total


class _FusedBase(torch.optim.Optimizer):
    def set_shadows(self, mapping):
        """mapping: {parameter: bf16 tensor of the same shape}.  The update kernel then also writes the rounded new
        value of that parameter there (2 B/param more traffic instead of a separate 6 B/param conversion pass)."""
        self._shadows = {id(p): t for p, t in mapping.items()}

    def _prepare(self, group):
        ps = _collect(group["params"])
        for p in ps:
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return ps

    def _lists(self, ps):
        for i in range(0, len(ps), ops._lib.MAX_TENSORS):
            chunk = ps[i:i + ops._lib.MAX_TENSORS]
            sh = getattr(self, "_shadows", None)
            yield chunk, ops.make_tensor_list([p.data for p in chunk], [p.grad for p in chunk],
                                              [self.state[p]["exp_avg"] for p in chunk],
                                              [self.state[p]["exp_avg_sq"] for p in chunk],
                                              [sh.get(id(p)) for p in chunk] if sh else None)


class FusedAdam(_FusedBase):
    """torch.optim.Adam(params, lr) semantics (betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, max_grad_norm=max_grad_norm))

    @torch.no_grad()
    def step(self, closure=None, grad_sumsq=None):
        """grad_sumsq: optional device double[1] with the global sum of squared gradients; together with the
        group's max_grad_norm the clip coefficient is applied inside the update kernel (fused clip + Adam)."""
        for group in self.param_groups:
            ps = self._prepare(group)
            if not ps:
                continue
            mg = group.get("max_grad_norm")
            sumsq = grad_sumsq
            if mg is not None and sumsq is None:
                sumsq = None
                for _, tl in self._lists(ps):
                    s = ops.grad_sumsq(tl, ps[0].device)
                    sumsq = s if sumsq is None else sumsq + s
            for chunk, tl in self._lists(ps):
                for p in chunk:
                    self.state[p]["step"] += 1
                step = int(self.state[chunk[0]]["step"].item())
                ops.adam_step(tl, float(group["lr"]), group["betas"][0], group["betas"][1], group["eps"], step,
                              sumsq if mg is not None else None, float(mg or 0.0))


class FusedRAdam(_FusedBase):
    """torch.optim.RAdam(params, eps, lr, weight_decay) semantics (non-decoupled weight decay)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      max_grad_norm=max_grad_norm))

    @torch.no_grad()
    def step(self, closure=None, grad_sumsq=None):
        for group in self.param_groups:
            ps = self._prepare(group)
            if not ps:
                continue
            mg = group.get("max_grad_norm")
            sumsq = grad_sumsq
            if mg is not None and sumsq is None:
                for _, tl in self._lists(ps):
                    s = ops.grad_sumsq(tl, ps[0].device)
                    sumsq = s if sumsq is None else sumsq + s
            for chunk, tl in self._lists(ps):
                for p in chunk:
                    self.state[p]["step"] += 1
                step = int(self.state[chunk[0]]["step"].item())
                ops.radam_step(tl, float(group["lr"]), group["betas"][0], group["betas"][1], group["eps"],
                               group["weight_decay"], step, sumsq if mg is not None else None, float(mg or 0.0))
