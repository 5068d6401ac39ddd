"""Fused data-parallel optimiser step over NVLink peer memory (csrc/collective.cu).

`FusedShardedAdam` replaces, for the TopK trainer under data parallelism,
    dist.all_reduce(flat gradient)  ->  grad-norm pass  ->  Adam replicated on every rank
(train_sae.py:448-450 made equal to the single-GPU step on the concatenated batch) by
    reduce-scatter over peer loads + slice norm  ->  barrier  ->  Adam on this rank's 1/G slice with the updated
    values stored straight into every rank's bf16 weight copies / fp32 biases  ->  barrier.

Host side (this file): the flat buffers live in `torch.distributed._symmetric_memory` (peer-mapped over
NVLink / NVSwitch, multicast address when the fabric offers one), the model's parameters and gradients are re-pointed
at views of them, and the cross-rank barriers are the symmetric-memory signal-pad barriers enqueued on the current
stream.  The fp32 master weights are authoritative only inside a rank's own slice between `consolidate()` calls;
`state_dict()` / `consolidate()` gather them (and the optimiser state) back into the reference layout
(`step` / `exp_avg` / `exp_avg_sq` per parameter, train_sae.py:232-248).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import ops
from ._lib import BF16, call

_ALIGN = 32  # elements; keeps every tensor (and every slice boundary) 128-byte aligned


def plan_flat_layout(numels, world):
    """Offsets of the tensors in the flat space, the padded total, and the per-rank slice bounds (pure host logic)."""
    offs, off = [], 0
    for n in numels:
        offs.append(off)
        off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
    per = (off + world - 1) // world
    per = (per + _ALIGN - 1) // _ALIGN * _ALIGN
    total = per * world
    return offs, total, [(r * per, (r + 1) * per) for r in range(world)]


class FusedShardedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr) semantics (betas (0.9, 0.999), eps 1e-8) + clip_grad_norm_(max_grad_norm), fused
    with the gradient exchange.  `named_params`: ordered {name: Parameter}; `weights`: names whose updated values are
    distributed as bf16 copies (the big matrices in bf16 mode) -- the rest is distributed as fp32."""

    def __init__(self, named_params: dict, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=None, group=None,
                 weights=()):
        import torch.distributed._symmetric_memory as symm_mem

        params = list(named_params.values())
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, max_grad_norm=max_grad_norm))
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.names = list(named_params.keys())
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedShardedAdam needs CUDA parameters (no CPU fallback)")
        numels = [p.numel() for p in params]
        self.offsets, self.total, slices = plan_flat_layout(numels, self.world)
        self.lo, self.hi = slices[self.rank]
        self.weight_names = tuple(weights)

        def symm(n, dtype):
            t = symm_mem.empty(n, dtype=dtype, device=dev)
            t.zero_()
            return t, symm_mem.rendezvous(t, self.group)

        self.flat_param, self.h_param = symm(self.total, torch.float32)
        self.flat_grad, self.h_grad = symm(self.total, torch.float32)
        self.flat_shadow, self.h_shadow = (symm(self.total, torch.bfloat16) if self.weight_names else (None, None))
        self.slots, self.h_slots = symm(16, torch.float64)
        # optimiser state: this rank's slice only
        self.exp_avg = torch.zeros(self.hi - self.lo, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros_like(self.exp_avg)
        self.partial = torch.zeros(1, dtype=torch.float64, device=dev)
        self.done_ctr = torch.zeros(1, dtype=torch.int32, device=dev)
        self.step_count = 0
        # re-point parameters / gradients at the flat buffers
        self.shadows = {}
        for name, p, off in zip(self.names, params, self.offsets):
            view = self.flat_param[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_grad[off:off + p.numel()].view_as(p)
            if name in self.weight_names:
                sh = self.flat_shadow[off:off + p.numel()].view_as(p)
                sh.copy_(ops.split_operand(p.data, BF16)[0])
                self.shadows[name] = sh
        self._regions = [(off, off + (n + 7) // 8 * 8, int(name in self.weight_names))
                         for name, off, n in zip(self.names, self.offsets, numels)]

        def ptr_array(handle):
            return (C.c_void_p * self.world)(*[int(x) for x in handle.buffer_ptrs])

        self._pg, self._pp, self._ps = ptr_array(self.h_grad), ptr_array(self.h_param), ptr_array(self.h_slots)
        self._psh = ptr_array(self.h_shadow) if self.h_shadow is not None else None
        use_mc = bool(int(self.h_grad.multicast_ptr or 0)) and self.world > 1
        self.multicast = use_mc
        self._mc_grad = C.c_void_p(int(self.h_grad.multicast_ptr)) if use_mc else None
        self._mc_param = C.c_void_p(int(self.h_param.multicast_ptr)) if use_mc else None
        self._mc_shadow = (C.c_void_p(int(self.h_shadow.multicast_ptr))
                           if use_mc and self.h_shadow is not None else None)
        rb = (C.c_int64 * len(self._regions))(*[r[0] for r in self._regions])
        re_ = (C.c_int64 * len(self._regions))(*[r[1] for r in self._regions])
        rw = (C.c_int32 * len(self._regions))(*[r[2] for r in self._regions])
        self._reg = (rb, re_, rw)
        self.stale = False
        self.h_grad.barrier(channel=0)  # every rank's buffers are initialised before anyone reads a peer's

    # ---------------------------------------------------------------------------------------------- the fused step
    @torch.no_grad()
    def step(self, closure=None):
        """Call after backward on every rank (the rank-local gradients sit in `flat_grad`)."""
        g = self.param_groups[0]
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.step_count += 1
        self.h_grad.barrier(channel=0)  # all ranks' gradients are complete
        call("freud_dp_reduce_scatter", self._pg, self._mc_grad, self.world, self.rank, self.lo, self.hi,
             C.c_void_p(self.partial.data_ptr()), C.c_void_p(self.done_ctr.data_ptr()), self._ps, stream)
        self.h_grad.barrier(channel=1)  # all partial norms posted; nobody still reads this rank's gradients
        mg = g.get("max_grad_norm")
        call("freud_dp_adam_allgather", self._pp, self._psh, self._mc_param, self._mc_shadow,
             C.c_void_p(self.flat_grad.data_ptr()), C.c_void_p(self.exp_avg.data_ptr()),
             C.c_void_p(self.exp_avg_sq.data_ptr()), self.world, self.rank, self.lo, self.hi, self._reg[0],
             self._reg[1], self._reg[2], len(self._regions), float(g["lr"]), g["betas"][0], g["betas"][1], g["eps"],
             self.step_count, C.c_void_p(self.slots.data_ptr()), float(mg or 0.0), int(mg is not None), stream)
        self.h_grad.barrier(channel=2)  # updated weights / biases have landed everywhere
        self.stale = self.world > 1 and bool(self.weight_names)

    def grad_sumsq(self):
        """Global sum of squared (summed) gradients of the LAST step: device double scalar."""
        return self.slots[: self.world].sum()

    # ---------------------------------------------------------------------------------------------- consolidation
    @torch.no_grad()
    def consolidate(self):
        """All-gather the fp32 master slices so that every rank holds the full, current fp32 parameters."""
        if self.world > 1 and self.stale:
            mine = self.flat_param[self.lo:self.hi].clone()
            dist.all_gather_into_tensor(self.flat_param, mine, group=self.group)
        self.stale = False

    def _gather_state(self, t):
        full = torch.empty(self.total, dtype=t.dtype, device=t.device)
        if self.world > 1:
            dist.all_gather_into_tensor(full, t.contiguous(), group=self.group)
        else:
            full.copy_(t)
        return full

    def state_dict(self):
        """Reference layout: {"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]} (collective)."""
        m, v = self._gather_state(self.exp_avg), self._gather_state(self.exp_avg_sq)
        state = {}
        for i, (p, off) in enumerate(zip(self.param_groups[0]["params"], self.offsets)):
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": m[off:off + p.numel()].view_as(p).clone(),
                        "exp_avg_sq": v[off:off + p.numel()].view_as(p).clone()}
        groups = [{k: v_ for k, v_ in g.items() if k != "params"} | {"params": list(range(len(g["params"])))}
                  for g in self.param_groups]
        return {"state": state, "param_groups": groups}

    def load_state_dict(self, sd):
        m = torch.zeros(self.total, dtype=torch.float32, device=self.exp_avg.device)
        v = torch.zeros_like(m)
        for i, (p, off) in enumerate(zip(self.param_groups[0]["params"], self.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            m[off:off + p.numel()] = st["exp_avg"].reshape(-1).to(m)
            v[off:off + p.numel()] = st["exp_avg_sq"].reshape(-1).to(v)
            self.step_count = int(st["step"])
        self.exp_avg.copy_(m[self.lo:self.hi])
        self.exp_avg_sq.copy_(v[self.lo:self.hi])
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            for k, val in sg.items():
                if k != "params":
                    g[k] = val
