"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink / NVSwitch on the B200 box,
gloo in the CPU tests).  The reference is single-device; parity for N ranks is defined as
"N-GPU result == 1-GPU result on the concatenated batch" (SURVEY.md section 8(e)), which needs exactly one real
exchange per step -- the gradient allreduce -- plus three scalar-sized ones (variance, SSE, did_fire)."""
from __future__ import annotations

import torch
import torch.distributed as dist


class DataParallel:
    def __init__(self, group=None, fused: bool = False):
        """fused: exchange the gradients inside the optimiser step (freud_b200.fused_dp.FusedShardedAdam: reduce-scatter
        over NVLink peer loads, Adam on a 1/G slice, updated weights stored to every rank) instead of one NCCL
        all-reduce followed by the replicated update.  TopK trainer with Adam only; needs symmetric memory."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.fused = bool(fused)
        self.group = group
        self.world_size = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._side = None

    def side_stream(self):
        """CUDA stream for the scalar-sized exchanges (variance, SSE, fire counts): they and their small torch
        kernels run beside the encoder / backward kernels instead of stalling the main stream for a collective's
        launch latency each.  None without CUDA (gloo tests)."""
        if self._side is None and torch.cuda.is_available():
            self._side = torch.cuda.Stream()
        return self._side

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def global_total_variance(self, tv_local: torch.Tensor, colmean_local: torch.Tensor, b_local: int):
        """sum_b (x_b - mean)^2 over the concatenated batch from per-rank pieces (equal per-rank batch)."""
        mean = colmean_local.clone()
        self.all_reduce_sum(mean)
        mean /= self.world_size
        between = (colmean_local - mean).double().pow(2).sum() * b_local
        tv = tv_local + between
        return self.all_reduce_sum(tv)

    def all_reduce_grads(self, grads, flat=None):
        """One fp32 gradient allreduce (sum) per step over a single flat bucket: NVSwitch makes collective cost
        latency- rather than link-bound, so one large message beats many small ones (overlapping four row-range
        reductions with the backward was measured: 4.01 vs 3.94 ms per C3 step at 2 GPUs, not adopted).
        `flat`: the buffer the gradients are views of, when the caller laid them out that way (no copies)."""
        if flat is not None:
            self.all_reduce_sum(flat)
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        self.all_reduce_sum(flat)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def all_reduce_max_(self, t: torch.Tensor):
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t
