"""The SAE train step -- body of the reference loop, src/scripts/train_sae.py:421-453 -- on the CUDA kernels.

`SAETrainer.step(activations)` = dead-mask, forward, loss, backward, global-norm clip, Adam/RAdam update,
LR-scheduler step and dead-latent bookkeeping, with the same hyper-parameter schema as `train(**config)`
(configs/train/*.json).  Both paths drive the kernels directly (no autograd graph: the train loop's loss is fixed);
the modules keep their autograd functions for callers that build their own loss.  With `dp` (freud_b200.parallel.DataParallel) the step equals the
single-GPU step on the batch concatenated over ranks.
"""
from __future__ import annotations

import torch
from torch.optim.lr_scheduler import CosineAnnealingLR, LambdaLR

from . import ops, topk_engine
from ._lib import BF16, FP32
from .models.l1autoencoder import L1AutoEncoder, l1_backward, l1_forward
from .models.topkautoencoder import TopKAutoEncoder
from .optim import FusedAdam, FusedRAdam

_TOPK_KEYS = ("encoder.weight", "encoder.bias", "W_dec", "b_dec")


def linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps):
    """Same lambda as transformers.get_linear_schedule_with_warmup (transformers/optimization.py:101-131),
    which train_sae.py:386-390 uses; restated to avoid importing transformers on the hot path."""

    def lr_lambda(current_step: int):
        if current_step < num_warmup_steps:
            return float(current_step) / float(max(1, num_warmup_steps))
        return max(0.0, float(num_training_steps - current_step) / float(max(1, num_training_steps - num_warmup_steps)))

    return LambdaLR(optimizer, lr_lambda)


def build_optimizer(model, optimizer: str, lr: float, weight_decay: float, clip_thresh):
    if optimizer == "radam":
        return FusedRAdam(model.parameters(), eps=1e-5, lr=lr, weight_decay=weight_decay, max_grad_norm=clip_thresh)
    if optimizer == "adam":
        return FusedAdam(model.parameters(), lr=lr, max_grad_norm=clip_thresh)
    raise ValueError(f"Invalid optimizer: {optimizer}, must be 'radam' or 'adam'")


def build_scheduler(opt, scheduler: str, scheduler_params: dict, steps: int):
    if scheduler == "cosine":
        return CosineAnnealingLR(opt, T_max=steps, eta_min=0)
    if scheduler == "linear":
        return linear_schedule_with_warmup(opt, scheduler_params["num_warmup_steps"], steps)
    raise ValueError(f"Invalid scheduler: {scheduler}, must be 'cosine' or 'linear'")


class SAETrainer:
    def __init__(self, model, *, lr, steps, clip_thresh=1.0, weight_decay=0.0, optimizer="adam", scheduler="linear",
                 scheduler_params=None, dead_feature_threshold=None, precision="bf16", dp=None,
                 materialize_outputs=True):
        """materialize_outputs (L1 SAE, bf16 mode): False skips the fp32 sae_out / latent copies the train loop does
        not need (step() then returns None for them)."""
        self.materialize_outputs = materialize_outputs
        if not next(model.parameters()).is_cuda:
            raise RuntimeError("SAETrainer needs the model on a CUDA device (no CPU fallback)")
        self.model = model
        self.is_topk = isinstance(model, TopKAutoEncoder)
        self.precision = {"bf16": BF16, "fp32": FP32}[precision]
        model.precision = precision
        self.clip_thresh = clip_thresh
        self.dp = dp
        self.fused_dp = bool(dp is not None and getattr(dp, "fused", False) and self.is_topk and optimizer == "adam")
        if self.fused_dp:
            from .fused_dp import FusedShardedAdam

            # model.parameters() order, so that optimizer.state_dict() indexes the parameters exactly like the
            # reference's Adam(model.parameters()) does (checkpoint layout, train_sae.py:232-248)
            named = dict(model.named_parameters())
            self.optimizer = FusedShardedAdam(named, lr=lr, max_grad_norm=clip_thresh, group=dp.group,
                                              weights=("encoder.weight", "W_dec") if precision == "bf16" else ())
            import weakref

            model._freud_sharded = weakref.ref(self.optimizer)  # module forward refuses stale fp32 weights
        else:
            self.optimizer = build_optimizer(model, optimizer, lr, weight_decay, clip_thresh)
        self.scheduler = build_scheduler(self.optimizer, scheduler, scheduler_params or {}, steps)
        self.step_count = 0
        self.tokens_seen = 0
        self.dead_feature_threshold = dead_feature_threshold
        dev = next(model.parameters()).device
        if self.is_topk:
            # train_sae.py:412-415 (recreated as zeros on resume, as upstream)
            self.num_frames_since_fired = torch.zeros(model.n_dict_components, device=dev, dtype=torch.long)
            self.params = {"encoder.weight": model.encoder.weight, "encoder.bias": model.encoder.bias,
                           "W_dec": model.W_dec, "b_dec": model.b_dec}
            if self.fused_dp:  # parameters and gradients already are views of the optimiser's symmetric flat buffers
                self._flat_grad = self.optimizer.flat_grad
                self._shadows, self._shadow_versions = self.optimizer.shadows, None
                self.last_state = None
                self._dead_next = None
                self._dead_pinned = torch.zeros(1, dtype=torch.int64).pin_memory()
                int((self.num_frames_since_fired > (dead_feature_threshold or 0)).sum())
                return
            # gradients are views of one flat fp32 buffer: data parallel reduces it in a single call, no copies
            # (every segment starts on a 128-byte boundary: the float4 paths of the clip / Adam / sparse-gradient
            # kernels need 16-byte-aligned tensors whatever n_dict_components is; the padding stays zero)
            seg = {key: (self.params[key].numel() + 31) // 32 * 32 for key in _TOPK_KEYS}
            flat = torch.zeros(sum(seg.values()), dtype=torch.float32, device=dev)
            off = 0
            for key in _TOPK_KEYS:
                p = self.params[key]
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += seg[key]
            self._flat_grad = flat
            # load torch's lazily-initialised kernels for the dead-mask read-back now, not in the first step that
            # crosses the threshold (a one-off ~60 ms module load otherwise lands inside the training loop)
            int((self.num_frames_since_fired > (dead_feature_threshold or 0)).sum())
            self._shadows, self._shadow_versions = {}, None
        self.last_state = None
        self._dead_next = None  # (mask, event, counter version) prefetched for the next step
        self._dead_pinned = torch.zeros(1, dtype=torch.int64).pin_memory() if dev.type == "cuda" else None

    def _fresh_shadows(self):
        """bf16 copies of W_enc / W_dec for the tensor-core and gather kernels.  Built once; afterwards the fused Adam
        kernel rewrites them together with the fp32 parameters.  In-place torch writes to a parameter
        (load_state_dict, `p.copy_`, `p.mul_` under no_grad) bump its version counter and the module's own
        set_decoder_norm_to_unit_norm bumps `_weights_epoch`: either triggers a rebuild.  Writes through `p.data`
        are invisible to both -- call `invalidate_weight_copies()` after editing weights that way."""
        if self.precision != BF16:
            return None
        if self.fused_dp:
            return self._shadows  # written by the fused optimiser step on every rank
        versions = tuple(self.params[k]._version for k in ("encoder.weight", "W_dec")) + \
            tuple(self.params[k].data_ptr() for k in ("encoder.weight", "W_dec")) + \
            (getattr(self.model, "_weights_epoch", 0),)
        if self._shadow_versions != versions or not self._shadows:
            for k in ("encoder.weight", "W_dec"):
                self._shadows[k] = ops.split_operand(self.params[k].data, BF16)[0]
            self.optimizer.set_shadows({self.params[k]: self._shadows[k] for k in self._shadows})
            self._shadow_versions = versions
        return self._shadows

    def invalidate_weight_copies(self):
        self._shadow_versions = None

    def consolidate(self):
        """Fused data-parallel mode keeps the fp32 master weights current only inside each rank's slice; this
        all-gathers them (collective: call on every rank) before the model is evaluated, saved or edited."""
        if self.fused_dp:
            self.optimizer.consolidate()

    # ------------------------------------------------------------------------------------------ TopK
    def _topk_step(self, x):
        m = self.model
        cfg = m.cfg
        B, T, _ = x.shape
        n_tokens = B * T * (self.dp.world_size if self.dp else 1)
        dead_mask, num_dead = None, None
        thr = self.dead_feature_threshold
        frames = self.num_frames_since_fired
        # a latent can only be dead once more than `thr` frames have been seen in total; before that the mask
        # is provably empty and the reference's per-step `dead_mask.sum()` read-back is skipped
        if thr is not None and self.tokens_seen > thr:
            if self._dead_next is not None and self._dead_next[2] == frames._version:
                # mask and count were produced right behind the previous step's CSC build and copied to pinned memory
                # ~1 ms before that step ended: the read-back returns at once instead of draining the device
                dead_mask, ev, _ = self._dead_next
                ev.synchronize()
                num_dead = int(self._dead_pinned[0])
            else:  # first step past the threshold, or the counters were edited from outside
                dead_mask = frames > thr
        self._dead_next = None
        res, st = topk_engine.topk_forward(x, m.encoder.weight.data, m.encoder.bias.data, m.W_dec.data,
                                           m.b_dec.data, cfg.k, precision=self.precision, dead_mask=dead_mask,
                                           auxk_alpha=float(cfg.auxk_alpha), multi_topk=bool(cfg.multi_topk),
                                           need_grad=True, dp=self.dp, defer_scal=self.dp is not None,
                                           num_dead=num_dead, shadows=self._fresh_shadows())
        # loss = fvu + auxk_loss + multi_topk_fvu / 8   (train_sae.py:441)
        grads = {k: p.grad for k, p in self.params.items()}
        fired_early = []

        def after_offsets(offsets):
            # did_fire / num_frames_since_fired (train_sae.py:442-446) as soon as the CSC index of the returned
            # encoding exists, beside the weight-gradient kernels; then next step's dead mask and its count
            if self.dp is None:
                ops.dead_latent_update(offsets, frames, n_tokens)
                self._prefetch_dead(n_tokens)
                fired_early.append(True)

        topk_engine.topk_backward(st, 1.0, 1.0, 1.0 / 8.0, out=grads, on_offsets=after_offsets)
        glist = [self.params[k].grad for k in _TOPK_KEYS]
        fired_done = None
        if self.dp is not None:
            side = self.dp.side_stream()
            if side is not None and st.csc_ready is not None:
                # did_fire over all ranks (on the concatenated batch): exchanged on the side stream
                main = torch.cuda.current_stream()
                side.wait_event(st.csc_ready)
                st.offsets.record_stream(side)
                with torch.cuda.stream(side):
                    self._update_fired_dp(st.offsets, n_tokens)
                    self._prefetch_dead(n_tokens)
                    if self._dead_next is not None:
                        self._dead_next[0].record_stream(main)  # the mask is consumed on the main stream next step
                    fired_done = side.record_event()
            if not self.fused_dp:
                self.dp.all_reduce_grads(glist, flat=self._flat_grad)
        if self.fused_dp:
            # gradient exchange + clip_grad_norm_ + optimizer.step (train_sae.py:449-450) in two peer-memory kernels
            self.optimizer.step()
            sumsq = self.optimizer.grad_sumsq().view(1)
        else:
            tl = ops.make_tensor_list([self.params[k].data for k in _TOPK_KEYS], glist)
            sumsq = ops.grad_sumsq(tl, x.device)
            self.optimizer.step(grad_sumsq=sumsq)  # clip_grad_norm_ + optimizer.step, fused
        self.scheduler.step()
        # did_fire / num_frames_since_fired (train_sae.py:442-446), from the CSC index of the returned encoding
        offsets = st.offsets
        if offsets is None:
            offsets, _ = ops.csc_build(res.top_idx, m.n_dict_components)
        if fired_early:
            pass
        elif self.dp is None:
            ops.dead_latent_update(offsets, self.num_frames_since_fired, n_tokens)
        elif fired_done is not None:
            torch.cuda.current_stream().wait_event(fired_done)
        else:
            self._update_fired_dp(offsets, n_tokens)
        self.tokens_seen += n_tokens
        self.last_state = st
        if st.generic:
            loss = res.fvu + res.auxk_loss + res.multi_topk_fvu / 8
        else:
            loss = res.fvu
        return {"loss": loss, "fvu": res.fvu, "auxk_loss": res.auxk_loss, "multi_topk_fvu": res.multi_topk_fvu,
                "grad_sumsq": sumsq, "top_idx": res.top_idx, "top_acts": res.top_acts, "sae_out": res.sae_out}

    def _prefetch_dead(self, n_tokens):
        """Next step's dead mask and `dead_mask.sum()` (topkautoencoder.py:109), copied to pinned memory on the
        current stream so that the next step's read-back does not have to wait for this step to finish."""
        thr = self.dead_feature_threshold
        if thr is None or self.tokens_seen + n_tokens <= thr:
            return
        mask = self.num_frames_since_fired > thr
        self._dead_pinned.copy_(mask.sum().view(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self._dead_next = (mask, ev, self.num_frames_since_fired._version)

    def _update_fired_dp(self, offsets, n_tokens):
        counts = (offsets[1:] - offsets[:-1]).contiguous()
        self.dp.all_reduce_sum(counts)
        f = self.num_frames_since_fired
        f.add_(n_tokens)
        f.masked_fill_(counts > 0, 0)

    # ------------------------------------------------------------------------------------------ L1
    def _l1_step(self, x):
        # model(x); loss = reconstruction_loss + l1_loss; loss.backward()  (train_sae.py:433-434, :448) -- the loss is
        # fixed, so forward and backward are called directly: no autograd graph, no engine round trip per step
        m = self.model
        m.dp, m.materialize_outputs = self.dp, self.materialize_outputs  # what a module call between steps would use
        d = x.shape[-1]
        W, b = m.decoder.weight, m.encoder_bias
        x_hat, latent, scal, saved = l1_forward(x.view(-1, d), W, b, float(m.recon_alpha), self.precision, self.dp,
                                                bool(self.materialize_outputs), True)
        W.grad, b.grad = l1_backward(saved, scal[3:5])  # both incoming gradients are 1
        if self.dp is not None:  # rank-local sums of gradients whose scales already are those of the global batch
            self.dp.all_reduce_grads([p.grad for p in m.parameters()])
        self.optimizer.step()
        self.scheduler.step()
        lead = x.shape[:-1]
        return {"loss": scal[0] + scal[1], "loss_recon": scal[1], "loss_l1": scal[0],
                "sae_out": x_hat.view(*lead, d) if x_hat is not None else None,  # None when materialize_outputs is
                "latent": latent.view(*lead, -1) if latent is not None else None}  # False (bf16 mode)

    def step(self, activations: torch.Tensor):
        """One optimisation step on a [B, T, d] CUDA batch (fp32, or fp16 / bf16 as stored).  Returns device tensors (no
        host sync)."""
        if not activations.is_cuda:
            raise RuntimeError("activations must already be on the CUDA device")
        if self.is_topk and self.precision == BF16 and activations.dtype in (torch.float16, torch.bfloat16):
            x = activations.contiguous()  # collected stores hold fp16: the fused bf16 path widens it in its own kernels
        else:
            x = activations.float().contiguous()
        out = self._topk_step(x) if self.is_topk else self._l1_step(x)
        self.step_count += 1
        return out
