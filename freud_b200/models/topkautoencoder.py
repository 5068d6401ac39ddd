"""Drop-in mirror of src/models/topkautoencoder.py (reference :15-175) on the freud_b200 CUDA kernels.

Same constructor, parameters (state_dict keys `W_dec`, `b_dec`, `encoder.weight`, `encoder.bias`), methods and
NamedTuple outputs; `forward` is a torch.autograd.Function over the C ABI, so `loss.backward()`,
`clip_grad_norm_` and `optimizer.step()` in train_sae.py:433-450 work unchanged.

Deviations, all documented in DESIGN.md:
  * CUDA only -- forward on CPU tensors raises (no CPU fallback).  Construction / load_state_dict happen on CPU
    exactly as in the reference (`model = TopKAutoEncoder(...); model.to(device)`, train_sae.py:358-362).
  * autocast (any dtype) selects the bf16 tensor-core mode; selection is always done on fp32 accumulators;
    top_acts are returned in the autocast dtype like the reference's (fp32 outside autocast).
  * differentiable outputs: the three loss scalars (what train_sae.py:441,448 uses), sae_out and top_acts (a
    caller's own loss on the reconstruction / the activations back-propagates to the parameters through the generic
    row-sparse route).  There is no gradient with respect to the input activations.
"""
from typing import NamedTuple

import torch
from torch import Tensor, nn

from .. import ops, topk_engine
from .._lib import BF16, FP32
from ..utils.models import get_n_dict_components
from .config import TopKAutoEncoderConfig


def eager_decode(top_indices: Tensor, top_acts: Tensor, W_dec: Tensor):
    """Reference signature (topkautoencoder.py:15-18): W_dec is passed as [d, n] (`self.W_dec.mT`).
    Computed as the sparse gather-sum over the k selected decoder rows; no [.., n] buffer is built."""
    Wd = W_dec.mT.contiguous()  # back to [n, d]
    lead = top_acts.shape[:-1]
    k = top_acts.shape[-1]
    zeros = torch.zeros(Wd.shape[1], dtype=torch.float32, device=Wd.device)
    out, _, _, _ = ops.topk_decode(top_acts.reshape(-1, k).float().contiguous(),
                                   top_indices.reshape(-1, k).to(torch.int32).contiguous(), Wd.float(), zeros)
    return out.view(*lead, Wd.shape[1])


class TopKEncoderOutput(NamedTuple):
    top_acts: Tensor
    """Activations of the top-k latents."""

    top_indices: Tensor
    """Indices of the top-k features."""


class TopKForwardOutput(NamedTuple):
    sae_out: Tensor

    encoded: TopKEncoderOutput

    fvu: Tensor
    """Fraction of variance unexplained."""

    auxk_loss: Tensor
    """AuxK loss, if applicable."""

    multi_topk_fvu: Tensor
    """Multi-TopK FVU, if applicable."""


def _precision(override: str) -> int:
    if override == "bf16":
        return BF16
    if override == "fp32":
        return FP32
    return BF16 if torch.is_autocast_enabled() else FP32


class _TopKForwardFn(torch.autograd.Function):
    """(x, params) -> (sae_out, top_acts, top_idx, fvu, auxk_loss, multi_topk_fvu, mse).  The loss scalars,
    sae_out and top_acts carry gradient to the parameters; outputs the caller never used arrive as None in
    backward (set_materialize_grads(False)), so the train-step case -- loss scalars only -- stays on the fused route."""

    @staticmethod
    def forward(ctx, x, W_enc, b_enc, W_dec, b_dec, dead_mask, k, auxk_alpha, multi_topk, precision):
        need_grad = any(ctx.needs_input_grad[1:5])
        res, st = topk_engine.topk_forward(x, W_enc, b_enc, W_dec, b_dec, k, precision=precision,
                                           dead_mask=dead_mask, auxk_alpha=auxk_alpha, multi_topk=multi_topk,
                                           need_grad=need_grad)
        ctx.st = st
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(res.top_idx, res.mse)
        return res.sae_out, res.top_acts, res.top_idx, res.fvu, res.auxk_loss, res.multi_topk_fvu, res.mse

    @staticmethod
    def backward(ctx, g_out, g_acts, g_idx, g_fvu, g_aux, g_multi, g_mse):
        grads = topk_engine.topk_backward(ctx.st, g_fvu if g_fvu is not None else 0.0, g_aux, g_multi,
                                          g_out=g_out, g_acts=g_acts)
        ctx.st = None
        return (None, grads["encoder.weight"], grads["encoder.bias"], grads["W_dec"], grads["b_dec"], None, None,
                None, None, None)


class TopKAutoEncoder(nn.Module):
    # class-level default so that modules unpickled from reference-written `torch.save(model)` files
    # (train_sae.py:594-595; __init__ is bypassed) still resolve it
    precision = "auto"  # "auto" (autocast -> bf16, else fp32) | "bf16" | "fp32"

    def __init__(self, activation_size: int, cfg: TopKAutoEncoderConfig, decoder: bool = True):
        """Same construction order as the reference (:45-70) so that a fixed torch seed yields the same init."""
        super().__init__()
        self.cfg = cfg
        self.d_in = activation_size
        self.n_dict_components = get_n_dict_components(activation_size, cfg.expansion_factor, cfg.n_dict_components)

        self.encoder = nn.Linear(self.d_in, self.n_dict_components)
        self.encoder.bias.data.zero_()

        self.W_dec = nn.Parameter(self.encoder.weight.data.clone()) if decoder else None
        if decoder and self.cfg.normalize_decoder:
            self.set_decoder_norm_to_unit_norm()

        self.b_dec = nn.Parameter(torch.zeros(self.d_in))

    # ------------------------------------------------------------------ helpers
    def _check(self, x: Tensor):
        if not x.is_cuda:
            raise RuntimeError("freud_b200 TopKAutoEncoder runs on CUDA only (no CPU fallback); move the model and "
                               "the activations to a CUDA device")
        opt = getattr(self, "_freud_sharded", None)
        if opt is not None and opt() is not None and opt().stale:
            raise RuntimeError("the fp32 weights of this model are sharded across ranks by the fused data-parallel "
                               "optimiser; call trainer.consolidate() on every rank before using the module directly")
        if x.dtype != torch.float32:
            x = x.float()
        return x.contiguous()

    def _as3d(self, x: Tensor, variance_axis: bool = False):
        """[B,T,d] as is.  A 2-D [T,d] input becomes [T,1,d] in forward (variance_axis) so that total_variance's
        `x.mean(0)` (topkautoencoder.py:104) is the mean over T, exactly as the reference computes it for 2-D input;
        encode / pre_acts have no variance and take it as one file [1,T,d]."""
        if x.dim() == 3:
            return x, None
        if x.dim() == 2:
            return (x.unsqueeze(1) if variance_axis else x.unsqueeze(0)), 2
        raise ValueError("expected [B, T, d] or [T, d] activations")

    # ------------------------------------------------------------------ reference API
    def pre_acts(self, x: Tensor) -> Tensor:
        """relu((x - b_dec) @ W_enc.T + b_enc), materialised [.., n] (reference :72-77)."""
        x = self._check(x)
        x3, squeeze = self._as3d(x)
        prec = _precision(self.precision)
        with torch.no_grad():
            xc_hi, xc_lo, we_hi, we_lo, _ = topk_engine.encode_operands(x3, self.encoder.weight, self.b_dec, prec)
            pre = ops.gemm_nt(xc_hi, xc_lo, we_hi, we_lo, self.encoder.bias, True, prec)
        return pre.view(*x.shape[:-1], self.n_dict_components)

    def select_topk(self, latents: Tensor) -> TopKEncoderOutput:
        """Select the top-k latents (reference :79-81); order: value desc, index asc."""
        lat = self._check(latents)
        vals, idx = ops.row_topk(lat.reshape(-1, lat.shape[-1]), self.cfg.k)
        lead = lat.shape[:-1]
        return TopKEncoderOutput(vals.view(*lead, -1), idx.view(*lead, -1).long())

    def encode(self, x: Tensor) -> TopKEncoderOutput:
        """Encode the input and select the top-k latents (reference :83-85); fused, [.., n] never materialised."""
        x = self._check(x)
        x3, squeeze = self._as3d(x)
        prec = _precision(self.precision)
        with torch.no_grad():
            if self.cfg.k == ops.K_FUSED:
                xc_hi, xc_lo, we_hi, we_lo, _ = topk_engine.encode_operands(x3, self.encoder.weight, self.b_dec, prec)
                vals, idx = ops.topk_encode(xc_hi, xc_lo, we_hi, we_lo, self.encoder.bias, prec)
                if prec == FP32:
                    ops.topk_refine(x3.view(-1, self.d_in), self.b_dec.data, self.encoder.weight.data,
                                    self.encoder.bias.data, idx, vals)
            else:
                return self.select_topk(self.pre_acts(x))
        lead = x.shape[:-1]
        return TopKEncoderOutput(vals.view(*lead, -1), idx.view(*lead, -1).long())

    def decode(self, top_acts: Tensor, top_indices: Tensor) -> Tensor:
        assert self.W_dec is not None, "Decoder weight was not initialized."
        lead = top_acts.shape[:-1]
        k = top_acts.shape[-1]
        with torch.no_grad():
            out, _, _, _ = ops.topk_decode(self._check(top_acts).reshape(-1, k),
                                           top_indices.reshape(-1, k).to(torch.int32).contiguous(), self.W_dec,
                                           self.b_dec)
        return out.view(*lead, self.d_in)

    def forward(self, x: Tensor, dead_mask: Tensor | None = None, return_mse: bool = False):
        assert self.W_dec is not None, "Decoder weight was not initialized."
        x = self._check(x)
        x3, squeeze = self._as3d(x, variance_axis=True)
        sae_out, top_acts, top_idx, fvu, auxk, mfvu, mse = _TopKForwardFn.apply(
            x3, self.encoder.weight, self.encoder.bias, self.W_dec, self.b_dec, dead_mask, self.cfg.k,
            float(self.cfg.auxk_alpha), bool(self.cfg.multi_topk), _precision(self.precision))
        if torch.is_autocast_enabled():  # the reference's top_acts come out of the autocast matmul (quirk 8)
            top_acts = top_acts.to(torch.get_autocast_dtype("cuda"))
        lead = x.shape[:-1]
        out = TopKForwardOutput(
            sae_out.view(*lead, self.d_in),
            TopKEncoderOutput(top_acts.view(*lead, -1), top_idx.view(*lead, -1).long()),
            fvu,
            auxk,
            mfvu,
        )
        if return_mse:
            return out, mse
        return out

    @torch.no_grad()
    def set_decoder_norm_to_unit_norm(self):
        assert self.W_dec is not None, "Decoder weight was not initialized."
        eps = torch.finfo(self.W_dec.dtype).eps
        if self.W_dec.is_cuda:
            ops.rownorm_project(self.W_dec.data, eps)
            self._weights_epoch = getattr(self, "_weights_epoch", 0) + 1  # SAETrainer rebuilds its bf16 copies
        else:
            # constructor-time only: the reference builds the module on the CPU and moves it afterwards
            # (train_sae.py:358-362); this is parameter initialisation, not the hot path.
            norm = torch.norm(self.W_dec.data, dim=1, keepdim=True)
            self.W_dec.data /= norm + eps

    @torch.no_grad()
    def remove_gradient_parallel_to_decoder_directions(self):
        assert self.W_dec is not None, "Decoder weight was not initialized."
        assert self.W_dec.grad is not None
        if not self.W_dec.is_cuda:
            raise RuntimeError("freud_b200 runs on CUDA only (no CPU fallback)")
        ops.remove_parallel_grad(self.W_dec.grad, self.W_dec.data)
