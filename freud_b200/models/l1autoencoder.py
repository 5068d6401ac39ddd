"""Drop-in mirror of src/models/l1autoencoder.py (reference :15-95) on the freud_b200 CUDA kernels.

Tied-weight L1 SAE: state_dict keys `encoder_bias` [n] and `decoder.weight` [d, n]; `encode` renormalises the
decoder columns IN PLACE on every call (reference :71-73), forward returns L1ForwardOutput.  CUDA only.
"""
from typing import NamedTuple

import torch
from torch import Tensor, nn

from .. import ops
from .._lib import BF16, FP32
from ..utils.models import get_n_dict_components
from .config import L1AutoEncoderConfig
from .topkautoencoder import _precision


class L1EncoderOutput(NamedTuple):
    latent: Tensor


class L1ForwardOutput(NamedTuple):
    sae_out: Tensor

    encoded: L1EncoderOutput

    l1_loss: Tensor

    reconstruction_loss: Tensor


def mse_loss(input, target, ignored_index, reduction):
    """mse_loss with ignored_index (reference :29-36), as one masked reduction kernel."""
    if reduction != "mean":
        raise NotImplementedError("freud_b200 mse_loss implements reduction='mean' (the only one the SAE uses)")
    if ignored_index != -1:
        raise NotImplementedError("freud_b200 mse_loss implements ignored_index=-1 (the only one the SAE uses)")
    x = target.contiguous().float()
    xh = input.contiguous().float()
    d = x.shape[-1]
    dummy = torch.zeros((x.numel() // d, 1), dtype=torch.float32, device=x.device)
    acc, _ = ops.l1_loss_reduce(dummy, xh.view(-1, d), x.view(-1, d), False)
    return (acc[1] / acc[2]).float()


def _gemm_operands(t: Tensor, precision: int):
    return ops.split_operand(t, precision)


def _pad_cols_bf16(t: Tensor, mult: int = 8) -> Tensor:
    """bf16 copy of a small fp32 matrix with the row pitch rounded up to `mult` elements (TMA rows are 16-byte multiples)."""
    r, c = t.shape
    ld = (c + mult - 1) // mult * mult
    if ld == c:
        return ops.split_operand(t, BF16)[0]
    out = torch.zeros((r, ld), dtype=torch.bfloat16, device=t.device)
    out[:, :c] = t.to(torch.bfloat16)
    return out


def l1_forward(x2, W, b, recon_alpha, precision, dp=None, want_outputs=True, need_grad=True):
    """Forward of the L1 SAE on [N, d] rows: returns (x_hat, latent, scal, saved) with
    scal = [l1_loss, reconstruction_loss, mse, d recon / d x_hat scale, d l1 / d c scale] (device) and `saved` what
    `l1_backward` needs.  Used by the autograd function below and, directly, by `SAETrainer` (whose loss is always
    reconstruction_loss + l1_loss, train_sae.py:433-434: no autograd graph, no engine round trip).

    bf16 mode runs the step in five tensor-core launches and no elementwise passes over [N, n] / [N, d] matrices:
      encode GEMM  -> epilogue relu + bf16 latent (next GEMM's operand) + sum|c|
      decode GEMM  -> epilogue residual against x, masked / unmasked SSE, count, bf16 masked residual
      dz GEMM      -> epilogue (c > 0) ? s_recon * (dxhat W) + s_l1 : 0 as bf16
      two products over the token axis (x^T dz, dxhat^T c) reading their operands through MN-major descriptors.
    fp32 outputs (x_hat, latent) are only materialised when `want_outputs` (the module API returns them; the trainer
    does not need them).  fp32 mode keeps the 3-pass split-TF32 GEMMs and separate reduction kernels."""
    N, d = x2.shape
    n = W.shape[1]
    Wt = ops.l1_colnorm(W.data)  # in place on decoder.weight.data + K-major transposed copy
    if precision == BF16:
        x16 = ops.split_operand(x2, BF16)[0]
        wt16 = ops.split_operand(Wt, BF16)[0]                                   # [n, d]
        w16 = _pad_cols_bf16(W.data)                                            # [d, ceil8(n)]
        acc = torch.zeros(4, dtype=torch.float64, device=x2.device)  # [sum|c|, masked sse, count, sse]
        c16, _, latent = ops.l1_encode_fused(x16, wt16, b, want_outputs, sums=acc[0:1])     # relu(x @ W + b)
        dxh16, _, x_hat = ops.l1_decode_fused(c16, w16, n, x2, want_outputs, sums=acc[1:4])  # c @ W.T vs x
        saved = (x16, c16, dxh16, wt16)
    else:
        x_ops = _gemm_operands(x2, precision)
        wt_ops = _gemm_operands(Wt, precision)
        w_ops = _gemm_operands(W.data, precision)
        latent = ops.gemm_nt(x_ops[0], x_ops[1], wt_ops[0], wt_ops[1], b, True, precision)   # relu(x @ W + b)
        c_ops = _gemm_operands(latent, precision)
        x_hat = ops.gemm_nt(c_ops[0], c_ops[1], w_ops[0], w_ops[1], None, False, precision)  # c @ W.T
        acc, dxhat = ops.l1_loss_reduce(latent, x_hat, x2, need_grad)
        saved = (x2, latent, dxhat, wt_ops)
    n_glob = N
    if dp is not None:  # losses (and with them the gradient scales) of the batch concatenated over ranks
        acc = dp.all_reduce_sum(acc)
        n_glob = N * dp.world_size
    # l1 = sum|c| / N, recon = alpha * masked mse, mse, and the two gradient scales: one launch (:85-86, :29-36)
    scal = ops.l1_loss_scalars(acc, n_glob, d, recon_alpha)
    return x_hat, latent, scal, saved + (precision, n, d)


def l1_backward(saved, scales):
    """(dW, db) of  g_recon * reconstruction_loss + g_l1 * l1_loss;  scales = [g_recon, g_l1] * scal[3:5] (device):
    d recon / d x_hat = 2*alpha*(x_hat-x)/N_unmasked, d l1 / d c = 1[c>0]/N, times the incoming gradients."""
    a0, a1, a2, a3, precision, n, d = saved
    s_recon = scales[0]
    if precision == BF16:
        x16, c16, dxh16, wt16 = a0, a1, a2, a3
        dz16 = ops.gemm_nt_mask(dxh16, wt16, c16, scales)  # (c>0) ? s_recon*(dxhat W) + s_l1 : 0
        db = ops.col_sum_bf16(dz16, n)
        Ga = ops.gemm_tn_splitk(x16, dz16, M=d, N=n)                             # x^T dz
        Gb = ops.gemm_tn_splitk(dxh16, c16, M=d, N=n)                            # dxhat^T c   (dxhat unscaled)
        dW = torch.addcmul(Ga, Gb, s_recon)
    else:
        x2, latent, dxhat, wt_ops = a0, a1, a2, a3
        # dc (unscaled) = dxhat_raw @ W  as  A = dxhat [N, K=d], B = W^T [n, K=d]
        dx_ops = _gemm_operands(dxhat, precision)
        dc = ops.gemm_nt(dx_ops[0], dx_ops[1], wt_ops[0], wt_ops[1], None, False, precision)
        db = ops.l1_dz(dc, latent, scales)          # dc becomes dz in place
        dW = ops.l1_weight_grad(x2, dc, dxhat, latent, torch.stack((torch.ones_like(s_recon), s_recon)))
    return dW, db


class _L1ForwardFn(torch.autograd.Function):
    """(x, W, b) -> (x_hat, latent, l1_loss, reconstruction_loss, mse); losses carry gradient to W and b
    (`l1_forward` / `l1_backward` behind autograd, for callers that build their own loss on the module's outputs)."""

    @staticmethod
    def forward(ctx, x2, W, b, recon_alpha, precision, dp=None, want_outputs=True):
        need_grad = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        x_hat, latent, scal, saved = l1_forward(x2, W, b, recon_alpha, precision, dp, want_outputs, need_grad)
        l1, recon, mse = scal[0], scal[1], scal[2]
        ctx.saved = (saved, scal)
        outs = [t for t in (x_hat, latent, mse) if t is not None]
        ctx.mark_non_differentiable(*outs)
        return x_hat, latent, l1, recon, mse

    @staticmethod
    def backward(ctx, g_xhat, g_latent, g_l1, g_recon, g_mse):
        saved, scal = ctx.saved
        zero = torch.zeros((), dtype=torch.float32, device=scal.device)
        g_l1 = zero if g_l1 is None else g_l1.float()
        g_recon = zero if g_recon is None else g_recon.float()
        dW, db = l1_backward(saved, torch.stack((g_recon, g_l1)) * scal[3:5])
        ctx.saved = None
        return None, dW, db, None, None, None, None


class L1AutoEncoder(nn.Module):
    # class-level defaults: modules unpickled from reference-written files bypass __init__
    precision = "auto"
    dp = None  # freud_b200.parallel.DataParallel: losses over the concatenated batch (set by SAETrainer)
    materialize_outputs = True  # False (set by SAETrainer): bf16 mode skips the fp32 sae_out / latent copies

    def __init__(self, activation_size: int, cfg: L1AutoEncoderConfig):
        """Same construction order as the reference (:40-67): decoder Linear, zero bias, orthogonal init."""
        super(L1AutoEncoder, self).__init__()
        self.cfg = cfg
        self.tied = True  # tie encoder and decoder weights
        self.activation_size = activation_size
        self.n_dict_components = get_n_dict_components(activation_size, cfg.expansion_factor, cfg.n_dict_components)
        self.recon_alpha = cfg.recon_alpha

        self.decoder = nn.Linear(self.n_dict_components, self.activation_size, bias=False)
        self.encoder_bias = nn.Parameter(torch.zeros(self.n_dict_components))
        nn.init.orthogonal_(self.decoder.weight)
        self.encoder = nn.Sequential(nn.ReLU())

    def _check(self, x: Tensor):
        if not x.is_cuda:
            raise RuntimeError("freud_b200 L1AutoEncoder runs on CUDA only (no CPU fallback); move the model and "
                               "the activations to a CUDA device")
        return x.float().contiguous()

    def encode(self, x: Tensor):
        x = self._check(x)
        d = x.shape[-1]
        prec = _precision(self.precision)
        with torch.no_grad():
            Wt = ops.l1_colnorm(self.decoder.weight.data)  # unit-norm constraint, in place (reference :71-73)
            x_ops = _gemm_operands(x.view(-1, d), prec)
            wt_ops = _gemm_operands(Wt, prec)
            c = ops.gemm_nt(x_ops[0], x_ops[1], wt_ops[0], wt_ops[1], self.encoder_bias, True, prec)
        return L1EncoderOutput(latent=c.view(*x.shape[:-1], self.n_dict_components))

    def decode(self, c: Tensor):
        c = self._check(c)
        n = c.shape[-1]
        prec = _precision(self.precision)
        with torch.no_grad():
            c_ops = _gemm_operands(c.view(-1, n), prec)
            w_ops = _gemm_operands(self.decoder.weight.data, prec)
            out = ops.gemm_nt(c_ops[0], c_ops[1], w_ops[0], w_ops[1], None, False, prec)
        return out.view(*c.shape[:-1], self.activation_size)

    def forward(self, x: Tensor, return_mse: bool = False):
        x = self._check(x)
        d = x.shape[-1]
        x_hat, latent, l1, recon, mse = _L1ForwardFn.apply(x.view(-1, d), self.decoder.weight, self.encoder_bias,
                                                           float(self.recon_alpha), _precision(self.precision),
                                                           self.dp, bool(self.materialize_outputs))
        lead = x.shape[:-1]
        c = latent.view(*lead, self.n_dict_components) if latent is not None else None
        forward_output = L1ForwardOutput(
            sae_out=x_hat.view(*lead, d) if x_hat is not None else None,
            encoded=L1EncoderOutput(c),
            l1_loss=l1,
            reconstruction_loss=recon,
        )
        if return_mse:
            return forward_output, mse
        return forward_output
