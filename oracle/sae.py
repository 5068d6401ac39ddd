"""CPU restatement of FREUD's SAE math (forward, losses, explicit backward).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain torch-CPU tensor
arithmetic with hand-derived gradients -- no autograd, no nn.Module -- so that a
CUDA kernel bug cannot hide behind the same autograd graph.  Every function
cites the reference lines it restates.  Pinned against the imported reference
by tests/golden/make_golden.py + tests/test_oracle_golden.py.

Precision modes
  "fp32": everything fp32 (or fp64 if the inputs are fp64).
  "bf16": restates the *CUDA path's* bf16 mode: GEMM/gather operands rounded to
          bf16 (x - b_dec, W_enc, W_dec, g_e), fp32 accumulation, fp32 selection.
          The reference's own CPU-autocast result (which also rounds GEMM
          outputs to bf16) is the golden this mode must match to 2e-2.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch


def _r(t: torch.Tensor, mode: str) -> torch.Tensor:
    """Round an operand the way the selected precision mode stores it."""
    if mode == "bf16":
        return t.to(torch.bfloat16).to(t.dtype)
    return t


# --------------------------------------------------------------------------- #
# selection
# --------------------------------------------------------------------------- #
FAST_TOPK = False  # bench.py's cpu_baseline sets this: time torch.topk (what the reference calls), not a full sort


def select_topk(latents: torch.Tensor, k: int):
    """topkautoencoder.py:79-81 ``latents.topk(k, sorted=False)``.

    torch.topk's order / tie-break is implementation defined; the oracle (and
    the CUDA kernels) define it: value descending, then index ascending.
    Returned sorted that way; compare with the reference as per-row *sets*.
    """
    if FAST_TOPK:
        return latents.topk(k, sorted=False)
    vals, idx = torch.sort(latents, dim=-1, descending=True, stable=True)
    return vals[..., :k].contiguous(), idx[..., :k].contiguous()


def tie_free_rows(latents: torch.Tensor, k: int) -> torch.Tensor:
    """Rows whose top-k set is unique (k-th value strictly above the (k+1)-th)."""
    vals, _ = torch.sort(latents, dim=-1, descending=True)
    if latents.shape[-1] == k:
        return torch.ones(latents.shape[:-1], dtype=torch.bool)
    return vals[..., k - 1] > vals[..., k]


# --------------------------------------------------------------------------- #
# TopK SAE
# --------------------------------------------------------------------------- #
class TopKOut(NamedTuple):
    sae_out: torch.Tensor
    top_acts: torch.Tensor
    top_indices: torch.Tensor
    fvu: torch.Tensor
    auxk_loss: torch.Tensor  # already multiplied by auxk_alpha (topkautoencoder.py:146)
    multi_topk_fvu: torch.Tensor
    mse: torch.Tensor
    # intermediates kept for the backward
    pre_acts: torch.Tensor
    e: torch.Tensor
    total_variance: torch.Tensor
    aux: Optional[tuple]
    multi: Optional[tuple]
    k_acts: torch.Tensor  # the k-selection behind fvu (== top_acts unless multi_topk)
    k_indices: torch.Tensor


def topk_pre_acts(x, W_enc, b_enc, b_dec, mode="fp32"):
    """topkautoencoder.py:72-77: relu((x - b_dec) @ W_enc.T + b_enc)."""
    xc = _r(x - b_dec, mode)
    return torch.relu(xc @ _r(W_enc, mode).T + b_enc)


def topk_decode(top_acts, top_indices, W_dec, b_dec, mode="fp32"):
    """topkautoencoder.py:15-18,87-91: scatter into zeros[..., n], dense @ W_dec, + b_dec
    (restated as the equivalent gather-sum over the k selected decoder rows)."""
    rows = _r(W_dec, mode)[top_indices]  # [..., k, d]
    return (top_acts.unsqueeze(-1) * rows).sum(-2) + b_dec


def total_variance(x):
    """topkautoencoder.py:104-106: sum((x - x.mean(0))**2), replaced by 1.0 if 0."""
    tv = (x - x.mean(0)).pow(2).sum()
    if tv == 0:
        tv = torch.ones_like(tv)
    return tv


def topk_forward(x, W_enc, b_enc, W_dec, b_dec, k, *, dead_mask=None, auxk_alpha=0.0,
                 multi_topk=False, mode="fp32", force_indices=None) -> TopKOut:
    """topkautoencoder.py:93-151 (x is [B,T,d]).

    force_indices ([..., k] int64): impose this main selection instead of the oracle's own (values are still the
    oracle's pre-activations at those indices).  Parity tests use it on the few rows whose k-th / (k+1)-th
    pre-activations tie within GEMM rounding, where either set is a valid top-k, so that sums over rows
    (losses, gradients) can still be compared on every row."""
    pre = topk_pre_acts(x, W_enc, b_enc, b_dec, mode)
    if force_indices is not None:
        top_idx = force_indices.reshape(*pre.shape[:-1], k).long()
        top_acts = torch.gather(pre, -1, top_idx)
    else:
        top_acts, top_idx = select_topk(pre, k)
    sae_out = topk_decode(top_acts, top_idx, W_dec, b_dec, mode)
    e = sae_out - x
    tv = total_variance(x)
    aux = None
    if dead_mask is not None and int(dead_mask.sum()) > 0:
        num_dead = int(dead_mask.sum())
        k_aux = x.shape[-1] // 2
        scale = min(num_dead / k_aux, 1.0)
        k_aux = min(k_aux, num_dead)
        aux_lat = torch.where(dead_mask[None], pre, torch.full_like(pre, -torch.inf))
        a_acts, a_idx = select_topk(aux_lat, k_aux)
        e_hat = topk_decode(a_acts, a_idx, W_dec, b_dec, mode)
        auxk = scale * (e_hat - e).pow(2).sum() / tv
        aux = (a_acts, a_idx, e_hat, scale)
    else:
        auxk = x.new_zeros(())
    fvu = e.pow(2).sum() / tv
    multi = None
    ret_out, ret_acts, ret_idx = sae_out, top_acts, top_idx
    if multi_topk:
        m_acts, m_idx = select_topk(pre, 4 * k)
        m_out = topk_decode(m_acts, m_idx, W_dec, b_dec, mode)
        mfvu = (m_out - x).pow(2).sum() / tv
        multi = (m_acts, m_idx, m_out)
        # reference rebinds the returned encoding to the 4k one (topkautoencoder.py:135-136)
        ret_out, ret_acts, ret_idx = m_out, m_acts, m_idx
    else:
        mfvu = x.new_zeros(())
    return TopKOut(ret_out, ret_acts, ret_idx, fvu, auxk * auxk_alpha, mfvu, e.pow(2).mean(),
                   pre, e, tv, aux, multi, top_acts, top_idx)


def topk_backward(x, W_enc, b_enc, W_dec, b_dec, out: TopKOut, k, *, auxk_alpha=0.0,
                  multi_topk=False, mode="fp32"):
    """Gradients of ``fvu + auxk_loss + multi_topk_fvu / 8`` (train_sae.py:441,448);
    SURVEY.md M8' formulas.  Returns dict of parameter gradients."""
    n, d = W_enc.shape
    tv = out.total_variance
    e = out.e
    xc = _r(x - b_dec, mode)
    Wd = _r(W_dec, mode)
    top_acts, top_idx, multi = out.k_acts, out.k_indices, out.multi

    g_e = 2.0 * e / tv
    decodes = []  # (acts, idx, grad wrt that decode's output)
    if out.aux is not None:
        a_acts, a_idx, e_hat, scale = out.aux
        g_ehat = auxk_alpha * scale * 2.0 * (e_hat - e) / tv
        g_e = g_e - g_ehat  # e is not detached (topkautoencoder.py:126)
        decodes.append((a_acts, a_idx, g_ehat))
    decodes.append((top_acts, top_idx, g_e))
    if multi is not None:
        m_acts, m_idx, m_out = multi
        decodes.append((m_acts, m_idx, 2.0 * (m_out - x) / tv / 8.0))

    dW_dec = torch.zeros_like(W_dec)
    db_dec = torch.zeros_like(b_dec)
    dpre = torch.zeros_like(out.pre_acts)
    N = x.numel() // d
    for acts, idx, g in decodes:
        gq = _r(g, mode)  # CUDA bf16 mode stores the decode-output gradient in bf16
        db_dec += g.reshape(N, d).sum(0)
        kk = idx.shape[-1]
        flat_idx = idx.reshape(N, kk)
        flat_acts = acts.reshape(N, kk)
        g2 = gq.reshape(N, d)
        # dW_dec[f] += act * g  (row scatter-add)
        dW_dec.index_add_(0, flat_idx.reshape(-1),
                          (flat_acts.unsqueeze(-1) * g2.unsqueeze(1)).reshape(-1, d))
        # d act = g . W_dec[f]
        dacts = (g2.unsqueeze(1) * Wd[flat_idx]).sum(-1)
        dpre.reshape(N, n).scatter_add_(1, flat_idx, dacts)
    dpre = dpre * (out.pre_acts > 0)
    dpre2 = dpre.reshape(N, n)
    dW_enc = dpre2.T @ xc.reshape(N, d)
    db_enc = dpre2.sum(0)
    db_dec = db_dec - (dpre2 @ _r(W_enc, mode)).sum(0)
    return {"encoder.weight": dW_enc, "encoder.bias": db_enc, "W_dec": dW_dec, "b_dec": db_dec}


def set_decoder_norm_to_unit_norm(W_dec):
    """topkautoencoder.py:153-159 (eps is *added* to the norm)."""
    eps = torch.finfo(W_dec.dtype).eps
    return W_dec / (torch.norm(W_dec, dim=1, keepdim=True) + eps)


def remove_gradient_parallel_to_decoder_directions(W_dec, W_dec_grad):
    """topkautoencoder.py:161-175."""
    par = (W_dec_grad * W_dec).sum(1, keepdim=True)
    return W_dec_grad - par * W_dec


# --------------------------------------------------------------------------- #
# L1 SAE
# --------------------------------------------------------------------------- #
class L1Out(NamedTuple):
    sae_out: torch.Tensor
    latent: torch.Tensor
    l1_loss: torch.Tensor
    reconstruction_loss: torch.Tensor
    mse: torch.Tensor
    W_normed: torch.Tensor
    n_unmasked: int


def l1_normalize_columns(W):
    """l1autoencoder.py:71-73: F.normalize(W.data, dim=0) on decoder.weight [d,n]."""
    return W / torch.clamp(torch.norm(W, dim=0, keepdim=True), min=1e-12)


def l1_forward(x, W, b, recon_alpha, mode="fp32") -> L1Out:
    """l1autoencoder.py:69-95 with W = decoder.weight [d,n], b = encoder_bias [n]."""
    Wn = l1_normalize_columns(W)
    c = torch.relu(_r(x, mode) @ _r(Wn, mode) + b)
    x_hat = _r(c, mode) @ _r(Wn, mode).T
    l1 = c.abs().sum(-1).mean()
    mask = x != -1  # mse_loss(..., ignored_index=-1) l1autoencoder.py:29-36
    n_unmasked = int(mask.sum())
    recon = recon_alpha * ((x_hat - x)[mask] ** 2).mean()
    return L1Out(x_hat, c, l1, recon, ((x_hat - x) ** 2).mean(), Wn, n_unmasked)


def l1_backward(x, W, b, out: L1Out, recon_alpha, mode="fp32"):
    """Gradients of reconstruction_loss + l1_loss (train_sae.py:434); SURVEY.md M3'.
    The in-place normalisation is outside autograd, so grads are w.r.t. the
    normalised weight the forward used."""
    d, n = W.shape
    N = x.numel() // d
    Wn, c = out.W_normed, out.latent.reshape(N, n)
    x2, xh = x.reshape(N, d), out.sae_out.reshape(N, d)
    mask = x2 != -1
    dxh = 2.0 * recon_alpha * (xh - x2) * mask / out.n_unmasked
    dxh_q = _r(dxh, mode)
    dc = dxh_q @ _r(Wn, mode) + (c > 0).to(x.dtype) / N
    dz = dc * (c > 0)
    dz_q = _r(dz, mode)
    dW = _r(x2, mode).T @ dz_q + dxh_q.T @ _r(c, mode)
    return {"decoder.weight": dW, "encoder_bias": dz.sum(0)}


# --------------------------------------------------------------------------- #
# validation feature statistics
# --------------------------------------------------------------------------- #
def topk_feature_absmax(top_acts, top_indices, n):
    """train_sae.py:70-118 topk_feature_extraction for one file: out[f] = max |act| over the (t, j) entries whose
    index is f, 0 where the feature never appears (the reference builds a [T,k,n] mask for this)."""
    out = torch.zeros(n, dtype=top_acts.dtype)
    flat_i = top_indices.reshape(-1).long()
    flat_a = top_acts.reshape(-1).abs()
    out.scatter_reduce_(0, flat_i, flat_a, reduce="amax", include_self=True)
    return out


def l1_feature_absmax(latent):
    """train_sae.py:175-178: per-file max over frames of |latent|."""
    return latent.reshape(-1, latent.shape[-1]).abs().max(0).values
