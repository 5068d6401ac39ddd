"""Torch-tensor front end of the C ABI: allocates outputs / workspaces with torch (the caller owns all memory)
and enqueues the CUDA kernels on torch's current stream.  No CPU path: every op requires CUDA tensors."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import BF16, FP32, TensorList, call

K_FUSED = 32  # selection width of the fused encoder epilogue


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream (as an integer: ctypes takes it for a void* parameter).  The short steps
    (C2, the L1 SAE) are bound by host enqueue time, so the per-call conversions are kept to the minimum."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("freud_b200 ops need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise RuntimeError("freud_b200 ops need contiguous tensors")
    return t.data_ptr()


def _f32(t, name):
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t


# ------------------------------------------------------------------------------------------------ TopK forward
def split_operand(w: torch.Tensor, precision: int):
    """W -> GEMM operand(s): bf16 copy, or (hi, lo) tf32 split."""
    _f32(w, "w")
    if precision == BF16:
        hi = torch.empty_like(w, dtype=torch.bfloat16)
        call("freud_split_operand", _ptr(w), _ptr(hi), None, w.numel(), precision, _stream())
        return hi, None
    hi, lo = torch.empty_like(w), torch.empty_like(w)
    call("freud_split_operand", _ptr(w), _ptr(hi), _ptr(lo), w.numel(), precision, _stream())
    return hi, lo


_X_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def _x_dtype(x: torch.Tensor) -> int:
    """Storage type code of an activation batch (fp32, or fp16 / bf16 as collected stores hold it)."""
    if x.dtype not in _X_DTYPES or not x.is_cuda or not x.is_contiguous():
        raise TypeError("activations must be a contiguous CUDA tensor of dtype float32, float16 or bfloat16")
    return _X_DTYPES[x.dtype]


def topk_prep_x(x: torch.Tensor, b_dec: torch.Tensor, precision: int, want_colmean: bool = False):
    """x [B,T,d] -> (xc_hi, xc_lo operand(s) of x - b_dec as [N,d], tv double[1][, x.mean(0) [T,d]])."""
    B, T, d = x.shape
    tv = torch.empty(1, dtype=torch.float64, device=x.device)
    colmean = torch.empty((T, d), dtype=torch.float32, device=x.device) if want_colmean else None
    if precision == BF16:
        hi = torch.empty((B * T, d), dtype=torch.bfloat16, device=x.device)
        lo = None
    else:
        hi = torch.empty((B * T, d), dtype=torch.float32, device=x.device)
        lo = torch.empty_like(hi)
    call("freud_topk_prep_x", _ptr(x), _x_dtype(x), _ptr(b_dec), _ptr(hi), _ptr(lo), _ptr(tv), _ptr(colmean), B, T, d,
         precision, _stream())
    if want_colmean:
        return hi, lo, tv, colmean
    return hi, lo, tv


def topk_encode_workspace_bytes(N: int, n: int) -> int:
    """Workspace freud_topk_encode wants for splitting the row blocks of a partial last wave (0: no split)."""
    need = C.c_int64(0)
    call("freud_topk_encode_workspace", N, n, C.byref(need))
    return need.value


def topk_encode(xc_hi, xc_lo, w_hi, w_lo, b_enc, precision: int, hist=None):
    """Fused GEMM + bias + ReLU + top-32.  Returns (top_vals fp32 [N,32], top_idx int32 [N,32]).
    hist: optional zeroed int32 [>= n] that receives the per-feature counts of the emitted indices (csc_build's start)."""
    N, d = xc_hi.shape
    n = w_hi.shape[0]
    vals = torch.empty((N, K_FUSED), dtype=torch.float32, device=xc_hi.device)
    idx = torch.empty((N, K_FUSED), dtype=torch.int32, device=xc_hi.device)
    need = topk_encode_workspace_bytes(N, n)
    ws = torch.empty(need, dtype=torch.uint8, device=xc_hi.device) if need else None
    call("freud_topk_encode", _ptr(xc_hi), _ptr(xc_lo), _ptr(w_hi), _ptr(w_lo), _ptr(b_enc), _ptr(vals), _ptr(idx),
         N, d, n, precision, _ptr(ws), need, _ptr(hist), _stream())
    return vals, idx


def topk_refine(x2, b_dec, W_enc, b_enc, idx, vals):
    """fp32 mode: selected pre-activations recomputed in place with an fp32 FMA chain (see freud_topk_refine)."""
    N, d = x2.shape
    call("freud_topk_refine", _ptr(_f32(x2, "x")), _ptr(b_dec), _ptr(_f32(W_enc, "W_enc")), _ptr(b_enc), _ptr(idx),
         _ptr(vals), N, d, idx.shape[1], _stream())
    return vals


def gemm_nt(a_hi, a_lo, b_hi, b_lo, bias, relu: bool, precision: int, out=None):
    """out[M,N] = act(A @ B^T + bias) on the tensor cores (operands as from split_operand / topk_prep_x)."""
    M, K = a_hi.shape
    Nn = b_hi.shape[0]
    if out is None:
        out = torch.empty((M, Nn), dtype=torch.float32, device=a_hi.device)
    call("freud_gemm_nt", _ptr(a_hi), _ptr(a_lo), _ptr(b_hi), _ptr(b_lo), _ptr(bias), _ptr(out), M, Nn, K,
         out.stride(0), int(relu), precision, _stream())
    return out


def row_topk(latents: torch.Tensor, k: int, col_mask: torch.Tensor | None = None):
    _f32(latents, "latents")
    rows, n = latents.shape
    vals = torch.empty((rows, k), dtype=torch.float32, device=latents.device)
    idx = torch.empty((rows, k), dtype=torch.int32, device=latents.device)
    m = None
    if col_mask is not None:
        m = col_mask.to(torch.uint8).contiguous()
    call("freud_row_topk", _ptr(latents), _ptr(m), _ptr(vals), _ptr(idx), rows, n, k, _stream())
    return vals, idx


def gather_rows(src: torch.Tensor, rows: torch.Tensor):
    """dst[r] = src[rows[r]] for a 1-D or 2-D source (rows int32)."""
    row_elems = 1 if src.dim() == 1 else src.shape[1]
    out = torch.empty((rows.numel(),) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    call("freud_gather_rows", _ptr(src), _ptr(rows), _ptr(out), rows.numel(), row_elems * src.element_size(),
         _stream())
    return out


def index_map(table: torch.Tensor, idx: torch.Tensor):
    out = torch.empty_like(idx)
    call("freud_index_map", _ptr(table), _ptr(idx), _ptr(out), idx.numel(), _stream())
    return out


def row_topk_mask(latents: torch.Tensor, k: int, ld: int, n: int = None, nonneg: bool = False):
    """Dense bf16 [rows, ld] with each row's exact top-k (among the first n columns) kept and everything else zero.
    `latents` may carry a padded row pitch (latents.shape[1] >= n).  nonneg: the caller guarantees latents >= 0
    (post-ReLU pre-activations), which allows the streaming warp-per-row kernel."""
    rows, ld_in = latents.shape
    n = ld_in if n is None else n
    out = torch.empty((rows, ld), dtype=torch.bfloat16, device=latents.device)
    call("freud_row_topk_mask", _ptr(_f32(latents, "latents")), _ptr(out), rows, n, k, ld_in, ld, int(nonneg),
         _stream())
    return out


def col_sum_bf16(x: torch.Tensor, n: int):
    rows, ld = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    call("freud_col_sum_bf16", _ptr(x), _ptr(out), rows, n, ld, _stream())
    return out


def scatter_add_rows(src: torch.Tensor, rows_idx: torch.Tensor, dst: torch.Tensor):
    row_elems = 1 if src.dim() == 1 else src.shape[1]
    call("freud_scatter_add_rows", _ptr(_f32(src, "src")), _ptr(rows_idx), _ptr(_f32(dst, "dst")), rows_idx.numel(),
         row_elems, _stream())


def gemm_nt_mask(a: torch.Tensor, b: torch.Tensor, act: torch.Tensor, affine: torch.Tensor = None):
    """bf16 [M, ld] = (act > 0) ? scale * (a[:, :K] @ b^T) + shift : 0 with a [M, lda], b [N, K] bf16, act bf16
    [M, ld >= N]; (scale, shift) = `affine` (2 device floats) or (1, 0)."""
    M, lda = a.shape
    N, K = b.shape
    out = torch.empty_like(act)
    call("freud_gemm_nt_mask", _ptr(a), _ptr(b), _ptr(act), _ptr(out), _ptr(affine), M, N, K, lda, act.shape[1],
         _stream())
    return out


def l1_loss_scalars(acc: torch.Tensor, n_glob: int, d: int, recon_alpha: float):
    """acc double[4] -> float[5] = (l1, recon, mse, 2*alpha/count, 1/n_glob) in one launch."""
    out = torch.empty(5, dtype=torch.float32, device=acc.device)
    call("freud_l1_loss_scalars", _ptr(acc), float(n_glob), float(d), float(recon_alpha), _ptr(out), _stream())
    return out


def l1_encode_fused(x16: torch.Tensor, wt16: torch.Tensor, bias: torch.Tensor, want_latent: bool, sums=None):
    """c = relu(x W + b) as bf16 [M, ceil8(n)] + sum(c) (double[1], accumulated into `sums` when given -- zeroed by the
    caller) [+ fp32 latent]."""
    M, K = x16.shape
    n = wt16.shape[0]
    ld = (n + 7) // 8 * 8
    c16 = torch.empty((M, ld), dtype=torch.bfloat16, device=x16.device)
    latent = torch.empty((M, n), dtype=torch.float32, device=x16.device) if want_latent else None
    if sums is None:
        sums = torch.zeros(1, dtype=torch.float64, device=x16.device)
    call("freud_l1_encode_fused", _ptr(x16), _ptr(wt16), _ptr(bias), _ptr(c16), _ptr(latent), _ptr(sums), M, n, K, ld,
         _stream())
    return c16, sums, latent


def l1_decode_fused(c16: torch.Tensor, w16: torch.Tensor, n: int, target: torch.Tensor, want_xhat: bool, sums=None):
    """x_hat = c W^T against `target`: (dxhat16 bf16 [M, ceil8(d)] unscaled masked residual, sums double[3] =
    (masked sse, count, sse) [, fp32 x_hat]).  c16 [M, lda] with n valid columns, w16 [d, ldb >= n]."""
    M, lda = c16.shape
    d, ldb = w16.shape
    ld = (d + 7) // 8 * 8
    r16 = torch.empty((M, ld), dtype=torch.bfloat16, device=c16.device)
    x_hat = torch.empty((M, d), dtype=torch.float32, device=c16.device) if want_xhat else None
    if sums is None:
        sums = torch.zeros(3, dtype=torch.float64, device=c16.device)
    call("freud_l1_decode_fused", _ptr(c16), _ptr(w16), _ptr(_f32(target, "target")), _ptr(r16), _ptr(x_hat),
         _ptr(sums), M, d, n, lda, ldb, ld, _stream())
    return r16, sums, x_hat


def _lib_sm_count() -> int:
    return torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count if torch.cuda.is_available() else 148


def _splits_for(M: int, K: int, splits=None):
    total_kb = (K + 63) // 64
    if splits is None:
        # one wave: row blocks x splits <= SM count (rounding UP put 160 CTAs on 148 SMs: a second, nearly empty wave
        # doubled the time of the [2458, 768] AuxK weight-gradient products)
        splits = max(1, min(total_kb, _lib_sm_count() // max(1, (M + 127) // 128)))
    per = (total_kb + splits - 1) // splits
    return (total_kb + per - 1) // per  # the split count the kernel will actually run (no empty partials)


def gemm_tn_splitk(a: torch.Tensor, b: torch.Tensor, M: int = None, N: int = None, splits: int = None):
    """a^T @ b for bf16 a [K, lda], b [K, ldb] stored row-major (K = tokens), first M / N columns used: fp32 [M, N].
    The operands are read through MN-major tensor-core descriptors -- nothing is transposed in memory."""
    K, lda = a.shape
    ldb = b.shape[1]
    M = lda if M is None else M
    N = ldb if N is None else N
    splits = _splits_for(M, K, splits)
    ws = torch.empty((splits, M, N), dtype=torch.float32, device=a.device)
    call("freud_gemm_tn_splitk", _ptr(a), _ptr(b), _ptr(ws), M, N, K, lda, ldb, splits, _stream())
    if splits == 1:
        return ws[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    call("freud_sum_splits", _ptr(ws), _ptr(out), splits, M * N, _stream())
    return out


def gemm_nn(a: torch.Tensor, b: torch.Tensor, bias=None, relu: bool = False, K: int = None, N: int = None):
    """act(a[:, :K] @ b[:K, :N] + bias) for bf16 a [M, lda] and b [K', ldb] stored row-major: fp32 [M, N]."""
    M, lda = a.shape
    ldb = b.shape[1]
    K = lda if K is None else K
    N = ldb if N is None else N
    ldo = (N + 3) // 4 * 4
    out = torch.empty((M, ldo), dtype=torch.float32, device=a.device)
    call("freud_gemm_nn", _ptr(a), _ptr(b), _ptr(bias), _ptr(out), M, N, K, lda, ldb, ldo, int(relu), _stream())
    return out[:, :N] if ldo != N else out


def topk_decode(top_vals, top_idx, W_dec, b_dec, target=None, *, resid_dtype=None, want_sse=False,
                want_colsum=False):
    """sae_out = sum_j a_j W_dec[i_j] + b_dec; optional residual / sse / column sums against `target`."""
    N, k = top_vals.shape
    d = W_dec.shape[1]
    dev = top_vals.device
    sae_out = torch.empty((N, d), dtype=torch.float32, device=dev)
    resid = torch.empty((N, d), dtype=resid_dtype, device=dev) if resid_dtype is not None else None
    sse = torch.zeros(1, dtype=torch.float64, device=dev) if want_sse else None
    colsum = torch.zeros(d, dtype=torch.float32, device=dev) if want_colsum else None
    call("freud_topk_decode", _ptr(top_vals), _ptr(top_idx), _ptr(W_dec), int(W_dec.dtype == torch.bfloat16),
         _ptr(b_dec), _ptr(target), _ptr(sae_out), _ptr(resid),
         int(resid is not None and resid.dtype == torch.bfloat16), _ptr(sse), _ptr(colsum), N, d, k, _stream())
    return sae_out, resid, sse, colsum


def decode_dacts_supported(d: int, k: int) -> bool:
    return bool(_lib.lib().freud_topk_decode_dacts_supported(d, k))


def topk_decode_dacts(top_vals, top_idx, W_dec, b_dec, target):
    """Fused bf16 fast path: (sae_out fp32, residual bf16, sse, colsum, dacts) with the decoder rows gathered once."""
    N, k = top_vals.shape
    d = W_dec.shape[1]
    dev = top_vals.device
    assert W_dec.dtype == torch.bfloat16
    sae_out = torch.empty((N, d), dtype=torch.float32, device=dev)
    resid = torch.empty((N, d), dtype=torch.bfloat16, device=dev)
    sse = torch.zeros(1, dtype=torch.float64, device=dev)
    colsum = torch.zeros(d, dtype=torch.float32, device=dev)
    dacts = torch.empty((N, k), dtype=torch.float32, device=dev)
    call("freud_topk_decode_dacts", _ptr(top_vals), _ptr(top_idx), _ptr(W_dec), _ptr(b_dec), _ptr(target),
         _x_dtype(target), _ptr(sae_out), _ptr(resid), _ptr(sse), _ptr(colsum), _ptr(dacts), N, d, k, _stream())
    return sae_out, resid, sse, colsum, dacts


def topk_dacts(g, top_idx, W_dec):
    N, k = top_idx.shape
    d = W_dec.shape[1]
    out = torch.empty((N, k), dtype=torch.float32, device=g.device)
    call("freud_topk_dacts", _ptr(g), int(g.dtype == torch.bfloat16), _ptr(top_idx), _ptr(W_dec),
         int(W_dec.dtype == torch.bfloat16), _ptr(out), N, d, k, _stream())
    return out


def axpby(a, b, coef, out_dtype):
    out = torch.empty(a.shape, dtype=out_dtype, device=a.device)
    call("freud_axpby", _ptr(a), _ptr(b), _ptr(coef), _ptr(out), int(out_dtype == torch.bfloat16), a.numel(),
         _stream())
    return out


# ------------------------------------------------------------------------------------------------ feature sharding
def shard_merge(all_vals, all_idx, n_local: int):
    """[G,N,32] all-gathered per-shard selections -> global (vals, dictionary indices) [N,32]."""
    G, N, k = all_vals.shape
    vals = torch.empty((N, k), dtype=torch.float32, device=all_vals.device)
    idx = torch.empty((N, k), dtype=torch.int32, device=all_vals.device)
    call("freud_shard_merge", _ptr(all_vals), _ptr(all_idx), _ptr(vals), _ptr(idx), N, G, n_local, _stream())
    return vals, idx


def shard_localize(vals, gidx, lo: int, n_local: int):
    lvals, lidx = torch.empty_like(vals), torch.empty_like(gidx)
    call("freud_shard_localize", _ptr(vals), _ptr(gidx), _ptr(lvals), _ptr(lidx), vals.numel(), lo, n_local, _stream())
    return lvals, lidx


def residual(sae_out, target, resid_dtype, want_colsum=True):
    N, d = sae_out.shape
    dev = sae_out.device
    resid = torch.empty((N, d), dtype=resid_dtype, device=dev) if resid_dtype is not None else None
    sse = torch.zeros(1, dtype=torch.float64, device=dev)
    colsum = torch.zeros(d, dtype=torch.float32, device=dev) if want_colsum else None
    call("freud_residual", _ptr(sae_out), _ptr(target), _ptr(resid), int(resid_dtype == torch.bfloat16), _ptr(sse),
         _ptr(colsum), N, d, _stream())
    return resid, sse, colsum


# ------------------------------------------------------------------------------------------------ TopK backward
def csc_build(top_idx: torch.Tensor, n: int, counts=None):
    """Feature-major index of the selected entries.  counts: int32 [n+1] already holding the per-feature counts
    (topk_encode's hist); it becomes the offsets array."""
    N, k = top_idx.shape
    dev = top_idx.device
    offsets = counts if counts is not None else torch.empty(n + 1, dtype=torch.int32, device=dev)
    entries = torch.empty(N * k, dtype=torch.int32, device=dev)
    cursor = torch.empty(n + 1, dtype=torch.int32, device=dev)  # fill cursors, then the long-list sort queue + its counter
    call("freud_csc_build", _ptr(top_idx), N, k, n, _ptr(offsets), _ptr(entries), _ptr(cursor),
         int(counts is not None), _stream())
    return offsets, entries


def topk_sparse_grads(offsets, entries, top_vals, dacts, g, xc, b_dec, scales, dW_dec, dW_enc, db_enc, k,
                      accumulate: bool):
    n, d = dW_dec.shape
    dev = dW_dec.device
    n_entries = entries.numel()
    meta = torch.empty(3 * n_entries, dtype=torch.int32, device=dev)
    chunk_off = torch.empty(n + 1, dtype=torch.int32, device=dev)
    call("freud_csc_meta", _ptr(offsets), _ptr(entries), _ptr(top_vals), _ptr(dacts), _ptr(scales), _ptr(meta),
         n_entries, n, k, _stream())
    call("freud_topk_sparse_grads", _ptr(offsets), _ptr(meta), _ptr(g), int(g.dtype == torch.bfloat16), _ptr(xc),
         int(xc.dtype == torch.bfloat16), _ptr(b_dec), _ptr(dW_dec), _ptr(dW_enc), _ptr(db_enc), _ptr(chunk_off),
         n_entries, n, d, k, int(accumulate), _stream())


def topk_bdec_grad(colsum, scales, db_enc, W_enc, db_dec, accumulate: bool):
    d = db_dec.numel()
    n = db_enc.numel() if db_enc is not None else 0
    call("freud_topk_bdec_grad", _ptr(colsum), _ptr(scales), _ptr(db_enc), _ptr(W_enc),
         int(W_enc is not None and W_enc.dtype == torch.bfloat16), _ptr(db_dec), n, d, int(accumulate), _stream())


def topk_loss_scalars(sse, tv, numel: int):
    out = torch.empty(5, dtype=torch.float32, device=sse.device)
    call("freud_topk_loss_scalars", _ptr(sse), _ptr(tv), _ptr(out), numel, _stream())
    return out


def dead_latent_update(offsets, frames, n_tokens: int):
    if frames.dtype != torch.int64:
        raise TypeError("num_frames_since_fired must be int64")
    call("freud_dead_latent_update", _ptr(offsets), _ptr(frames), frames.numel(), n_tokens, _stream())


def rownorm_project(W, eps: float):
    call("freud_rownorm_project", _ptr(_f32(W, "W")), W.shape[0], W.shape[1], eps, _stream())


def remove_parallel_grad(G, W):
    call("freud_remove_parallel_grad", _ptr(_f32(G, "G")), _ptr(_f32(W, "W")), W.shape[0], W.shape[1], _stream())


# ------------------------------------------------------------------------------------------------ L1
def l1_colnorm(W):
    d, n = W.shape
    Wt = torch.empty((n, d), dtype=torch.float32, device=W.device)
    call("freud_l1_colnorm", _ptr(_f32(W, "W")), _ptr(Wt), d, n, _stream())
    return Wt


def l1_loss_reduce(latent, x_hat, x, want_dxhat: bool):
    N, n = latent.shape
    d = x.shape[-1]
    acc = torch.zeros(4, dtype=torch.float64, device=x.device)
    dxhat = torch.empty((N, d), dtype=torch.float32, device=x.device) if want_dxhat else None
    call("freud_l1_loss_reduce", _ptr(latent), _ptr(x_hat), _ptr(x), _ptr(dxhat), _ptr(acc), N, d, n, _stream())
    return acc, dxhat


def l1_dz(dc, latent, scales):
    N, n = latent.shape
    db = torch.zeros(n, dtype=torch.float32, device=dc.device)
    call("freud_l1_dz", _ptr(dc), _ptr(latent), _ptr(scales), _ptr(db), N, n, _stream())
    return db


def l1_weight_grad(x, dz, dxhat, latent, scales):
    N, d = x.shape
    n = latent.shape[1]
    dW = torch.empty((d, n), dtype=torch.float32, device=x.device)
    call("freud_l1_weight_grad", _ptr(x), _ptr(dz), _ptr(dxhat), _ptr(latent), _ptr(scales), _ptr(dW), N, d, n,
         _stream())
    return dW


# ------------------------------------------------------------------------------------------------ optimiser
def make_tensor_list(params, grads, exp_avgs=None, exp_avg_sqs=None, shadows=None) -> TensorList:
    if len(params) > _lib.MAX_TENSORS:
        raise ValueError(f"at most {_lib.MAX_TENSORS} tensors per list")
    tl = TensorList()
    tl.count = len(params)
    for i, p in enumerate(params):
        for t in (p, grads[i]):
            if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                raise RuntimeError("optimizer tensors must be contiguous float32 CUDA tensors")
        tl.param[i] = p.data_ptr()
        tl.grad[i] = grads[i].data_ptr()
        tl.exp_avg[i] = exp_avgs[i].data_ptr() if exp_avgs is not None else None
        tl.exp_avg_sq[i] = exp_avg_sqs[i].data_ptr() if exp_avg_sqs is not None else None
        tl.bf16_shadow[i] = shadows[i].data_ptr() if shadows is not None and shadows[i] is not None else None
        tl.numel[i] = p.numel()
    return tl


def grad_sumsq(tl: TensorList, device):
    out = torch.empty(1, dtype=torch.float64, device=device)
    call("freud_grad_sumsq", C.byref(tl), _ptr(out), _stream())
    return out


def clip_grads(tl: TensorList, sumsq, max_norm: float):
    norm = torch.empty((), dtype=torch.float32, device=sumsq.device)
    call("freud_clip_grads", C.byref(tl), _ptr(sumsq), max_norm, _ptr(norm), _stream())
    return norm


def adam_step(tl: TensorList, lr, beta1, beta2, eps, step: int, sumsq=None, max_norm: float = 0.0):
    call("freud_adam_step", C.byref(tl), lr, beta1, beta2, eps, step, _ptr(sumsq), max_norm, _stream())


def radam_step(tl: TensorList, lr, beta1, beta2, eps, weight_decay, step: int, sumsq=None, max_norm: float = 0.0):
    call("freud_radam_step", C.byref(tl), lr, beta1, beta2, eps, weight_decay, step, _ptr(sumsq), max_norm,
         _stream())


# ------------------------------------------------------------------------------------------------ validation stats
def feature_absmax(acts, idx, n: int):
    out = torch.empty(n, dtype=torch.float32, device=acts.device)
    if idx.dtype not in (torch.int64, torch.int32):
        raise TypeError("indices must be int64 or int32")
    call("freud_feature_absmax", _ptr(_f32(acts, "acts")), _ptr(idx), int(idx.dtype == torch.int64), acts.numel(),
         _ptr(out), n, _stream())
    return out


def col_absmax(x):
    rows, n = x.shape
    out = torch.empty(n, dtype=torch.float32, device=x.device)
    call("freud_col_absmax", _ptr(_f32(x, "x")), rows, n, _ptr(out), _stream())
    return out


# ------------------------------------------------------------------------------------------------ search
def search_dense(acts, n_frames, feature: int, want_trace: bool):
    n_files, T, F = acts.shape
    dev = acts.device
    vmax = torch.empty(n_files, dtype=torch.float32, device=dev)
    amax = torch.empty(n_files, dtype=torch.int32, device=dev)
    vabs = torch.empty(n_files, dtype=torch.float32, device=dev)
    trace = torch.empty((n_files, T), dtype=torch.float32, device=dev) if want_trace else None
    if acts.dtype not in (torch.float32, torch.float16):
        raise TypeError("dense activations must be float32 or float16")
    call("freud_search_dense", _ptr(acts), int(acts.dtype == torch.float16), _ptr(n_frames), n_files, T, F, feature,
         _ptr(vmax), _ptr(amax), _ptr(vabs), _ptr(trace), _stream())
    return vmax, amax, vabs, trace


def search_table_dense(acts, n_frames):
    """(vmax, amax, vabs) tables [n_files, F] for every feature column in one pass over the store."""
    n_files, T, F = acts.shape
    dev = acts.device
    vmax = torch.empty((n_files, F), dtype=torch.float32, device=dev)
    amax = torch.empty((n_files, F), dtype=torch.int32, device=dev)
    vabs = torch.empty((n_files, F), dtype=torch.float32, device=dev)
    if acts.dtype not in (torch.float32, torch.float16):
        raise TypeError("dense activations must be float32 or float16")
    call("freud_search_table_dense", _ptr(acts), int(acts.dtype == torch.float16), _ptr(n_frames), n_files, T, F,
         _ptr(vmax), _ptr(amax), _ptr(vabs), _stream())
    return vmax, amax, vabs


def search_indexed(vals, idx, n_frames, feature: int, want_trace: bool):
    n_files, T, k = vals.shape
    dev = vals.device
    vmax = torch.empty(n_files, dtype=torch.float32, device=dev)
    amax = torch.empty(n_files, dtype=torch.int32, device=dev)
    vabs = torch.empty(n_files, dtype=torch.float32, device=dev)
    trace = torch.empty((n_files, T), dtype=torch.float32, device=dev) if want_trace else None
    if idx.dtype not in (torch.int64, torch.int32):
        raise TypeError("feature indices must be int64 or int32")
    call("freud_search_indexed", _ptr(_f32(vals, "vals")), _ptr(idx), int(idx.dtype == torch.int64), _ptr(n_frames),
         n_files, T, k, feature, _ptr(vmax), _ptr(amax), _ptr(vabs), _ptr(trace), _stream())
    return vmax, amax, vabs, trace


def search_topn(vmax, vabs, absolute: bool, min_val, max_val, n_top: int):
    dev = vmax.device
    out = torch.empty(n_top, dtype=torch.int32, device=dev)
    cnt = torch.empty(1, dtype=torch.int32, device=dev)
    call("freud_search_topn", _ptr(vmax), _ptr(vabs), vmax.numel(), int(absolute), int(min_val is not None),
         float(min_val if min_val is not None else 0.0), int(max_val is not None),
         float(max_val if max_val is not None else 0.0), n_top, _ptr(out), _ptr(cnt), _stream())
    return out, cnt
