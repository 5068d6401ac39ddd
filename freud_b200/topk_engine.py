"""Kernel orchestration of the TopK SAE forward / backward (reference: TopKAutoEncoder.forward,
src/models/topkautoencoder.py:93-151, and its autograd; formulas in SURVEY.md M5-M8').

Two routes through the same C ABI:
  fast    : k == 32, no live dead mask, no multi-TopK -- fused tcgen05 GEMM + top-k epilogue, the [N,n]
            pre-activations never exist.
  generic : AuxK and/or multi-TopK and/or k != 32 -- pre-activations materialised by the store-epilogue GEMM,
            selections by the radix row top-k kernel, one sparse backward per decode.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import ops
from ._lib import BF16, FP32


@dataclass
class TopKState:
    """Everything the backward needs (all device tensors)."""
    precision: int
    x2: torch.Tensor            # [N,d] fp32 view of the input
    xc_hi: torch.Tensor         # encoder A operand (bf16, or tf32-hi)
    wd: torch.Tensor            # decoder weight as gathered (bf16 copy or the fp32 parameter)
    W_enc: torch.Tensor         # encoder weight as the db_dec term reads it (bf16 copy in bf16 mode, else fp32)
    b_dec: torch.Tensor
    k: int
    n: int
    scal: torch.Tensor          # [fvu, mse, 2/tv, 2/tv, tv] fp32
    generic: bool
    vals: torch.Tensor = None   # k-selection behind fvu
    idx: torch.Tensor = None
    e: torch.Tensor = None      # residual sae_out - x (bf16 in the bf16 fast path, else fp32)
    colsum_e: torch.Tensor = None
    aux: Optional[tuple] = None     # (a_vals, a_idx, resid_aux fp32, scale)   row-sparse AuxK
    aux_dense: Optional[tuple] = None  # (dead_idx, S, Sp, A bf16 [N,Sp], W_dec[dead] bf16, resid_aux fp32, scale)
    multi: Optional[tuple] = None   # (m_vals, m_idx, resid_m fp32, colsum_m)
    auxk_alpha: float = 0.0
    offsets: Optional[torch.Tensor] = None  # CSC offsets of the returned encoding (did_fire bookkeeping)
    scal_ready: Optional[torch.cuda.Event] = None  # set when `scal` is still being produced on a side stream
    csc_ready: Optional[torch.cuda.Event] = None   # recorded once `offsets` is complete on the main stream
    csc: Optional[tuple] = None  # (offsets, entries, event) of the main selection, built on the side stream
    dacts: Optional[torch.Tensor] = None  # <e, W_dec[idx]> when the forward already formed it (fused decode + dacts)
    extra: dict = field(default_factory=dict)


@dataclass
class TopKResult:
    sae_out: torch.Tensor       # [N,d] fp32 (of the 4k selection when multi_topk, as the reference rebinds it)
    top_acts: torch.Tensor      # [N,k'] fp32
    top_idx: torch.Tensor       # [N,k'] int32
    fvu: torch.Tensor           # 0-d fp32
    auxk_loss: torch.Tensor     # 0-d fp32, already times auxk_alpha
    multi_topk_fvu: torch.Tensor
    mse: torch.Tensor


_aux_streams = {}


def aux_stream(device):
    """Side stream for work that depends only on the selected indices (the CSC index of the backward): it runs beside
    the decode / activation-gradient kernels instead of between them."""
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    s = _aux_streams.get(key)
    if s is None:
        s = _aux_streams[key] = torch.cuda.Stream(torch.device("cuda", key))
    return s


def encode_operands(x, W_enc, b_dec, precision, dp=None, we_hi=None):
    """we_hi: the bf16 copy of W_enc when the caller maintains one (the fused Adam kernel writes it next to the fp32
    parameter, so the per-step conversion pass disappears); bf16 mode only."""
    if dp is None:
        xc_hi, xc_lo, tv = ops.topk_prep_x(x, b_dec, precision)
    else:
        # data parallel: variance of the CONCATENATED batch = sum_r tv_r + sum_r B_r * sum (mean_r - mean)^2
        xc_hi, xc_lo, tv, colmean = ops.topk_prep_x(x, b_dec, precision, want_colmean=True)
        side = dp.side_stream()
        if side is None:
            tv = dp.global_total_variance(tv, colmean, x.shape[0])
        else:  # the two small allreduces + their torch kernels run beside the encoder; joined before the loss scalars
            side.wait_stream(torch.cuda.current_stream())
            tv.record_stream(side)
            colmean.record_stream(side)
            with torch.cuda.stream(side):
                tv = dp.global_total_variance(tv, colmean, x.shape[0])
    if we_hi is not None and precision == BF16:
        we_lo = None
    else:
        we_hi, we_lo = ops.split_operand(W_enc, precision)
    return xc_hi, xc_lo, we_hi, we_lo, tv


def topk_forward(x, W_enc, b_enc, W_dec, b_dec, k, *, precision, dead_mask=None, auxk_alpha=0.0, multi_topk=False,
                 need_grad=True, dp=None, defer_scal=False, num_dead=None, shadows=None):
    """dp: optional freud_b200.parallel.DataParallel -- makes tv / sse (and with them every loss and gradient
    scale) those of the batch concatenated over ranks; gradients stay rank-local sums for the caller to allreduce.
    defer_scal (data parallel, fused main path only): the loss scalars are produced on dp's side stream and the
    main stream is NOT joined; st.scal_ready is the event to wait for before reading them (topk_backward does).
    num_dead: `int(dead_mask.sum())` when the caller already holds it (the trainer reads it back asynchronously during
    the previous step); otherwise it is read here, which synchronises with the device like the reference (:109).
    shadows: {"encoder.weight": bf16 [n,d], "W_dec": bf16 [n,d]} up-to-date bf16 copies of the weights (bf16 mode)."""
    if x.dim() != 3:
        raise ValueError("x must be [B, T, d]")
    x = x.contiguous()
    B, T, d = x.shape
    N, n = B * T, W_enc.shape[0]
    # reference: `int(dead_mask.sum())` (topkautoencoder.py:109) -- one 8-byte device->host read, as upstream
    if num_dead is None:
        num_dead = int(dead_mask.sum()) if dead_mask is not None else 0
    fused_main = (k == ops.K_FUSED) and not multi_topk
    generic = not fused_main or num_dead > 0  # backward needs materialised gradient seeds (AuxK couples e_hat and e)
    fused_decode = (fused_main and need_grad and not generic and precision == BF16 and ops.decode_dacts_supported(d, k))
    # fp16 / bf16 activations (what collected stores hold) feed the fused bf16 path as they are -- the two kernels that
    # read x widen it exactly -- every other route works on an fp32 copy
    if x.dtype != torch.float32 and not fused_decode:
        x = x.float()
    x2 = x.view(N, d)
    shadows = shadows if (shadows and precision == BF16) else {}
    xc_hi, xc_lo, we_hi, we_lo, tv = encode_operands(x, W_enc, b_dec, precision, dp, shadows.get("encoder.weight"))
    if precision == BF16:
        wd = shadows["W_dec"] if "W_dec" in shadows else ops.split_operand(W_dec, BF16)[0]
    else:
        wd = W_dec
    zero = torch.zeros((), dtype=torch.float32, device=x.device)

    pre = None
    side_csc = (fused_main and need_grad and not generic and os.environ.get("FREUD_CSC_MODE", "side") != "serial")
    counts = None
    if fused_main:
        if side_csc:  # the encoder counts the entries per feature while it writes its rows: the CSC build starts there
            counts = torch.zeros(n + 1, dtype=torch.int32, device=x.device)
        vals, idx = ops.topk_encode(xc_hi, xc_lo, we_hi, we_lo, b_enc, precision, hist=counts)
        if precision == FP32:
            ops.topk_refine(x2, b_dec, W_enc, b_enc, idx, vals)
    else:
        pre = ops.gemm_nt(xc_hi, xc_lo, we_hi, we_lo, b_enc, True, precision)
        vals, idx = ops.row_topk(pre, k)

    csc = None
    if side_csc:
        # the feature-major (CSC) index of the backward depends on the indices alone: built on a side stream while
        # the main stream decodes and forms the activation gradients
        cur, side = torch.cuda.current_stream(), aux_stream(x.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            offsets, entries = ops.csc_build(idx, n, counts=counts)
            csc_done = side.record_event()
        idx.record_stream(side)
        offsets.record_stream(cur)
        entries.record_stream(cur)
        csc = (offsets, entries, csc_done)

    resid_dtype = None
    if generic:
        resid_dtype = torch.float32
    elif need_grad:
        resid_dtype = torch.bfloat16 if precision == BF16 else torch.float32
    dacts = None
    if fused_decode:
        # the decoder rows of a token are gathered once for the reconstruction AND the activation gradients
        sae_out, e, sse, colsum_e, dacts = ops.topk_decode_dacts(vals, idx, wd, b_dec, x2)
    else:
        sae_out, e, sse, colsum_e = ops.topk_decode(vals, idx, wd, b_dec, x2, resid_dtype=resid_dtype, want_sse=True,
                                                    want_colsum=need_grad)
    numel = N * d
    scal_ready = None
    side = dp.side_stream() if dp is not None else None
    if dp is not None:
        numel *= dp.world_size
    if side is None:
        if dp is not None:
            sse = dp.all_reduce_sum(sse)
        scal = ops.topk_loss_scalars(sse, tv, numel)
    else:
        # SSE allreduce + loss scalars on the side stream (which already holds the global variance); the main
        # stream joins here, or -- defer_scal -- only where the backward first needs the gradient scale
        cur = torch.cuda.current_stream()
        side.wait_stream(cur)
        sse.record_stream(side)
        with torch.cuda.stream(side):
            sse = dp.all_reduce_sum(sse)
            scal = ops.topk_loss_scalars(sse, tv, numel)
            scal_ready = side.record_event()
        scal.record_stream(cur)
        tv.record_stream(cur)
        if not (defer_scal and not generic):
            cur.wait_event(scal_ready)
            scal_ready = None
    # the db_dec term back-propagates through the matmul operand: the bf16 copy in bf16 mode, as under the
    # reference's autocast (and the only copy of W_enc that is current on every rank under the fused DP optimiser)
    st = TopKState(precision, x2, xc_hi, wd, we_hi if precision == BF16 else W_enc, b_dec, k, n, scal, generic, vals,
                   idx, e, colsum_e, auxk_alpha=auxk_alpha)
    st.scal_ready = scal_ready
    st.csc = csc
    st.dacts = dacts

    auxk = zero
    if num_dead > 0:
        k_aux = d // 2
        scale = min(num_dead / k_aux, 1.0)
        k_aux = min(k_aux, num_dead)
        if pre is not None:
            a_vals, a_idx = ops.row_topk(pre, k_aux, col_mask=dead_mask)
        elif precision == BF16:
            # Dense AuxK on the compacted dead-latent subset (S columns): k_aux = d/2 selections per token make
            # the row-sparse kernels ~12x the main path's work, but as dense [N,S] x [S,d] products it is cheap.
            # No operand is transposed in memory: products over the token axis read their operands through
            # MN-major tensor-core descriptors (gemm_nn / gemm_tn_splitk).
            dead_idx = torch.nonzero_static(dead_mask, size=num_dead).squeeze(1).to(torch.int32)  # no host sync
            S = num_dead
            Sp = (S + 7) // 8 * 8
            ws = ops.gather_rows(we_hi, dead_idx)
            bs = ops.gather_rows(b_enc, dead_idx)
            pre_dead = torch.empty((N, Sp), dtype=torch.float32, device=x.device)          # row pitch Sp, S valid columns
            ops.gemm_nt(xc_hi, None, ws, None, bs, True, precision, out=pre_dead)
            A = ops.row_topk_mask(pre_dead, k_aux, Sp, n=S, nonneg=True)                              # [N,Sp] bf16, top-k_aux kept
            del pre_dead
            wd_sub = ops.gather_rows(wd, dead_idx)                                       # [S,d] bf16
            e_hat = ops.gemm_nn(A, wd_sub, b_dec, False, K=S)                            # A @ W_dec[dead] + b_dec
            r_aux, sse_aux, _ = ops.residual(e_hat, e, torch.float32, want_colsum=False)
            st.aux_dense = (dead_idx, S, Sp, A, wd_sub, r_aux, scale)
            a_vals = a_idx = None
        else:
            # dead latents are a column subset: GEMM against their compacted encoder rows, select among them, map
            # the subset-local indices back (same (value desc, index asc) order as the masked full-width top-k)
            dead_idx = torch.nonzero(dead_mask).squeeze(1).to(torch.int32)
            ws_hi = ops.gather_rows(we_hi, dead_idx)
            ws_lo = ops.gather_rows(we_lo, dead_idx) if we_lo is not None else None
            bs = ops.gather_rows(b_enc, dead_idx)
            pre_dead = ops.gemm_nt(xc_hi, xc_lo, ws_hi, ws_lo, bs, True, precision)
            a_vals, a_loc = ops.row_topk(pre_dead, k_aux)
            a_idx = ops.index_map(dead_idx, a_loc)
        if a_vals is not None:
            _, r_aux, sse_aux, _ = ops.topk_decode(a_vals, a_idx, wd, b_dec, e, resid_dtype=torch.float32,
                                                   want_sse=True)
            st.aux = (a_vals, a_idx, r_aux, scale)
        if dp is not None:
            sse_aux = dp.all_reduce_sum(sse_aux)
        auxk = (scale * sse_aux[0] / scal[4].double()).float() * auxk_alpha

    mfvu = zero
    ret_out, ret_vals, ret_idx = sae_out, vals, idx
    if multi_topk:
        m_vals, m_idx = ops.row_topk(pre, 4 * k)
        m_out, r_m, sse_m, colsum_m = ops.topk_decode(m_vals, m_idx, wd, b_dec, x2, resid_dtype=torch.float32,
                                                      want_sse=True, want_colsum=need_grad)
        if dp is not None:
            sse_m = dp.all_reduce_sum(sse_m)
        mfvu = (sse_m[0] / scal[4].double()).float()
        st.multi = (m_vals, m_idx, r_m, colsum_m)
        ret_out, ret_vals, ret_idx = m_out, m_vals, m_idx
    res = TopKResult(ret_out, ret_vals, ret_idx, scal[0], auxk, mfvu, scal[1])
    return res, st


def topk_backward(st: TopKState, g_fvu, g_aux=None, g_multi=None, *, out=None, on_offsets=None, g_out=None,
                  g_acts=None):
    """Parameter gradients of g_fvu*fvu + g_aux*auxk_loss + g_multi*multi_topk_fvu (+ <g_out, sae_out> +
    <g_acts, top_acts> when the caller built its own loss on the returned reconstruction / activations: those two
    go through the generic row-sparse route and belong to the RETURNED encoding, i.e. the 4k one under multi-TopK).
    g_* are 0-d device tensors (or python floats).  Returns dict name -> gradient tensor.
    `out` may supply preallocated gradient buffers {name: tensor}.  on_offsets(offsets) is called as soon as the CSC
    offsets of the RETURNED encoding are enqueued (the did_fire bookkeeping can then run beside the rest of the
    backward instead of after it)."""
    dev = st.x2.device
    n, d, k = st.n, st.x2.shape[1], st.k
    out = out or {}
    dW_enc = out.get("encoder.weight") if "encoder.weight" in out else torch.empty((n, d), dtype=torch.float32, device=dev)
    dW_dec = out.get("W_dec") if "W_dec" in out else torch.empty((n, d), dtype=torch.float32, device=dev)
    db_enc = out.get("encoder.bias") if "encoder.bias" in out else torch.empty(n, dtype=torch.float32, device=dev)
    db_dec = out.get("b_dec") if "b_dec" in out else torch.empty(d, dtype=torch.float32, device=dev)
    bf16 = st.precision == BF16
    xc = st.xc_hi if bf16 else st.x2
    two_over_tv = st.scal[2]

    extra_seed = g_out is not None or g_acts is not None
    if not st.generic and not extra_seed:
        if st.scal_ready is not None and not (isinstance(g_fvu, float) and g_fvu == 1.0):
            torch.cuda.current_stream().wait_event(st.scal_ready)  # the scale is about to be read by a torch kernel
            st.scal_ready = None
        if isinstance(g_fvu, torch.Tensor):
            scales = st.scal[2:4] * g_fvu.to(torch.float32)
        elif g_fvu == 1.0:
            scales = st.scal[2:4]  # (2/tv, 2/tv) straight from the loss-scalar kernel, no extra launch
        else:
            scales = st.scal[2:4] * float(g_fvu)
        dacts = st.dacts if st.dacts is not None else ops.topk_dacts(st.e, st.idx, st.wd)
        if st.csc is not None:
            offsets, entries, csc_done = st.csc
            torch.cuda.current_stream().wait_event(csc_done)
            st.csc_ready = csc_done
        else:
            offsets, entries = ops.csc_build(st.idx, n)
            st.csc_ready = torch.cuda.current_stream().record_event()
        st.offsets = offsets
        if on_offsets is not None:
            on_offsets(offsets)
        if st.scal_ready is not None:  # gradient scale produced on the data-parallel side stream
            torch.cuda.current_stream().wait_event(st.scal_ready)
            st.scal_ready = None
        ops.topk_sparse_grads(offsets, entries, st.vals, dacts, st.e, xc, st.b_dec, scales, dW_dec, dW_enc, db_enc,
                              k, False)
        ops.topk_bdec_grad(st.colsum_e, scales, db_enc, st.W_enc, db_dec, False)
    else:
        def as_t(g):
            if g is None:
                return torch.zeros((), dtype=torch.float32, device=dev)
            return g.to(torch.float32) if isinstance(g, torch.Tensor) else torch.tensor(float(g), device=dev)

        if st.scal_ready is not None:
            torch.cuda.current_stream().wait_event(st.scal_ready)
            st.scal_ready = None
        if st.e.dtype != torch.float32:  # fused forward kept the residual in bf16; this route seeds from fp32
            st.e = st.e.float()
        gdt = torch.bfloat16 if bf16 else torch.float32
        c_main = as_t(g_fvu) * two_over_tv
        zero = torch.zeros((), dtype=torch.float32, device=dev)
        one = torch.ones((), dtype=torch.float32, device=dev)
        decodes = []  # (tag, vals, idx, G)
        db_direct = c_main * st.colsum_e
        dense_aux = None
        if st.aux_dense is not None:
            dead_idx, S, Sp, A, wd_sub, r_aux, scale = st.aux_dense
            c_aux = as_t(g_aux) * (st.auxk_alpha * scale) * two_over_tv
            G_main = ops.axpby(st.e, r_aux, torch.stack((c_main, -c_aux)), gdt)  # e is not detached (:126)
            dense_aux = (dead_idx, S, Sp, A, wd_sub, ops.axpby(r_aux, None, torch.stack((c_aux, zero)), gdt))
        elif st.aux is not None:
            a_vals, a_idx, r_aux, scale = st.aux
            # auxk_loss already carries auxk_alpha, so d auxk_loss / d e_hat = alpha*scale*2(e_hat - e)/tv
            c_aux = as_t(g_aux) * (st.auxk_alpha * scale) * two_over_tv
            G_main = ops.axpby(st.e, r_aux, torch.stack((c_main, -c_aux)), gdt)  # e is not detached (:126)
            decodes.append(("aux", a_vals, a_idx, ops.axpby(r_aux, None, torch.stack((c_aux, zero)), gdt)))
        else:
            G_main = ops.axpby(st.e, None, torch.stack((c_main, zero)), gdt)
        decodes.insert(0, ("main", st.vals, st.idx, G_main))
        if st.multi is not None:
            m_vals, m_idx, r_m, colsum_m = st.multi
            c_m = as_t(g_multi) * two_over_tv
            decodes.append(("multi", m_vals, m_idx, ops.axpby(r_m, None, torch.stack((c_m, zero)), gdt)))
            db_direct = db_direct + c_m * colsum_m
        ones = torch.ones(2, dtype=torch.float32, device=dev)
        first = True
        returned = "multi" if st.multi is not None else "main"
        if g_out is not None:  # caller's own loss on sae_out: one more seed on the returned decode's output
            g2 = g_out.reshape(-1, d).to(torch.float32).contiguous()
            db_direct = db_direct + g2.sum(0)
            decodes = [(tag, vals, idx, ops.axpby(G.float() if G.dtype != torch.float32 else G, g2,
                                                  torch.stack((one, one)), gdt) if tag == returned else G)
                       for tag, vals, idx, G in decodes]
        for tag, vals, idx, G in decodes:
            dacts = ops.topk_dacts(G, idx, st.wd)
            if g_acts is not None and tag == returned:
                dacts += g_acts.reshape(dacts.shape).to(torch.float32)
            offsets, entries = ops.csc_build(idx, n)
            if tag == returned:
                st.offsets = offsets  # did_fire is taken from the RETURNED encoding (train_sae.py:442)
                st.csc_ready = torch.cuda.current_stream().record_event()
                if on_offsets is not None:
                    on_offsets(offsets)
            ops.topk_sparse_grads(offsets, entries, vals, dacts, G, xc, st.b_dec, ones, dW_dec, dW_enc, db_enc,
                                  idx.shape[1], not first)
            first = False
        if dense_aux is not None:
            # dense backward of the AuxK branch on the dead subset: three tensor-core products, then a row scatter
            dead_idx, S, Sp, A, wd_sub, G_hat = dense_aux
            dpre = ops.gemm_nt_mask(G_hat, wd_sub, A)               # (A > 0) ? G_hat @ W_dec[dead]^T : 0   bf16 [N,Sp]
            db_sub = ops.col_sum_bf16(dpre, S)
            dWdec_sub = ops.gemm_tn_splitk(A, G_hat, M=S, N=d)      # A^T @ G_hat            [S,d]
            dWenc_sub = ops.gemm_tn_splitk(dpre, st.xc_hi, M=S, N=d)  # dpre^T @ (x - b_dec)   [S,d]
            ops.scatter_add_rows(dWdec_sub, dead_idx, dW_dec)
            ops.scatter_add_rows(dWenc_sub, dead_idx, dW_enc)
            ops.scatter_add_rows(db_sub, dead_idx, db_enc)
        db_dec.copy_(db_direct)
        ops.topk_bdec_grad(None, None, db_enc, st.W_enc, db_dec, True)
    return {"encoder.weight": dW_enc, "encoder.bias": db_enc, "W_dec": dW_dec, "b_dec": db_dec}
