"""GPU parity of the TopK SAE path: CUDA kernels (through the C ABI) vs the reference goldens and the oracle."""
import pytest
import torch

from oracle import sae as osae
from tests.util import load_golden, rel_err, sets_equal_rows, t

pytestmark = pytest.mark.gpu

TOPK_KEYS = ["encoder.weight", "encoder.bias", "W_dec", "b_dec"]


def _model_from_golden(z, meta, prefix="init"):
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder

    cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": meta["n"], "k": meta["k"],
                                           "multi_topk": meta["multi_topk"], "auxk_alpha": meta["auxk_alpha"]})
    model = TopKAutoEncoder(meta["d"], cfg)
    model.load_state_dict({k: t(z[f"{prefix}.{k}"]) for k in TOPK_KEYS})
    return model.cuda()


@pytest.mark.parametrize("name", ["topk_fp32", "topk_fp32_auxk", "topk_fp32_auxk_few", "topk_fp32_multi",
                                  "topk_fp32_b1"])
def test_trainer_follows_reference_trajectory_fp32(name):
    """SAETrainer.step == body of train_sae.py:421-453 run by the reference (fp32, 1e-5 relative)."""
    from freud_b200.trainer import SAETrainer

    z, meta = load_golden(name)
    model = _model_from_golden(z, meta)
    tr = SAETrainer(model, lr=meta["lr"], steps=meta["total_steps"], clip_thresh=meta["clip"], optimizer="adam",
                    scheduler="linear", scheduler_params={"num_warmup_steps": meta["warmup"]},
                    dead_feature_threshold=meta["dead_thresh"], precision="fp32")
    tr.tokens_seen = 10 ** 12  # the goldens start with pre-aged counters
    for s in range(meta["steps"]):
        # step 0 is held to the north_star's 1e-5; later steps inherit last-bit parameter differences that Adam's
        # m/sqrt(v) normalisation amplifies (an update is ~lr*sign(g) wherever |g| is tiny), hence 1e-4 on params
        gtol, ptol = (1e-5, 1e-5) if s == 0 else (3e-5, 1e-4)
        x = t(z[f"s{s}.x"]).cuda()
        tr.num_frames_since_fired.copy_(t(z[f"s{s}.frames_in"]))
        assert abs(tr.optimizer.param_groups[0]["lr"] - float(z[f"s{s}.lr"])) < 1e-12
        out = tr.step(x)
        torch.cuda.synchronize()
        for key in ("fvu", "auxk_loss", "multi_topk_fvu"):
            assert rel_err(out[key].cpu(), z[f"s{s}.{key}"]) < gtol or abs(float(out[key]) - float(z[f"s{s}.{key}"])) < 1e-9, key
        assert rel_err(out["sae_out"].cpu().view(z[f"s{s}.sae_out"].shape), z[f"s{s}.sae_out"]) < gtol
        for k in TOPK_KEYS:
            assert rel_err(tr.params[k].grad.cpu(), z[f"s{s}.grad.{k}"]) < gtol, f"grad {k} step {s}"
        assert rel_err(out["grad_sumsq"].sqrt().cpu(), z[f"s{s}.grad_norm"]) < gtol
        for k in TOPK_KEYS:
            assert rel_err(tr.params[k].data.cpu(), z[f"s{s}.param.{k}"]) < ptol, f"param {k} step {s}"
        assert torch.equal(tr.num_frames_since_fired.cpu(), t(z[f"s{s}.frames_out"])), "dead-latent counters"


def test_module_autograd_dropin_fp32():
    """model(x, dead_mask) -> loss.backward() -> clip_grad_norm_ -> FusedAdam.step, as train_sae.py calls them."""
    from freud_b200.optim import FusedAdam, clip_grad_norm_

    z, meta = load_golden("topk_fp32_auxk")
    model = _model_from_golden(z, meta)
    opt = FusedAdam(model.parameters(), lr=float(z["s0.lr"]))
    x = t(z["s0.x"]).cuda()
    dead = (t(z["s0.frames_in"]) > meta["dead_thresh"]).cuda()
    out, mse = model(x, dead_mask=dead, return_mse=True)
    loss = out.fvu + out.auxk_loss + out.multi_topk_fvu / 8
    loss.backward()
    assert out.encoded.top_indices.dtype == torch.int64
    assert rel_err(loss.detach().cpu(), z["s0.loss"]) < 1e-5
    assert rel_err(mse.cpu(), z["s0.mse"]) < 1e-5
    free = osae.tie_free_rows(t(z["s0.pre_acts"]), meta["k"]).reshape(-1)
    same = sets_equal_rows(out.encoded.top_indices.cpu(), z["s0.top_indices"])
    assert bool(same[free].all())
    named = dict(model.named_parameters())
    for k in TOPK_KEYS:
        assert rel_err(named[k].grad.cpu(), z[f"s0.grad.{k}"]) < 1e-5, k
    total = clip_grad_norm_(model.parameters(), meta["clip"])
    assert rel_err(total.cpu(), z["s0.grad_norm"]) < 1e-5
    opt.step()
    for k in TOPK_KEYS:
        assert rel_err(named[k].data.cpu(), z[f"s0.param.{k}"]) < 1e-5, k


def test_bf16_mode_vs_reference_autocast():
    """bf16 tensor-core mode vs the reference under autocast('cpu'): 2e-2 (north_star), same selected sets."""
    z, meta = load_golden("topk_bf16")
    model = _model_from_golden(z, meta)
    model.precision = "bf16"
    x = t(z["s0.x"]).cuda()
    out = model(x)
    loss = out.fvu + out.auxk_loss + out.multi_topk_fvu / 8
    loss.backward()
    assert bool(sets_equal_rows(out.encoded.top_indices.cpu(), z["s0.top_indices"]).all())
    assert rel_err(out.fvu.detach().cpu(), z["s0.fvu"]) < 2e-2
    assert rel_err(out.sae_out.cpu(), z["s0.sae_out"]) < 2e-2
    named = dict(model.named_parameters())
    for k in TOPK_KEYS:
        assert rel_err(named[k].grad.cpu(), z[f"s0.grad.{k}"]) < 2e-2, k


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 4e-3)])
@pytest.mark.parametrize("shape", [(3, 200, 64, 512), (2, 129, 384, 6144)])
def test_fused_fast_path_vs_oracle(precision, tol, shape):
    """k == 32 fused tcgen05 GEMM + top-k epilogue, sparse decode and CSC backward vs the oracle in the same
    precision mode: index sets equal on every row (rows that tie within GEMM rounding are reported, see
    tests/test_gpu_bench_shapes.py), values, loss, reconstruction and all gradients compared on every row."""
    from tests.test_gpu_bench_shapes import check_fused_step

    check_fused_step(shape, precision, tol)


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 4e-3)])
@pytest.mark.parametrize("n_dead", [7, 150])
def test_fused_main_with_live_auxk_vs_oracle(precision, tol, n_dead):
    """k == 32 with dead latents: fused encoder for the main selection, AuxK on the compacted dead-latent
    subset (GEMM against gathered encoder rows), coupled backward (e not detached, topkautoencoder.py:126)."""
    from freud_b200 import topk_engine
    from freud_b200._lib import BF16, FP32

    B, T, d, n, k = 2, 150, 64, 512, 32
    g = torch.Generator().manual_seed(9)
    x = torch.randn(B, T, d, generator=g) * (0.5 + torch.rand(T, 1, generator=g))
    W_enc = torch.randn(n, d, generator=g) / d ** 0.5
    W_dec = osae.set_decoder_norm_to_unit_norm(W_enc.clone() + 0.1 * torch.randn(n, d, generator=g))
    b_enc = 0.05 * torch.randn(n, generator=g)
    b_dec = 0.1 * torch.randn(d, generator=g)
    dead = torch.zeros(n, dtype=torch.bool)
    dead[torch.randperm(n, generator=g)[:n_dead]] = True
    alpha = 1 / 32
    ref = osae.topk_forward(x, W_enc, b_enc, W_dec, b_dec, k, dead_mask=dead, auxk_alpha=alpha, mode=precision)
    rg = osae.topk_backward(x, W_enc, b_enc, W_dec, b_dec, ref, k, auxk_alpha=alpha, mode=precision)
    prec = BF16 if precision == "bf16" else FP32
    cu = [v.cuda() for v in (x, W_enc, b_enc, W_dec, b_dec)]
    res, st = topk_engine.topk_forward(*cu, k, precision=prec, dead_mask=dead.cuda(), auxk_alpha=alpha)
    assert st.aux is not None or st.aux_dense is not None
    grads = topk_engine.topk_backward(st, 1.0, 1.0, 0.125)
    torch.cuda.synchronize()
    same_main = sets_equal_rows(res.top_idx.cpu(), ref.top_indices.reshape(-1, k))
    ka = ref.aux[1].shape[-1]
    if st.aux is not None:
        same_aux = sets_equal_rows(st.aux[1].cpu(), ref.aux[1].reshape(-1, ka))
    else:
        # dense AuxK (bf16 mode): the selection lives in the non-zero pattern of A [N, S]; compare the sets of
        # POSITIVE selected latents (zero-valued selections are indistinguishable and contribute nothing)
        dead_idx, S, Sp, A = st.aux_dense[:4]
        got = torch.zeros(B * T, n, dtype=torch.bool)
        got[:, dead_idx.cpu().long()] = A[:, :S].float().cpu() > 0
        want = torch.zeros(B * T, n, dtype=torch.bool)
        ra, ri = ref.aux[0].reshape(-1, ka), ref.aux[1].reshape(-1, ka)
        want.scatter_(1, ri, ra > 0)
        same_aux = (got == want).all(-1)
    # seeded inputs without ties at either selection boundary: every row must agree, nothing is skipped
    assert bool(same_main.all()), f"main selection differs on {int((~same_main).sum())} rows"
    assert bool(same_aux.all()), f"AuxK selection differs on {int((~same_aux).sum())} rows"
    assert rel_err(res.fvu.cpu(), ref.fvu) < max(tol, 1e-5)
    assert rel_err(res.auxk_loss.cpu(), ref.auxk_loss) < tol
    for key in TOPK_KEYS:
        assert rel_err(grads[key].cpu(), rg[key]) < tol, key


def test_short_rows_and_ragged_sizes():
    """Rows with fewer than 32 positive pre-activations are completed with zeros at the lowest free indices
    (oracle order); N not a multiple of 128 and n not a multiple of 256 exercise the TMA out-of-bounds fill."""
    from freud_b200 import ops
    from freud_b200._lib import FP32

    g = torch.Generator().manual_seed(11)
    N, d, n = 77, 40, 600
    x = torch.randn(1, N, d, generator=g)
    W = torch.randn(n, d, generator=g) / d ** 0.5
    b_enc = torch.full((n,), -2.5)  # strongly negative bias -> only a handful of positives per row
    b_dec = torch.zeros(d)
    ref = osae.topk_pre_acts(x, W, b_enc, b_dec)[0]
    rv, ri = osae.select_topk(ref, 32)
    assert int((rv > 0).sum(-1).min()) < 32
    xc_hi, xc_lo, tv = ops.topk_prep_x(x.cuda(), b_dec.cuda(), FP32)
    w_hi, w_lo = ops.split_operand(W.cuda(), FP32)
    vals, idx = ops.topk_encode(xc_hi, xc_lo, w_hi, w_lo, b_enc.cuda(), FP32)
    assert torch.equal(torch.sort(idx.cpu().long(), -1).values, torch.sort(ri, -1).values)
    assert rel_err(torch.sort(vals.cpu(), -1, descending=True).values, rv) < 1e-5


@pytest.mark.parametrize("shape,bias", [((300, 64, 2560), -2.5), ((300, 64, 2560), 0.0), ((19109, 64, 8192), 0.0),
                                        ((1000, 128, 4000), -3.0)])
def test_tail_wave_column_split_is_exact(shape, bias, monkeypatch):
    """Row blocks of a partial last wave are scanned in column pieces by several CTAs (shared thresholds, partial
    lists, merge kernel).  The result must be identical -- values, indices and order -- to the unsplit scan, and
    equal to the oracle's selection, including short rows whose zero fillers come from different pieces."""
    from freud_b200 import ops
    from freud_b200._lib import BF16

    N, d, n = shape
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, N, d, generator=g)
    W = torch.randn(n, d, generator=g) / d ** 0.5
    b_enc = (bias + 0.05 * torch.randn(n, generator=g)).cuda()
    b_dec = torch.zeros(d)
    xc_hi, _, _ = ops.topk_prep_x(x.cuda(), b_dec.cuda(), BF16)
    w_hi, _ = ops.split_operand(W.cuda(), BF16)
    need = ops.topk_encode_workspace_bytes(N, n)
    assert need > 0, "shape does not exercise the split"
    vals, idx = ops.topk_encode(xc_hi, None, w_hi, None, b_enc, BF16)
    monkeypatch.setenv("FREUD_NO_TAIL_SPLIT", "1")
    assert ops.topk_encode_workspace_bytes(N, n) == 0
    vals1, idx1 = ops.topk_encode(xc_hi, None, w_hi, None, b_enc, BF16)
    torch.cuda.synchronize()
    assert torch.equal(vals, vals1) and torch.equal(idx, idx1)
    # oracle selection on the same bf16 operands (fp32 accumulate): exact wherever the k-th gap is clear
    # (oracle.sae.select_topk semantics -- stable descending sort -- evaluated with torch on the device: the
    #  [19109, 8192] case is too large for the CPU suite's time budget)
    pre = torch.relu(xc_hi.float() @ w_hi.float().T + b_enc)
    srt, order = torch.sort(pre, dim=-1, descending=True, stable=True)
    rv, ri = srt[:, :32].cpu(), order[:, :32].cpu()
    srt = srt[:, :33].cpu()
    clear = ((srt[:, 31] - srt[:, 32]) > 1e-4 * srt[:, 31].abs().clamp_min(1e-3)) | (srt[:, 31] == 0)
    same = sets_equal_rows(idx.cpu(), ri)
    assert bool(same[clear].all()) and clear.float().mean() > 0.9
    if bias < 0:
        assert int((rv > 0).sum(-1).min()) < 32  # short rows are present


@pytest.mark.parametrize("shape", [(300, 64, 200), (129, 64, 256), (77, 40, 64), (1000, 64, 257)])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_fused_encoder_edge_shapes(shape, precision):
    """Dictionaries of one column tile (set 1 never scans), exactly one tile, the minimum width, and a second tile
    with a single valid column: selected sets equal a stable descending sort wherever the k-th gap is clear."""
    from freud_b200 import ops
    from freud_b200._lib import BF16, FP32

    N, d, n = shape
    prec = BF16 if precision == "bf16" else FP32
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, N, d, generator=g).cuda()
    W = (torch.randn(n, d, generator=g) / d ** 0.5).cuda()
    b_enc = (0.1 * torch.randn(n, generator=g)).cuda()
    b_dec = (0.1 * torch.randn(d, generator=g)).cuda()
    xc, xl, _ = ops.topk_prep_x(x, b_dec, prec)
    w, wl = ops.split_operand(W, prec)
    vals, idx = ops.topk_encode(xc, xl, w, wl, b_enc, prec)
    pre = torch.relu(xc.float() @ w.float().T + b_enc) if prec == BF16 else torch.relu((x[0] - b_dec) @ W.T + b_enc)
    srt, order = torch.sort(pre, dim=-1, descending=True, stable=True)
    same = (torch.sort(idx.long(), -1).values == torch.sort(order[:, :32], -1).values).all(-1)
    clear = (srt[:, 31] - srt[:, 32] > 1e-4 * srt[:, 31].abs().clamp_min(1e-3)) | (srt[:, 31] == 0)
    assert bool(same[clear].all()) and float(clear.float().mean()) > 0.98
    assert float((torch.sort(vals, -1, descending=True).values - srt[:, :32]).abs().max()) < 1e-5


def test_row_topk_masked_exact():
    from freud_b200 import ops

    g = torch.Generator().manual_seed(3)
    lat = torch.relu(torch.randn(50, 1000, generator=g))
    mask = torch.rand(1000, generator=g) < 0.3
    for k, m in ((8, None), (128, None), (192, mask)):
        src = lat if m is None else torch.where(m[None], lat, torch.full_like(lat, -torch.inf))
        rv, ri = osae.select_topk(src, k)
        vals, idx = ops.row_topk(lat.cuda(), k, None if m is None else m.cuda())
        assert torch.equal(idx.cpu().long(), ri), (k, m is not None)  # identical order: value desc, index asc
        assert torch.equal(vals.cpu(), rv)


def test_decoder_norm_helpers():
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder

    z, _ = load_golden("topk_decoder_norm")
    cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": 96, "k": 4, "normalize_decoder": False})
    model = TopKAutoEncoder(24, cfg).cuda()
    model.W_dec.data.copy_(t(z["W_dec_in"]))
    model.set_decoder_norm_to_unit_norm()
    assert rel_err(model.W_dec.data.cpu(), z["W_dec_unit"]) < 1e-6
    model.W_dec.grad = t(z["grad_in"]).cuda()
    model.remove_gradient_parallel_to_decoder_directions()
    assert rel_err(model.W_dec.grad.cpu(), z["grad_out"]) < 1e-5


def test_cpu_input_fails_loudly():
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder

    model = TopKAutoEncoder(32, TopKAutoEncoderConfig.from_dict({"n_dict_components": 256}))
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.randn(2, 3, 32))


def test_two_dimensional_input_matches_reference_semantics():
    """[T, d] input: total_variance subtracts the mean over axis 0 = over T (topkautoencoder.py:104), not zero."""
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder

    torch.manual_seed(0)
    d, n, k, T = 64, 512, 32, 300
    model = TopKAutoEncoder(d, TopKAutoEncoderConfig.from_dict({"n_dict_components": n, "k": k}))
    g = torch.Generator().manual_seed(2)
    model.b_dec.data = 0.1 * torch.randn(d, generator=g)
    x = torch.randn(T, d, generator=g) + 0.5 * torch.randn(d, generator=g)
    sd = {k_: v.clone() for k_, v in model.state_dict().items()}
    ref = osae.topk_forward(x, sd["encoder.weight"], sd["encoder.bias"], sd["W_dec"], sd["b_dec"], k)
    assert float(ref.total_variance) > 1.0
    rg = osae.topk_backward(x, sd["encoder.weight"], sd["encoder.bias"], sd["W_dec"], sd["b_dec"], ref, k)
    model = model.cuda()
    model.precision = "fp32"
    out, mse = model(x.cuda(), return_mse=True)
    out.fvu.backward()
    assert out.sae_out.shape == (T, d) and out.encoded.top_indices.shape == (T, k)
    assert bool(sets_equal_rows(out.encoded.top_indices.cpu(), ref.top_indices).all())
    assert rel_err(out.fvu.detach().cpu(), ref.fvu) < 1e-5
    assert rel_err(mse.cpu(), ref.mse) < 1e-5
    named = dict(model.named_parameters())
    for key in TOPK_KEYS:
        assert rel_err(named[key].grad.cpu(), rg[key]) < 1e-5, key


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
def test_gradients_through_sae_out_and_top_acts(precision, tol):
    """A caller's own loss on the returned reconstruction / activations back-propagates to the parameters
    (the reference's outputs are ordinary autograd tensors): checked against torch autograd on the CPU over the
    same selection."""
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder

    torch.manual_seed(0)
    d, n, k, B, T = 64, 512, 32, 2, 150
    model = TopKAutoEncoder(d, TopKAutoEncoderConfig.from_dict({"n_dict_components": n, "k": k}))
    g = torch.Generator().manual_seed(4)
    model.b_dec.data = 0.1 * torch.randn(d, generator=g)
    model.encoder.bias.data = 0.05 * torch.randn(n, generator=g)
    x = torch.randn(B, T, d, generator=g)
    R = torch.randn(B, T, d, generator=g)
    S = torch.randn(B, T, k, generator=g)
    cpu = {k_: v.clone().requires_grad_(True) for k_, v in model.state_dict().items()}
    model = model.cuda()
    model.precision = precision
    out = model(x.cuda())
    loss = (out.sae_out * R.cuda()).sum() + (out.encoded.top_acts.float() * S.cuda()).sum() + 3.0 * out.fvu
    loss.backward()
    idx = out.encoded.top_indices.cpu()
    pre = torch.relu((x - cpu["b_dec"]) @ cpu["encoder.weight"].T + cpu["encoder.bias"])
    acts = torch.gather(pre, -1, idx)
    sae = (acts.unsqueeze(-1) * cpu["W_dec"][idx]).sum(-2) + cpu["b_dec"]
    tv = (x - x.mean(0)).pow(2).sum()
    ref_loss = (sae * R).sum() + (acts * S).sum() + 3.0 * (sae - x).pow(2).sum() / tv
    ref_loss.backward()
    assert rel_err(loss.detach().cpu(), ref_loss.detach()) < tol
    named = dict(model.named_parameters())
    for key in TOPK_KEYS:
        assert rel_err(named[key].grad.cpu(), cpu[key].grad) < tol, key


@pytest.mark.parametrize("d,N", [(384, 1000), (512, 77), (768, 1500), (1024, 130), (1280, 301)])
@pytest.mark.parametrize("sharded", [False, True])
def test_fused_decode_dacts_matches_separate_kernels(d, N, sharded):
    """freud_topk_decode_dacts (rows gathered once through shared memory) == freud_topk_decode followed by
    freud_topk_dacts: reconstruction / residual / column sums bit-for-bit (same FMA order), dacts to fp32 rounding of
    a differently-ordered dot, and both against an fp64 evaluation of eager_decode (topkautoencoder.py:15-18) and its
    autograd; `sharded` marks a third of the entries as another shard's (index -1)."""
    from freud_b200 import ops

    g = torch.Generator().manual_seed(d + N)
    n, k = 4096, 32
    W = (torch.randn(n, d, generator=g) / d ** 0.5).to(torch.bfloat16).cuda()
    b_dec = (0.1 * torch.randn(d, generator=g)).cuda()
    x = torch.randn(N, d, generator=g).cuda()
    idx = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(N)]).to(torch.int32)
    vals = torch.rand(N, k, generator=g)
    if sharded:
        drop = torch.rand(N, k, generator=g) < 0.33
        drop[0] = True  # a token none of whose winners this shard owns
        idx[drop] = -1
        vals[drop] = 0.0
    idx, vals = idx.cuda(), vals.cuda()
    out_f, e_f, sse_f, cs_f, da_f = ops.topk_decode_dacts(vals, idx, W, b_dec, x)
    out_s, e_s, sse_s, cs_s = ops.topk_decode(vals, idx, W, b_dec, x, resid_dtype=torch.bfloat16, want_sse=True,
                                              want_colsum=True)
    da_s = ops.topk_dacts(e_s, idx, W)
    torch.cuda.synchronize()
    assert torch.equal(out_f, out_s)
    assert torch.equal(e_f.view(torch.int16), e_s.view(torch.int16))
    assert rel_err(sse_f.cpu(), sse_s.cpu()) < 1e-12
    assert rel_err(cs_f.cpu(), cs_s.cpu()) < 1e-5
    assert rel_err(da_f.cpu(), da_s.cpu()) < 2e-6
    # fp64 evaluation
    Wd = W.double().cpu()
    safe = idx.cpu().long().clamp_min(0)
    live = (idx.cpu() >= 0).double()
    rows = Wd[safe]                                                    # [N,k,d]
    out64 = ((vals.cpu().double() * live).unsqueeze(-1) * rows).sum(1) + b_dec.double().cpu()
    assert rel_err(out_f.cpu(), out64) < 1e-6
    da64 = (rows * e_f.double().cpu().unsqueeze(1)).sum(-1) * live
    assert rel_err(da_f.cpu(), da64) < 2e-6
    assert bool((da_f.cpu()[idx.cpu() < 0] == 0).all())


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_half_precision_activations_feed_the_fused_path_exactly(dtype):
    """Activation stores collected on CUDA hold fp16 (hooked_model.py:106-108 decodes with fp16=True): SAETrainer.step
    takes them as stored -- freud_topk_prep_x / freud_topk_decode_dacts widen exactly -- and must produce bit-for-bit
    what it produces for the same values widened to fp32 first."""
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder
    from freud_b200.trainer import SAETrainer

    d, n = 384, 2048
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(3, 200, d, generator=g) * 0.7).to(dtype).cuda()
    outs = []
    for feed in (x, x.float()):
        torch.manual_seed(1)
        model = TopKAutoEncoder(d, TopKAutoEncoderConfig.from_dict({"n_dict_components": n, "k": 32})).cuda()
        tr = SAETrainer(model, lr=1e-3, steps=10, clip_thresh=1.0, optimizer="adam", scheduler="linear",
                        scheduler_params={"num_warmup_steps": 2}, dead_feature_threshold=10 ** 9, precision="bf16")
        o = None
        for _ in range(2):
            o = tr.step(feed)
        torch.cuda.synchronize()
        outs.append((float(o["loss"]), o["sae_out"].clone(), o["top_idx"].clone(), model.W_dec.data.clone(),
                     model.encoder.weight.data.clone()))
    assert outs[0][0] == outs[1][0]
    for a, b in zip(outs[0][1:], outs[1][1:]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("N,n", [(3000, 64), (3000, 256), (3000, 2048), (700, 32), (5000, 40000), (257, 300)])
def test_csc_lists_are_complete_and_token_sorted(N, n):
    """freud_csc_build: offsets = exclusive scan of the per-feature counts, and every feature's list holds exactly its
    flat positions p = t*k + j in ascending order (the order that makes the gradient sums run-to-run deterministic):
    warp-level bitonic sorts for lists up to 1024 entries, the CTA-wide network up to 4096; entries of other shards
    (index -1) are left out."""
    from freud_b200 import ops

    k = 32
    g = torch.Generator().manual_seed(N + n)
    idx = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(N)]).to(torch.int32) if n >= k else None
    if idx is None:
        pytest.skip("needs n >= k")
    idx[torch.rand(N, k, generator=g) < 0.05] = -1
    offsets, entries = ops.csc_build(idx.cuda(), n)
    torch.cuda.synchronize()
    offsets, entries = offsets.cpu().long(), entries.cpu().long()
    flat = idx.reshape(-1).long()
    counts = torch.bincount(flat[flat >= 0], minlength=n)
    assert torch.equal(offsets[1:] - offsets[:-1], counts)
    order = torch.sort(flat[flat >= 0], stable=True)
    want = torch.nonzero(flat >= 0).flatten()[order.indices]  # positions grouped by feature, ascending inside a group
    total = int(offsets[-1])
    lens = counts
    sortable = torch.repeat_interleave(lens <= 4096, lens)
    assert torch.equal(entries[:total][sortable], want[sortable])
    # longer lists: complete, order unspecified
    assert torch.equal(torch.sort(flat[entries[:total]]).values, torch.sort(flat[flat >= 0]).values)
