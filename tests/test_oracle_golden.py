"""Pin the CPU oracle against the golden outputs of the real reference
(tests/golden/*.npz, written by tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from oracle import optim as ooptim
from oracle import sae as osae
from oracle import search as osearch
from tests.util import load_golden, rel_err, sets_equal_rows, t

TOPK_KEYS = ["encoder.weight", "encoder.bias", "W_dec", "b_dec"]


def _topk_params(z, prefix):
    return [t(z[f"{prefix}.{k}"]) for k in TOPK_KEYS]


@pytest.mark.parametrize("name", ["topk_fp32", "topk_fp32_auxk", "topk_fp32_auxk_few", "topk_fp32_multi",
                                  "topk_fp32_b1"])
def test_topk_fp32_trajectory(name):
    """Forward, losses, explicit grads, clip, Adam and the linear-warmup LR follow
    the reference trajectory step by step (fp32: 1e-5 relative)."""
    z, meta = load_golden(name)
    W_enc, b_enc, W_dec, b_dec = _topk_params(z, "init")
    m = {k: torch.zeros_like(p) for k, p in zip(TOPK_KEYS, (W_enc, b_enc, W_dec, b_dec))}
    v = {k: torch.zeros_like(p) for k, p in zip(TOPK_KEYS, (W_enc, b_enc, W_dec, b_dec))}
    for s in range(meta["steps"]):
        x = t(z[f"s{s}.x"])
        frames = t(z[f"s{s}.frames_in"])
        dead = frames > meta["dead_thresh"]
        out = osae.topk_forward(x, W_enc, b_enc, W_dec, b_dec, meta["k"], dead_mask=dead,
                                auxk_alpha=meta["auxk_alpha"], multi_topk=meta["multi_topk"])
        if s == 0:
            assert rel_err(out.pre_acts, z["s0.pre_acts"]) < 1e-5
        free = osae.tie_free_rows(out.pre_acts, out.top_indices.shape[-1]).reshape(-1)
        same = sets_equal_rows(out.top_indices, z[f"s{s}.top_indices"])
        assert bool(same[free].all()), "top-k index sets differ on tie-free rows"
        assert free.float().mean() > 0.9
        assert rel_err(out.sae_out, z[f"s{s}.sae_out"]) < 1e-5
        for key in ("fvu", "auxk_loss", "multi_topk_fvu", "mse"):
            assert rel_err(getattr(out, key), z[f"s{s}.{key}"]) < 1e-5, key
        grads = osae.topk_backward(x, W_enc, b_enc, W_dec, b_dec, out, meta["k"],
                                   auxk_alpha=meta["auxk_alpha"], multi_topk=meta["multi_topk"])
        for k in TOPK_KEYS:
            assert rel_err(grads[k], z[f"s{s}.grad.{k}"]) < 1e-5, k
        clipped, total = ooptim.clip_grad_norm([grads[k] for k in TOPK_KEYS], meta["clip"])
        assert rel_err(total, z[f"s{s}.grad_norm"]) < 1e-5
        lr = ooptim.linear_warmup_lr(meta["lr"], s, meta["warmup"], meta["total_steps"])
        assert abs(lr - float(z[f"s{s}.lr"])) <= 1e-12 * max(1.0, lr)
        params = dict(zip(TOPK_KEYS, (W_enc, b_enc, W_dec, b_dec)))
        for k, g in zip(TOPK_KEYS, clipped):
            params[k], m[k], v[k] = ooptim.adam_step(params[k], g, m[k], v[k], s + 1, lr)
            assert rel_err(params[k], z[f"s{s}.param.{k}"]) < 1e-5, k
        W_enc, b_enc, W_dec, b_dec = (params[k] for k in TOPK_KEYS)
        frames_out = ooptim.dead_latent_update(frames, t(z[f"s{s}.top_indices"]), x.shape[0] * x.shape[1])
        assert torch.equal(frames_out, t(z[f"s{s}.frames_out"]))


def test_topk_bf16_mode_vs_reference_autocast():
    """The oracle's bf16 mode (what the CUDA bf16 path computes) stays within the
    north_star's 2e-2 of the reference run under autocast('cpu')."""
    z, meta = load_golden("topk_bf16")
    W_enc, b_enc, W_dec, b_dec = _topk_params(z, "init")
    x = t(z["s0.x"])
    out = osae.topk_forward(x, W_enc, b_enc, W_dec, b_dec, meta["k"], auxk_alpha=meta["auxk_alpha"], mode="bf16")
    # golden inputs were drawn with a k-th/(k+1)-th gap above bf16 rounding, so the selection is unambiguous
    assert bool(sets_equal_rows(out.top_indices, z["s0.top_indices"]).all())
    assert rel_err(out.fvu, z["s0.fvu"]) < 2e-2
    assert rel_err(out.sae_out, z["s0.sae_out"]) < 2e-2
    grads = osae.topk_backward(x, W_enc, b_enc, W_dec, b_dec, out, meta["k"], auxk_alpha=meta["auxk_alpha"],
                               mode="bf16")
    for k in TOPK_KEYS:
        assert rel_err(grads[k], z[f"s0.grad.{k}"]) < 2e-2, k


@pytest.mark.parametrize("name", ["l1_fp32", "l1_fp32_wd"])
def test_l1_fp32_trajectory(name):
    z, meta = load_golden(name)
    W, b = t(z["init.decoder.weight"]), t(z["init.encoder_bias"])
    keys = ["encoder_bias", "decoder.weight"]
    m = {"decoder.weight": torch.zeros_like(W), "encoder_bias": torch.zeros_like(b)}
    v = {"decoder.weight": torch.zeros_like(W), "encoder_bias": torch.zeros_like(b)}
    for s in range(meta["steps"]):
        x = t(z[f"s{s}.x"])
        out = osae.l1_forward(x, W, b, meta["recon_alpha"])
        assert rel_err(out.W_normed, z[f"s{s}.W_normed"]) < 1e-6
        assert rel_err(out.latent, z[f"s{s}.latent"]) < 1e-5
        assert rel_err(out.sae_out, z[f"s{s}.sae_out"]) < 1e-5
        for key in ("l1_loss", "reconstruction_loss", "mse"):
            assert rel_err(getattr(out, key), z[f"s{s}.{key}"]) < 1e-5, key
        grads = osae.l1_backward(x, W, b, out, meta["recon_alpha"])
        for k in keys:
            assert rel_err(grads[k], z[f"s{s}.grad.{k}"]) < 1e-5, k
        clipped, total = ooptim.clip_grad_norm([grads[k] for k in keys], meta["clip"])
        assert rel_err(total, z[f"s{s}.grad_norm"]) < 1e-5
        lr = ooptim.cosine_lr(meta["lr"], s, meta["total_steps"])
        assert abs(lr - float(z[f"s{s}.lr"])) <= 1e-9 * meta["lr"]
        params = {"decoder.weight": out.W_normed, "encoder_bias": b}  # encode() left W normalised in place
        for k, g in zip(keys, clipped):
            params[k], m[k], v[k] = ooptim.radam_step(params[k], g, m[k], v[k], s + 1, lr,
                                                      weight_decay=meta["weight_decay"])
            assert rel_err(params[k], z[f"s{s}.param.{k}"]) < 1e-5, k
        W, b = params["decoder.weight"], params["encoder_bias"]


def test_l1_bf16_mode_vs_reference_autocast():
    z, meta = load_golden("l1_bf16")
    W, b = t(z["init.decoder.weight"]), t(z["init.encoder_bias"])
    x = t(z["s0.x"])
    out = osae.l1_forward(x, W, b, meta["recon_alpha"], mode="bf16")
    assert rel_err(out.reconstruction_loss, z["s0.reconstruction_loss"]) < 2e-2
    assert rel_err(out.l1_loss, z["s0.l1_loss"]) < 2e-2
    grads = osae.l1_backward(x, W, b, out, meta["recon_alpha"], mode="bf16")
    for k in ("decoder.weight", "encoder_bias"):
        assert rel_err(grads[k], z[f"s0.grad.{k}"]) < 2e-2, k


def test_decoder_norm_helpers():
    z, _ = load_golden("topk_decoder_norm")
    assert rel_err(osae.set_decoder_norm_to_unit_norm(t(z["W_dec_in"])), z["W_dec_unit"]) < 1e-6
    out = osae.remove_gradient_parallel_to_decoder_directions(t(z["W_dec_unit"]), t(z["grad_in"]))
    assert rel_err(out, z["grad_out"]) < 1e-5


def test_search_rankings_exact():
    """top_activations: same files, same order, same values / times / max_per_file (exact)."""
    z, meta = load_golden("search")
    filenames = meta["filenames"]
    n_frames = [osearch.n_frames_from_samples(int(s)) for s in z["num_samples"]]
    for q, query in enumerate(meta["queries"]):
        if query["kind"] == "dense":
            acts = z["dense"][:, :, query["feature"]]
        else:
            acts = osearch.dense_from_indexed(z["vals"], z["idx"], query["feature"])
        pq, mpf = osearch.top_activations(acts, filenames, n_frames, query["n_files"], query["max_val"],
                                          query["min_val"], query["abs"], True)
        assert [filenames.index(p[0]) for p in pq] == z[f"q{q}.files"].tolist(), query
        assert np.array_equal(np.array([p[2] for p in pq]), z[f"q{q}.values"]), query
        assert np.array_equal(np.array([p[3] for p in pq]), z[f"q{q}.times"]), query
        assert [len(p[1]) for p in pq] == z[f"q{q}.lens"].tolist()
        assert np.array_equal(np.array(mpf, dtype=np.float64), z[f"q{q}.max_per_file"]), query
        if pq:
            assert np.array_equal(pq[0][1], z[f"q{q}.trace0"])


def test_validation_feature_statistics():
    """topk_feature_extraction / L1 abs-max (train_sae.py:70-118,175-178), exact."""
    from tests.util import GOLDEN

    z = np.load(f"{GOLDEN}/feature_stats.npz")
    got = osae.topk_feature_absmax(torch.from_numpy(z["acts"]), torch.from_numpy(z["idx"]), 96)
    assert np.array_equal(got.numpy(), z["topk_max"])
    assert np.array_equal(osae.l1_feature_absmax(torch.from_numpy(z["latent"])).numpy(), z["l1_max"])
