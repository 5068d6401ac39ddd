"""Upload-clip helpers: vectorised versions vs the reference's per-frame loops."""
import pytest
import torch

from oracle import clip as oclip


def _encoding(T, k, n, seed, ties=False):
    g = torch.Generator().manual_seed(seed)
    idx = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(T)])
    vals = torch.rand(T, k, generator=g)
    if ties:  # quantised values: many exact ties between features and across frames
        vals = (vals * 6).round() / 6
    return vals, idx


def test_top_features_of_clip_matches_reference_loop():
    from freud_b200.utils.clip import top_features_of_clip

    for seed, (T, k, n, top_n, ties) in enumerate([(40, 8, 50, 10, False), (60, 8, 30, 10, True), (25, 4, 12, 20, True),
                                                   (1, 8, 64, 5, False), (30, 6, 200, 10, True)]):
        vals, idx = _encoding(T, k, n, seed, ties)
        ref = oclip.top_features_loop(vals, idx, top_n)
        feats, values, traces = top_features_of_clip(vals, idx, top_n)
        assert feats == [f for f, _ in ref], (seed, feats, ref)
        assert values == [v for _, v in ref]
        for f, v, tr in zip(feats, values, traces):
            assert torch.equal(tr, oclip.activation_tensor_from_indexed(vals, idx, f))
            assert float(tr.max()) == v  # the reference's own sanity check (:203-205)


def test_top_features_of_dense_clip_matches_reference_loop():
    from freud_b200.utils.clip import top_features_of_dense_clip

    g = torch.Generator().manual_seed(3)
    acts = torch.randn(50, 40, generator=g)
    top_n = 7
    res = acts.topk(top_n)
    ref = oclip.top_features_loop(res.values, res.indices, top_n)
    feats, values, traces = top_features_of_dense_clip(acts, top_n)
    assert feats == [f for f, _ in ref] and values == [v for _, v in ref]
    assert torch.equal(traces, acts[:, feats].T)


def test_manipulate_topk_encoding_matches_reference_loop():
    from freud_b200.utils.clip import manipulate_topk_encoding

    vals, idx = _encoding(80, 8, 40, 11)
    for feat in (0, 7, 39, 1000):
        assert torch.equal(manipulate_topk_encoding(vals, idx, feat, 2.5), oclip.manipulate_topk_loop(vals, idx, feat, 2.5))


def _golden_encoding(z):
    """TopK encoding of the golden clip by the oracle's SAE forward (pinned elsewhere), trimmed like the reference."""
    from oracle import sae as osae
    from tests.util import t

    acts = t(z["acts"])
    n = int((int(z["audio_len"]) / 16000) / (30 / 1500))
    out = osae.topk_forward(acts, t(z["topk.encoder.weight"]), t(z["topk.encoder.bias"]), t(z["topk.W_dec"]),
                            t(z["topk.b_dec"]), 8)
    return out.top_acts.reshape(-1, 8)[:n], out.top_indices.reshape(-1, 8)[:n], n


def test_oracle_loops_pinned_to_reference_golden():
    """oracle/clip.py against the reference's own top_activations_for_audio / manipulate_latent run on a stand-in
    Whisper (tests/golden/make_golden.py clip_case): feature order, traces, manipulated values."""
    from tests.util import load_golden, t

    z, _ = load_golden("clip")
    top_acts, top_idx, n = _golden_encoding(z)
    ref = oclip.top_features_loop(top_acts, top_idx, int(z["top_n"]))
    assert [f for f, _ in ref] == z["topk.top.features"].tolist()
    for (f, _), tr in zip(ref, t(z["topk.top.traces"])):
        assert torch.equal(oclip.activation_tensor_from_indexed(top_acts, top_idx, f), tr)
    feat = int(z["topk.man.feature"])
    man = oclip.manipulate_topk_loop(top_acts, top_idx, feat, 3.5)
    assert torch.equal(oclip.activation_tensor_from_indexed(man, top_idx, feat), t(z["topk.man.value"]))
    assert torch.equal(oclip.activation_tensor_from_indexed(top_acts, top_idx, feat), t(z["topk.man.pre"]))
    # dense variant (no SAE): per-frame topk(top_n) first, then the same loop (:167-171)
    dense = t(z["acts"])[0, :n]
    res = dense.topk(int(z["top_n"]))
    ref = oclip.top_features_loop(res.values, res.indices, int(z["top_n"]))
    assert [f for f, _ in ref] == z["none.top.features"].tolist()


class _Result:
    def __init__(self, text):
        self.text = text


class _Cache:
    model_name = "tiny"

    def __init__(self, acts, device):
        self._acts, self.device = acts, device

    def forward(self, mel):
        self.activations = self._acts.clone()
        return _Result("baseline")


class _Subbed:
    def __init__(self):
        self.seen = []

    def forward(self, mel, sub):
        self.seen.append(sub.detach().clone())
        return _Result(f"subbed{len(self.seen)}")


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["topk", "l1", "none"])
def test_upload_clip_entry_points_match_reference(tag):
    """freud_b200.utils.activations.top_activations_for_audio / manipulate_latent (reference signatures, SAE on the
    CUDA kernels) against the reference's outputs for the same clip: same features in the same order, traces,
    texts, manipulated values and the tensors handed to the activation-substituted Whisper."""
    import numpy as np

    from freud_b200.models.config import L1AutoEncoderConfig, TopKAutoEncoderConfig
    from freud_b200.models.l1autoencoder import L1AutoEncoder
    from freud_b200.models.topkautoencoder import TopKAutoEncoder
    from freud_b200.utils import activations as ua
    from tests.util import load_golden, rel_err, t

    z, _ = load_golden("clip")
    acts = t(z["acts"]).cuda()
    audio = np.zeros(int(z["audio_len"]), dtype=np.float32)
    top_n = int(z["top_n"])
    sae = None
    if tag == "topk":
        sae = TopKAutoEncoder(32, TopKAutoEncoderConfig.from_dict({"n_dict_components": 256, "k": 8}))
    elif tag == "l1":
        sae = L1AutoEncoder(32, L1AutoEncoderConfig.from_dict({"n_dict_components": 40}))
    if sae is not None:
        sae.load_state_dict({k[len(tag) + 1:]: t(z[k]) for k in z.files if k.startswith(tag + ".") and
                             k.split(".")[1] not in ("top", "man")})
        sae = sae.cuda()
        sae.precision = "fp32"
    ua.set_mel_frontend(lambda device, a, n_mels: ("mel", len(a), n_mels))
    try:
        feats, traces = ua.top_activations_for_audio(audio, _Cache(acts, "cuda"), sae, top_n)
        assert list(feats) == z[f"{tag}.top.features"].tolist()
        assert rel_err(torch.stack([tr.float().cpu() for tr in traces]), z[f"{tag}.top.traces"]) < 1e-5
        sub = _Subbed()
        feat = int(z[f"{tag}.man.feature"])
        base, man_text, std_text, pre, man = ua.manipulate_latent(audio, _Cache(acts, "cuda"), sae, sub, feat, 3.5)
        assert [str(base), man_text, std_text] == z[f"{tag}.man.texts"].tolist()
        assert rel_err(pre, z[f"{tag}.man.pre"]) < 1e-5 and rel_err(man, z[f"{tag}.man.value"]) < 1e-5
        assert rel_err(sub.seen[0].float().cpu(), z[f"{tag}.man.sub_manipulated"]) < 1e-5
        assert rel_err(sub.seen[1].float().cpu(), z[f"{tag}.man.sub_standard"]) < 1e-5
    finally:
        ua.set_mel_frontend(None)
