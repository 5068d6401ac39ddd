"""Upload-clip helpers: vectorised versions vs the reference's per-frame loops."""
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
