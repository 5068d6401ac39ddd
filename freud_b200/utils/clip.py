"""Upload-clip helpers of the GUI server without their per-frame Python loops.

The reference computes both from one clip's Whisper activations (`top_activations_for_audio`, `manipulate_latent`,
src/utils/activations.py:135-275); the Whisper forward is out of scope here, so these functions start from the SAE
encoding (or dense activations) of the clip -- what `sae_model.forward(activations)` returned -- and run as a handful
of tensor ops on whatever device that encoding lives on.
"""
from __future__ import annotations

import torch


def top_features_of_clip(top_acts: torch.Tensor, top_indices: torch.Tensor, top_n: int):
    """Result of the loop at utils/activations.py:173-189 for one clip's encoding ([T, k] each, already trimmed to
    the clip's true length): the `top_n` features with the largest activation anywhere in the clip, as
    (feature indices, values, per-feature traces [T]).

    The loop keeps a stably re-sorted, de-duplicated, truncated list frame after frame; because the list only ever
    improves, its final content is: every feature at its maximum over the clip, ordered by (value descending, the
    first frame / slot where that maximum occurs ascending), cut to `top_n`."""
    T, k = top_acts.shape
    vals = top_acts.reshape(-1)
    feats = top_indices.reshape(-1).long()
    order = torch.sort(vals, descending=True, stable=True).indices  # ties keep (frame, slot) order
    f_sorted = feats[order]
    # first occurrence of every feature in that order = its maximum, at the earliest place it is reached
    uniq, inv = torch.unique(f_sorted, return_inverse=True)
    first = torch.full((uniq.numel(),), f_sorted.numel(), dtype=torch.long, device=vals.device)
    first.scatter_reduce_(0, inv, torch.arange(f_sorted.numel(), device=vals.device), reduce="amin")
    best = torch.sort(first).values[:top_n]  # ranks (in the sorted entry list) of the winners, best first
    win_feats = f_sorted[best]
    win_vals = vals[order][best]
    traces = (top_acts.unsqueeze(0) * (top_indices.unsqueeze(0) == win_feats.view(-1, 1, 1))).sum(-1)  # [n, T]
    return win_feats.tolist(), win_vals.tolist(), traces


def top_features_of_dense_clip(activations: torch.Tensor, top_n: int):
    """Dense variant (no SAE, or an L1 SAE's latents): utils/activations.py:167-171 takes a per-frame
    `topk(top_n)` first and then runs the same loop."""
    res = activations.topk(top_n)
    feats, vals, _ = top_features_of_clip(res.values, res.indices, top_n)
    return feats, vals, activations[:, feats].T.contiguous()


def manipulate_topk_encoding(top_acts: torch.Tensor, top_indices: torch.Tensor, feat_idx: int, factor: float):
    """utils/activations.py:256-262: the encoding with feature `feat_idx` scaled by `factor` wherever it was selected."""
    return torch.where(top_indices == feat_idx, top_acts * factor, top_acts)
