"""CPU restatement of the upload-clip helpers of src/utils/activations.py (TEST INFRASTRUCTURE ONLY).

`top_features_loop` follows the per-frame Python loop of top_activations_for_audio (:173-189) literally;
`manipulate_topk_loop` the per-frame loop of manipulate_latent (:258-262).  Both start from the SAE encoding /
activations of one clip -- the Whisper forward that produces them is out of scope.
"""
import torch


def top_features_loop(top_acts: torch.Tensor, top_indices: torch.Tensor, top_n: int):
    """utils/activations.py:173-189.  top_acts / top_indices: [T, k].  Returns [(feature, value)] of length <= top_n."""
    unique_top_activations = []
    for top_acts_at_t, top_indices_at_t in zip(top_acts, top_indices):
        unique_top_activations.extend(
            [(idx.item(), value.item()) for idx, value in zip(top_indices_at_t, top_acts_at_t)])
        unique_top_activations = sorted(unique_top_activations, key=lambda x: x[1], reverse=True)  # stable
        new_unique = []
        for idx, value in unique_top_activations:
            if idx not in [i for i, _ in new_unique] and len(new_unique) < top_n:
                new_unique.append((idx, value))
        unique_top_activations = new_unique
    return unique_top_activations


def activation_tensor_from_indexed(top_acts: torch.Tensor, top_indices: torch.Tensor, feature: int):
    """utils/activations.py:41-57 for one clip: dense [T] trace of `feature`."""
    out = torch.zeros(top_acts.shape[0], dtype=top_acts.dtype)
    for t in range(top_acts.shape[0]):
        for j in range(top_acts.shape[1]):
            if int(top_indices[t, j]) == feature:
                out[t] = top_acts[t, j]
    return out


def manipulate_topk_loop(top_acts: torch.Tensor, top_indices: torch.Tensor, feat_idx: int, factor: float):
    """utils/activations.py:256-262: scale the activation of `feat_idx` wherever it was selected."""
    manipulated = top_acts.clone()
    for i, (idx_at_t, act_at_t) in enumerate(zip(top_indices, top_acts)):
        if feat_idx in idx_at_t:
            idx = (idx_at_t == feat_idx).nonzero().item()
            manipulated[i, idx] = act_at_t[idx] * factor
    return manipulated
