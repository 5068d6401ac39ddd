"""Validation feature statistics of src/scripts/train_sae.py.

`topk_feature_extraction(out, mag_vals_dim, batch_idx, device)` keeps the reference signature (:70) and returns the
same [mag_vals_dim] vector -- the per-feature maximum |activation| of one file -- from one scatter-max kernel instead
of a [T, k, n] boolean mask.  `l1_feature_extraction` is the L1 branch of validate() (:175-178)."""
import torch

from .. import ops


def topk_feature_extraction(out, mag_vals_dim, batch_idx=None, device=None):
    acts = out.encoded.top_acts.detach()
    idx = out.encoded.top_indices
    if not acts.is_cuda:
        raise RuntimeError("freud_b200 feature statistics run on CUDA tensors (no CPU fallback)")
    k = acts.shape[-1]
    return ops.feature_absmax(acts.reshape(-1, k).float().contiguous(), idx.reshape(-1, k).contiguous(),
                              int(mag_vals_dim))


def l1_feature_extraction(out):
    latent = out.encoded.latent.detach()
    if not latent.is_cuda:
        raise RuntimeError("freud_b200 feature statistics run on CUDA tensors (no CPU fallback)")
    return ops.col_absmax(latent.reshape(-1, latent.shape[-1]).float().contiguous())
