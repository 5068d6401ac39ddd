"""Mirror of the search entry point of src/utils/activations.py:61-132 on the CUDA search kernels.

`top_activations(dataloader, feature_idx, n_files, max_val, min_val, absolute_magnitude, return_max_per_file)`
keeps the reference's seven positional arguments (it is the `top_fn` lambda of gui_server.py:91-99) and its return
value `(pq, max_per_file)` with pq[i] = (filename, trimmed_activation, value, time).  Instead of re-reading every
file through the DataLoader (and re-decoding its audio) on every query, the first call uploads the whole store to
HBM once (DeviceActivationStore) and each query is one streaming kernel + one ranking kernel.
"""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from ..dataset.activations import DeviceActivationStore
from .constants import TIMESTEP_S

_STORE_ATTR = "_freud_b200_store"


def attach_store(dataloader, store: DeviceActivationStore):
    setattr(dataloader, _STORE_ATTR, store)
    return store


def get_store(dataloader, **store_kwargs) -> DeviceActivationStore:
    store = getattr(dataloader, _STORE_ATTR, None)
    if store is None:
        if not hasattr(dataloader, "dataset") or not hasattr(dataloader.dataset, "metadata"):
            raise RuntimeError("freud_b200.top_activations needs a MemoryMappedActivationDataLoader (--from_disk)")
        store = attach_store(dataloader, DeviceActivationStore(dataloader.dataset, **store_kwargs))
    return store


@torch.no_grad()
def _gather_files(store, t: torch.Tensor, fill):
    """Per-file vector of this rank's block -> the full [n_total] vector on every rank (file order)."""
    import torch.distributed as dist

    world = store.shard[1]
    part = torch.full((store.per,), fill, dtype=t.dtype, device=t.device)
    part[: t.numel()] = t
    full = torch.empty(world * store.per, dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(full, part)
    return full[: store.n_total].contiguous()


def top_activations(dataloader, feature_idx: int, n_files: int, max_val: Optional[float], min_val: Optional[float],
                    absolute_magnitude: bool, return_max_per_file: bool):
    store = get_store(dataloader)
    if store.activation_type == "tensor":
        F = store.acts.shape[-1]
        if not 0 <= int(feature_idx) < F:
            raise IndexError(f"feature_idx {feature_idx} out of range for {F} features")
        if F % 4 == 0:
            # all-feature table (one pass over the store, first query only): this query is a column of it
            tv, ta, tb = store.feature_table()
            f = int(feature_idx)
            vmax, amax, vabs = tv[:, f].contiguous(), ta[:, f].contiguous(), tb[:, f].contiguous()
        elif getattr(store, "acts_fm", None) is not None:
            # contiguous [N_files, T] slab of this feature: the same kernel with a unit column stride
            vmax, amax, vabs, _ = ops.search_dense(store.acts_fm[int(feature_idx)].unsqueeze(-1), store.n_frames, 0, False)
        else:
            vmax, amax, vabs, _ = ops.search_dense(store.acts, store.n_frames, int(feature_idx), False)
    else:
        vmax, amax, vabs, _ = ops.search_indexed(store.vals, store.idx, store.n_frames, int(feature_idx), False)
    sharded = getattr(store, "shard", None) is not None and store.shard[1] > 1
    if sharded:  # files are sharded over the ranks: exchange the per-file results, rank globally on every rank
        vmax, amax, vabs = _gather_files(store, vmax, 0.0), _gather_files(store, amax, 0), _gather_files(store, vabs, 0.0)
    files, count = ops.search_topn(vmax, vabs, bool(absolute_magnitude), min_val, max_val, int(n_files))
    stat = vabs if absolute_magnitude else vmax
    n_found = int(count.item())
    files_h = files[:n_found].tolist()
    pq = []
    if n_found:
        sel = files[:n_found].long()
        stat_h = stat[sel].tolist()
        amax_h = amax[sel].tolist()
        own = [r for r, fi in enumerate(files_h) if store.lo <= fi < store.hi]
        loc = torch.tensor([files_h[r] - store.lo for r in own], dtype=torch.long, device=sel.device)
        traces = torch.zeros((n_found, store.T), dtype=torch.float32, device=sel.device)
        if own:
            if store.activation_type == "tensor":
                tr = store.acts[loc, :, int(feature_idx)].float()
            else:
                _, _, _, tr = ops.search_indexed(store.vals[loc].contiguous(), store.idx[loc].contiguous(),
                                                 store.n_frames[loc].contiguous(), int(feature_idx), True)
            traces[torch.tensor(own, dtype=torch.long, device=sel.device)] = tr
        if sharded:  # every winner's trace lives on exactly one rank
            import torch.distributed as dist

            dist.all_reduce(traces)
        traces = traces.cpu()
        for r, fi in enumerate(files_h):
            nf = store.n_frames_host[fi]
            signed = stat_h[r]
            value = abs(signed) if absolute_magnitude else signed
            pq.append((store.filenames[fi], traces[r, :nf].clone(), value, amax_h[r] * TIMESTEP_S))
    max_per_file = None
    if return_max_per_file:
        # the reference appends one entry per file (utils/activations.py:110-117)
        max_per_file = stat.tolist()
    return pq, max_per_file


# ------------------------------------------------------------------------------------------ upload-clip helpers
# Same signatures as the reference (utils/activations.py:135-300; called by gui_server.py for an uploaded clip).
# The Whisper side stays the caller's: `whisper_cache` / `whisper_subbed` are the reference's hooked-model objects
# (anything with .model_name, .device, .forward(mel) -> result with .text, .activations), and the mel front end is
# the reference's own `src.utils.audio_utils.get_mels_from_np_array` unless `set_mel_frontend` installed another.
_mel_frontend = None


def set_mel_frontend(fn):
    """fn(device, audio_array, n_mels) -> mel; None restores the default (the reference's audio_utils)."""
    global _mel_frontend
    _mel_frontend = fn


def _mels(whisper_cache, audio_array):
    fn = _mel_frontend
    if fn is None:
        try:
            from src.utils.audio_utils import get_mels_from_np_array as fn  # the reference's (Whisper is out of scope)
        except ImportError as ex:
            raise RuntimeError("no mel front end: run inside the reference checkout or call "
                               "freud_b200.utils.activations.set_mel_frontend(fn)") from ex
    n_mels = 128 if "v3" in whisper_cache.model_name else 80  # utils/constants.py get_n_mels
    return fn(whisper_cache.device, audio_array, n_mels)


def activation_length_from_audio_array(audio_array) -> int:
    """utils/activations.py:32-38."""
    from .constants import SAMPLE_RATE

    return int((len(audio_array) / SAMPLE_RATE) / TIMESTEP_S)


@torch.no_grad()
def top_activations_for_audio(audio_array, whisper_cache, sae_model, top_n: int):
    """utils/activations.py:135-216: (feature indices, their [true_length] traces) of the `top_n` features with the
    largest activation anywhere in the clip.  The per-frame sort / de-duplicate loop (:173-189) runs as a few tensor
    ops on the device (utils/clip.py); traces come back on the CPU like the reference's."""
    from ..models.l1autoencoder import L1EncoderOutput
    from .clip import top_features_of_clip, top_features_of_dense_clip

    mel = _mels(whisper_cache, audio_array)
    whisper_cache.forward(mel)
    activations = whisper_cache.activations
    true_length = activation_length_from_audio_array(audio_array)
    if sae_model:
        output = sae_model.forward(activations.to(whisper_cache.device))
        if isinstance(output.encoded, L1EncoderOutput):
            activations = output.encoded.latent
        else:
            top_acts = output.encoded.top_acts.squeeze()[:true_length, :]
            top_indices = output.encoded.top_indices.squeeze()[:true_length, :]
            feats, _, traces = top_features_of_clip(top_acts.float(), top_indices, top_n)
            return feats, [tr.cpu() for tr in traces]
    activations = activations.squeeze()[:true_length, :]
    feats, _, traces = top_features_of_dense_clip(activations, top_n)
    return feats, [tr for tr in traces]


@torch.no_grad()
def manipulate_latent(audio_array, whisper_cache, sae_model, whisper_subbed, feat_idx: int,
                      manipulation_factor: float):
    """utils/activations.py:219-300: scale one feature of the clip's encoding, decode with and without the change,
    hand both to the activation-substituted Whisper; returns (baseline text | None, manipulated text, standard text,
    feature value before [true_length], after [true_length])."""
    from ..models.l1autoencoder import L1EncoderOutput
    from .clip import manipulate_topk_encoding

    mel = _mels(whisper_cache, audio_array)
    baseline_result = whisper_cache.forward(mel)
    activations = whisper_cache.activations.to(whisper_cache.device)
    if sae_model:
        output = sae_model.forward(activations)
        if isinstance(output.encoded, L1EncoderOutput):
            latent = output.encoded.latent
            value_pre_activation = latent[:, :, feat_idx]
            manipulated_value = value_pre_activation * manipulation_factor
            manipulated_encoding = latent.clone()
            manipulated_encoding[:, :, feat_idx] = manipulated_value
            manipulated_decoded = sae_model.decode(manipulated_encoding)
            standard_decoded = sae_model.decode(latent)
        else:
            top_acts = output.encoded.top_acts.squeeze()
            top_indices = output.encoded.top_indices.squeeze()
            manipulated_top_acts = manipulate_topk_encoding(top_acts, top_indices, feat_idx, manipulation_factor)
            manipulated_decoded = sae_model.decode(manipulated_top_acts.unsqueeze(0), top_indices.unsqueeze(0))
            standard_decoded = sae_model.decode(top_acts.unsqueeze(0), top_indices.unsqueeze(0))
            # activation_tensor_from_indexed (:41-57) for one clip: the feature's value where selected, else 0
            value_pre_activation = (top_acts.float() * (top_indices == feat_idx)).sum(-1).unsqueeze(0)
            manipulated_value = value_pre_activation * manipulation_factor
    else:
        value_pre_activation = activations[:, :, feat_idx]
        manipulated_value = value_pre_activation * manipulation_factor
        manipulated_encoding = activations.clone()
        manipulated_encoding[:, :, feat_idx] = manipulated_value
        manipulated_decoded = manipulated_encoding
        standard_decoded = activations
    manipulated_subbed_result = whisper_subbed.forward(mel, manipulated_decoded)
    standard_subbed_result = whisper_subbed.forward(mel, standard_decoded)
    baseline_text = None if sae_model is None else baseline_result.text
    n = activation_length_from_audio_array(audio_array)
    return (baseline_text, manipulated_subbed_result.text, standard_subbed_result.text,
            value_pre_activation.squeeze()[:n].cpu(), manipulated_value.squeeze()[:n].cpu())
