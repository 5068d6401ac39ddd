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
