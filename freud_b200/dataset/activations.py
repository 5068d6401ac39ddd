"""Mirror of src/dataset/activations.py:16-31,116-206 (checkpoint loader + memory-mapped activation loader) plus
a device-resident activation store used by the CUDA feature search.

On-disk format (src/scripts/collect_activations.py:37-41,60-63,101-108):
  {layer}_metadata.json = {"tensor_shape": [T, F or k], "activation_shape": [T, n_features], "filenames": [...]}
  dense   : {layer}_tensors.npy            [N_files, T*F]
  indexed : {layer}_activation_values.npy + {layer}_feature_indices.npy   [N_files, T*k]
"""
from __future__ import annotations

import json
import os
from typing import Callable, Optional

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from ..models.config import L1AutoEncoderConfig, TopKAutoEncoderConfig
from ..models.l1autoencoder import L1AutoEncoder
from ..models.topkautoencoder import TopKAutoEncoder
from ..utils.constants import SAMPLE_RATE, TIMESTEP_S


def init_sae_from_checkpoint(checkpoint: str, device: Optional[str | torch.device] = None):
    """reference :16-31 -- same checkpoint layout ({model, optimizer, scheduler, step, best_val_loss, hparams})."""
    checkpoint = torch.load(checkpoint, map_location=device)
    activation_size = checkpoint["hparams"]["activation_size"]
    if checkpoint["hparams"]["autoencoder_variant"] == "l1":
        cfg = L1AutoEncoderConfig.from_dict(checkpoint["hparams"]["autoencoder_config"])
        model = L1AutoEncoder(activation_size, cfg)
    else:
        cfg = TopKAutoEncoderConfig.from_dict(checkpoint["hparams"]["autoencoder_config"])
        model = TopKAutoEncoder(activation_size, cfg)
    model.load_state_dict(checkpoint["model"])
    model.eval().to(device)
    return model


class MemoryMappedActivationsDataset(Dataset):
    """reference :116-174."""

    def __init__(self, data_path: str, layer_name: str, subset_size: Optional[int] = None):
        self.data_path = data_path
        self.layer_name = layer_name
        self.metadata_file = os.path.join(data_path, f"{layer_name}_metadata.json")
        with open(self.metadata_file, "r") as f:
            self.metadata = json.load(f)
        self.tensor_file = os.path.join(data_path, f"{layer_name}_tensors.npy")
        if not os.path.exists(self.tensor_file):
            self.activation_value_file = os.path.join(data_path, f"{layer_name}_activation_values.npy")
            self.feature_index_file = os.path.join(data_path, f"{layer_name}_feature_indices.npy")
            self.activation_type = "indexed"
            self.act_mmap = np.load(self.activation_value_file, mmap_mode="r")
            self.idx_mmap = np.load(self.feature_index_file, mmap_mode="r")
        else:
            self.activation_type = "tensor"
            self.mmap = np.load(self.tensor_file, mmap_mode="r")
        if subset_size is not None:
            self.metadata["filenames"] = self.metadata["filenames"][:subset_size]
            if self.activation_type == "indexed":
                self.act_mmap = self.act_mmap[:subset_size]
                self.idx_mmap = self.idx_mmap[:subset_size]
            else:
                self.mmap = self.mmap[:subset_size]
        self.activation_shape = self.metadata["activation_shape"]

    def __len__(self):
        return len(self.metadata["filenames"])

    def __getitem__(self, idx):
        filename = self.metadata["filenames"][idx]
        shape = self.metadata["tensor_shape"]
        if self.activation_type == "indexed":
            act = torch.from_numpy(np.array(self.act_mmap[idx]).reshape(shape))
            ind = torch.from_numpy(np.array(self.idx_mmap[idx]).reshape(shape))
            return act, ind, filename
        return torch.from_numpy(np.array(self.mmap[idx]).reshape(shape)), filename


class MemoryMappedActivationDataLoader(DataLoader):
    """reference :177-206 (including its `len() == dataset // batch_size` quirk)."""

    def __init__(self, data_path: str, layer_name: str, batch_size: int, dl_max_workers: int,
                 subset_size: Optional[int] = None, dl_kwargs: dict = {}):
        self._dataset = MemoryMappedActivationsDataset(data_path, layer_name, subset_size)
        dl_kwargs = {"batch_size": batch_size, "num_workers": dl_max_workers, **dl_kwargs}
        super().__init__(self._dataset, **dl_kwargs)
        self.activation_shape = self.dataset.activation_shape
        self.activation_type = self.dataset.activation_type
        self.dataset_length = len(self._dataset)

    def __len__(self):
        return len(self._dataset) // self.batch_size


def audio_num_samples(filename: str) -> tuple[int, int]:
    """(num_samples, sample_rate) of an audio file, for the trim arithmetic of utils/activations.py:19-29.
    Read once per file when a device store is built (the reference decodes every file on every query)."""
    try:
        import soundfile as sf

        info = sf.info(filename)
        return int(info.frames), int(info.samplerate)
    except ImportError:
        pass
    import torchaudio

    audio, sr = torchaudio.load(filename)
    return int(audio.shape[-1]), int(sr)


def n_frames_from_samples(num_samples: int, sample_rate: int = SAMPLE_RATE) -> int:
    """utils/activations.py:26-28 in the same python-float arithmetic."""
    return int((num_samples / sample_rate) / TIMESTEP_S)


class DeviceActivationStore:
    """All stored activations of a MemoryMappedActivationDataLoader resident in HBM (C5: 23 GB dense fp32, or
    1.9 GB values + 3.8 GB int64 indices), plus the per-file trimmed length."""

    def __init__(self, dataset: MemoryMappedActivationsDataset, device="cuda",
                 num_samples: Optional[dict] = None, frames_fn: Optional[Callable[[str], int]] = None,
                 chunk_files: int = 256, feature_major: Optional[bool] = None,
                 shard: Optional[tuple] = None):
        """shard=(rank, world): keep only this rank's contiguous block of files on the device (files
        are independent, so the search shards by files; `top_activations` then exchanges the per-file maxima)."""
        self.filenames = list(dataset.metadata["filenames"])
        self.activation_type = dataset.activation_type
        T, F = dataset.metadata["tensor_shape"]
        self.T = T
        self.n_total = len(self.filenames)
        self.shard = shard
        if shard is None:
            self.lo, self.hi = 0, self.n_total
        else:
            rank, world = shard
            self.per = -(-self.n_total // world)
            self.lo, self.hi = min(self.n_total, rank * self.per), min(self.n_total, (rank + 1) * self.per)
        n = self.hi - self.lo
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceActivationStore lives in GPU memory (no CPU fallback)")

        def upload(mm, dtype_out=None):
            out = torch.empty((n, T, F), dtype=dtype_out or torch.from_numpy(np.zeros(1, mm.dtype)).dtype, device=dev)
            for s in range(0, n, chunk_files):
                blk = torch.from_numpy(np.ascontiguousarray(mm[self.lo + s:self.lo + min(n, s + chunk_files)]))
                blk = blk.view(-1, T, F)
                if dtype_out is not None and blk.dtype != dtype_out:
                    out[s:s + blk.shape[0]].copy_(blk.pin_memory().to(dev, non_blocking=True))  # converts on the device
                else:
                    out[s:s + blk.shape[0]].copy_(blk.pin_memory(), non_blocking=True)
            torch.cuda.synchronize(dev)
            return out

        self.acts_fm = None
        self._table = None
        if self.activation_type == "indexed":
            self.vals = upload(dataset.act_mmap).float()
            # the .npy holds int64 indices (torch.topk); a dictionary never has 2^31 features, so they are narrowed
            # chunk by chunk on the way in: the per-query index scan reads half the bytes
            self.idx = upload(dataset.idx_mmap, torch.int32)
        else:
            acts = upload(dataset.mmap)
            self.acts = acts if acts.dtype in (torch.float32, torch.float16) else acts.float()
            # (feature_major=True keeps a second, [F, N_files, T] copy so that a single-feature scan is contiguous; the
            #  all-feature table below makes scans unnecessary, so it is no longer built by default)
            if feature_major:
                self.acts_fm = self.acts.permute(2, 0, 1).contiguous()
        if frames_fn is None:
            if num_samples is not None:
                frames_fn = lambda f: n_frames_from_samples(num_samples[f])  # noqa: E731
            else:
                def frames_fn(f):
                    ns, sr = audio_num_samples(f)
                    return n_frames_from_samples(ns, sr)
        self.n_frames_host = [min(int(frames_fn(f)), T) for f in self.filenames]
        self.n_frames = torch.tensor(self.n_frames_host[self.lo:self.hi], dtype=torch.int32, device=dev)


    def feature_table(self):
        """(vmax, amax, vabs) [n_files_local, F] of EVERY feature of a dense store, built by one pass over it on first
        use (freud_search_table_dense); every later query is a column gather + the ranking kernel."""
        if self._table is None:
            from .. import ops

            self._table = ops.search_table_dense(self.acts, self.n_frames)
        return self._table


class DeviceResidentActivationLoader:
    """Training feed with the whole dense activation set resident in HBM.

    Drop-in for `MemoryMappedActivationDataLoader(train_folder, layer, batch_size, workers, dl_kwargs={"shuffle": True,
    "drop_last": True})` (train_sae.py:322-334): same `(activations [B,T,d], filenames)` batches in the SAME order --
    the index batches come from a real torch `DataLoader` over the file indices, so the global RNG is consumed
    exactly as by the loader the reference builds -- but a batch is one device-side row gather instead of B page-cache
    reads, a collate and a 147 MB host->device copy per step (which would otherwise bound a 3 ms step)."""

    def __init__(self, data_path: str, layer_name: str, batch_size: int, dl_max_workers: int = 0,
                 subset_size: Optional[int] = None, dl_kwargs: dict = {}, device="cuda", chunk_files: int = 256):
        del dl_max_workers  # no worker processes: the data never leaves the device
        self._dataset = MemoryMappedActivationsDataset(data_path, layer_name, subset_size)
        if self._dataset.activation_type != "tensor":
            raise ValueError("training reads dense activations ({layer}_tensors.npy)")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceResidentActivationLoader lives in GPU memory (no CPU fallback)")
        self.batch_size = batch_size
        self.shuffle = bool(dl_kwargs.get("shuffle", False))
        self.drop_last = bool(dl_kwargs.get("drop_last", False))
        self.generator = dl_kwargs.get("generator")
        self.activation_shape = self._dataset.activation_shape
        self.activation_type = "tensor"
        self.dataset_length = len(self._dataset)
        self.filenames = list(self._dataset.metadata["filenames"])
        T, F = self._dataset.metadata["tensor_shape"]
        n = self.dataset_length
        mm = self._dataset.mmap
        self.acts = torch.empty((n, T, F), dtype=torch.from_numpy(np.zeros(1, mm.dtype)).dtype, device=dev)
        for s in range(0, n, chunk_files):
            blk = torch.from_numpy(np.ascontiguousarray(mm[s:s + chunk_files])).view(-1, T, F)
            self.acts[s:s + blk.shape[0]].copy_(blk.pin_memory(), non_blocking=True)
        torch.cuda.synchronize(dev)

    @property
    def dataset(self):
        return self._dataset

    def __len__(self):  # the reference loader's own quirk: len // batch_size even without drop_last (:205-206)
        return self.dataset_length // self.batch_size

    def __iter__(self):
        index_batches = DataLoader(range(self.dataset_length), batch_size=self.batch_size, shuffle=self.shuffle,
                                   drop_last=self.drop_last, generator=self.generator, num_workers=0,
                                   collate_fn=list)
        for idx in index_batches:
            sel = torch.tensor(idx, dtype=torch.long, device=self.acts.device)
            yield self.acts.index_select(0, sel), [self.filenames[i] for i in idx]
