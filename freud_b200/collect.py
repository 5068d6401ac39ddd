"""SAE-latent collection over a stored activation set.

Mirror of the SAE half of src/scripts/collect_activations.py:12-136: the reference runs Whisper and the SAE on the
fly (`FlyActivationDataLoader`, dataset/activations.py:90-110) and appends one row per file to memory-mappable .npy
files.  Whisper is out of scope here, so the input is a stored dense activation set (the files Whisper collection
wrote); the SAE half -- `sae.encode` on every file, written in the reference's own layout -- runs on the fused
encoder + top-k kernel.  Output (collect_activations.py:37-41,101-108):

  {layer}_metadata.json           {"tensor_shape": [T, k or n], "activation_shape": [T, n], "filenames": [...]}
  TopK SAE: {layer}_activation_values.npy fp32 + {layer}_feature_indices.npy int64, both [N_files, T*k]
  L1 SAE  : {layer}_tensors.npy fp32 [N_files, T*n]
"""
from __future__ import annotations

import json
import os
from pathlib import Path
from typing import List, Optional

import numpy as np
import torch

from .dataset.activations import MemoryMappedActivationDataLoader
from .models.topkautoencoder import TopKAutoEncoder


class NpyRowAppender:
    """Append [rows, row_len] blocks to a C-order .npy file (what `NpyAppendArray` does for the reference; that
    package is not a dependency here).  The v1 header is padded to a fixed 128 bytes so the growing row count can
    be rewritten in place; the file is a plain .npy at every point and loads with `np.load(..., mmap_mode="r")`."""
    HEADER_BYTES = 128

    def __init__(self, path):
        self.path = str(path)
        self.rows = 0
        self.row_len = None
        self.dtype = None
        self.f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _header(self):
        d = "{'descr': '%s', 'fortran_order': False, 'shape': (%d, %d), }" % (
            np.lib.format.dtype_to_descr(self.dtype), self.rows, self.row_len)
        body = d.encode("latin1")
        pad = self.HEADER_BYTES - 10 - len(body) - 1
        if pad < 0:
            raise ValueError("npy header does not fit the reserved space")
        hlen = self.HEADER_BYTES - 10
        return b"\x93NUMPY\x01\x00" + bytes((hlen & 0xff, hlen >> 8)) + body + b" " * pad + b"\n"

    def append(self, block: np.ndarray):
        block = np.ascontiguousarray(block)
        if block.ndim != 2:
            raise ValueError("append expects a [rows, row_len] block")
        if self.f is None:
            if os.path.exists(self.path):  # continue an existing file written by this class
                arr = np.load(self.path, mmap_mode="r")
                self.rows, self.row_len, self.dtype = arr.shape[0], arr.shape[1], arr.dtype
                del arr
                self.f = open(self.path, "r+b")
            else:
                self.row_len, self.dtype = block.shape[1], block.dtype
                self.f = open(self.path, "w+b")
                self.f.write(self._header())
        if block.shape[1] != self.row_len or block.dtype != self.dtype:
            raise ValueError("all appended rows must share one length and dtype")
        self.f.seek(0, os.SEEK_END)
        self.f.write(block.tobytes())
        self.rows += block.shape[0]
        self.f.seek(0)
        self.f.write(self._header())

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None


def save_data_for_memory_mapping(metadata_file: Path, data_files: List[Path], data: List[torch.Tensor],
                                 filenames: List[str], tensor_shape: List[int], activation_shape: List[int]):
    """collect_activations.py:12-63: append one row per file to every data file and rewrite the metadata."""
    assert len(data[0]) == len(filenames), "Number of data tensors and filenames must match"
    if os.path.exists(metadata_file):
        with open(metadata_file, "r") as f:
            metadata = json.load(f)
    else:
        metadata = {"tensor_shape": list(tensor_shape), "activation_shape": list(activation_shape), "filenames": []}
    for t in data:
        if metadata["tensor_shape"] != list(t.shape[1:]):
            raise ValueError(f"All tensors must have the same shape as the first tensor. "
                             f"Expected {metadata['tensor_shape']}, got {list(t.shape[1:])}")
    metadata["filenames"].extend(filenames)
    with open(metadata_file, "w") as f:
        json.dump(metadata, f)
    for file, t in zip(data_files, data):
        with NpyRowAppender(file) as ap:
            ap.append(t.detach().cpu().numpy().reshape(t.shape[0], -1))


def collect_sae_latents(data_path: str, layer_name: str, sae, batch_size: int, out_folder: str,
                        max_workers: int = 0, collect_max: Optional[int] = None, device="cuda"):
    """SAE activations of every stored file, written in the reference's memory-mappable layout.  `sae` is a
    TopKAutoEncoder / L1AutoEncoder (e.g. from init_sae_from_checkpoint) already on `device`."""
    loader = MemoryMappedActivationDataLoader(data_path, layer_name, batch_size, max_workers, collect_max)
    if loader.activation_type != "tensor":
        raise ValueError("SAE latents are computed from stored dense activations")
    out = Path(out_folder)
    metadata_file = out / f"{layer_name}_metadata.json"
    indexed = isinstance(sae, TopKAutoEncoder)
    data_files = ([out / f"{layer_name}_activation_values.npy", out / f"{layer_name}_feature_indices.npy"]
                  if indexed else [out / f"{layer_name}_tensors.npy"])
    for file in [metadata_file] + data_files:  # collect_activations.py:110-113
        if file.exists():
            file.unlink()
    out.mkdir(parents=True, exist_ok=True)
    T = loader.activation_shape[0]
    activation_shape = [T, sae.n_dict_components]
    with torch.no_grad():
        for acts, names in loader:
            x = acts.to(device, non_blocking=True).float()
            if indexed:
                enc = sae.encode(x)
                data = [enc.top_acts.float(), enc.top_indices]
            else:
                data = [sae.encode(x).latent]
            save_data_for_memory_mapping(metadata_file, data_files, data, list(names), list(data[0].shape[1:]),
                                         activation_shape)
    return metadata_file, data_files
