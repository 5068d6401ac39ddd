"""Device-resident training feed and SAE-latent collection on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dense_store(tmp, n_files=23, T=40, d=64, seed=0):
    rng = np.random.default_rng(seed)
    acts = rng.standard_normal((n_files, T, d)).astype(np.float32)
    names = [f"/audio/clip_{i:03d}.flac" for i in range(n_files)]
    os.makedirs(tmp, exist_ok=True)
    np.save(f"{tmp}/layer_tensors.npy", acts.reshape(n_files, -1))
    json.dump({"tensor_shape": [T, d], "activation_shape": [T, d], "filenames": names},
              open(f"{tmp}/layer_metadata.json", "w"))
    return acts, names


def test_device_resident_loader_yields_the_reference_loaders_batches(tmp_path):
    """Same batches, in the same shuffled order, as the DataLoader train_sae.py:322-334 builds (shuffle, drop_last)."""
    from freud_b200.dataset.activations import DeviceResidentActivationLoader, MemoryMappedActivationDataLoader

    _dense_store(str(tmp_path))
    kw = {"shuffle": True, "drop_last": True}
    for epoch_seed in (0, 5):
        torch.manual_seed(epoch_seed)
        ref = [(a.clone(), list(f)) for a, f in MemoryMappedActivationDataLoader(str(tmp_path), "layer", 4, 0,
                                                                               dl_kwargs=kw)]
        dev_loader = DeviceResidentActivationLoader(str(tmp_path), "layer", 4, 0, dl_kwargs=kw)
        torch.manual_seed(epoch_seed)
        got = [(a.cpu(), f) for a, f in dev_loader]
        assert len(got) == len(ref) == 5 == len(dev_loader)
        for (ra, rf), (ga, gf) in zip(ref, got):
            assert rf == gf and torch.equal(ra, ga)
    # sequential, keep the ragged tail (drop_last False): 6 batches while len() keeps the reference's quirk
    seq = list(DeviceResidentActivationLoader(str(tmp_path), "layer", 4, 0))
    assert len(seq) == 6 and seq[-1][0].shape[0] == 3


def test_collect_sae_latents_writes_the_reference_layout(tmp_path):
    from freud_b200.collect import collect_sae_latents
    from freud_b200.dataset.activations import MemoryMappedActivationsDataset
    from freud_b200.models.config import L1AutoEncoderConfig, TopKAutoEncoderConfig
    from freud_b200.models.l1autoencoder import L1AutoEncoder
    from freud_b200.models.topkautoencoder import TopKAutoEncoder

    acts, names = _dense_store(f"{tmp_path}/whisper", n_files=11)
    torch.manual_seed(0)
    sae = TopKAutoEncoder(64, TopKAutoEncoderConfig.from_dict({"n_dict_components": 512, "k": 32})).cuda()
    sae.precision = "fp32"
    collect_sae_latents(f"{tmp_path}/whisper", "layer", sae, 4, f"{tmp_path}/sae_topk")
    ds = MemoryMappedActivationsDataset(f"{tmp_path}/sae_topk", "layer")
    assert ds.activation_type == "indexed" and ds.metadata["filenames"] == names
    assert ds.metadata["tensor_shape"] == [40, 32] and ds.metadata["activation_shape"] == [40, 512]
    assert ds.act_mmap.shape == (11, 40 * 32) and ds.act_mmap.dtype == np.float32
    assert ds.idx_mmap.dtype == np.int64
    enc = sae.encode(torch.from_numpy(acts).cuda())
    v, i, f = ds[7]
    assert f == names[7]
    assert torch.equal(v, enc.top_acts[7].cpu()) and torch.equal(i, enc.top_indices[7].cpu())
    # a second run replaces the files instead of appending to them (collect_activations.py:110-113)
    collect_sae_latents(f"{tmp_path}/whisper", "layer", sae, 4, f"{tmp_path}/sae_topk", collect_max=5)
    assert len(MemoryMappedActivationsDataset(f"{tmp_path}/sae_topk", "layer")) == 5

    l1 = L1AutoEncoder(64, L1AutoEncoderConfig.from_dict({"n_dict_components": 96})).cuda()
    l1.precision = "fp32"
    collect_sae_latents(f"{tmp_path}/whisper", "layer", l1, 4, f"{tmp_path}/sae_l1")
    dl = MemoryMappedActivationsDataset(f"{tmp_path}/sae_l1", "layer")
    assert dl.activation_type == "tensor" and dl.metadata["tensor_shape"] == [40, 96]
    lat = l1.encode(torch.from_numpy(acts).cuda()).latent
    # not bit-equal: every L1 encode() re-normalises decoder.weight in place (reference quirk, l1autoencoder.py:71-73)
    assert float((dl[3][0] - lat[3].cpu()).abs().max()) < 1e-5


def test_example_training_loop_writes_a_reference_checkpoint(tmp_path):
    """examples/train_from_store.py: reference config schema in, reference checkpoint layout out, loss going down."""
    import importlib.util
    import sys

    from freud_b200.dataset.activations import init_sae_from_checkpoint

    _dense_store(f"{tmp_path}/train", n_files=24, T=50, d=64)
    cfg = {"seed": 0, "train_folder": f"{tmp_path}/train", "val_folder": f"{tmp_path}/train", "device": "cuda",
           "run_dir": f"{tmp_path}/run", "lr": 2e-3, "weight_decay": 0.0, "steps": 40, "clip_thresh": 1.0,
           "batch_size": 8, "dl_max_workers": 0, "log_tb_every": 20, "save_every": 20, "val_every": 100,
           "start_checkpoint": None, "whisper_config": {"model": "tiny", "layer_name": "layer"},
           "optimizer": "adam", "scheduler": "linear", "scheduler_params": {"num_warmup_steps": 5}, "from_disk": True,
           "autoencoder_variant": "topk",
           "autoencoder_config": {"n_dict_components": 256, "k": 32, "auxk_alpha": 1 / 32, "dead_feature_threshold": 1e6}}
    spec = importlib.util.spec_from_file_location(
        "train_from_store", os.path.join(os.path.dirname(os.path.dirname(__file__)), "examples", "train_from_store.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model, trainer = mod.train(cfg, precision="fp32")
    ckpt = torch.load(f"{tmp_path}/run/checkpoints/step40.pth", map_location="cpu")
    assert set(ckpt) == {"model", "optimizer", "scheduler", "step", "best_val_loss", "hparams"} and ckpt["step"] == 40
    assert set(ckpt["model"]) == {"W_dec", "b_dec", "encoder.weight", "encoder.bias"}
    assert ckpt["hparams"]["activation_size"] == 64
    loaded = init_sae_from_checkpoint(f"{tmp_path}/run/checkpoints/step40.pth", device="cuda")
    x = torch.randn(2, 50, 64, device="cuda")
    first = init_sae_from_checkpoint(f"{tmp_path}/run/checkpoints/step20.pth", device="cuda")
    with torch.no_grad():
        assert float(loaded(x).fvu) < 1.0
    assert torch.equal(loaded.W_dec.detach().cpu(), model.W_dec.detach().cpu())
    assert not torch.equal(first.W_dec.cpu(), loaded.W_dec.cpu())
    # resume (train_sae.py:408-410): from the step-20 checkpoint to step 40.  The dead-latent counters restart at zero
    # as upstream, and no latent can die within 40 x 400 tokens, so the resumed run retraces the first one
    cfg2 = dict(cfg, start_checkpoint=f"{tmp_path}/run/checkpoints/step20.pth", run_dir=f"{tmp_path}/run2")
    model2, trainer2 = mod.train(cfg2, precision="fp32")
    assert trainer2.step_count == 40
    assert os.path.exists(f"{tmp_path}/run2/checkpoints/step40.pth")
