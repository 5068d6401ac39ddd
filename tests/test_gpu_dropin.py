"""End-to-end drop-in proof: the reference's OWN training loop (`train()`, src/scripts/train_sae.py:297-601) runs
unmodified on this build's kernels after `freud_b200.compat.install_as_src()`, writes the reference's checkpoint
layout, and the result loads back into the UNMODIFIED reference classes.

Needs the reference tree (oracle/_ref, populated by oracle/make_ref.sh; it travels to the GPU box) -- skipped without.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import ref_shims

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_shims.reference_available(), reason="reference tree (oracle/_ref) absent")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T, F, N_FILES = 100, 64, 24


def _write_store(folder, seed):
    g = torch.Generator().manual_seed(seed)
    mix = torch.randn(16, F, generator=g)
    x = (torch.randn(N_FILES, T, 16, generator=g) @ mix + 0.1 * torch.randn(N_FILES, T, F, generator=g)).float()
    os.makedirs(folder, exist_ok=True)
    np.save(os.path.join(folder, "layer_tensors.npy"), x.reshape(N_FILES, -1).numpy())
    json.dump({"tensor_shape": [T, F], "activation_shape": [T, F], "filenames": [f"f{i}.wav" for i in range(N_FILES)]},
              open(os.path.join(folder, "layer_metadata.json"), "w"))
    return x


def _config(folder, run_dir, device, variant):
    cfg = dict(seed=0, train_folder=folder, val_folder=folder, device=device, run_dir=run_dir, lr=2e-3, weight_decay=0.0,
               steps=20, clip_thresh=1.0, batch_size=4, dl_max_workers=0, log_tb_every=5, save_every=10,
               val_every=10 ** 9, start_checkpoint=None, whisper_config={"model": "tiny", "layer_name": "layer"},
               optimizer="adam", scheduler="linear", scheduler_params={"num_warmup_steps": 4}, from_disk=True,
               autoencoder_variant=variant)
    if variant == "topk":
        cfg["autoencoder_config"] = {"n_dict_components": 512, "k": 32, "auxk_alpha": 1 / 32, "multi_topk": False,
                                     "normalize_decoder": True, "dead_feature_threshold": 1200}
    else:
        cfg.update(optimizer="radam", scheduler="cosine")
        cfg["autoencoder_config"] = {"n_dict_components": 96, "recon_alpha": 100.0}
    return cfg


_PURE_REFERENCE = r"""
import json, sys, torch
sys.path.insert(0, {root!r})
from oracle import ref_shims
ref_shims.install()
import src.scripts.train_sae as ts
cfg = json.load(open({cfg!r}))
cfg["device"] = torch.device("cpu")
ts.train(**cfg)
"""


@pytest.mark.parametrize("variant", ["topk", "l1"])
def test_reference_train_loop_runs_on_this_build(tmp_path, variant):
    folder = str(tmp_path / "store")
    x = _write_store(folder, 3)
    # (1) the pure reference on the CPU, in a separate process (its modules must not see the aliases)
    ref_cfg = _config(folder, str(tmp_path / "run_ref"), "cpu", variant)
    cfg_file = str(tmp_path / "ref_cfg.json")
    json.dump(ref_cfg, open(cfg_file, "w"))
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    subprocess.run([sys.executable, "-c", _PURE_REFERENCE.format(root=ROOT, cfg=cfg_file)], check=True, env=env,
                   cwd=ROOT, timeout=600)
    # (2) the same train() on this build: reference loop, reference loader, our modules / optimiser / kernels
    ref_shims.install()
    import src.models.l1autoencoder as ref_l1_mod          # handles on the unmodified classes, taken BEFORE aliasing
    import src.models.topkautoencoder as ref_topk_mod
    import src.models.config as ref_cfg_mod
    RefTopK, RefL1 = ref_topk_mod.TopKAutoEncoder, ref_l1_mod.L1AutoEncoder
    from freud_b200 import _lib, compat

    saved = {k: sys.modules.get(k) for k in ("src.models.config", "src.models.l1autoencoder",
                                             "src.models.topkautoencoder", "src.utils.models")}
    try:
        compat.install_as_src()
        import src.scripts.train_sae as ts
        import src.dataset.activations as ref_ds

        from freud_b200.models.topkautoencoder import TopKAutoEncoder as OurTopK
        assert ts.TopKAutoEncoder is OurTopK and ref_ds.TopKAutoEncoder is OurTopK
        k0 = _lib.kernel_launches
        ours_cfg = _config(folder, str(tmp_path / "run_ours"), torch.device("cuda"), variant)
        ts.train(**ours_cfg)
        assert _lib.kernel_launches - k0 > 20 * 5, "the reference loop did not reach this build's kernels"
        ckpt = str(tmp_path / "run_ours" / "checkpoints" / "step20.pth")
        assert os.path.exists(ckpt) and os.path.exists(str(tmp_path / "run_ours" / "checkpoints" / "step10.pth"))
        # (3) reload through the reference's own init_sae_from_checkpoint (now bound to our classes) ...
        ours = ref_ds.init_sae_from_checkpoint(ckpt, "cuda")
        state = torch.load(ckpt, map_location="cpu")
        assert set(state) == {"model", "optimizer", "scheduler", "step", "best_val_loss", "hparams"}
        assert state["step"] == 20
        opt_state = state["optimizer"]["state"]
        assert all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in opt_state.values())
    finally:
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    # ... and into the UNMODIFIED reference class on the CPU: same forward on a held-out batch
    hp = state["hparams"]
    xb = x[:4]
    if variant == "topk":
        ref_model = RefTopK(hp["activation_size"], ref_cfg_mod.TopKAutoEncoderConfig.from_dict(hp["autoencoder_config"]))
        ref_model.load_state_dict(state["model"])
        ours.precision = "fp32"
        with torch.no_grad():
            ro = ref_model(xb)
            oo = ours(xb.cuda())
        assert abs(float(oo.fvu) - float(ro.fvu)) / float(ro.fvu) < 1e-4
        assert float((oo.sae_out.cpu() - ro.sae_out).abs().max() / ro.sae_out.abs().max()) < 1e-4
        # the pure-reference CPU run of the same config (bf16 autocast there, bf16 tensor-core mode here) ends close by
        pure = torch.load(str(tmp_path / "run_ref" / "checkpoints" / "step20.pth"), map_location="cpu")
        pure_model = RefTopK(hp["activation_size"], ref_cfg_mod.TopKAutoEncoderConfig.from_dict(hp["autoencoder_config"]))
        pure_model.load_state_dict(pure["model"])
        with torch.no_grad():
            po = pure_model(xb)
        assert float(ro.fvu) < 0.9, "training did not reduce the loss"
        assert abs(float(ro.fvu) - float(po.fvu)) / float(po.fvu) < 0.05
    else:
        ref_model = RefL1(hp["activation_size"], ref_cfg_mod.L1AutoEncoderConfig.from_dict(hp["autoencoder_config"]))
        ref_model.load_state_dict(state["model"])
        ours.precision = "fp32"
        with torch.no_grad():
            ro = ref_model(xb)
            oo = ours(xb.cuda())
        assert abs(float(oo.reconstruction_loss) - float(ro.reconstruction_loss)) / float(ro.reconstruction_loss) < 1e-4
        assert abs(float(oo.l1_loss) - float(ro.l1_loss)) / float(ro.l1_loss) < 1e-4
        pure = torch.load(str(tmp_path / "run_ref" / "checkpoints" / "step20.pth"), map_location="cpu")
        pure_model = RefL1(hp["activation_size"], ref_cfg_mod.L1AutoEncoderConfig.from_dict(hp["autoencoder_config"]))
        pure_model.load_state_dict(pure["model"])
        with torch.no_grad():
            po = pure_model(xb)
        lo, lp = float(ro.reconstruction_loss + ro.l1_loss), float(po.reconstruction_loss + po.l1_loss)
        assert abs(lo - lp) / lp < 0.05
