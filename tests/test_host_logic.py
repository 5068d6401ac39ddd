"""CPU-side tests: C-ABI surface, config / state_dict / checkpoint layout, schedules, loader, DP arithmetic."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    from freud_b200 import build

    return build.build()


def test_c_abi_exports_every_declared_symbol(built_lib):
    """Every prototype in include/freud_b200.h is exported by the shared library and bound in _lib.SIGNATURES."""
    from freud_b200 import _lib

    header = open(os.path.join(ROOT, "include", "freud_b200.h")).read()
    declared = set(re.findall(r"\b(freud_[a-z0-9_]+)\s*\(", header)) - {"freud_tensor_list"}
    handle = ctypes.CDLL(built_lib)
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in the header but not exported"
    assert declared - {"freud_last_error", "freud_version"} == set(_lib.SIGNATURES)
    assert set(_lib.KERNELS_PER_CALL) == set(_lib.SIGNATURES)
    assert _lib.lib().freud_version() >= 100


def test_c_abi_argument_errors_are_reported_not_thrown(built_lib):
    from freud_b200 import _lib

    with pytest.raises(RuntimeError, match="n >= 64"):
        _lib.call("freud_topk_encode", None, None, None, None, None, None, None, 128, 64, 8, 0, None, 0, None, None)
    assert b"n >= 64" in _lib.lib().freud_last_error()


def test_config_from_dict_drops_unknown_keys_and_defaults():
    from freud_b200.models.config import L1AutoEncoderConfig, TopKAutoEncoderConfig

    # the autoencoder_config block of configs/train/tiny_topk.json:6-14 (carries the non-dataclass key)
    cfg = TopKAutoEncoderConfig.from_dict({"expansion_factor": 16, "normalize_decoder": True, "k": 32,
                                           "multi_topk": False, "auxk_alpha": 0.03125,
                                           "dead_feature_threshold": 1000000.0})
    assert (cfg.expansion_factor, cfg.k, cfg.auxk_alpha, cfg.multi_topk, cfg.normalize_decoder) == (16, 32, 0.03125,
                                                                                                     False, True)
    assert not hasattr(cfg, "dead_feature_threshold")
    l1 = L1AutoEncoderConfig.from_dict({"n_dict_components": 200, "recon_alpha": 1e4})
    assert (l1.n_dict_components, l1.recon_alpha, l1.expansion_factor) == (200, 1e4, 32)
    assert TopKAutoEncoderConfig().to_dict()["k"] == 32


def test_state_dict_layout_and_seeded_init_match_reference_contract():
    """key names/shapes; the same torch seed gives the same init as the reference (golden)."""
    from freud_b200.models.config import L1AutoEncoderConfig, TopKAutoEncoderConfig
    from freud_b200.models.l1autoencoder import L1AutoEncoder
    from freud_b200.models.topkautoencoder import TopKAutoEncoder
    from tests.util import load_golden

    torch.manual_seed(3)
    m = TopKAutoEncoder(24, TopKAutoEncoderConfig.from_dict({"n_dict_components": 96, "k": 4,
                                                             "normalize_decoder": False}))
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {
        "W_dec": (96, 24), "b_dec": (24,), "encoder.weight": (96, 24), "encoder.bias": (96,)}
    z, _ = load_golden("topk_decoder_norm")
    assert np.array_equal(m.W_dec.data.numpy(), z["W_dec_in"])  # reference init under torch.manual_seed(3)
    m.set_decoder_norm_to_unit_norm()
    assert np.allclose(m.W_dec.data.numpy(), z["W_dec_unit"], rtol=1e-6, atol=0)
    l1 = L1AutoEncoder(32, L1AutoEncoderConfig.from_dict({"n_dict_components": 40}))
    assert {k: tuple(v.shape) for k, v in l1.state_dict().items()} == {"encoder_bias": (40,),
                                                                       "decoder.weight": (32, 40)}
    assert l1.n_dict_components == 40 and m.n_dict_components == 96 and m.d_in == 24


def test_checkpoint_roundtrip_through_init_sae_from_checkpoint(tmp_path):
    """The reference checkpoint layout (train_sae.py:232-248,336-351) loads through init_sae_from_checkpoint."""
    from freud_b200.dataset.activations import init_sae_from_checkpoint
    from freud_b200.models.config import TopKAutoEncoderConfig
    from freud_b200.models.topkautoencoder import TopKAutoEncoder

    acfg = {"expansion_factor": 4, "k": 8, "auxk_alpha": 0.03125, "dead_feature_threshold": 1e6}
    model = TopKAutoEncoder(16, TopKAutoEncoderConfig.from_dict(acfg))
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    ckpt = {"model": model.state_dict(), "optimizer": opt.state_dict(), "scheduler": {}, "step": 7,
            "best_val_loss": 0.5, "hparams": {"autoencoder_variant": "topk", "autoencoder_config": acfg,
                                               "activation_size": 16}}
    path = str(tmp_path / "step7.pth")
    torch.save(ckpt, path)
    loaded = init_sae_from_checkpoint(path, device="cpu")
    assert isinstance(loaded, TopKAutoEncoder) and not loaded.training
    for k, v in model.state_dict().items():
        assert torch.equal(loaded.state_dict()[k], v)


def test_fused_optimizer_state_layout_matches_torch():
    """exp_avg / exp_avg_sq / step keys as torch.optim.Adam, so optimizer.state_dict() checkpoints interchange."""
    from freud_b200.optim import FusedAdam, FusedRAdam

    p = torch.nn.Parameter(torch.zeros(4))
    for cls, ref in ((FusedAdam, torch.optim.Adam), (FusedRAdam, torch.optim.RAdam)):
        ours, theirs = cls([p], lr=1e-3), ref([p], lr=1e-3)
        assert set(ours.state_dict()["param_groups"][0]) >= {"lr", "betas", "eps", "params"}
        theirs_state = {"step": torch.tensor(3.0), "exp_avg": torch.ones(4), "exp_avg_sq": torch.ones(4)}
        sd = theirs.state_dict()
        sd["state"] = {0: theirs_state}
        ours_sd = ours.state_dict()
        ours_sd["state"] = sd["state"]
        ours.load_state_dict(ours_sd)
        assert set(ours.state[p]) == {"step", "exp_avg", "exp_avg_sq"}


def test_lr_schedules_match_reference_schedulers():
    from transformers import get_linear_schedule_with_warmup

    from freud_b200.trainer import linear_schedule_with_warmup

    p = [torch.nn.Parameter(torch.zeros(1))]
    a, b = torch.optim.SGD(p, lr=1e-4), torch.optim.SGD(p, lr=1e-4)
    sa, sb = linear_schedule_with_warmup(a, 5, 40), get_linear_schedule_with_warmup(b, 5, 40)
    for _ in range(45):
        assert a.param_groups[0]["lr"] == b.param_groups[0]["lr"]
        a.step(); b.step(); sa.step(); sb.step()


def test_memory_mapped_loader_contract(tmp_path):
    """On-disk format + loader attributes of src/dataset/activations.py:116-206 (incl. the len() quirk)."""
    from freud_b200.dataset.activations import MemoryMappedActivationDataLoader, n_frames_from_samples

    n_files, T, F = 7, 10, 4
    arr = np.arange(n_files * T * F, dtype=np.float32).reshape(n_files, T * F)
    np.save(tmp_path / "layer_tensors.npy", arr)
    names = [f"/a/{i}.flac" for i in range(n_files)]
    json.dump({"tensor_shape": [T, F], "activation_shape": [T, F], "filenames": names},
              open(tmp_path / "layer_metadata.json", "w"))
    dl = MemoryMappedActivationDataLoader(str(tmp_path), "layer", batch_size=3, dl_max_workers=0)
    assert dl.activation_type == "tensor" and dl.dataset_length == 7 and list(dl.activation_shape) == [T, F]
    batches = list(dl)
    assert len(dl) == 2 and len(batches) == 3  # reference quirk 11: len() floors, iteration yields the tail
    acts, fn = batches[0]
    assert acts.shape == (3, T, F) and list(fn) == names[:3]
    assert torch.equal(acts[1], torch.from_numpy(arr[1].reshape(T, F)))
    sub = MemoryMappedActivationDataLoader(str(tmp_path), "layer", batch_size=2, dl_max_workers=0, subset_size=4)
    assert sub.dataset_length == 4
    assert n_frames_from_samples(16000) == 50 and n_frames_from_samples(479999) == 1499


def test_install_as_src_aliases_reference_module_paths():
    import sys

    from freud_b200.compat import install_as_src

    saved = {k: sys.modules.get(k) for k in ("src.models.topkautoencoder", "src.models.l1autoencoder",
                                             "src.models.config", "src.utils.models")}
    try:
        install_as_src(patch_search=False)
        from freud_b200.models.topkautoencoder import TopKAutoEncoder

        assert sys.modules["src.models.topkautoencoder"].TopKAutoEncoder is TopKAutoEncoder
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_npy_row_appender_and_collection_metadata(tmp_path):
    """The appender keeps a valid, memory-mappable .npy after every append (collect_activations.py:60-63)."""
    import json

    import numpy as np
    import torch

    from freud_b200.collect import NpyRowAppender, save_data_for_memory_mapping

    p = tmp_path / "rows.npy"
    blocks = [np.arange(12, dtype=np.int64).reshape(2, 6), np.arange(12, 30, dtype=np.int64).reshape(3, 6)]
    for b in blocks:
        with NpyRowAppender(p) as ap:
            ap.append(b)
        assert np.load(p, mmap_mode="r").shape[1] == 6
    assert np.array_equal(np.load(p), np.concatenate(blocks))
    with pytest.raises(ValueError):
        with NpyRowAppender(p) as ap:
            ap.append(np.zeros((1, 5), dtype=np.int64))
    meta = tmp_path / "layer_metadata.json"
    files = [tmp_path / "layer_activation_values.npy", tmp_path / "layer_feature_indices.npy"]
    for s in range(2):
        vals = torch.rand(3, 5, 4)
        idx = torch.randint(0, 100, (3, 5, 4))
        save_data_for_memory_mapping(meta, files, [vals, idx], [f"f{s}{j}" for j in range(3)], [5, 4], [5, 100])
    m = json.load(open(meta))
    assert m["tensor_shape"] == [5, 4] and m["activation_shape"] == [5, 100] and len(m["filenames"]) == 6
    assert np.load(files[0]).shape == (6, 20) and np.load(files[1]).dtype == np.int64
    with pytest.raises(ValueError):
        save_data_for_memory_mapping(meta, files, [torch.rand(1, 5, 3), torch.zeros(1, 5, 3, dtype=torch.long)],
                                     ["x"], [5, 3], [5, 100])


def test_encoder_tail_split_plan_sizes():
    """freud_topk_encode_workspace is a pure host-side plan (no GPU needed): nothing to split when the row blocks fill
    whole waves of 148 SMs or the dictionary has too few tiles; otherwise S partial lists + the per-row thresholds."""
    from freud_b200 import _lib

    def need(N, n):
        out = ctypes.c_int64(-1)
        _lib.call("freud_topk_encode_workspace", N, n, ctypes.byref(out))
        return out.value

    assert need(148 * 128, 24576) == 0 and need(2 * 148 * 128, 24576) == 0   # whole waves
    assert need(77, 600) == 0                                                # 3 column tiles: not worth splitting
    for N, n in ((48000, 24576), (24000, 81920), (300, 2560), (19109, 8192)):
        b = need(N, n)
        rows = -(-N // 128) * 128
        assert b > 0 and (b - rows * 4) % (rows * 256) == 0                  # S * rows * 32 * 8 + rows * 4
        S = (b - rows * 4) // (rows * 256)
        assert 2 <= S <= 16 and 2 * S <= -(-n // 256)
    with pytest.raises(RuntimeError):
        need(128, 8)                                                         # n < 64 is rejected like the encoder


def test_fused_dp_flat_layout_plan():
    """freud_b200.fused_dp.plan_flat_layout (pure host logic of the fused data-parallel optimiser): every tensor starts
    on a 128-byte boundary, the per-rank slices tile the padded space without gaps or overlap and have equal,
    128-byte-multiple lengths (the kernels' 16-byte vector paths and the 8-element bf16 stores rely on it)."""
    from freud_b200.fused_dp import plan_flat_layout

    for numels, world in (([24576 * 768, 24576, 24576 * 768, 768], 8), ([1000 * 33, 1001, 1000 * 33, 33], 2),
                          ([7, 5, 3], 4), ([81920 * 1280, 81920, 81920 * 1280, 1280], 3)):
        offs, total, slices = plan_flat_layout(numels, world)
        assert all(o % 32 == 0 for o in offs)
        assert all(offs[i] + numels[i] <= offs[i + 1] for i in range(len(numels) - 1))
        assert offs[-1] + numels[-1] <= total
        assert len(slices) == world and slices[0][0] == 0 and slices[-1][1] == total
        assert all(slices[r][1] == slices[r + 1][0] for r in range(world - 1))
        lens = {hi - lo for lo, hi in slices}
        assert len(lens) == 1 and lens.pop() % 32 == 0
