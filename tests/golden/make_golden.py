"""Generate golden fixtures by running the UNMODIFIED reference (ksadov/FREUD).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/*.npz (small; committed).  The reference has no tests or
fixtures of its own, so these outputs of its real classes --
L1AutoEncoder / TopKAutoEncoder under torch autograd, torch.optim.Adam/RAdam,
clip_grad_norm_, the two LR schedulers, MemoryMappedActivationDataLoader and
top_activations -- are what pins the oracle (tests/test_oracle_golden.py) and,
through it, the CUDA path.
"""
import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_shims  # noqa: E402

ref_shims.install()

from src.models.config import L1AutoEncoderConfig, TopKAutoEncoderConfig  # noqa: E402
from src.models.l1autoencoder import L1AutoEncoder  # noqa: E402
from src.models.topkautoencoder import TopKAutoEncoder  # noqa: E402
from torch.optim import Adam, RAdam  # noqa: E402
from torch.optim.lr_scheduler import CosineAnnealingLR  # noqa: E402
from transformers import get_linear_schedule_with_warmup  # noqa: E402


def npify(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            v = v.detach()
            if v.dtype == torch.bfloat16:
                v = v.float()
            v = v.numpy()
        out[k] = np.asarray(v)
    return out


def synth_x(B, T, d, seed, with_mean=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, d, generator=g)
    if with_mean:
        x = x * (0.5 + torch.rand(T, 1, generator=g)) + torch.randn(d, generator=g)
    return x


def synth_x_margin(model, B, T, d, seed, k, margin):
    """Rows whose k-th / (k+1)-th pre-activation gap exceeds `margin` (relative), so that the
    selected set is stable under bf16 rounding and the bf16 golden is selection-independent."""
    cand = synth_x(B * 64, T, d, seed).reshape(-1, d)
    with torch.no_grad():
        v = torch.sort(model.pre_acts(cand), dim=-1, descending=True).values
    ok = (v[:, k - 1] - v[:, k]) > margin * v[:, k - 1]
    assert int(ok.sum()) >= B * T
    return cand[ok][: B * T].reshape(B, T, d).contiguous()


def topk_case(name, B, T, d, n, k, *, auxk_alpha, multi_topk, n_dead, autocast, steps, seed=0, margin=0.0):
    """One reference TopK train trajectory: body of train_sae.py:421-453, Adam + linear warmup."""
    torch.manual_seed(seed)
    cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": n, "k": k, "multi_topk": multi_topk,
                                           "auxk_alpha": auxk_alpha, "dead_feature_threshold": 5})
    model = TopKAutoEncoder(d, cfg)
    # non-trivial biases so the b_dec / b_enc paths are exercised
    g = torch.Generator().manual_seed(seed + 1)
    model.b_dec.data = 0.1 * torch.randn(d, generator=g)
    model.encoder.bias.data = 0.05 * torch.randn(n, generator=g)
    rec = {f"init.{k_}": v.clone() for k_, v in model.state_dict().items()}
    lr, clip, total_steps, warm = 1e-3, 1.0, 10, 2
    opt = Adam(model.parameters(), lr=lr)
    sched = get_linear_schedule_with_warmup(opt, num_warmup_steps=warm, num_training_steps=total_steps)
    frames = torch.zeros(n, dtype=torch.long)
    if n_dead:
        frames[torch.randperm(n, generator=g)[:n_dead]] = 10 ** 9
    dead_thresh = 10 ** 6
    for s in range(steps):
        x = synth_x_margin(model, B, T, d, seed + 10 + s, k, margin) if margin else synth_x(B, T, d, seed + 10 + s)
        rec[f"s{s}.x"] = x
        rec[f"s{s}.frames_in"] = frames.clone()
        rec[f"s{s}.lr"] = torch.tensor(opt.param_groups[0]["lr"], dtype=torch.float64)
        did_fire = torch.zeros(n, dtype=torch.bool)
        opt.zero_grad()
        from torch.amp import autocast as ac
        import contextlib
        ctx = ac("cpu") if autocast else contextlib.nullcontext()
        with ctx:
            dead_mask = frames > dead_thresh
            out, mse = model(x, dead_mask=dead_mask, return_mse=True)
            loss = out.fvu + out.auxk_loss + out.multi_topk_fvu / 8
            did_fire[out.encoded.top_indices.flatten()] = True
            frames += x.shape[0] * x.shape[1]
            frames[did_fire] = 0
            if s == 0:
                rec["s0.pre_acts"] = model.pre_acts(x)
        loss.backward()
        for k_, p in model.named_parameters():
            rec[f"s{s}.grad.{k_}"] = p.grad.clone()
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
        opt.step()
        sched.step()
        rec[f"s{s}.sae_out"] = out.sae_out
        rec[f"s{s}.top_acts"] = out.encoded.top_acts
        rec[f"s{s}.top_indices"] = out.encoded.top_indices
        rec[f"s{s}.fvu"] = out.fvu
        rec[f"s{s}.auxk_loss"] = out.auxk_loss
        rec[f"s{s}.multi_topk_fvu"] = out.multi_topk_fvu
        rec[f"s{s}.mse"] = mse
        rec[f"s{s}.loss"] = loss
        rec[f"s{s}.grad_norm"] = total
        rec[f"s{s}.frames_out"] = frames.clone()
        for k_, v in model.state_dict().items():
            rec[f"s{s}.param.{k_}"] = v.clone()
    rec["meta"] = json.dumps(dict(B=B, T=T, d=d, n=n, k=k, auxk_alpha=auxk_alpha, multi_topk=multi_topk,
                                  n_dead=n_dead, autocast=autocast, steps=steps, lr=lr, clip=clip,
                                  total_steps=total_steps, warmup=warm, dead_thresh=dead_thresh))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npify(rec))
    print("wrote", name)


def l1_case(name, B, T, d, n, *, recon_alpha, autocast, steps, weight_decay=0.0, seed=0):
    """Reference L1 train trajectory: RAdam(eps=1e-5) + cosine, train_sae.py:374-386,421-453."""
    torch.manual_seed(seed)
    cfg = L1AutoEncoderConfig.from_dict({"n_dict_components": n, "recon_alpha": recon_alpha})
    model = L1AutoEncoder(d, cfg)
    g = torch.Generator().manual_seed(seed + 1)
    model.encoder_bias.data = 0.05 * torch.randn(n, generator=g)
    model.decoder.weight.data *= 1.0 + torch.rand(n, generator=g)  # un-normalised columns
    rec = {f"init.{k_}": v.clone() for k_, v in model.state_dict().items()}
    lr, clip, total_steps = 4e-4, 1.0, 10
    opt = RAdam(model.parameters(), eps=1e-5, lr=lr, weight_decay=weight_decay)
    sched = CosineAnnealingLR(opt, T_max=total_steps, eta_min=0)
    for s in range(steps):
        x = synth_x(B, T, d, seed + 10 + s)
        x.view(-1)[torch.randperm(x.numel(), generator=g)[: x.numel() // 50]] = -1.0  # masked targets
        rec[f"s{s}.x"] = x
        rec[f"s{s}.lr"] = torch.tensor(opt.param_groups[0]["lr"], dtype=torch.float64)
        opt.zero_grad()
        from torch.amp import autocast as ac
        import contextlib
        ctx = ac("cpu") if autocast else contextlib.nullcontext()
        with ctx:
            out, mse = model(x, return_mse=True)
            loss = out.reconstruction_loss + out.l1_loss
        rec[f"s{s}.W_normed"] = model.decoder.weight.data.clone()
        loss.backward()
        for k_, p in model.named_parameters():
            rec[f"s{s}.grad.{k_}"] = p.grad.clone()
        total = torch.nn.utils.clip_grad_norm_(model.parameters(), clip)
        opt.step()
        sched.step()
        rec[f"s{s}.sae_out"] = out.sae_out
        rec[f"s{s}.latent"] = out.encoded.latent
        rec[f"s{s}.l1_loss"] = out.l1_loss
        rec[f"s{s}.reconstruction_loss"] = out.reconstruction_loss
        rec[f"s{s}.mse"] = mse
        rec[f"s{s}.grad_norm"] = total
        for k_, v in model.state_dict().items():
            rec[f"s{s}.param.{k_}"] = v.clone()
    rec["meta"] = json.dumps(dict(B=B, T=T, d=d, n=n, recon_alpha=recon_alpha, autocast=autocast,
                                  steps=steps, lr=lr, clip=clip, total_steps=total_steps,
                                  weight_decay=weight_decay))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **npify(rec))
    print("wrote", name)


def misc_case():
    """set_decoder_norm_to_unit_norm / remove_gradient_parallel_to_decoder_directions
    (topkautoencoder.py:153-175) on a random decoder + gradient."""
    torch.manual_seed(3)
    cfg = TopKAutoEncoderConfig.from_dict({"n_dict_components": 96, "k": 4, "normalize_decoder": False})
    model = TopKAutoEncoder(24, cfg)
    rec = {"W_dec_in": model.W_dec.data.clone()}
    model.set_decoder_norm_to_unit_norm()
    rec["W_dec_unit"] = model.W_dec.data.clone()
    model.W_dec.grad = torch.randn_like(model.W_dec)
    rec["grad_in"] = model.W_dec.grad.clone()
    model.remove_gradient_parallel_to_decoder_directions()
    rec["grad_out"] = model.W_dec.grad.clone()
    np.savez_compressed(os.path.join(HERE, "topk_decoder_norm.npz"), **npify(rec))
    print("wrote topk_decoder_norm")


def feature_stats_case():
    """topk_feature_extraction (train_sae.py:70-118) and the L1 per-file abs-max (:175-178) on one file's encoding."""
    from collections import namedtuple

    from src.scripts.train_sae import topk_feature_extraction

    g = torch.Generator().manual_seed(11)
    T, k, n = 50, 8, 96
    acts = torch.randn(1, T, k, generator=g)
    idx = torch.stack([torch.randperm(n, generator=g)[:k] for _ in range(T)]).unsqueeze(0)
    Enc = namedtuple("Enc", "top_acts top_indices")
    Out = namedtuple("Out", "encoded")
    res = topk_feature_extraction(Out(Enc(acts, idx)), n, 1, "cpu")
    latent = torch.randn(1, T, n, generator=g)
    l1 = torch.max(torch.abs(latent).squeeze(), dim=0).values
    np.savez_compressed(os.path.join(HERE, "feature_stats.npz"), acts=acts.numpy(), idx=idx.numpy(), topk_max=res.numpy(),
                        latent=latent.numpy(), l1_max=l1.numpy())
    print("wrote feature_stats")


def search_case():
    """top_activations (utils/activations.py:61-132) through the reference's own
    MemoryMappedActivationDataLoader on a dense and an indexed on-disk set."""
    from src.dataset.activations import MemoryMappedActivationDataLoader

    rng = np.random.default_rng(7)
    n_files, T, F, k, n = 37, 50, 24, 6, 64
    filenames = [f"/synthetic/audio_{i:04d}.flac" for i in range(n_files)]
    num_samples = {f: int(rng.integers(1600, T * 320 + 1))
                   for f in filenames}
    ua = ref_shims.patch_trim(num_samples)
    rec = {"num_samples": np.array([num_samples[f] for f in filenames])}
    with tempfile.TemporaryDirectory() as tmp:
        # dense
        dense = rng.standard_normal((n_files, T, F)).astype(np.float32)
        dense[5, :, 3] = dense[9, :, 3]  # exact tie between two files on feature 3
        os.makedirs(f"{tmp}/dense")
        np.save(f"{tmp}/dense/layer_tensors.npy", dense.reshape(n_files, -1))
        json.dump({"tensor_shape": [T, F], "activation_shape": [T, F], "filenames": filenames},
                  open(f"{tmp}/dense/layer_metadata.json", "w"))
        # indexed
        idx = np.stack([np.stack([rng.permutation(n)[:k] for _ in range(T)]) for _ in range(n_files)]).astype(np.int64)
        vals = np.abs(rng.standard_normal((n_files, T, k))).astype(np.float32)
        os.makedirs(f"{tmp}/indexed")
        np.save(f"{tmp}/indexed/layer_activation_values.npy", vals.reshape(n_files, -1))
        np.save(f"{tmp}/indexed/layer_feature_indices.npy", idx.reshape(n_files, -1))
        json.dump({"tensor_shape": [T, k], "activation_shape": [T, n], "filenames": filenames},
                  open(f"{tmp}/indexed/layer_metadata.json", "w"))
        rec.update(dense=dense, idx=idx, vals=vals)
        queries = []
        for kind, feats in (("dense", [0, 3, 7, 23]), ("indexed", [0, 5, 63])):
            dl = MemoryMappedActivationDataLoader(f"{tmp}/{kind}", "layer", batch_size=8, dl_max_workers=0)
            for f in feats:
                for (mx, mn, ab) in ((None, None, False), (None, None, True), (1.5, 0.2, False), (1.0, -1.0, True)):
                    pq, mpf = ua.top_activations(dl, f, 5, mx, mn, ab, True)
                    q = len(queries)
                    queries.append(dict(kind=kind, feature=f, max_val=mx, min_val=mn, abs=ab, n_files=5))
                    rec[f"q{q}.files"] = np.array([filenames.index(p[0]) for p in pq], dtype=np.int64)
                    rec[f"q{q}.values"] = np.array([p[2] for p in pq], dtype=np.float64)
                    rec[f"q{q}.times"] = np.array([p[3] for p in pq], dtype=np.float64)
                    rec[f"q{q}.lens"] = np.array([len(p[1]) for p in pq], dtype=np.int64)
                    rec[f"q{q}.trace0"] = pq[0][1].numpy() if pq else np.zeros(0, np.float32)
                    rec[f"q{q}.max_per_file"] = np.array(mpf, dtype=np.float64)
    rec["meta"] = json.dumps(dict(queries=queries, filenames=filenames))
    np.savez_compressed(os.path.join(HERE, "search.npz"), **rec)
    print("wrote search")


def clip_case():
    """top_activations_for_audio / manipulate_latent (utils/activations.py:135-300) with a stand-in Whisper: the
    cache's forward() just exposes a seeded [1, 1500, d] activation tensor, the substituted model records what it is
    handed.  Everything after the Whisper forward -- SAE forward, per-frame loops, trims, decode -- is the reference's."""
    import src.utils.activations as ua

    ua.get_mels_from_np_array = lambda device, audio, n_mels: ("mel", len(audio), n_mels)

    class Result:
        def __init__(self, text):
            self.text = text

    class Cache:
        model_name, device = "tiny", "cpu"

        def __init__(self, acts):
            self._acts = acts

        def forward(self, mel):
            self.activations = self._acts.clone()
            return Result("baseline")

    class Subbed:
        def __init__(self):
            self.seen = []

        def forward(self, mel, sub):
            self.seen.append(sub.detach().clone())
            return Result(f"subbed{len(self.seen)}")

    torch.manual_seed(11)
    d, T, top_n = 32, 1500, 6
    audio = np.zeros(139_777, dtype=np.float32)  # 8.736 s -> 436 frames kept of the padded 1500
    acts = torch.randn(1, T, d) * 0.7
    acts[0, 100:110] = acts[0, 90:100]  # repeated frames: exact value ties across frames
    topk = TopKAutoEncoder(d, TopKAutoEncoderConfig.from_dict({"n_dict_components": 256, "k": 8}))
    l1 = L1AutoEncoder(d, L1AutoEncoderConfig.from_dict({"n_dict_components": 40}))
    with torch.no_grad():
        l1.encoder_bias.add_(0.05)
    rec = {"acts": acts, "audio_len": np.int64(len(audio)), "top_n": np.int64(top_n)}
    for k_, v in topk.state_dict().items():
        rec[f"topk.{k_}"] = v.clone()
    for k_, v in l1.state_dict().items():
        rec[f"l1.{k_}"] = v.clone()
    for tag, sae in (("topk", topk), ("l1", l1), ("none", None)):
        feats, traces = ua.top_activations_for_audio(audio, Cache(acts), sae, top_n)
        rec[f"{tag}.top.features"] = np.asarray(feats, dtype=np.int64)
        rec[f"{tag}.top.traces"] = torch.stack([torch.as_tensor(t).float() for t in traces])
        feat = int(feats[1])
        sub = Subbed()
        base, man_text, std_text, pre, man = ua.manipulate_latent(audio, Cache(acts), sae, sub, feat, 3.5)
        rec[f"{tag}.man.feature"] = np.int64(feat)
        rec[f"{tag}.man.texts"] = np.asarray([str(base), man_text, std_text])
        rec[f"{tag}.man.pre"] = pre.float()
        rec[f"{tag}.man.value"] = man.float()
        rec[f"{tag}.man.sub_manipulated"] = sub.seen[0].float()
        rec[f"{tag}.man.sub_standard"] = sub.seen[1].float()
    out = {k_: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k_, v in rec.items()}
    np.savez_compressed(os.path.join(HERE, "clip.npz"), **out)
    print("wrote clip")


if __name__ == "__main__":
    if not ref_shims.reference_available():
        sys.exit("reference tree not found")
    if sys.argv[1:] == ["clip"]:
        clip_case()
        sys.exit(0)
    topk_case("topk_fp32", 4, 6, 32, 256, 8, auxk_alpha=1 / 32, multi_topk=False, n_dead=0, autocast=False, steps=3)
    topk_case("topk_fp32_auxk", 4, 6, 32, 256, 8, auxk_alpha=1 / 32, multi_topk=False, n_dead=40, autocast=False, steps=2)
    topk_case("topk_fp32_auxk_few", 4, 6, 32, 256, 8, auxk_alpha=1 / 32, multi_topk=False, n_dead=5, autocast=False, steps=1)
    topk_case("topk_fp32_multi", 4, 6, 32, 256, 8, auxk_alpha=1 / 32, multi_topk=True, n_dead=40, autocast=False, steps=2)
    topk_case("topk_bf16", 8, 16, 32, 256, 8, auxk_alpha=1 / 32, multi_topk=False, n_dead=0, autocast=True, steps=1,
              margin=0.05)
    topk_case("topk_fp32_b1", 1, 6, 32, 256, 8, auxk_alpha=0.0, multi_topk=False, n_dead=0, autocast=False, steps=1)
    l1_case("l1_fp32", 4, 6, 32, 40, recon_alpha=1e4, autocast=False, steps=7)
    l1_case("l1_fp32_wd", 4, 6, 32, 40, recon_alpha=1.0, autocast=False, steps=7, weight_decay=0.01)
    l1_case("l1_bf16", 4, 6, 32, 40, recon_alpha=1e4, autocast=True, steps=2)
    misc_case()
    feature_stats_case()
    search_case()
    clip_case()
