"""Train an SAE from a stored activation set with the reference's own config schema (configs/train/*.json).

    python examples/train_from_store.py --config my_topk.json [--max-steps 200]

A compact stand-in for the loop of src/scripts/train_sae.py:297-600 for users who do not have the reference checked
out (with it, `freud_b200.compat.install_as_src()` runs the reference's own script on these kernels -- see
INTEGRATION.md).  It uses the same config keys, builds the same model classes from `autoencoder_config`, draws the
same shuffled batches (DeviceResidentActivationLoader), runs `SAETrainer.step` and writes checkpoints in the
reference layout and place (`<run_dir>/checkpoints/step<N>.pth` holding {"model","optimizer","scheduler","step",
"best_val_loss","hparams"}, train_sae.py:232-248,365,492,600) that `init_sae_from_checkpoint` and the reference's GUI
server load; `start_checkpoint` in the config resumes from one (train_sae.py:408-410: model, optimizer, scheduler and
step; the dead-latent counters restart at zero, as upstream).  Validation, TensorBoard and plots are left out.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200.dataset.activations import DeviceResidentActivationLoader  # noqa: E402
from freud_b200.models.config import L1AutoEncoderConfig, TopKAutoEncoderConfig  # noqa: E402
from freud_b200.models.l1autoencoder import L1AutoEncoder  # noqa: E402
from freud_b200.models.topkautoencoder import TopKAutoEncoder  # noqa: E402
from freud_b200.trainer import SAETrainer  # noqa: E402


def train(cfg: dict, max_steps=None, precision="bf16"):
    torch.manual_seed(cfg["seed"])  # train_sae.py:225-229
    layer = cfg["whisper_config"]["layer_name"]
    loader = DeviceResidentActivationLoader(cfg["train_folder"], layer, cfg["batch_size"], 0,
                                            dl_kwargs={"shuffle": True, "drop_last": True}, device=cfg["device"])
    activation_size = loader.activation_shape[-1]
    ae_cfg = cfg["autoencoder_config"]
    if cfg["autoencoder_variant"] == "topk":
        model = TopKAutoEncoder(activation_size, TopKAutoEncoderConfig.from_dict(ae_cfg))
    else:
        model = L1AutoEncoder(activation_size, L1AutoEncoderConfig.from_dict(ae_cfg))
    model = model.to(cfg["device"])
    steps = cfg["steps"]
    trainer = SAETrainer(model, lr=cfg["lr"], steps=steps, clip_thresh=cfg["clip_thresh"],
                         weight_decay=cfg["weight_decay"], optimizer=cfg["optimizer"], scheduler=cfg["scheduler"],
                         scheduler_params=cfg.get("scheduler_params", {}),
                         dead_feature_threshold=ae_cfg.get("dead_feature_threshold"), precision=precision)
    hparams = {k: cfg[k] for k in ("autoencoder_variant", "autoencoder_config", "lr", "weight_decay", "steps",
                                   "clip_thresh", "batch_size", "whisper_config", "train_folder", "val_folder",
                                   "optimizer", "scheduler", "scheduler_params") if k in cfg}
    hparams["activation_size"] = activation_size
    ckpt_dir = os.path.join(cfg["run_dir"], "checkpoints")  # train_sae.py:365
    os.makedirs(ckpt_dir, exist_ok=True)
    step, losses = 0, []
    if cfg.get("start_checkpoint"):  # train_sae.py:408-410
        ck = torch.load(cfg["start_checkpoint"], map_location=cfg["device"], weights_only=False)
        model.load_state_dict(ck["model"])
        trainer.optimizer.load_state_dict(ck["optimizer"])
        trainer.scheduler.load_state_dict(ck["scheduler"])
        trainer.invalidate_weight_copies()
        step = int(ck["step"])
        trainer.step_count = step
    limit = min(steps, max_steps) if max_steps else steps

    def save():
        torch.save({"model": model.state_dict(), "optimizer": trainer.optimizer.state_dict(),
                    "scheduler": trainer.scheduler.state_dict(), "step": step, "best_val_loss": float("inf"),
                    "hparams": hparams}, os.path.join(ckpt_dir, f"step{step}.pth"))

    while step < limit:
        for acts, _ in loader:
            out = trainer.step(acts)
            losses.append(out["loss"])  # device tensors: no host sync inside the loop
            step += 1
            if step % cfg["save_every"] == 0 or step == limit:
                save()
            if step % cfg["log_tb_every"] == 0 or step == limit:
                print(f"step {step}: loss {float(torch.stack(losses).mean()):.6f}", flush=True)
                losses = []
            if step >= limit:
                break
    return model, trainer


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True)
    ap.add_argument("--max-steps", type=int, default=None)
    ap.add_argument("--precision", default="bf16", choices=("bf16", "fp32"))
    a = ap.parse_args()
    with open(a.config) as f:
        train(json.load(f), a.max_steps, a.precision)
