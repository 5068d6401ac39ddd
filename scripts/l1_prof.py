"""Per-C-ABI-call CUDA-event breakdown of the L1 SAE step (C1: d=384, n=200, 150 000 tokens)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from freud_b200 import _lib  # noqa: E402
from freud_b200.models.config import L1AutoEncoderConfig  # noqa: E402
from freud_b200.models.l1autoencoder import L1AutoEncoder  # noqa: E402
from freud_b200.trainer import SAETrainer  # noqa: E402

dev = torch.device("cuda:0")
n_feat = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for prec in ("bf16", "fp32"):
    torch.manual_seed(0)
    m = L1AutoEncoder(384, L1AutoEncoderConfig.from_dict({"n_dict_components": n_feat, "recon_alpha": 1e4})).to(dev)
    tr = SAETrainer(m, lr=4e-4, steps=100000, clip_thresh=1.0, optimizer="radam", scheduler="cosine", precision=prec,
                    materialize_outputs=False)
    xs = [bench.synth_batch(100, 1500, 384, 70 + i).to(dev) for i in range(3)]
    for i in range(4):
        tr.step(xs[i % 3])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(8):
        tr.step(xs[i % 3])
    e1.record()
    torch.cuda.synchronize()
    print(prec, "ms/step", e0.elapsed_time(e1) / 8)
    _lib.profile = {}
    for i in range(4):
        tr.step(xs[i % 3])
    torch.cuda.synchronize()
    summ = _lib.profile_summary()
    _lib.profile = None
    for k, v in sorted(summ.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:28s} calls {v[0]:3d}  ms/step {v[1] / 4:.4f}")
    del tr, xs, m
    torch.cuda.empty_cache()
