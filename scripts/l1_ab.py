"""A/B of the generic GEMM's TMA input / output boxes on the L1 step (C1) and the AuxK-live step (C3), warm GPU."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from freud_b200.models.config import L1AutoEncoderConfig  # noqa: E402
from freud_b200.models.l1autoencoder import L1AutoEncoder  # noqa: E402
from freud_b200.trainer import SAETrainer  # noqa: E402

dev = torch.device("cuda:0")
x = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
for _ in range(200):
    y = x @ x
torch.cuda.synchronize()


def timed(fn, iters=30, warm=5):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(iters):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


torch.manual_seed(0)
m = L1AutoEncoder(384, L1AutoEncoderConfig.from_dict({"n_dict_components": 200, "recon_alpha": 1e4})).to(dev)
tr = SAETrainer(m, lr=4e-4, steps=100000, clip_thresh=1.0, optimizer="radam", scheduler="cosine", precision="bf16",
                materialize_outputs=False)
xs = [bench.synth_batch(100, 1500, 384, 70 + i).to(dev) for i in range(3)]
w = bench.WORKLOADS["c3"]
tra = bench.build_trainer(w, "bf16", None, dev)
xa = [bench.synth_batch(w["B"], w["T"], w["d"], 50 + i).to(dev) for i in range(3)]
dead = torch.randperm(w["n"], device=dev)[: w["n"] // 10]
tra.tokens_seen = 10 ** 12


def aux_step(i):
    tra.num_frames_since_fired[dead] = 10 ** 9
    tra.step(xa[i % 3])


for rep in range(2):
    for name, env in (("tma in+out", {}), ("tma out only", {"FREUD_NO_TMA_IN": "1"}), ("direct", {"FREUD_NO_TMA_OUT": "1"})):
        for k in ("FREUD_NO_TMA_IN", "FREUD_NO_TMA_OUT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        print(f"{name:14s} L1 {timed(lambda i: tr.step(xs[i % 3])):.4f} ms   AuxK-live C3 {timed(aux_step, 10, 3):.4f} ms", flush=True)
