import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from freud_b200 import _lib
wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
w = bench.WORKLOADS[wl]; dev = torch.device("cuda", 0)
tr = bench.build_trainer(w, "bf16", None, dev)
xs = [bench.synth_batch(w["B"], w["T"], w["d"], 50 + i).to(dev) for i in range(3)]
n = w["n"]; dead = torch.randperm(n, device=dev)[: n // 10]; tr.tokens_seen = 10 ** 12
def step(i):
    tr.num_frames_since_fired[dead] = 10 ** 9
    tr.step(xs[i % 3])
for i in range(3): step(i)
torch.cuda.synchronize()
_lib.profile = {}
for i in range(5): step(i)
torch.cuda.synchronize()
tot = 0
for k, (c, ms) in sorted(_lib.profile_summary().items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} calls/step {c/5:4.1f}  {ms/5:7.3f} ms/step"); tot += ms / 5
print("sum", tot)
