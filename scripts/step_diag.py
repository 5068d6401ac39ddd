import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from freud_b200 import _lib
w = bench.WORKLOADS["c3"]; dev = torch.device("cuda", 0)
tr = bench.build_trainer(w, "bf16", None, dev)
xs = [bench.synth_batch(w["B"], w["T"], w["d"], 1000 + i).to(dev) for i in range(3)]
for mode in ("plain", "profile"):
    _lib.profile = {} if mode == "profile" else None
    times = []
    for i in range(14):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); o = tr.step(xs[i % 3]); e.record()
        t1 = time.perf_counter(); torch.cuda.synchronize()
        times.append((round(s.elapsed_time(e), 2), round((t1 - t0) * 1e3, 2), tr.tokens_seen))
    print(mode, times)
    _lib.profile = None
print("num dead now:", int((tr.num_frames_since_fired > 1e6).sum()), "max frames", int(tr.num_frames_since_fired.max()))
