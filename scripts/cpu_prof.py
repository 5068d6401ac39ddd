import os, sys, time, torch, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS["c3"]
dev = torch.device("cuda", 0)
tr = bench.build_trainer(w, "bf16", None, dev)
x = bench.synth_batch(w["B"], w["T"], w["d"], 1000).to(dev)
for _ in range(3):
    tr.step(x)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    tr.step(x)["loss"].item()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
