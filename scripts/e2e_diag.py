import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c3"]
dev = torch.device("cuda", 0)
tr = bench.build_trainer(w, "bf16", None, dev)
B, T, d = w["B"], w["T"], w["d"]
host = [bench.synth_batch(B, T, d, 1000 + i, pin=True) for i in range(3)]
print("pinned:", [h.is_pinned() for h in host])
dev_x = [h.to(dev) for h in host]
stage = torch.empty_like(dev_x[0])
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("copy only          ms:", t(lambda i=0: stage.copy_(host[i % 3], non_blocking=True)))
print("step async         ms:", t(lambda i=0: tr.step(dev_x[i % 3])))
print("step + item sync   ms:", t(lambda i=0: tr.step(dev_x[i % 3])["loss"].item()))
def both(i=0):
    stage.copy_(host[i % 3], non_blocking=True)
    return tr.step(stage)["loss"].item()
print("copy+step+item ser ms:", t(both))
# CPU-side cost of enqueueing one step
torch.cuda.synchronize(); t0 = time.perf_counter(); tr.step(dev_x[0]); t1 = time.perf_counter(); torch.cuda.synchronize()
print("cpu enqueue one step ms:", (t1 - t0) * 1e3)
