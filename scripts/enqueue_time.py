"""Host time to ENQUEUE one training step vs the device time to run it (is the loop launch-bound?)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
for wl in ("c3", "c2"):
    w = bench.WORKLOADS[wl]
    dev = torch.device("cuda", 0)
    s = torch.cuda.Stream(dev); torch.cuda.set_stream(s)
    tr = bench.build_trainer(w, "bf16", None, dev)
    xs = [bench.synth_batch(w["B"], w["T"], w["d"], 1000 + i).to(dev) for i in range(3)]
    for i in range(5):
        tr.step(xs[i % 3])
    torch.cuda.synchronize()
    # enqueue-only: GPU kept busy by a long kernel first so launches never wait on the device
    big = torch.randn(8192, 8192, device=dev)
    for _ in range(3):
        big = big @ big * 1e-4
    t0 = time.perf_counter()
    for i in range(20):
        tr.step(xs[i % 3])
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        tr.step(xs[i % 3])
    e1.record(); torch.cuda.synchronize()
    print(f"{wl}: host enqueue {1e3*(t1-t0)/20:.3f} ms/step; device {e0.elapsed_time(e1)/20:.3f} ms/step", flush=True)
    del tr, xs
    torch.cuda.empty_cache()
