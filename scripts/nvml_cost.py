import time, pynvml, torch
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for name, fn in (("clock", lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                 ("reasons", lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)),
                 ("power", lambda: pynvml.nvmlDeviceGetPowerUsage(h))):
    for busy in (False, True):
        if busy:
            for _ in range(50): y = x @ x
        t0 = time.perf_counter()
        for _ in range(10): fn()
        dt = (time.perf_counter() - t0) / 10
        torch.cuda.synchronize()
        print(f"{name:8s} busy={busy}: {dt*1e3:.3f} ms per call")
# effect on launch-bound loop
import threading
stop = threading.Event()
def poll(period):
    while not stop.is_set():
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        time.sleep(period)
a = torch.randn(1 << 20, device="cuda")
def loop():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3000): a.add_(1.0)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
print("launch loop no poll:", loop())
for period in (0.02, 0.1):
    stop.clear(); t = threading.Thread(target=poll, args=(period,)); t.start(); print(f"launch loop poll {period}:", loop()); stop.set(); t.join()
