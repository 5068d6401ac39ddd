import torch, time
x = torch.empty(147456000 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device="cuda")
for _ in range(3):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 10
print(f"H2D pinned 147 MB: {dt*1e3:.2f} ms -> {x.numel()*4/dt/1e9:.1f} GB/s")
y = torch.empty(147456000 // 4, dtype=torch.float32)
t0 = time.perf_counter(); d.copy_(y); torch.cuda.synchronize(); print("pageable", (time.perf_counter()-t0)*1e3, "ms")
