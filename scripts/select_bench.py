"""Time the masked row top-k kernel on the AuxK shape of C3 (48000 x 2458 dead candidates, k_aux = 384)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import ops
rows, n, k = 48000, 2458, 384
ld = (n + 7) // 8 * 8
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.relu(torch.randn((rows, ld), device="cuda", generator=g))
for _ in range(3):
    out = ops.row_topk_mask(x, k, ld, n=n, nonneg=True)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    out = ops.row_topk_mask(x, k, ld, n=n, nonneg=True)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
print(f"row_topk_mask {rows}x{n} k={k}: {ms:.3f} ms, {(rows*ld*4 + rows*ld*2)/ms/1e6:.0f} GB/s; kept per row {(out[:, :n] > 0).sum(1).float().mean().item():.1f}")
