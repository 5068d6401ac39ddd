"""Time the fused encoder variants (FREUD_ENC_VARIANT is read once per process, so run one process per variant)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import ops
from freud_b200._lib import BF16
torch.manual_seed(0)
def run(N, d, n):
    x = torch.randn(1, N, d, device="cuda")
    W = torch.randn(n, d, device="cuda") / d ** 0.5
    b_enc = 0.1 * torch.randn(n, device="cuda"); b_dec = 0.1 * torch.randn(d, device="cuda")
    xc, _, _ = ops.topk_prep_x(x, b_dec, BF16); w, _ = ops.split_operand(W, BF16)
    vals, idx = ops.topk_encode(xc, None, w, None, b_enc, BF16)
    pre = torch.relu(xc.float() @ w.float().T + b_enc)
    rv, ri = pre.topk(32, dim=-1)
    same = (torch.sort(idx.long(), -1).values == torch.sort(ri, -1).values).all(-1).float().mean().item()
    for _ in range(3): ops.topk_encode(xc, None, w, None, b_enc, BF16)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): ops.topk_encode(xc, None, w, None, b_enc, BF16)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f"variant={os.environ.get('FREUD_ENC_VARIANT','0')} N={N} d={d} n={n}: match={same:.4f} {ms:.3f} ms {2.0*N*d*n/ms/1e9:.0f} TF/s", flush=True)
if os.environ.get("ONLY_C3"):
    run(48000, 768, 24576)
    sys.exit(0)
if os.environ.get("ONLY_SHAPE"):
    run(*[int(v) for v in os.environ["ONLY_SHAPE"].split(",")])
    sys.exit(0)
run(1000, 64, 512)
run(75000, 384, 6144)
run(48000, 768, 24576)
run(24000, 1280, 81920 // 8)   # C4 per-GPU shard at 8 GPUs
run(24000, 1280, 81920)        # C4 on one GPU
