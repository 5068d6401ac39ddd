"""Measured ceilings for the gather kernels (DESIGN.md section 5 grades them against these, not against a quoted figure):
  * HBM stream  : read a buffer far larger than L2 with 16-byte loads
  * L2 -> SM    : every warp gathers random 1536-byte rows (a bf16 d=768 decoder row) out of a 38 MB table that
                  stays resident in the 126 MB L2 -- the access pattern of freud_topk_decode / dacts / sparse_grads.
Uses torch ops only (index_select / sum), so the numbers are an independent reference, not this library timing itself.
Prints one JSON document."""
import json
import torch

dev = torch.device("cuda", 0)
out = {}


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


big = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev).normal_()  # 2 GiB
ms = timed(lambda: big.sum())
out["hbm_stream_read_GBps"] = big.numel() * 2 / ms / 1e6
dst = torch.empty_like(big)
ms = timed(lambda: dst.copy_(big))
out["hbm_copy_GBps"] = 2 * big.numel() * 2 / ms / 1e6
del dst, big
table = torch.randn((24576, 768), device=dev).to(torch.bfloat16)  # 37.7 MB: W_dec of C3 as gathered
g = torch.Generator(device=dev).manual_seed(0)
idx = torch.randint(0, 24576, (48000 * 32,), device=dev, generator=g)
buf = torch.empty((idx.numel(), 768), dtype=torch.bfloat16, device=dev)  # 2.36 GB written: an upper bound on time
ms = timed(lambda: torch.index_select(table, 0, idx, out=buf))
out["l2_gather_plus_hbm_write_GBps"] = idx.numel() * 768 * 2 / ms / 1e6
# gather + reduce without the big write: embedding_bag sums 32 gathered rows per token (the decode pattern)
off = torch.arange(0, idx.numel(), 32, device=dev)
ms = timed(lambda: torch.nn.functional.embedding_bag(idx, table, off, mode="sum"))
out["l2_gather_reduce_GBps_embedding_bag"] = idx.numel() * 768 * 2 / ms / 1e6
out["note"] = ("gathered bytes / time; torch kernels (index_select, embedding_bag) as an independent reference for the L2->SM "
               "gather rate that freud_topk_decode / dacts / sparse_grads are graded against")
print(json.dumps(out, indent=1))
