"""Where do the fused encoder's epilogue warps spend their cycles?  (instrumented build, FREUD_ENC_STATS=1)

    FREUD_ENC_STATS=1 python scripts/enc_stats.py

Per shape: the un-instrumented launch time, then for the instrumented build the share of a scanner warp's lifetime
spent waiting for an accumulator tile (= MMA-bound) or for a free candidate buffer (= compactor-bound), the compactor's
busy share, and hand-overs / compactions per tile."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import _lib, ops  # noqa: E402
from freud_b200._lib import BF16  # noqa: E402

torch.manual_seed(0)


def run(N, d, n):
    x = torch.randn(1, N, d, device="cuda")
    W = torch.randn(n, d, device="cuda") / d ** 0.5
    b_enc = 0.1 * torch.randn(n, device="cuda")
    b_dec = 0.1 * torch.randn(d, device="cuda")
    xc, _, _ = ops.topk_prep_x(x, b_dec, BF16)
    w, _ = ops.split_operand(W, BF16)
    for _ in range(3):
        ops.topk_encode(xc, None, w, None, b_enc, BF16)
    torch.cuda.synchronize()
    out = (C.c_ulonglong * 9)()
    _lib.lib().freud_topk_encode_stats(out, 1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        ops.topk_encode(xc, None, w, None, b_enc, BF16)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 5
    _lib.lib().freud_topk_encode_stats(out, 1)
    v = [int(t) for t in out]
    if v[7] == 0:
        print(f"N={N} d={d} n={n}: {ms:.3f} ms (no statistics: run with FREUD_ENC_STATS=1)")
        return
    scan, wt, wc, ho, comp, cw, nc, warps, tiles = v
    print(f"N={N} d={d} n={n}: {ms:.3f} ms/launch ({2.0 * N * d * n / ms / 1e9:.0f} TF/s); per scanner warp-tile "
          f"{scan / tiles:.0f} cycles, waiting for the MMA {100 * wt / scan:.1f} %, for a candidate buffer "
          f"{100 * wc / scan:.1f} %, scanning {100 * (scan - wt - wc) / scan:.1f} %; hand-overs / tile {ho / tiles:.2f}; "
          f"compactor busy {100 * (comp - cw) / comp:.1f} % of its lifetime, {(comp - cw) / max(nc, 1):.0f} cycles per "
          f"compaction, {nc / tiles:.2f} compactions / tile", flush=True)


run(75000, 384, 6144)
run(48000, 768, 24576)
run(24000, 1280, 81920)
