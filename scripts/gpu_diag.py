"""Bring-up diagnostics for the CUDA kernels (run on the GPU box): prints errors instead of asserting."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from freud_b200 import ops  # noqa: E402
from freud_b200._lib import BF16, FP32  # noqa: E402

torch.manual_seed(0)
dev = "cuda"


def relerr(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def test_gemm(M, N, K, precision):
    a = torch.randn(M, K, device=dev)
    b = torch.randn(N, K, device=dev) / K ** 0.5
    bias = torch.randn(N, device=dev)
    a_ops = ops.split_operand(a, precision)
    b_ops = ops.split_operand(b, precision)
    out = ops.gemm_nt(a_ops[0], a_ops[1], b_ops[0], b_ops[1], bias, True, precision)
    torch.cuda.synchronize()
    if precision == BF16:
        ref = torch.relu(a.bfloat16().double() @ b.bfloat16().double().T + bias.double())
    else:
        ref = torch.relu(a.double() @ b.double().T + bias.double())
    print(f"gemm M={M} N={N} K={K} prec={precision}: relerr={relerr(out, ref):.3e}", flush=True)


def test_encode(N, d, n, precision, timing=False):
    x = torch.randn(1, N, d, device=dev)
    W = torch.randn(n, d, device=dev) / d ** 0.5
    b_enc = 0.1 * torch.randn(n, device=dev)
    b_dec = 0.1 * torch.randn(d, device=dev)
    xc_hi, xc_lo, tv = ops.topk_prep_x(x, b_dec, precision)
    w_hi, w_lo = ops.split_operand(W, precision)
    vals, idx = ops.topk_encode(xc_hi, xc_lo, w_hi, w_lo, b_enc, precision)
    torch.cuda.synchronize()
    if precision == BF16:
        pre = torch.relu(xc_hi.double() @ w_hi.double().T + b_enc.double())
    else:
        pre = torch.relu((x[0] - b_dec).double() @ W.double().T + b_enc.double())
    rv, ri = pre.topk(32, dim=-1)
    same = (torch.sort(idx.long(), -1).values == torch.sort(ri, -1).values).all(-1).float().mean().item()
    sv = torch.sort(vals, -1, descending=True).values
    print(f"encode N={N} d={d} n={n} prec={precision}: set-match={same:.4f} val relerr={relerr(sv, rv):.3e}",
          flush=True)
    if timing:
        ms = timeit(lambda: ops.topk_encode(xc_hi, xc_lo, w_hi, w_lo, b_enc, precision))
        flops = 2.0 * N * d * n * (3 if precision == FP32 else 1)
        print(f"   encode time {ms:.3f} ms  -> {flops / ms / 1e9:.1f} TFLOP/s (executed)", flush=True)
        ms2 = timeit(lambda: ops.gemm_nt(xc_hi, xc_lo, w_hi, w_lo, b_enc, True, precision))
        print(f"   gemm_nt(store) time {ms2:.3f} ms -> {flops / ms2 / 1e9:.1f} TFLOP/s", flush=True)
        a16 = xc_hi if precision == BF16 else x[0]
        w16 = w_hi if precision == BF16 else W
        ms3 = timeit(lambda: torch.relu(a16 @ w16.T))
        print(f"   torch matmul+relu time {ms3:.3f} ms", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    for prec in (BF16, FP32):
        for (M, N, K) in [(128, 256, 64), (128, 256, 128), (256, 512, 384), (100, 200, 200), (1000, 6144, 384)]:
            try:
                test_gemm(M, N, K, prec)
            except Exception as ex:  # noqa: BLE001
                print("GEMM FAIL", M, N, K, prec, ex, flush=True)
                sys.exit(1)
    for prec in (BF16, FP32):
        for (N, d, n) in [(128, 64, 256), (200, 384, 6144), (1000, 768, 24576)]:
            try:
                test_encode(N, d, n, prec)
            except Exception as ex:  # noqa: BLE001
                print("ENCODE FAIL", N, d, n, prec, ex, flush=True)
                sys.exit(1)
    test_encode(75000, 384, 6144, BF16, timing=True)
    test_encode(48000, 768, 24576, BF16, timing=True)
    test_encode(75000, 384, 6144, FP32, timing=True)
