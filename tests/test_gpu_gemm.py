"""tcgen05 GEMM entry points against a plain fp32 torch product of the same bf16-rounded operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())


@pytest.mark.parametrize("K,M,N,lda,ldb", [(1000, 128, 256, 128, 256), (4096, 200, 384, 200, 384),
                                           (21000, 384, 200, 384, 200), (3000, 2458, 768, 2464, 768),
                                           (777, 70, 64, 72, 64)])
def test_gemm_tn_splitk_reads_mn_major_operands(K, M, N, lda, ldb):
    """out = A^T B with A [K, lda] and B [K, ldb] row-major (the weight-gradient shape, K = tokens): MN-major UMMA
    descriptors over 64 x 64 TMA boxes, ragged M / N / K, padded pitches, split-K partial sums."""
    from freud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn((K, lda), device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn((K, ldb), device="cuda", generator=g).to(torch.bfloat16)
    if lda != M:
        a[:, M:] = 7.0  # padding columns must not leak into the product
    out = ops.gemm_tn_splitk(a, b, M, N)
    ref = a[:, :M].float().T @ b[:, :N].float()
    assert out.shape == (M, N)
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("M,K,N,relu", [(300, 256, 768, False), (1000, 2458, 768, True), (5000, 200, 384, False)])
def test_gemm_nn_reads_b_row_major(M, K, N, relu):
    """out = act(A B + bias) with B stored [K, N] row-major (MN-major B operand), A K-major with a padded pitch."""
    from freud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(1)
    lda = (K + 7) // 8 * 8
    a = torch.zeros((M, lda), device="cuda", dtype=torch.bfloat16)
    a[:, :K] = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn((K, N), device="cuda", generator=g).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    out = ops.gemm_nn(a, b, bias, relu, K=K)
    ref = a[:, :K].float() @ b.float() + bias
    if relu:
        ref = torch.relu(ref)
    assert _rel(out, ref) < 2e-5


@pytest.mark.parametrize("rows,n,k,pad", [(300, 2458, 384, True), (64, 200, 100, True), (50, 4096, 640, False),
                                          (33, 70, 70, True), (20, 5000, 100, False), (129, 1000, 1, True)])
def test_row_topk_mask_is_exact_with_ties(rows, n, k, pad):
    """Dense masked top-k (AuxK selection on the dead subset, topkautoencoder.py:118-121): warp-per-row radix select
    (register-resident rows) and the CTA-per-row fallback for wide rows; values quantised so that ties at the k-th
    value are common -- the lower column index must win, as in the oracle's stable sort."""
    from freud_b200 import ops
    from oracle import sae as osae

    g = torch.Generator().manual_seed(rows + n)
    lat = torch.relu(torch.randn(rows, n, generator=g))
    lat = torch.round(lat * 8) / 8
    ld_in = (n + 7) // 8 * 8 if pad else n
    buf = torch.full((rows, ld_in), 123.0)
    buf[:, :n] = lat
    ld_out = (n + 7) // 8 * 8
    out = ops.row_topk_mask(buf.cuda(), k, ld_out, n=n, nonneg=True).float().cpu()
    rv, ri = osae.select_topk(lat, k)
    want = torch.zeros(rows, ld_out)
    want.scatter_(1, ri, rv)
    assert torch.equal(out, want.to(torch.bfloat16).float())


@pytest.mark.parametrize("M,N,K", [(300, 2458, 768), (1000, 200, 384), (5000, 96, 64)])
def test_gemm_nt_mask_epilogue(M, N, K):
    from freud_b200 import ops

    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn((M, K), device="cuda", generator=g).to(torch.bfloat16)
    b = torch.randn((N, K), device="cuda", generator=g).to(torch.bfloat16)
    ld = (N + 7) // 8 * 8
    act = torch.zeros((M, ld), device="cuda", dtype=torch.bfloat16)
    act[:, :N] = torch.relu(torch.randn((M, N), device="cuda", generator=g) - 0.5).to(torch.bfloat16)
    out = ops.gemm_nt_mask(a, b, act)
    ref = (a.float() @ b.float().T) * (act[:, :N] > 0)
    assert _rel(out[:, :N].float(), ref.to(torch.bfloat16).float()) < 1e-2  # bf16 output rounding
    assert float(out[:, N:].float().abs().max()) == 0.0 if ld > N else True
    s = ops.col_sum_bf16(out, N)
    assert _rel(s, out[:, :N].float().sum(0)) < 1e-5


@pytest.mark.parametrize("M,N,K,ldo,relu", [(20000, 200, 384, 200, True),     # persistent CTAs over >1 row block, TMA boxes
                                            (1000, 2458, 768, 2464, True),    # ragged N inside a padded pitch (dense AuxK)
                                            (300, 201, 64, 201, False),       # pitch not a multiple of 16 bytes: direct stores
                                            (129, 64, 128, 64, False), (5, 96, 32, 96, True)])
@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_gemm_nt_store_epilogue(M, N, K, ldo, relu, precision):
    """out = act(A B^T + bias), fp32 output through the TMA-box epilogue (or direct stores when the pitch rules it
    out): ragged M / N, padded pitch whose padding columns must stay untouched, one CTA walking several row blocks."""
    from freud_b200 import ops
    from freud_b200._lib import BF16, FP32

    prec = BF16 if precision == "bf16" else FP32
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = torch.randn((M, K), device="cuda", generator=g)
    b = torch.randn((N, K), device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    a_ops, b_ops = ops.split_operand(a, prec), ops.split_operand(b, prec)
    out = torch.full((M, ldo), 77.0, device="cuda")
    ops.gemm_nt(a_ops[0], a_ops[1], b_ops[0], b_ops[1], bias, relu, prec, out=out)
    if precision == "bf16":
        ref = a.to(torch.bfloat16).double() @ b.to(torch.bfloat16).double().T + bias.double()
        tol = 2e-5
    else:
        ref = a.double() @ b.double().T + bias.double()
        tol = 1e-5
    if relu:
        ref = torch.relu(ref)
    assert _rel(out[:, :N], ref) < tol
    if ldo > N:
        # padding columns stay untouched, except that the TMA path may zero the (up to 3) padding columns sharing a
        # 16-byte granule with the last valid column
        g4 = (N + 3) // 4 * 4
        assert bool((out[:, g4:] == 77.0).all()), "padding columns of the output pitch were written"
        assert bool(((out[:, N:g4] == 77.0) | (out[:, N:g4] == 0.0)).all())
