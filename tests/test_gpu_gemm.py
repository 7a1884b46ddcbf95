"""regat_gemm through the C ABI: the tcgen05/TMA bf16 kernel and the fp32 SIMT kernel against a plain torch fp32
matmul of the same (bf16-rounded) operands, for every operand-major combination, ragged sizes, the fused
epilogue and split-K."""
import ctypes as C

import numpy as np
import pytest
import torch

from tf_vqa_regat_b200 import _lib

pytestmark = pytest.mark.gpu


def _gemm(dtype, tA, tB, M, N, K, A, B, Cout, c_dtype, epi=None):
    l = _lib.lib()
    lda, ldb, ldc = A.stride(0), B.stride(0), Cout.stride(0)
    _lib.check(l.regat_gemm(dtype, tA, tB, M, N, K, A.data_ptr(), lda, B.data_ptr(), ldb, Cout.data_ptr(), ldc, c_dtype,
                            C.byref(epi) if epi is not None else None, torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()


def _operands(tA, tB, M, N, K, dt, pad=0, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    # leading dimension: a multiple of 8 elements (16 B for bf16), deliberately larger than the row when pad > 0
    mk = lambda r, c: torch.randn(r, (c + 7) // 8 * 8 + pad, generator=g, device="cuda", dtype=torch.float32).to(dt)[:, :c]
    A = mk(K, M) if tA else mk(M, K)
    B = mk(N, K) if tB else mk(K, N)
    Af = (A.float().t() if tA else A.float())
    Bf = (B.float().t() if tB else B.float())
    return A, B, Af @ Bf


SHAPES = [(128, 128, 64), (256, 256, 512), (300, 200, 136), (128, 264, 1000), (72, 3129, 1536), (1024, 768, 4608), (384, 3136, 256)]
# enough 128 x 256 tiles for the 256-wide configuration, which runs on CTA pairs (tcgen05.mma.cta_group::2, 256-row units):
# a last unit whose second CTA has no rows at all (4736 = 18.5 x 256), ragged M and N with a column half fully out of range
# (1100 = 4 x 256 + 76), several units per pair with a long K loop (ring wrap-around, both accumulator stages)
SHAPES_PAIRS = [(4736, 1024, 320), (5000, 1100, 200), (2304, 4096, 1024), (9216, 1024, 2048)]


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_bf16_tcgen05_matches_torch(tA, tB, M, N, K):
    A, B, ref = _operands(tA, tB, M, N, K, torch.bfloat16, pad=8)
    out = torch.full((M, N + 3), float("nan"), device="cuda", dtype=torch.float32)[:, :N]
    _gemm(_lib.BF16, tA, tB, M, N, K, A, B, out, _lib.F32)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-3, err          # fp32 accumulation of exact bf16 products: only summation-order noise


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", SHAPES_PAIRS)
@pytest.mark.parametrize("out_dtype", ["f32", "f32_aligned", "bf16"])
def test_bf16_cta_pair_matches_torch(tA, tB, M, N, K, out_dtype):
    # "f32": odd row pitch -> the generic epilogue kind; "f32_aligned" / "bf16": the lean kinds when N is a multiple of 32
    A, B, ref = _operands(tA, tB, M, N, K, torch.bfloat16, pad=8)
    if out_dtype.startswith("f32"):
        out = torch.full((M, N + (3 if out_dtype == "f32" else 8)), float("nan"), device="cuda", dtype=torch.float32)[:, :N]
        _gemm(_lib.BF16, tA, tB, M, N, K, A, B, out, _lib.F32)
        tol = 2e-3
    else:
        out = torch.full((M, (N + 7) // 8 * 8 + 8), float("nan"), device="cuda", dtype=torch.bfloat16)[:, :N]
        _gemm(_lib.BF16, tA, tB, M, N, K, A, B, out, _lib.BF16)
        tol = 1e-2
    assert torch.isfinite(out.float()).all()
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, err
    # a second call on the same operands: the persistent pipeline state (ring phase, accumulator stage) restarts cleanly
    out2 = torch.zeros_like(out)
    _gemm(_lib.BF16, tA, tB, M, N, K, A, B, out2, _lib.F32 if out_dtype.startswith("f32") else _lib.BF16)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_fp32_simt_matches_torch(tA, tB):
    M, N, K = 200, 301, 777
    A, B, ref = _operands(tA, tB, M, N, K, torch.float32)
    ref = (A.double().t() if tA else A.double()) @ (B.double().t() if tB else B.double())
    out = torch.empty(M, N, device="cuda", dtype=torch.float32)
    _gemm(_lib.F32, tA, tB, M, N, K, A, B, out, _lib.F32)
    assert ((out.double() - ref).abs().max() / ref.abs().max()).item() < 1e-5


@pytest.mark.parametrize("dtype,G,N", [("bf16", 32, 256), ("fp32", 32, 256), ("bf16", 600, 1024)])
def test_fused_epilogue(dtype, G, N):
    # G = 600, N = 1024: 43 x 4 tiles of 128 x 256 -> the CTA-pair configuration, every fused term at once
    dt, code = (torch.bfloat16, _lib.BF16) if dtype == "bf16" else (torch.float32, _lib.F32)
    rows_in, keep = 9, 5
    M, K = rows_in * G, 192
    A, B, acc = _operands(0, 0, M, N, K, dt, seed=3)
    g = torch.Generator(device="cuda").manual_seed(5)
    alpha = torch.tensor([0.7, -1.3], device="cuda").repeat(N // 256)
    bias = torch.randn(N, device="cuda", generator=g)
    addend = torch.randn(G, N, device="cuda", generator=g)
    row_scale = (torch.rand(M, device="cuda", generator=g) > 0.3).float()
    cold = torch.randn(M, N, device="cuda", generator=g).to(dt)
    gate = torch.randn(M, N, device="cuda", generator=g).to(dt)
    out = cold.clone()
    c2 = torch.zeros(G * keep, N, device="cuda", dtype=dt)
    epi = _lib.Epilogue(alpha.data_ptr(), 128, bias.data_ptr(), addend.data_ptr(), N, rows_in, row_scale.data_ptr(), 1, 1,
                        gate.data_ptr(), N, c2.data_ptr(), N, rows_in, keep, 1)
    _gemm(code, 0, 0, M, N, K, A, B, out, code, epi)
    x = acc + row_scale[:, None] * addend.repeat_interleave(rows_in, 0)
    x = x * alpha.repeat_interleave(128)[None, :] + bias
    x = torch.relu(x) + cold.float()
    x = torch.where(gate.float() > 0, x, torch.zeros_like(x))
    tol = 2e-2 if dtype == "bf16" else 1e-5
    assert ((out.float() - x).abs().max() / x.abs().max()).item() < tol
    ref2 = x.view(G, rows_in, N)[:, :keep].reshape(G * keep, N)
    assert ((c2.float() - ref2).abs().max() / x.abs().max()).item() < tol


def test_split_k_weight_gradient_shape():
    # wgrad-like: few output tiles, long K, both operands MN-major
    M, N, K = 256, 384, 9216
    A, B, ref = _operands(1, 0, M, N, K, torch.bfloat16, seed=9)
    out = torch.full((M, N), 7.0, device="cuda", dtype=torch.float32)      # must be overwritten, not accumulated into
    epi = _lib.Epilogue()
    epi.split_k = 6
    _gemm(_lib.BF16, 1, 0, M, N, K, A, B, out, _lib.F32, epi)
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 2e-3


@pytest.mark.parametrize("M,N,K", [(1024, 2048, 9216), (1024, 4096, 5120), (1024, 1024, 9216), (2048, 1024, 4608)])
def test_long_k_weight_gradient_on_cta_pairs(M, N, K):
    """Weight-gradient shapes (both operands MN-major, plain fp32 output, K >= 2048, few output tiles): 256 x 256 units on CTA pairs
    with the split-K factor that fills one round of pairs (1, 2 or 4 here) -- vector red into a destination the call zeroes itself."""
    A, B, ref = _operands(1, 0, M, N, K, torch.bfloat16, seed=11)
    out = torch.full((M, N), 3.0, device="cuda", dtype=torch.float32)      # must be overwritten, not accumulated into
    _gemm(_lib.BF16, 1, 0, M, N, K, A, B, out, _lib.F32)
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 2e-3
    out2 = torch.full((M, N), -1.0, device="cuda", dtype=torch.float32)
    _gemm(_lib.BF16, 1, 0, M, N, K, A, B, out2, _lib.F32)
    assert ((out2 - out).abs().max() / ref.abs().max()).item() < 1e-5        # atomics: order noise only


def test_gemm_argument_errors():
    a = torch.zeros(8, 8, device="cuda")
    l = _lib.lib()
    assert l.regat_gemm(0, 0, 0, 8, 8, 8, None, 8, a.data_ptr(), 8, a.data_ptr(), 8, 0, None, None) == -1
    assert l.regat_gemm(0, 0, 0, 8, 8, 8, a.data_ptr(), 4, a.data_ptr(), 8, a.data_ptr(), 8, 0, None, None) == -2
    assert l.regat_gemm(7, 0, 0, 8, 8, 8, a.data_ptr(), 8, a.data_ptr(), 8, a.data_ptr(), 8, 0, None, None) == -3
    h = torch.zeros(8, 12, device="cuda", dtype=torch.bfloat16)
    assert l.regat_gemm(1, 0, 0, 8, 8, 8, h.data_ptr(), 12, h.data_ptr(), 12, a.data_ptr(), 8, 0, None, None) == -5
    assert "16-byte" in _lib.last_error()
