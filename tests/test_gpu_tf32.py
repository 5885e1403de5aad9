"""The TMA-fed tf32 tcgen05 GEMM engine of the training pipeline (csrc/gemm_tf32.cu), against torch fp64 on the
same operands.  tf32 keeps 10 mantissa bits of each operand (truncation): a K-term dot product of O(1) values is off
by <= ~K * 2^-10 in the worst case and ~sqrt(K) * 2^-11 typically; the tests allow 8e-3 * sqrt(K) * scale (max over ~1e6 outputs)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lib():
    from cikm2020_dmt_b200 import abi
    return abi, abi.load()


def _tol(K, scale):
    return 8e-3 * math.sqrt(K) * scale


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (1000, 192, 64), (4099, 256, 64), (777, 64, 256), (130, 128, 64),
                                   (3000, 80, 80), (2050, 240, 80), (1500, 160, 80), (900, 80, 320), (50000, 64, 192)])
def test_tf32_rows_plain(M, N, K):
    abi, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bt = torch.randn(N, K, device="cuda", generator=g)
    C = torch.full((M, N), 7.0, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    abi.check(lib.dmt_selftest_tf32_rows(A.data_ptr(), K, Bt.data_ptr(), K, M, N, K, C.data_ptr(), N, None, None, 0,
                                         None, 0, 1.0, 0, 0, st))
    torch.cuda.synchronize()
    want = A.double() @ Bt.double().t()
    err = (C.double() - want).abs().max().item()
    assert err <= _tol(K, 1.0), err


def test_tf32_rows_epilogue_and_strides():
    """bias, addend, alpha, ReLU, mask, accumulate; operands / outputs that are column slices of wider matrices."""
    abi, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 2600, 64, 64
    wideA = torch.randn(M, 192, device="cuda", generator=g)
    A = wideA[:, 64:128]
    Bt = torch.randn(N, K, device="cuda", generator=g)
    bias = torch.randn(N, device="cuda", generator=g)
    addend = torch.randn(M, N, device="cuda", generator=g)
    mask = torch.randn(M, 256, device="cuda", generator=g)
    wideC = torch.randn(M, 128, device="cuda", generator=g)
    old = wideC.clone()
    Cv = wideC[:, 64:]
    st = torch.cuda.current_stream().cuda_stream
    abi.check(lib.dmt_selftest_tf32_rows(A.data_ptr(), 192, Bt.data_ptr(), K, M, N, K, Cv.data_ptr(), 128,
                                         bias.data_ptr(), addend.data_ptr(), N, mask.data_ptr(), 256, 0.5, 1, 1, st))
    torch.cuda.synchronize()
    v = (A.double() @ Bt.double().t() + addend.double()) * 0.5 + bias.double()
    v = torch.relu(v) * (mask[:, :N] > 0).double() + old[:, 64:].double()
    assert (wideC[:, 64:].double() - v).abs().max().item() <= _tol(K, 1.0)
    assert torch.equal(wideC[:, :64], old[:, :64])               # the neighbouring columns are untouched


@pytest.mark.parametrize("T,MA,NB,transposed", [(64, 128, 64, 0), (5000, 256, 64, 1), (12345, 192, 64, 1),
                                                (3001, 128, 64, 0), (7000, 320, 80, 1), (4100, 240, 80, 1),
                                                (333, 160, 80, 0), (0, 128, 64, 0)])
def test_tf32_wgrad(T, MA, NB, transposed):
    abi, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(T + MA)
    P = torch.randn(max(T, 1), MA, device="cuda", generator=g)[:T]
    Q = torch.randn(max(T, 1), NB, device="cuda", generator=g)[:T]
    shape = (NB, MA) if transposed else (MA, NB)
    C = torch.randn(shape, device="cuda", generator=g)
    old = C.clone()
    cs = torch.randn(MA, device="cuda", generator=g)            # bias gradient riding along: column sums of P
    ws = torch.empty(lib.dmt_selftest_tf32_wgrad_bytes(T, MA, NB), dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for acc in (0, 1):
        abi.check(lib.dmt_selftest_tf32_wgrad(P.data_ptr() if T else ws.data_ptr(), MA, Q.data_ptr() if T else ws.data_ptr(),
                                              NB, T, MA, NB, C.data_ptr(), shape[1], transposed, acc, cs.data_ptr(),
                                              ws.data_ptr(), st))
    torch.cuda.synchronize()
    D = P.double().t() @ Q.double()
    want = 2 * (D.t() if transposed else D)                    # written once, accumulated once
    err = (C.double() - want).abs().max().item()
    assert err <= 2 * _tol(max(T, 1), 1.0), err
    err_cs = (cs.double() - 2 * P.double().sum(0)).abs().max().item()
    assert err_cs <= 2 * _tol(max(T, 1), 1.0), err_cs
    # deterministic: the same call twice gives the same bits
    C2 = old.clone()
    for acc in (0, 1):
        abi.check(lib.dmt_selftest_tf32_wgrad(P.data_ptr() if T else ws.data_ptr(), MA, Q.data_ptr() if T else ws.data_ptr(),
                                              NB, T, MA, NB, C2.data_ptr(), shape[1], transposed, acc, None,
                                              ws.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(C, C2)


@pytest.mark.parametrize("T,W", [(1, 64), (5000, 64), (33333, 192), (2000, 256), (777, 80), (0, 128)])
def test_tf32_colsum(T, W):
    abi, lib = _lib()
    X = torch.randn(max(T, 1), W + 64, device="cuda")[:T]
    out = torch.ones(W, device="cuda")
    scratch = torch.empty(296 * W, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    abi.check(lib.dmt_selftest_tf32_colsum(X.data_ptr(), W + 64, T, W, out.data_ptr(), 1, scratch.data_ptr(), st))
    torch.cuda.synchronize()
    want = 1.0 + X[:, :W].double().sum(0)
    assert (out.double() - want).abs().max().item() <= 1e-4 * max(1.0, math.sqrt(max(T, 1)))


@pytest.mark.parametrize("a_mn,b_mn", [(0, 1), (0, 0), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(8192, 512, 1087), (1000, 128, 256), (1087, 512, 3000), (300, 472, 512), (130, 4, 40)])
def test_tf32_gemm_all_operand_orders(a_mn, b_mn, M, N, K):
    """The general tiled kernel through every operand storage order (K-major / MN-major, i.e. SWIZZLE_128B /
    SWIZZLE_128B_ATOM_32B boxes), ragged M / N / K, bias + ReLU + mask + accumulate, misaligned output columns."""
    abi, lib = _lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K + a_mn * 2 + b_mn)
    pad = lambda n: (n + 3) // 4 * 4
    A = torch.randn(K, pad(M), device="cuda", generator=g)[:, :M] if a_mn else torch.randn(M, pad(K), device="cuda", generator=g)[:, :K]
    B = torch.randn(K, pad(N), device="cuda", generator=g)[:, :N] if b_mn else torch.randn(N, pad(K), device="cuda", generator=g)[:, :K]
    bias = torch.randn(N, device="cuda", generator=g)
    mask = torch.randn(M, pad(N), device="cuda", generator=g)
    wide = torch.randn(M, pad(N) + 8, device="cuda", generator=g)
    old = wide.clone()
    Cv = wide[:, 3:3 + N]                                   # misaligned first column: scalar store path
    st = torch.cuda.current_stream().cuda_stream
    abi.check(lib.dmt_selftest_tf32_gemm(A.data_ptr(), A.stride(0), a_mn, B.data_ptr(), B.stride(0), b_mn, M, N, K,
                                         Cv.data_ptr(), wide.stride(0), bias.data_ptr(), mask.data_ptr(), mask.stride(0),
                                         1, 1, st))
    torch.cuda.synchronize()
    Am = A.double().t() if a_mn else A.double()
    Bm = B.double() if b_mn else B.double().t()
    want = torch.relu(Am @ Bm + bias.double()) * (mask[:, :N] > 0).double() + old[:, 3:3 + N].double()
    assert (wide[:, 3:3 + N].double() - want).abs().max().item() <= _tol(K, 1.0)
    assert torch.equal(wide[:, :3], old[:, :3]) and torch.equal(wide[:, 3 + N:], old[:, 3 + N:])
    # aligned output: vector store path
    C2 = torch.zeros(M, pad(N), device="cuda")
    abi.check(lib.dmt_selftest_tf32_gemm(A.data_ptr(), A.stride(0), a_mn, B.data_ptr(), B.stride(0), b_mn, M, N, K,
                                         C2.data_ptr(), C2.stride(0), None, None, 0, 0, 0, st))
    torch.cuda.synchronize()
    assert (C2[:, :N].double() - Am @ Bm).abs().max().item() <= _tol(K, 1.0)
