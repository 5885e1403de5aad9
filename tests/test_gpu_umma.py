"""tcgen05 building blocks: every shared-memory operand layout / descriptor encoding the bf16 kernels
rely on, checked against a plain fp32 matmul of the same bf16-rounded operands (exact products, fp32
accumulation: tolerance 1e-3 relative to the row scale)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("N,K", [(64, 64), (192, 64), (256, 64), (64, 256), (128, 32), (32, 128), (240, 80), (80, 320)])
def test_umma_selftest(mode, N, K):
    from cikm2020_dmt_b200 import abi
    if mode == 2 and K % 64:
        pytest.skip("SWIZZLE_128B tiles are 64 elements wide")
    lib = abi.load()
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    A = (torch.randn(128, K, generator=g)).to(torch.bfloat16)
    Bt = (torch.randn(N, K, generator=g)).to(torch.bfloat16)          # B^T: [N, K]
    want = A.float() @ Bt.float().t()
    Ad = A.cuda()
    Bd = (Bt.t().contiguous() if mode == 1 else Bt).cuda()            # mode 1 takes [K, N]
    C = torch.full((128, N), float("nan"), device="cuda")
    abi.check(lib.dmt_selftest_umma(mode, Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), N, K,
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (C.cpu() - want).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), "mode %d N %d K %d: max err %g" % (mode, N, K, err)


@pytest.mark.parametrize("mode", [3, 4, 5])
@pytest.mark.parametrize("N,K", [(64, 64), (32, 128), (64, 256), (80, 128), (128, 32), (16, 64)])
def test_umma_tmem_operand_selftest(mode, N, K):
    """A operand read from tensor memory (modes 3, 4: P.V, H.W2 in the sequence kernel) and the compact
    16-row A image of the decoder context MMA (mode 5)."""
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    g = torch.Generator().manual_seed(N * 1000 + K + mode)
    A = (torch.randn(128, K, generator=g)).to(torch.bfloat16)
    Bt = (torch.randn(N, K, generator=g)).to(torch.bfloat16)
    want = A.float() @ Bt.float().t()
    Ad = A.cuda()
    Bd = (Bt if mode == 3 else Bt.t().contiguous()).cuda()
    C = torch.full((128, N), float("nan"), device="cuda")
    abi.check(lib.dmt_selftest_umma(mode, Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), N, K,
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    rows = 16 if mode == 5 else 128
    err = (C.cpu()[:rows] - want[:rows]).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), "mode %d N %d K %d: max err %g" % (mode, N, K, err)


@pytest.mark.parametrize("N,K", [(16, 128), (16, 64), (32, 128), (64, 32)])
def test_umma_mn_major_a_selftest(N, K):
    """A operand read MN-major from shared memory (the memory rows of the transposed decoder-context MMA)."""
    from cikm2020_dmt_b200 import abi
    lib = abi.load()
    g = torch.Generator().manual_seed(N * 1000 + K + 6)
    A = (torch.randn(128, K, generator=g)).to(torch.bfloat16)
    Bt = (torch.randn(N, K, generator=g)).to(torch.bfloat16)
    want = A.float() @ Bt.float().t()
    Ad = A.t().contiguous().cuda()            # [K][128]
    Bd = Bt.cuda()
    C = torch.full((128, N), float("nan"), device="cuda")
    abi.check(lib.dmt_selftest_umma(6, Ad.data_ptr(), Bd.data_ptr(), C.data_ptr(), N, K,
                                    torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    err = (C.cpu() - want).abs().max().item()
    assert err < 1e-3 * max(1.0, want.abs().max().item()), "N %d K %d: max err %g" % (N, K, err)
