"""Achieved HBM GB/s of the tf32 GEMM engine on the shapes of the training pipeline (config 3: ~300 k tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cikm2020_dmt_b200 import abi

lib = abi.load()
st = torch.cuda.current_stream().cuda_stream
T = 300000


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print("rows: M=%d" % T)
for (K, N, extra) in [(64, 192, ""), (64, 256, ""), (256, 64, "addend"), (64, 256, "mask"), (192, 64, "addend"),
                      (128, 64, ""), (64, 128, ""), (80, 240, ""), (320, 80, "addend"), (80, 160, "")]:
    A = torch.randn(T, K, device="cuda")
    Bt = torch.randn(N, K, device="cuda")
    C = torch.empty(T, N, device="cuda")
    ad = torch.randn(T, N, device="cuda") if extra == "addend" else None
    mk = torch.randn(T, N, device="cuda") if extra == "mask" else None
    fn = lambda: abi.check(lib.dmt_selftest_tf32_rows(A.data_ptr(), K, Bt.data_ptr(), K, T, N, K, C.data_ptr(), N, None,
                                                      abi.ptr(ad), N, abi.ptr(mk), N, 1.0, 0, 0, st))
    ms = timeit(fn)
    gb = T * (K + N + (N if extra else 0)) * 4 / 1e9
    print("  K=%3d N=%3d %-6s %7.3f ms  %7.1f GB/s" % (K, N, extra, ms, gb / (ms / 1e3)))
print("wgrad: T=%d" % T)
for (MA, NB) in [(256, 64), (192, 64), (128, 64), (320, 80), (240, 80)]:
    P = torch.randn(T, MA, device="cuda")
    Q = torch.randn(T, NB, device="cuda")
    C = torch.zeros(NB, MA, device="cuda")
    ws = torch.empty(lib.dmt_selftest_tf32_wgrad_bytes(T, MA, NB), dtype=torch.uint8, device="cuda")
    fn = lambda: abi.check(lib.dmt_selftest_tf32_wgrad(P.data_ptr(), MA, Q.data_ptr(), NB, T, MA, NB, C.data_ptr(), MA, 1, 0,
                                                       ws.data_ptr(), st))
    ms = timeit(fn)
    gb = T * (MA + NB) * 4 / 1e9
    print("  MA=%3d NB=%3d %7.3f ms  %7.1f GB/s" % (MA, NB, ms, gb / (ms / 1e3)))
X = torch.randn(T, 256, device="cuda")
out = torch.zeros(256, device="cuda")
scr = torch.empty(296 * 256, device="cuda")
ms = timeit(lambda: abi.check(lib.dmt_selftest_tf32_colsum(X.data_ptr(), 256, T, 256, out.data_ptr(), 0, scr.data_ptr(), st)))
print("colsum W=256: %.3f ms %.1f GB/s" % (ms, T * 256 * 4 / 1e9 / (ms / 1e3)))
