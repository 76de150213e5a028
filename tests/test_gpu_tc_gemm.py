"""GPU: pins the tcgen05 building blocks (descriptor encodings, 128-byte swizzle, TMEM addressing, commit/mbarrier) with a
128 x N x K split-precision TF32 GEMM against an fp64 reference."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def run(pkg, N, K, passes, seed=0):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.randn(128, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    C = torch.zeros(128, N, device="cuda")
    rc = lib.b200_debug_tc_gemm(N, K, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(W.data_ptr()),
                                ctypes.c_void_p(C.data_ptr()), passes, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    ref = (A.double() @ W.double().t())
    return (C.double() - ref).abs().max().item(), ref.abs().max().item()


@pytest.mark.parametrize("N,K", [(128, 32), (128, 128), (256, 128), (128, 288), (256, 160)])
def test_split_tf32_gemm(pkg, N, K):
    err1, scale = run(pkg, N, K, passes=1)
    err3, _ = run(pkg, N, K, passes=3)
    assert err1 < 2e-2 * scale, "plain TF32 result is wrong (layout/descriptor error): %g" % err1
    assert err3 < 3e-6 * max(scale, 1.0), "split-precision result not fp32-accurate: %g" % err3
    assert err3 < err1 / 20
