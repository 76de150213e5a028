"""tcgen05 issue-mode micro-benchmark + TMEM-A-operand self-test (3dioumatch_b200/csrc/tc_bench.cu)."""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("3dioumatch_b200")
lib = ctypes.CDLL(pkg.LIB_PATH)
names = {0: "SS N=128", 1: "SS N=256", 2: "SS N=64", 3: "TS N=128", 4: "TS N=256", 5: "TS N=64"}
# TS correctness first
for (N, K) in ((128, 32), (128, 96), (64, 64), (32, 96)):
    g = torch.Generator(device="cuda").manual_seed(N + K)
    A = torch.randn(128, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    C = torch.zeros(128, N, device="cuda")
    rc = lib.b200_debug_tc_gemm_ts(N, K, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(W.data_ptr()), ctypes.c_void_p(C.data_ptr()), None)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    print("TS gemm N=%d K=%d rc=%d max err %.3g (scale %.3g)" % (N, K, rc, (C.double() - ref).abs().max().item(), ref.abs().max().item()))
out = torch.zeros(3, dtype=torch.int64, device="cuda")
for stress in (0, 4, 8):
    for mode in range(6):
        out.zero_()
        rc = lib.b200_debug_tc_rate(mode, 512, stress, ctypes.c_void_p(out.data_ptr()), None)
        torch.cuda.synchronize()
        cyc, n, sb = out.tolist()
        print("stress=%d %-9s rc=%d: %.1f cycles/MMA (%d MMAs, %d cycles), stress stores %.1f B/cycle" % (stress, names[mode], rc, cyc / max(n, 1), n, cyc, sb / max(cyc, 1)))
