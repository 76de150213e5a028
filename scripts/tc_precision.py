import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("3dioumatch_b200")
lib = ctypes.CDLL(pkg.LIB_PATH)
torch.backends.cuda.matmul.allow_tf32 = False
for K in (128, 288):
    for dist in ("randn", "relu"):
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(128, K, device="cuda", generator=g)
        if dist == "relu":
            A = A.clamp_min(0) * 2
        W = torch.randn(128, K, device="cuda", generator=g) / K ** 0.5
        ref = A.double() @ W.double().t()
        scale = ref.abs().max().item()
        out = {}
        for passes in (1, 3, 4):
            C = torch.zeros(128, 128, device="cuda")
            lib.b200_debug_tc_gemm(128, K, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(W.data_ptr()), ctypes.c_void_p(C.data_ptr()), passes, None)
            torch.cuda.synchronize()
            e = (C.double() - ref).abs()
            out[passes] = (e.max().item(), e.mean().item(), ((C.double() - ref).mean().item()))
        e = ((A @ W.t()).double() - ref).abs()
        print("K=%d %s scale=%.2f | fp32 torch: max %.2e mean %.2e | " % (K, dist, scale, e.max().item(), e.mean().item()) +
              " | ".join("passes=%d max %.2e mean %.2e bias %.1e" % (p, *out[p]) for p in (1, 3, 4)))
