"""Issue rates of the FPS update's instruction kinds (csrc/pipe_bench.cu).  On the GPU box:  python scripts/pipe_rate.py"""
import ctypes, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = ctypes.CDLL(os.environ.get("B200_LIB_PATH") or os.path.join(ROOT, "3dioumatch_b200", "lib", "libb200pc.so"))
out = (ctypes.c_float * 18)()
rc = L.b200_debug_pipe_rates(2000, out)
assert rc == 0, rc
print("cycles per warp-instruction per scheduler (1.0 = full rate)")
print("%-16s %8s %8s %8s" % ("kind", "4 warps", "8 warps", "16 warps"))
for k, name in enumerate(("FFMA", "FFMA2 (f32x2)", "FADD2 (f32x2)", "FMNMX", "FSETP+FSEL+SEL", "IADD")):
    print("%-16s %8.2f %8.2f %8.2f" % (name, out[k * 3], out[k * 3 + 1], out[k * 3 + 2]))
