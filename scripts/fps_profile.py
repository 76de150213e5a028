"""Per-stage cycle breakdown of the FPS kernel (needs a library built with EXTRA=-DB200_FPS_PROFILE)."""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, cases
pkg = importlib.import_module("3dioumatch_b200"); pkg.install_dropin()
import pointnet2._ext as ext
L = ctypes.CDLL(pkg.LIB_PATH)
names = ["update", "warp argmax", "cta argmax", "st.async", "mbar wait", "cluster argmax"]
for (B, N, m) in ((8, 40000, 2048), (8, 2048, 1024), (8, 512, 256)):
    x = torch.from_numpy(cases.scene_cloud(0, B, N)[:, :, :3].copy()).cuda()
    ext.furthest_point_sampling(x, m)
    buf = (ctypes.c_ulonglong * 8)()
    L.b200_debug_fps_profile(buf)
    ext.furthest_point_sampling(x, m)
    L.b200_debug_fps_profile(buf)
    tot = sum(buf[:6])
    print("B=%d N=%d m=%d cluster=%s threads=%s: %.0f cycles/iter" % (B, N, m, os.environ.get("B200_FPS_CLUSTER", "auto"),
          os.environ.get("B200_FPS_THREADS", "auto"), tot / (m - 1)))
    print("   " + "  ".join("%s %.0f" % (n, buf[i] / (m - 1)) for i, n in enumerate(names)))
