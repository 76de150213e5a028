"""FPS launch-shape sweep (in-process, b200pn2_fps_force_shape): time per launch for every kernel generation and every
(cluster size, threads per CTA) that can hold the cloud, at the shapes of a VoteNet step.  CUDA events, B=8.
  python scripts/fps_shapes.py [quick]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import cases  # noqa: E402

pkg = importlib.import_module("3dioumatch_b200")
pkg.install_dropin()
cabi = importlib.import_module("3dioumatch_b200._cabi")
import pointnet2._ext as ext  # noqa: E402


def timeit(fn, iters=10, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
shapes = ((8, 40000, 2048), (16, 20000, 2048)) if quick else \
    ((8, 40000, 2048), (8, 2048, 1024), (8, 1024, 512), (8, 512, 256), (8, 1024, 256), (16, 20000, 2048))
names = {-1: "auto", 0: "owner", 2: "v1"}
for (B, N, m) in shapes:
    x = torch.from_numpy(cases.scene_cloud(0, B, N)[:, :, :3].copy()).cuda()
    cabi.force_fps_shape(2, 0, 0)
    ref = ext.furthest_point_sampling(x, m)
    print("B=%d N=%d m=%d" % (B, N, m), flush=True)
    for kern in (-1, 0, 2):
        for pol in ("latency", "throughput"):
            cabi.force_fps_shape(kern, 0, 0)
            cabi.set_fps_policy(pol)
            ms = timeit(lambda: ext.furthest_point_sampling(x, m))
            same = torch.equal(ext.furthest_point_sampling(x, m), ref)
            print("  %-11s auto/%-10s %.3f ms (%.3f us/iter) %s" % (names[kern], pol, ms, ms * 1e3 / (m - 1), "" if same else "MISMATCH"), flush=True)
        cabi.set_fps_policy("latency")
        if kern == -1:
            continue
        for cs in (1, 2, 4, 8, 16):
            row = []
            for th in (32, 64, 128, 256, 512):
                if N >= 8192 and cs * th < 1024:
                    continue
                cabi.force_fps_shape(kern, cs, th)
                try:
                    ms = timeit(lambda: ext.furthest_point_sampling(x, m), iters=5, warm=1)
                    same = torch.equal(ext.furthest_point_sampling(x, m), ref)
                    row.append("%dx%d: %.3f%s" % (cs, th, ms, "" if same else " MISMATCH"))
                except RuntimeError:
                    pass
            if row:
                print("  %-11s %s" % (names[kern], "   ".join(row)), flush=True)
cabi.force_fps_shape(-1, 0, 0)
