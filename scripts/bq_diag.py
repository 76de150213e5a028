"""Developer check: device time vs host enqueue time of the standalone ball query at SA1 shape."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import cases  # noqa: E402

pkg = importlib.import_module("3dioumatch_b200")
pkg.install_dropin()
import pointnet2._ext as ext  # noqa: E402

x = torch.from_numpy(cases.scene_cloud(0, 8, 40000)[:, :, :3].copy()).cuda()
inds = ext.furthest_point_sampling(x, 2048)
new_xyz = ext.gather_points(x.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
for label, env in (("grid", None), ("brute", "0")):
    if env is None:
        os.environ.pop("B200_BQ_GRID", None)
    else:
        os.environ["B200_BQ_GRID"] = env
    for _ in range(3):
        ext.ball_query(new_xyz, x, 0.2, 64)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        ext.ball_query(new_xyz, x, 0.2, 64)
    e1.record()
    t_host = (time.perf_counter() - t0) / 20 * 1e3
    torch.cuda.synchronize()
    print("%s: device %.3f ms/call, host enqueue %.3f ms/call" % (label, e0.elapsed_time(e1) / 20, t_host))
