"""Developer micro-benchmarks (CUDA events, device-resident inputs): FPS cluster/thread sweep and the split of the fused
SA call into ball query vs MLP+max.  Usage on the GPU box: python scripts/op_sweep.py [fps|fps_one|sa|nms|all]"""
import importlib
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def timeit(torch, fn, iters=20, warm=3):
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


def fps_one():
    import torch
    import cases
    pkg = importlib.import_module("3dioumatch_b200")
    pkg.install_dropin()
    import pointnet2._ext as ext
    for (B, N, m) in ((8, 40000, 2048), (8, 2048, 1024), (8, 1024, 512), (8, 512, 256), (8, 1024, 256), (16, 20000, 2048)):
        x = torch.from_numpy(cases.scene_cloud(0, B, N)[:, :, :3].copy()).cuda()
        try:
            ms = timeit(torch, lambda: ext.furthest_point_sampling(x, m))
            print("  B=%d N=%d m=%d: %.3f ms  (%.3f us/iter)" % (B, N, m, ms, ms * 1e3 / (m - 1)))
        except RuntimeError as e:
            print("  B=%d N=%d m=%d: n/a (%s)" % (B, N, m, str(e)[:60]))


def fps_sweep():
    for cs in (0, 1, 2, 4, 8, 16):
        for th in (0, 32, 128, 256, 512):
            env = dict(os.environ)
            if cs:
                env["B200_FPS_CLUSTER"] = str(cs)
            if th:
                env["B200_FPS_THREADS"] = str(th)
            print("cluster=%s threads=%s" % (cs or "auto", th or "auto"), flush=True)
            subprocess.run([sys.executable, os.path.abspath(__file__), "fps_one"], env=env)


def nms_ops():
    """Device-side suppression against the numpy restatement of the reference's host loops."""
    import time
    import numpy as np
    import torch
    import cases
    pkg = importlib.import_module("3dioumatch_b200")
    pkg.install_dropin()
    nms = importlib.import_module("utils.nms")
    from oracle import oracle as orc
    for (B, K) in ((8, 64), (8, 256), (12, 128), (8, 1024)):
        b = np.stack([cases.aabb_boxes(i, K, 18) for i in range(B)])
        t = torch.from_numpy(b).cuda()
        ms = timeit(torch, lambda: nms.suppress_batch(t, 0.25, True, True))
        t0 = time.perf_counter()
        for i in range(B):
            orc.aabb_suppress(b[i], 0.25, True, True)
        cpu = (time.perf_counter() - t0) * 1e3
        print("  lhs_3d_faster_samecls B=%d K=%d: device %.3f ms (incl. torch wrapper), numpy host loop %.2f ms" % (B, K, ms, cpu))
        c = torch.rand((B, K, 3), device="cuda")
        s = torch.rand((B, K, 3), device="cuda", dtype=torch.float64) + 0.1
        h = torch.zeros((B, K), device="cuda", dtype=torch.float64)
        print("  box_extents B=%d K=%d: %.3f ms" % (B, K, timeit(torch, lambda: nms.box_extents_batch(c, s, h))))


def sa_split():
    import numpy as np
    import torch
    import cases
    pkg = importlib.import_module("3dioumatch_b200")
    pkg.install_dropin()
    import pointnet2._ext as ext
    cfgs = [(8, 40000, 2048, 1, 0.2, 64, [4, 64, 64, 128]), (8, 2048, 1024, 128, 0.4, 32, [131, 128, 128, 256]),
            (8, 1024, 512, 256, 0.8, 16, [259, 128, 128, 256]), (8, 512, 256, 256, 1.2, 16, [259, 128, 128, 256]),
            (8, 1024, 256, 256, 0.3, 16, [259, 128, 128, 128])]
    for (B, N, M, C, r, ns, spec) in cfgs:
        xyz = torch.from_numpy(cases.scene_cloud(0, B, N)[:, :, :3].copy()).cuda()
        feats = torch.randn(B, C, N, device="cuda")
        inds = ext.furthest_point_sampling(xyz, M)
        new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        layers = []
        for ly in cases.mlp_params(0, spec):
            layers.append((torch.from_numpy(ly["weight"]).cuda(), torch.from_numpy(ly["gamma"]).cuda(),
                           torch.from_numpy(ly["beta"]).cuda()))
        idx = ext.ball_query(new_xyz, xyz, r, ns)
        fpm = feats.transpose(1, 2).contiguous()
        t_bq = timeit(torch, lambda: ext.ball_query(new_xyz, xyz, r, ns))
        t_all = timeit(torch, lambda: ext.sa_forward(xyz, feats, new_xyz, r, ns, layers, normalize_xyz=True))
        t_mlp = timeit(torch, lambda: ext.sa_forward(xyz, None, new_xyz, r, ns, layers, normalize_xyz=True, idx=idx,
                                                      features_pm=fpm))
        flops = 2.0 * B * M * ns * sum(a * b for a, b in zip(spec[:-1], spec[1:]))
        print("SA N=%d M=%d ns=%d %s: ball_query %.3f ms | fused %.3f ms | mlp+max only %.3f ms (%.1f TFLOP/s fp32)" %
              (N, M, ns, spec, t_bq, t_all, t_mlp, flops / t_mlp / 1e9))


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "fps_one":
        fps_one()
    if what in ("fps", "all"):
        fps_sweep()
    if what in ("sa", "all"):
        sa_split()
    if what in ("nms", "all"):
        nms_ops()
