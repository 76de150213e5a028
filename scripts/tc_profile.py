"""Stage cycle breakdown of sa_tc_kernel (library built with EXTRA=-DB200_TC_PROFILE): CTA (0,0), worker thread 0."""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, cases
pkg = importlib.import_module("3dioumatch_b200"); pkg.install_dropin()
import pointnet2._ext as ext
L = ctypes.CDLL(pkg.LIB_PATH)
names = ["gather", "wait L1 mma", "epi L1", "wait L2 mma", "epi L2", "wait L3 mma", "epi L3(max)"]
for (B, N, M, C, r, ns, spec) in [(8, 2048, 1024, 128, 0.4, 32, [131, 128, 128, 256]), (8, 1024, 512, 256, 0.8, 16, [259, 128, 128, 256])]:
    xyz = torch.from_numpy(cases.scene_cloud(0, B, N)[:, :, :3].copy()).cuda()
    feats = torch.randn(B, C, N, device="cuda")
    inds = ext.furthest_point_sampling(xyz, M)
    new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    layers = [(torch.from_numpy(l["weight"]).cuda(), torch.from_numpy(l["gamma"]).cuda(), torch.from_numpy(l["beta"]).cuda()) for l in cases.mlp_params(0, spec)]
    idx = ext.ball_query(new_xyz, xyz, r, ns)
    fpm = feats.transpose(1, 2).contiguous()
    buf = (ctypes.c_ulonglong * 16)()
    ext.sa_forward(xyz, None, new_xyz, r, ns, layers, normalize_xyz=True, idx=idx, features_pm=fpm)
    L.b200_debug_tc_profile(buf)
    n = 5
    for _ in range(n):
        ext.sa_forward(xyz, None, new_xyz, r, ns, layers, normalize_xyz=True, idx=idx, features_pm=fpm)
    L.b200_debug_tc_profile(buf)
    print(spec, "ns", ns, " total %.0f cycles/tile: " % (sum(buf[:7]) / n) + "  ".join("%s %.0f" % (nm, buf[i] / n) for i, nm in enumerate(names)))
