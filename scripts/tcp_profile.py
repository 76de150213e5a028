"""Wait-cycle breakdown per role of the persistent tensor-core SA kernel (sa_tcp_kernel), CTA 0.
Build:  make -C 3dioumatch_b200/csrc prof      Run:  B200_LIB_PATH=3dioumatch_b200/lib/libb200pc_prof.so python scripts/tcp_profile.py"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("B200_LIB_PATH", os.path.join(ROOT, "3dioumatch_b200", "lib", "libb200pc_prof.so"))
import torch, cases
pkg = importlib.import_module("3dioumatch_b200"); pkg.install_dropin()
import pointnet2._ext as ext
L = ctypes.CDLL(os.environ["B200_LIB_PATH"])
cats = {0: "ldr:tq_empty", 1: "ldr:empty_w", 2: "mma:tile", 3: "mma:x_ready", 4: "mma:d_free", 5: "mma:full_a", 6: "mma:full_w",
        8: "prod:tile", 9: "prod:empty_a", 7: "mma:issue W_hi groups", 10: "prod:r1_free", 15: "epi:hidden body", 24: "epi:final ld", 25: "epi:final math+shfl", 26: "epi:final store", 27: "epi:combine", 28: "epi:hidden ld", 29: "epi:hidden math", 12: "epi:tile", 13: "epi:accum_full", 14: "epi:accum_half",
        16: "TOTAL loader", 17: "TOTAL mma", 18: "TOTAL producer", 19: "TOTAL epilogue"}
cfgs = [(8, 40000, 2048, 1, 0.2, 64, [4, 64, 64, 128]), (8, 2048, 1024, 128, 0.4, 32, [131, 128, 128, 256]),
        (8, 1024, 512, 256, 0.8, 16, [259, 128, 128, 256]), (8, 1024, 256, 256, 0.3, 16, [259, 128, 128, 128])]
for (B, N, M, C, r, ns, spec) in cfgs:
    xyz = torch.from_numpy(cases.scene_cloud(0, B, N)[:, :, :3].copy()).cuda()
    feats = torch.randn(B, C, N, device="cuda")
    inds = ext.furthest_point_sampling(xyz, M)
    new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    layers = [(torch.from_numpy(l["weight"]).cuda(), torch.from_numpy(l["gamma"]).cuda(), torch.from_numpy(l["beta"]).cuda())
              for l in cases.mlp_params(0, spec)]
    idx = ext.ball_query(new_xyz, xyz, r, ns)
    fpm = feats.transpose(1, 2).contiguous()
    buf = (ctypes.c_ulonglong * 48)()
    for _ in range(2):
        ext.sa_forward(xyz, None, new_xyz, r, ns, layers, normalize_xyz=True, idx=idx, features_pm=fpm)
    L.b200_debug_tcp_profile(buf)
    n = 5
    for _ in range(n):
        ext.sa_forward(xyz, None, new_xyz, r, ns, layers, normalize_xyz=True, idx=idx, features_pm=fpm)
    L.b200_debug_tcp_profile(buf)
    tiles = max(buf[21], 1)
    print("%s ns=%d: CTA0 ran %.1f tiles/launch; cycles per tile:" % (spec, ns, buf[21] / n))
    print("   " + "  ".join("%s %.0f" % (cats[c], buf[c] / tiles) for c in sorted(cats)))

# ---- row MLPs (ROWOUT kernels): the GridConv mlp_before_iou of the unmodified reference callers (channel-major input),
#      a per-point row GEMM (pass 1 of a factorised first layer) and a head-sized stack
rows_cfgs = [("mlp_before_iou (8,259,16384) cm", 8, 16384, 259, [128, 128, 128], True),
             ("mlp_before_iou (8,16384,260) pm", 8, 16384, 259, [128, 128, 128], False),
             ("row GEMM 16384 x 128 -> 128", 1, 16384, 128, [128], False),
             ("head 8192 rows 256 -> 256 -> 256", 8, 1024, 256, [128, 128], False)]
for (name, S, R, C, spec, cm) in rows_cfgs:
    g = torch.Generator(device="cuda").manual_seed(1)
    layers, cin = [], C
    for co in spec:
        layers.append((torch.randn(co, cin, device="cuda", generator=g) / cin ** 0.5, torch.ones(co, device="cuda"),
                       torch.zeros(co, device="cuda")))
        cin = co
    if cm:
        x = torch.randn(S, C, R, device="cuda", generator=g)
        call = lambda: ext.row_mlp_forward_cm(x, layers, relu_last=True, want_cm=True)
    else:
        ld = (C + 3) // 4 * 4
        x = torch.randn(S, R, ld, device="cuda", generator=g)
        call = lambda: ext.row_mlp_forward(x, layers, relu_last=True, want_cm=True, channels=C)
    buf = (ctypes.c_ulonglong * 48)()
    for _ in range(2):
        call()
    L.b200_debug_tcp_profile(buf)
    n = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    L.b200_debug_tcp_profile(buf)
    tiles = max(buf[21], 1)
    print("%s: %.3f ms/launch (profile build), CTA0 ran %.1f tiles/launch; cycles per tile:" % (name, e0.elapsed_time(e1) / n, buf[21] / n))
    print("   " + "  ".join("%s %.0f" % (cats[c], buf[c] / tiles) for c in sorted(cats)))
