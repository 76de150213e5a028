"""Every libb200pc kernel once at small shapes, for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, cases
pkg = importlib.import_module("3dioumatch_b200"); pkg.install_dropin()
import pointnet2._ext as ext
from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
def lay(spec):
    return [(torch.from_numpy(l["weight"]).cuda(), torch.from_numpy(l["gamma"]).cuda(), torch.from_numpy(l["beta"]).cuda()) for l in cases.mlp_params(0, spec)]
for (B, N, m) in ((2, 700, 64), (2, 9000, 128)):          # single-CTA and cluster FPS
    xyz = torch.from_numpy(cases.cloud(0, B, N, dup_frac=0.1)).cuda()
    inds = ext.furthest_point_sampling(xyz, m)
    new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
    for grid in ("0", "1"):
        os.environ["B200_BQ_GRID"] = grid
        idx = ext.ball_query(new_xyz, xyz, 0.4, 16)
    feats = torch.randn(B, 32, N, device="cuda")
    g = ext.group_points(feats, idx); ext.group_points_grad(g, idx, N)
    ext.gather_points_grad(ext.gather_points(feats, inds), inds, N)
    d2, nn = ext.three_nn(xyz, new_xyz)
    w = torch.full_like(d2, 1 / 3)
    kf = ext.gather_points(feats, inds)
    o = ext.three_interpolate(kf, nn, w); ext.three_interpolate_grad(o, nn, w, m)
    ext.sa_forward(xyz, feats, new_xyz, 0.4, 16, lay([35, 64, 128]), normalize_xyz=True)          # tensor-core kernel
    ext.sa_forward(xyz, feats, new_xyz, 0.4, 16, lay([35, 128, 128, 256]), normalize_xyz=True)    # 256-wide, two halves
    ext.sa_forward(xyz, feats[:, :1].contiguous(), new_xyz, 0.4, 64, lay([4, 64, 64, 128]), normalize_xyz=True)  # compact
    ext.sa_forward(xyz, feats, new_xyz, 0.4, 12, lay([35, 20, 33]), normalize_xyz=True)           # fp32 FFMA kernel
    rows = torch.randint(0, m, (B, 4 * 64, 3), device="cuda", dtype=torch.int32)
    ext.interp_mlp_forward(kf, rows, torch.rand(B, 4 * 64, 3, device="cuda"), torch.rand(B, 4 * 64, 3, device="cuda"), 64, lay([35, 128, 128]))
a = torch.from_numpy(cases.boxes(0, 100)).cuda(); b = torch.from_numpy(cases.boxes(1, 70)).cuda()
iu.boxes_iou3d_gpu(a, b); iu.boxes_iou_bev(a, b); iu.nms_gpu(a, torch.rand(100, device="cuda"), 0.25); iu.nms_normal_gpu(a, torch.rand(100, device="cuda"), 0.25)
iu.boxes_iou3d_batched(a.view(2, 50, 7), b[:60].reshape(2, 30, 7).contiguous())
torch.cuda.synchronize(); print("sanitize run complete")
