"""Generates the golden fixtures of tests/golden/ from the UNMODIFIED reference built into oracle/_ref
(oracle/build_ref.py).  Inputs come from tests/cases.py (seeded); outputs are what the reference's own code returned.

  python tests/golden/make_golden.py --cpu            # here: reference CPU entry boxes_iou_bev_cpu -> iou_bev_cpu.npz
  python tests/golden/make_golden.py --gpu --out DIR  # on the B200 box: reference CUDA ops -> ref_cuda_*.npz
  python tests/golden/make_golden.py --nms            # here: the reference's numpy suppression loops and corner code
                                                      # imported from /root/reference -> ref_aabb_nms.npz
"""
import argparse
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402


def load_ref(name, rel):
    import torch  # noqa: F401
    path = os.path.join(ROOT, "oracle", "_ref", rel)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def iou_inputs():
    a = cases.boxes(0, 64)
    b = cases.boxes(1, 64, jitter_of=a)
    d = cases.degenerate_boxes()
    return {"rand_a": a, "rand_b": b, "deg": d}


def make_cpu(out_dir):
    import torch
    iou = load_ref("iou3d_nms_cuda", "pcdet/ops/iou3d_nms/iou3d_nms_cuda.so")
    inp = iou_inputs()
    res = {}
    for key, (a, b) in {"rand": (inp["rand_a"], inp["rand_b"]), "deg": (inp["deg"], inp["deg"])}.items():
        ta, tb = torch.from_numpy(a), torch.from_numpy(b)
        ans = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32)
        iou.boxes_iou_bev_cpu(ta, tb, ans)
        res[key + "_a"], res[key + "_b"], res[key + "_iou_bev"] = a, b, ans.numpy()
    np.savez_compressed(os.path.join(out_dir, "iou_bev_cpu.npz"), **res)
    print("wrote iou_bev_cpu.npz")


def pointops_inputs():
    """name -> dict of inputs (small enough to commit)."""
    c = {}
    c["c1"] = dict(xyz=cases.cloud(0, 1, 2000, centre=False), npoint=128, radius=0.2, nsample=32)
    c["ragged"] = dict(xyz=cases.cloud(1, 3, 1531, dup_frac=0.05, origin_frac=0.02), npoint=200, radius=0.35, nsample=16)
    c["small"] = dict(xyz=cases.cloud(2, 2, 300, dup_frac=0.3), npoint=64, radius=0.5, nsample=8)
    c["dups"] = dict(xyz=cases.cloud(3, 2, 4096, dup_frac=0.5, origin_frac=0.01), npoint=512, radius=0.15, nsample=64)
    c["tiny"] = dict(xyz=cases.cloud(4, 2, 9, extent=(1, 1, 1)), npoint=9, radius=5.0, nsample=6)
    return c


def make_gpu(out_dir):
    import torch
    ext = load_ref("_ext", "pointnet2/_ext.so")
    iou = load_ref("iou3d_nms_cuda", "pcdet/ops/iou3d_nms/iou3d_nms_cuda.so")
    dev = "cuda:0"
    res = {}
    for name, c in pointops_inputs().items():
        xyz = torch.from_numpy(c["xyz"]).to(dev)
        B, N, _ = xyz.shape
        inds = ext.furthest_point_sampling(xyz, c["npoint"])
        new_xyz = ext.gather_points(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        bq = ext.ball_query(new_xyz, xyz, c["radius"], c["nsample"])
        rng = np.random.default_rng(100)
        feats = torch.from_numpy(rng.standard_normal((B, 5, N)).astype(np.float32)).to(dev)
        grouped = ext.group_points(feats, bq)
        dist2, nn_idx = ext.three_nn(xyz, new_xyz)
        w = 1.0 / (torch.sqrt(dist2) + 1e-8)
        w = (w / w.sum(2, keepdim=True)).contiguous()
        known_feats = ext.gather_points(feats, inds)
        interp = ext.three_interpolate(known_feats, nn_idx, w)
        res.update({name + "_xyz": c["xyz"], name + "_fps": inds.cpu().numpy(), name + "_bq": bq.cpu().numpy(),
                    name + "_feats": feats.cpu().numpy(), name + "_grouped_sum": grouped.sum((2, 3)).cpu().numpy(),
                    name + "_nn_dist2": dist2.cpu().numpy(), name + "_nn_idx": nn_idx.cpu().numpy(),
                    name + "_w": w.cpu().numpy(), name + "_interp": interp.cpu().numpy(),
                    name + "_cfg": np.asarray([c["npoint"], c["nsample"]], np.int32),
                    name + "_radius": np.asarray([c["radius"]], np.float32)})
    np.savez_compressed(os.path.join(out_dir, "ref_cuda_pointops.npz"), **res)

    inp = iou_inputs()
    res = {}
    for key, (a, b) in {"rand": (inp["rand_a"], inp["rand_b"]), "deg": (inp["deg"], inp["deg"])}.items():
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        ov = torch.zeros((a.shape[0], b.shape[0]), device=dev)
        iou.boxes_overlap_bev_gpu(ta, tb, ov)
        ib = torch.zeros_like(ov)
        iou.boxes_iou_bev_gpu(ta, tb, ib)
        res[key + "_a"], res[key + "_b"] = a, b
        res[key + "_overlap"], res[key + "_iou_bev"] = ov.cpu().numpy(), ib.cpu().numpy()
    # NMS on score-sorted boxes (int32 keep: see SURVEY 2a quirk)
    big = cases.boxes(5, 300, extent=(4.0, 4.0, 1.0))
    scores = np.random.default_rng(6).random(300).astype(np.float32)
    order = np.argsort(-scores, kind="stable")
    sb = torch.from_numpy(big[order]).to(dev).contiguous()
    for thr in (0.25, 0.05):
        keep = torch.zeros(300, dtype=torch.int32)
        n = iou.nms_gpu(sb, keep, thr)
        res["nms_keep_%g" % thr] = keep[:n].numpy().copy()
        keep2 = torch.zeros(300, dtype=torch.int32)
        n2 = iou.nms_normal_gpu(sb, keep2, thr)
        res["nms_normal_keep_%g" % thr] = keep2[:n2].numpy().copy()
    res["nms_boxes_sorted"] = big[order]
    np.savez_compressed(os.path.join(out_dir, "ref_cuda_iou.npz"), **res)
    print("wrote ref_cuda_pointops.npz, ref_cuda_iou.npz to", out_dir)


def load_reference_py(name, rel, stubs=()):
    """Import one UNMODIFIED python module of the reference by path (never copied into the repo)."""
    import types
    for st in stubs:
        sys.modules.setdefault(st, types.ModuleType(st))
    spec = importlib.util.spec_from_file_location(name, os.path.join("/root/reference", rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_nms(out_dir):
    """utils/nms.py (nms_2d_faster, nms_3d_faster, nms_3d_faster_samecls, lhs_3d_faster_samecls) and
    utils/box_util.py get_3d_box + models/ap_helper.py flip_axis_to_camera on seeded inputs with distinct scores."""
    import types
    stub = types.ModuleType("pc_util")
    stub.bbox_corner_dist_measure = None
    sys.modules["pc_util"] = stub
    ref = load_reference_py("ref_nms", "utils/nms.py")
    for name in ("pcdet", "pcdet.ops", "pcdet.ops.iou3d_nms", "pcdet.ops.iou3d_nms.iou3d_nms_utils"):
        sys.modules.setdefault(name, types.ModuleType(name))  # box_util imports (and here never calls) the IoU op
    sys.modules["pcdet.ops.iou3d_nms.iou3d_nms_utils"].boxes_iou3d_gpu = None
    box_util = load_reference_py("ref_box_util", "utils/box_util.py")
    res = {}
    for name, (seed, K, ncls) in {"k64": (0, 64, 3), "k256": (1, 256, 18), "k37": (2, 37, 1), "k1": (3, 1, 2)}.items():
        b = cases.aabb_boxes(seed, K, ncls)
        res[name + "_boxes"] = b
        for thr in (0.25, 0.5):
            for old in (0, 1):
                tag = "%s_t%g_o%d" % (name, thr, old)
                res[tag + "_nms3d"] = np.asarray(ref.nms_3d_faster(b[:, :7], thr, bool(old)), np.int32)
                res[tag + "_nms3d_cls"] = np.asarray(ref.nms_3d_faster_samecls(b, thr, bool(old)), np.int32)
                res[tag + "_lhs_cls"] = np.asarray(ref.lhs_3d_faster_samecls(b, thr, bool(old)), np.int32)
                res[tag + "_nms2d"] = np.asarray(ref.nms_2d_faster(b[:, [0, 2, 3, 5, 6]], thr, bool(old)), np.int32)
    # corners: the loop body of predictions2corners3d (models/ap_helper.py:82-91)
    rng = np.random.default_rng(9)
    K = 200
    center = (rng.random((K, 3)) * [8, 8, 3] - [4, 4, 0]).astype(np.float32)
    size = rng.random((K, 3)) * 2 + 0.05 + rng.standard_normal((K, 3)).astype(np.float32) * 0.01
    heading = (rng.random(K) - 0.5) * 2 * np.pi
    heading[:20] = 0.0  # ScanNet: class2angle returns zeros
    cam = center.copy()
    cam[..., [0, 1, 2]] = cam[..., [0, 2, 1]]  # flip_axis_to_camera (models/ap_helper.py:28-35)
    cam[..., 1] *= -1
    corners = np.zeros((K, 8, 3), np.float32)
    for j in range(K):
        corners[j] = box_util.get_3d_box(size[j], heading[j], cam[j])
    res.update(corner_center=center, corner_size=size, corner_heading=heading, corners=corners)
    np.savez_compressed(os.path.join(out_dir, "ref_aabb_nms.npz"), **res)
    print("wrote ref_aabb_nms.npz")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--nms", action="store_true")
    ap.add_argument("--out", default=HERE)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    if a.cpu:
        make_cpu(a.out)
    if a.gpu:
        make_gpu(a.out)
    if a.nms:
        make_nms(a.out)
