"""numpy front-end of the CPU oracle (oracle/*.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
may import this module.  The product package never does.

Each function mirrors one entry of the reference's native operator surface
(pointnet2/_ext_src/src/bindings.cpp:11-24 and
OpenPCDet/pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17) on numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)
_int = ctypes.c_int
_float = ctypes.c_float


def build(force=False):
    """Compile oracle/*.c with gcc (oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("pointnet2_oracle.c", "iou3d_oracle.c", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.orc_opt_n_threads.argtypes = [_int]
        L.orc_opt_n_threads.restype = _int
        L.orc_num_threads.restype = _int
        L.orc_furthest_point_sampling.argtypes = [_int, _int, _int, _f32p, _i32p]
        L.orc_gather_points.argtypes = [_int, _int, _int, _int, _f32p, _i32p, _f32p]
        L.orc_gather_points_grad.argtypes = [_int, _int, _int, _int, _f32p, _i32p, _f32p]
        L.orc_ball_query.argtypes = [_int, _int, _int, _float, _int, _f32p, _f32p, _i32p]
        L.orc_group_points.argtypes = [_int, _int, _int, _int, _int, _f32p, _i32p, _f32p]
        L.orc_group_points_grad.argtypes = [_int, _int, _int, _int, _int, _f32p, _i32p, _f32p]
        L.orc_three_nn.argtypes = [_int, _int, _int, _f32p, _f32p, _f32p, _i32p]
        L.orc_three_interpolate.argtypes = [_int, _int, _int, _int, _f32p, _i32p, _f32p, _f32p]
        L.orc_three_interpolate_grad.argtypes = [_int, _int, _int, _int, _f32p, _i32p, _f32p, _f32p]
        L.orc_box_overlap.argtypes = [_f32p, _f32p]
        L.orc_box_overlap.restype = _float
        for name in ("orc_boxes_overlap_bev", "orc_boxes_iou_bev", "orc_boxes_iou3d"):
            getattr(L, name).argtypes = [_int, _f32p, _int, _f32p, _f32p]
        L.orc_nms.argtypes = [_int, _f32p, _float, _int, _i32p]
        L.orc_nms.restype = _int
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(_f32p if a.dtype == np.float32 else _i32p)


def num_threads():
    return int(lib().orc_num_threads())


def opt_n_threads(n):
    return int(lib().orc_opt_n_threads(int(n)))


# ---- pointnet2._ext surface --------------------------------------------------------------------

def furthest_point_sampling(points, nsamples):
    """points (B,N,3) f32 -> idx (B,nsamples) int32   [sampling.cpp:70-91]"""
    points = _f32(points)
    B, N, _ = points.shape
    idx = np.zeros((B, nsamples), np.int32)
    lib().orc_furthest_point_sampling(B, N, int(nsamples), _p(points), _p(idx))
    return idx


def gather_points(points, idx):
    """points (B,C,N), idx (B,m) -> (B,C,m)   [sampling.cpp:20-44]"""
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    m = idx.shape[1]
    out = np.zeros((B, C, m), np.float32)
    lib().orc_gather_points(B, C, N, m, _p(points), _p(idx), _p(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, m = grad_out.shape
    out = np.zeros((B, C, n), np.float32)
    lib().orc_gather_points_grad(B, C, int(n), m, _p(grad_out), _p(idx), _p(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) int32   [ball_query.cpp:13-37]"""
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    B, M, _ = new_xyz.shape
    N = xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query(B, N, M, float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx))
    return idx


def group_points(points, idx):
    """points (B,C,N), idx (B,M,ns) -> (B,C,M,ns)   [group_points.cpp:17-39]"""
    points, idx = _f32(points), _i32(idx)
    B, C, N = points.shape
    _, M, ns = idx.shape
    out = np.zeros((B, C, M, ns), np.float32)
    lib().orc_group_points(B, C, N, M, ns, _p(points), _p(idx), _p(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    B, C, M, ns = grad_out.shape
    out = np.zeros((B, C, n), np.float32)
    lib().orc_group_points_grad(B, C, int(n), M, ns, _p(grad_out), _p(idx), _p(out))
    return out


def three_nn(unknown, known):
    """unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) f32, idx (B,n,3) int32   [interpolate.cpp:19-45]"""
    unknown, known = _f32(unknown), _f32(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    dist2 = np.zeros((B, n, 3), np.float32)
    idx = np.zeros((B, n, 3), np.int32)
    lib().orc_three_nn(B, n, m, _p(unknown), _p(known), _p(dist2), _p(idx))
    return dist2, idx


def three_interpolate(points, idx, weight):
    """points (B,C,m), idx/weight (B,n,3) -> (B,C,n)   [interpolate.cpp:47-74]"""
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), np.float32)
    lib().orc_three_interpolate(B, C, m, n, _p(points), _p(idx), _p(weight), _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """True scatter-add gradient (interpolate_gpu.cu:121-148), not the reference host bug."""
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    B, C, n = grad_out.shape
    out = np.zeros((B, C, m), np.float32)
    lib().orc_three_interpolate_grad(B, C, n, int(m), _p(grad_out), _p(idx), _p(weight), _p(out))
    return out


# ---- iou3d_nms_cuda surface ----------------------------------------------------------------------

def _pairwise(fn, boxes_a, boxes_b):
    boxes_a, boxes_b = _f32(boxes_a), _f32(boxes_b)
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    out = np.zeros((boxes_a.shape[0], boxes_b.shape[0]), np.float32)
    fn(boxes_a.shape[0], _p(boxes_a), boxes_b.shape[0], _p(boxes_b), _p(out))
    return out


def boxes_overlap_bev(boxes_a, boxes_b):
    """iou3d_nms.cpp:49-68"""
    return _pairwise(lib().orc_boxes_overlap_bev, boxes_a, boxes_b)


def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_nms.cpp:70-88"""
    return _pairwise(lib().orc_boxes_iou_bev, boxes_a, boxes_b)


def boxes_iou3d(boxes_a, boxes_b):
    """iou3d_nms_utils.py:48-81 (BEV overlap kernel + torch height/volume epilogue)"""
    return _pairwise(lib().orc_boxes_iou3d, boxes_a, boxes_b)


def nms(boxes_sorted, thresh, normal=False):
    """iou3d_nms.cpp:90-138 / :141-190 on score-sorted boxes -> kept positions (int32)."""
    boxes_sorted = _f32(boxes_sorted)
    n = boxes_sorted.shape[0]
    keep = np.zeros((max(n, 1),), np.int32)
    num = lib().orc_nms(n, _p(boxes_sorted), float(thresh), 1 if normal else 0, _p(keep))
    return keep[:num].copy()


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, normal=False):
    """iou3d_nms_utils.py:84-116: sort by score (descending), NMS, map back."""
    boxes = _f32(boxes)
    scores = _f32(scores)
    # torch.sort(descending) is not guaranteed stable; tests use distinct scores.
    order = np.argsort(-scores, kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    keep = nms(boxes[order], thresh, normal=normal)
    return order[keep].astype(np.int64)


# ---- axis-aligned box suppression -- the reference here is numpy itself -----------------
def aabb_suppress(boxes, thresh, use_cls=False, lhs=False, old_type=False, valid=None):
    """Restatement of utils/nms.py:84-122 (nms_3d_faster), :125-165 (nms_3d_faster_samecls; use_cls), :168-215
    (lhs_3d_faster_samecls; use_cls + lhs) and, with z1 = 0 / z2 = 1 columns, :52-81 (nms_2d_faster).

    boxes (K,8) [x1,y1,z1,x2,y2,z2,score,class] -> the reference's `pick` list (indices into `boxes`).
    `valid` (K) plays nonempty_box_mask (models/ap_helper.py:198-201): masked rows are dropped first and the picks
    are mapped back.  float64 throughout, numpy's operation order; equal scores are ordered by index (stable sort),
    the one thing numpy.argsort's default leaves unspecified.  Formulated as a full K x K relation + one sweep (the
    device kernel's structure) instead of the reference's shrinking-array loop."""
    bx = np.asarray(boxes, np.float64)
    keep = np.arange(bx.shape[0]) if valid is None else np.nonzero(np.asarray(valid) != 0)[0]
    bx = bx[keep]
    n = bx.shape[0]
    lo, hi, score, cls = bx[:, 0:3], bx[:, 3:6], bx[:, 6], bx[:, 7]
    ext = hi - lo
    vol = ext[:, 0] * ext[:, 1] * ext[:, 2]
    if lhs:
        vol = vol + 1e-8
    order = np.argsort(score, kind="stable")                    # ascending; NaN last
    with np.errstate(all="ignore"):
        d = np.maximum(0, np.minimum(hi[:, None, :], hi[None, :, :]) - np.maximum(lo[:, None, :], lo[None, :, :]))
        inter = d[..., 0] * d[..., 1] * d[..., 2]
        ov = inter / vol[None, :] if old_type else inter / (vol[:, None] + vol[None, :] - inter)   # [i picks, j tested]
        if use_cls:
            ov = ov * (cls[:, None] == cls[None, :])
        rel = ov > thresh
    alive = np.ones(n, bool)
    pick = []
    for p in range(n - 1, -1, -1):
        i = order[p]
        if not alive[i]:
            continue
        alive[i] = False
        pick.append(int(i))
        lower = order[:p]
        hit = lower[alive[lower] & rel[i, lower]]                # ascending score
        alive[hit] = False
        if lhs:
            pick.extend(int(j) for j in hit[::-1][: len(hit) // 2])
    return [int(keep[i]) for i in pick]


def box_extents(center, size, heading):
    """predictions2corners3d + the min/max loops (models/ap_helper.py:76-93,187-197; utils/box_util.py:266-272,335-358):
    center (K,3) f32 upright-depth, size (K,3) f64 (l,w,h), heading (K) f64 -> corners (K,8,3) f32 upright-camera,
    extents (K,6) f32."""
    center = np.asarray(center, np.float32)
    size = np.asarray(size, np.float64)
    heading = np.asarray(heading, np.float64)
    cam = np.stack([center[:, 0], -center[:, 2], center[:, 1]], 1)          # flip_axis_to_camera, float32
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1], np.float64)
    sy = np.array([1, 1, 1, 1, -1, -1, -1, -1], np.float64)
    sz = np.array([1, -1, -1, 1, 1, -1, -1, 1], np.float64)
    c, s = np.cos(heading)[:, None], np.sin(heading)[:, None]
    xc, yc, zc = sx * (size[:, 0:1] / 2), sy * (size[:, 2:3] / 2), sz * (size[:, 1:2] / 2)
    with np.errstate(all="ignore"):
        rx = (c * xc + 0.0 * yc) + s * zc
        ry = (0.0 * xc + 1.0 * yc) + 0.0 * zc
        rz = (-s * xc + 0.0 * yc) + c * zc
        corners = np.stack([rx + cam[:, 0:1], ry + cam[:, 1:2], rz + cam[:, 2:3]], -1).astype(np.float32)
    return corners, np.concatenate([corners.min(1), corners.max(1)], 1)
