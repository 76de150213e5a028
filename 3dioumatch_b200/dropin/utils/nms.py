"""Drop-in for the reference's `utils/nms.py` (:52-215) on the device kernels of libb200pc.so.

Same four public functions, argument orders and return values as the reference -- numpy (n,5)/(n,7)/(n,8) boxes in,
Python list `pick` out -- so `models/ap_helper.py:23` and `models/loss_helper_unlabeled.py:14` import it unchanged
(`utils` is a namespace package in the reference: with this tree first on sys.path only `utils.nms` is replaced).
The reference runs these loops on the host after copying every head output back; the `*_batch` functions below take
device tensors for a whole batch, never synchronise, and are what a GPU-resident caller should use.

There is no CPU fallback: the functions need a CUDA device and libb200pc.so.
"""
import ctypes

import numpy as np
import torch

from _b200_bridge import cabi, stream_ptr

_L = cabi.lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def suppress_batch(boxes, overlap_threshold, use_cls=False, lhs=False, old_type=False, valid=None):
    """boxes (B,K,8) CUDA tensor [x1,y1,z1,x2,y2,z2,score,class] (any float dtype; evaluated in float64 like numpy),
    valid (B,K) bool/uint8 or None (the reference's nonempty_box_mask).
    Returns pick (B,K) int32 (the reference's pick list per scene, -1 padded), num_pick (B) int32,
    picked (B,K) bool (= pred_mask of models/ap_helper.py:201).  No host synchronisation."""
    if not boxes.is_cuda:
        raise RuntimeError("boxes must be a CUDA tensor")
    if boxes.dim() != 3 or boxes.size(2) != 8:
        raise RuntimeError("boxes must be (B, K, 8) [x1, y1, z1, x2, y2, z2, score, class]")
    B, K, _ = boxes.shape
    b64 = boxes.detach().to(torch.float64).contiguous()
    v8 = None
    if valid is not None:
        if tuple(valid.shape) != (B, K):
            raise RuntimeError("valid must be (B, K)")
        v8 = valid.to(device=boxes.device, dtype=torch.uint8).contiguous()
    pick = torch.empty((B, K), dtype=torch.int32, device=boxes.device)
    num = torch.empty((B,), dtype=torch.int32, device=boxes.device)
    picked = torch.empty((B, K), dtype=torch.uint8, device=boxes.device)
    with torch.cuda.device(boxes.device):
        cabi.check(_L().b200nms_aabb_suppress(B, K, int(bool(use_cls)), int(bool(lhs)), int(bool(old_type)),
                                              float(overlap_threshold), _p(b64), _p(v8), _p(pick), _p(num), _p(picked),
                                              stream_ptr()), "aabb_suppress")
    return pick, num, picked.bool()


def box_extents_batch(center, size, heading, return_corners=True):
    """predictions2corners3d + the per-box min/max loops of the reference (models/ap_helper.py:76-93,187-197) for a
    batch: center (B,K,3) upright-depth, size (B,K,3) = class2size(...), heading (B,K) = class2angle(...), CUDA
    tensors.  Returns (corners (B,K,8,3) f32 upright-camera or None, extents (B,K,6) f32 [min xyz, max xyz])."""
    if not center.is_cuda:
        raise RuntimeError("center must be a CUDA tensor")
    B, K, _ = center.shape
    c32 = center.detach().to(torch.float32).contiguous()
    s64 = size.detach().to(device=center.device, dtype=torch.float64).contiguous()
    h64 = heading.detach().to(device=center.device, dtype=torch.float64).contiguous()
    if tuple(s64.shape) != (B, K, 3) or tuple(h64.shape) != (B, K):
        raise RuntimeError("size must be (B, K, 3) and heading (B, K)")
    corners = torch.empty((B, K, 8, 3), dtype=torch.float32, device=center.device) if return_corners else None
    extents = torch.empty((B, K, 6), dtype=torch.float32, device=center.device)
    with torch.cuda.device(center.device):
        cabi.check(_L().b200nms_box_extents(B, K, _p(c32), _p(s64), _p(h64), _p(corners), _p(extents), stream_ptr()),
                   "box_extents")
    return corners, extents


def _single(boxes8, overlap_threshold, use_cls, lhs, old_type):
    """numpy (n,8) -> Python list, through the device (the reference's calling convention)."""
    n = boxes8.shape[0]
    if n == 0:
        return []
    t = torch.from_numpy(np.ascontiguousarray(boxes8, dtype=np.float64)).cuda().unsqueeze(0)
    pick, num, _ = suppress_batch(t, overlap_threshold, use_cls, lhs, old_type)
    return pick[0, : int(num[0])].cpu().tolist()


def nms_2d_faster(boxes, overlap_threshold, old_type=False):
    """boxes (n,5) [x1,y1,x2,y2,score] (reference :52-81)."""
    b = np.asarray(boxes, np.float64)
    b8 = np.zeros((b.shape[0], 8))
    b8[:, 0], b8[:, 1], b8[:, 3], b8[:, 4], b8[:, 6] = b[:, 0], b[:, 1], b[:, 2], b[:, 3], b[:, 4]
    b8[:, 5] = 1.0  # unit height: areas and intersections are unchanged
    return _single(b8, overlap_threshold, False, False, old_type)


def nms_3d_faster(boxes, overlap_threshold, old_type=False):
    """boxes (n,7) [x1,y1,z1,x2,y2,z2,score] (reference :84-122)."""
    b = np.asarray(boxes, np.float64)
    b8 = np.zeros((b.shape[0], 8))
    b8[:, :7] = b[:, :7]
    return _single(b8, overlap_threshold, False, False, old_type)


def nms_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """boxes (n,8) [x1,y1,z1,x2,y2,z2,score,class]; only same-class boxes suppress each other (reference :125-165)."""
    return _single(np.asarray(boxes, np.float64)[:, :8], overlap_threshold, True, False, old_type)


def lhs_3d_faster_samecls(boxes, overlap_threshold, old_type=False):
    """Lower-half suppression (reference :168-215): every pick also keeps the better half of what it suppresses."""
    return _single(np.asarray(boxes, np.float64)[:, :8], overlap_threshold, True, True, old_type)
