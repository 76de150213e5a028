"""Drop-in for the reference's native module `pcdet.ops.iou3d_nms.iou3d_nms_cuda`
(OpenPCDet/pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:11-17): same five functions and argument orders, on the
sm_100a kernels of libb200pc.so.  Caller allocates the outputs, the callee fills them in place and returns an int
(1, or num_to_keep), like the reference.  Argument errors raise RuntimeError instead of exit(-1)
(iou3d_nms.cpp:14-26).  Launches go to the current torch stream (the reference uses the legacy default stream)."""
import ctypes

import torch

from _b200_bridge import cabi, stream_ptr

_L = cabi.lib


def _check_input(t, name, cuda=True):
    if cuda and not t.is_cuda:
        raise RuntimeError("%s must be CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous tensor" % name)
    if t.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t.numel() > 0 else ctypes.c_void_p(0)


def _pairwise(fn_name, boxes_a, boxes_b, ans):
    for t, nm in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans, "ans")):
        _check_input(t, nm)
    if boxes_a.dim() != 2 or boxes_b.dim() != 2 or boxes_a.size(1) != 7 or boxes_b.size(1) != 7:
        raise RuntimeError("boxes must be (N, 7) [x, y, z, dx, dy, dz, heading]")
    na, nb = boxes_a.size(0), boxes_b.size(0)
    if ans.numel() != na * nb:
        raise RuntimeError("ans must hold (N, M) = (%d, %d) values" % (na, nb))
    with torch.cuda.device(boxes_a.device):
        cabi.check(getattr(_L(), fn_name)(na, _p(boxes_a), nb, _p(boxes_b), _p(ans), stream_ptr()), fn_name)
    return 1


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    """(N,7), (M,7) -> ans_overlap (N,M): rotated BEV overlap area   [iou3d_nms.cpp:49-68]"""
    return _pairwise("b200iou_boxes_overlap_bev", boxes_a, boxes_b, ans_overlap)


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    """(N,7), (M,7) -> ans_iou (N,M): rotated BEV IoU   [iou3d_nms.cpp:70-88]"""
    return _pairwise("b200iou_boxes_iou_bev", boxes_a, boxes_b, ans_iou)


def boxes_iou3d_gpu_fused(boxes_a, boxes_b, ans_iou):
    """Extension: the whole of iou3d_nms_utils.boxes_iou3d_gpu (:48-81) in one launch."""
    return _pairwise("b200iou_boxes_iou3d", boxes_a, boxes_b, ans_iou)


def boxes_iou3d_batched(boxes_a, boxes_b, ans_iou):
    """Extension: (S,K,7) x (S,G,7) -> (S,K,G), same-scene pairs only (models/loss_helper_iou.py:95-111)."""
    for t, nm in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        _check_input(t, nm)
    S, K, G = boxes_a.size(0), boxes_a.size(1), boxes_b.size(1)
    if boxes_b.size(0) != S or boxes_a.size(2) != 7 or boxes_b.size(2) != 7 or ans_iou.numel() != S * K * G:
        raise RuntimeError("boxes_iou3d_batched: shapes must be (S,K,7), (S,G,7), (S,K,G)")
    with torch.cuda.device(boxes_a.device):
        cabi.check(_L().b200iou_boxes_iou3d_batched(S, K, _p(boxes_a), G, _p(boxes_b), _p(ans_iou), stream_ptr()),
                   "boxes_iou3d_batched")
    return 1


def _nms(boxes, keep, thresh, mode):
    _check_input(boxes, "boxes")
    if not keep.is_contiguous():
        raise RuntimeError("keep must be contiguous tensor")
    if keep.is_cuda or keep.dtype not in (torch.int32, torch.int64):
        # the reference reads `keep` as int32 (iou3d_nms.cpp:98) while its own Python wrapper allocates a
        # LongTensor (iou3d_nms_utils.py:97); both are accepted here
        raise RuntimeError("keep must be a CPU int32 or int64 tensor")
    n = boxes.size(0)
    if keep.numel() < n:
        raise RuntimeError("keep must hold at least N entries")
    buf = keep if keep.dtype == torch.int32 else torch.empty((max(n, 1),), dtype=torch.int32)
    num = ctypes.c_int(0)
    with torch.cuda.device(boxes.device):
        cabi.check(_L().b200iou_nms(n, _p(boxes), float(thresh), mode, ctypes.c_void_p(buf.data_ptr()),
                                    ctypes.byref(num), stream_ptr()), "nms_gpu")
    if buf is not keep:
        keep[:num.value] = buf[:num.value].to(torch.int64)
    return int(num.value)


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """boxes (N,7) CUDA sorted by score, keep (N) CPU -> num_to_keep; 3D-IoU criterion   [iou3d_nms.cpp:90-138]"""
    return _nms(boxes, keep, nms_overlap_thresh, 0)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    """axis-aligned BEV IoU criterion   [iou3d_nms.cpp:141-190]"""
    return _nms(boxes, keep, nms_overlap_thresh, 1)


def nms_device(boxes, thresh, normal=False):
    """Extension: fully asynchronous NMS; returns (keep (N) int32 CUDA, num (1) int32 CUDA)."""
    _check_input(boxes, "boxes")
    n = boxes.size(0)
    dev = boxes.device
    ws = torch.empty((max(n, 1) * ((n + 63) // 64 + 1),), dtype=torch.int64, device=dev)
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=dev)
    num = torch.empty((1,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        cabi.check(_L().b200iou_nms_device(n, _p(boxes), float(thresh), 1 if normal else 0,
                                           ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(keep.data_ptr()),
                                           ctypes.c_void_p(num.data_ptr()), stream_ptr()), "nms_device")
    return keep, num


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    """CPU tensors (N,7), (M,7) -> ans_iou (N,M)   [iou3d_cpu.cpp:232-252]"""
    for t, nm in ((boxes_a, "boxes_a"), (boxes_b, "boxes_b"), (ans_iou, "ans_iou")):
        if t.is_cuda:
            raise RuntimeError("%s must be a CPU tensor" % nm)
        _check_input(t, nm, cuda=False)
    na, nb = boxes_a.size(0), boxes_b.size(0)
    cabi.check(_L().b200iou_boxes_iou_bev_cpu(na, _p(boxes_a), nb, _p(boxes_b), _p(ans_iou)), "boxes_iou_bev_cpu")
    return 1
