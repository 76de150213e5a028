"""Python operator surface of OpenPCDet/pcdet/ops/iou3d_nms/iou3d_nms_utils.py (:12-116) on the sm_100a kernels.

Same functions, arguments and return values: boxes_bev_iou_cpu, boxes_iou_bev, boxes_iou3d_gpu, nms_gpu,
nms_normal_gpu.  boxes_iou3d_gpu runs as ONE kernel (BEV overlap + height overlap + volume ratio) instead of one
kernel + ten torch elementwise launches; utils/box_util.py:140-149 (box3d_iou_batch_gpu) calls it unchanged.
"""
import torch

from ...utils import common_utils
from . import iou3d_nms_cuda


def boxes_bev_iou_cpu(boxes_a, boxes_b):
    """(N,7), (M,7) CPU tensors or numpy arrays -> (N,M) BEV IoU."""
    boxes_a, is_numpy = common_utils.check_numpy_to_torch(boxes_a)
    boxes_b, is_numpy = common_utils.check_numpy_to_torch(boxes_b)
    assert not (boxes_a.is_cuda or boxes_b.is_cuda), 'Only support CPU tensors'
    assert boxes_a.shape[1] == 7 and boxes_b.shape[1] == 7
    ans_iou = boxes_a.new_zeros(torch.Size((boxes_a.shape[0], boxes_b.shape[0])))
    iou3d_nms_cuda.boxes_iou_bev_cpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou.numpy() if is_numpy else ans_iou


def boxes_iou_bev(boxes_a, boxes_b):
    """(N,7), (M,7) CUDA -> (N,M) rotated BEV IoU."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    ans_iou = torch.empty((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou_bev_gpu(boxes_a.contiguous(), boxes_b.contiguous(), ans_iou)
    return ans_iou


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7), (M,7) CUDA [x, y, z, dx, dy, dz, heading] -> (N,M) 3D IoU."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    iou3d = torch.empty((boxes_a.shape[0], boxes_b.shape[0]), dtype=torch.float32, device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou3d_gpu_fused(boxes_a.contiguous(), boxes_b.contiguous(), iou3d)
    return iou3d


def boxes_iou3d_batched(boxes_a, boxes_b):
    """Extension: (S,K,7), (S,G,7) -> (S,K,G); the block-diagonal of the all-pairs call in loss_helper_iou.py."""
    out = torch.empty((boxes_a.shape[0], boxes_a.shape[1], boxes_b.shape[1]), dtype=torch.float32,
                      device=boxes_a.device)
    iou3d_nms_cuda.boxes_iou3d_batched(boxes_a.contiguous(), boxes_b.contiguous(), out)
    return out


def _nms(normal, boxes, scores, thresh, pre_maxsize=None):
    """Sort -> suppression mask -> greedy sweep, all on the device (b200iou_nms_device); the keep list never visits the
    host.  The only host interaction is reading the 4-byte count, because the RESULT'S SHAPE depends on it (the reference
    copies the whole N x N/64 mask to the host and sweeps it there, iou3d_nms.cpp:111-134)."""
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    keep, num = iou3d_nms_cuda.nms_device(boxes, thresh, normal=normal)
    num_out = int(num.item())
    return order[keep[:num_out].long()].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """boxes (N,7), scores (N) -> (kept indices (LongTensor, CUDA), None); 3D-IoU criterion."""
    return _nms(False, boxes, scores, thresh, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    """axis-aligned BEV IoU criterion."""
    return _nms(True, boxes, scores, thresh)
