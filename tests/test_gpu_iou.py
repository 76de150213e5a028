"""GPU: rotated BEV overlap / IoU / 3D IoU / NMS against the CPU oracle (float parity <= 1e-5, NMS keep lists exact)."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
TOL = 1e-5


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, ref, scale=1.0):
    ok = np.isfinite(ref)
    assert np.array_equal(np.isnan(got), np.isnan(ref))
    return np.abs(got[ok] - ref[ok]).max() <= TOL * scale


def test_c3_256x256(pkg, orc):
    from pcdet.ops.iou3d_nms import iou3d_nms_cuda as native
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
    a = cases.boxes(0, 256)
    b = cases.boxes(1, 256, jitter_of=a)
    ov = torch.zeros(256, 256, device="cuda")
    assert native.boxes_overlap_bev_gpu(dev(a), dev(b), ov) == 1
    ref_ov = orc.boxes_overlap_bev(a, b)
    assert close(ov.cpu().numpy(), ref_ov, max(1.0, ref_ov.max()))
    assert close(iu.boxes_iou_bev(dev(a), dev(b)).cpu().numpy(), orc.boxes_iou_bev(a, b))
    iou3d = iu.boxes_iou3d_gpu(dev(a), dev(b)).cpu().numpy()
    ref = orc.boxes_iou3d(a, b)
    assert close(iou3d, ref)
    assert (ref > 0.05).sum() >= 200                 # the case has real overlaps
    assert (ref == 0).mean() > 0.5 and np.array_equal(iou3d == 0, ref == 0)


def test_degenerate_suite(pkg, orc):
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
    d = cases.degenerate_boxes()
    assert close(iu.boxes_iou_bev(dev(d), dev(d)).cpu().numpy(), orc.boxes_iou_bev(d, d))
    assert close(iu.boxes_iou3d_gpu(dev(d), dev(d)).cpu().numpy(), orc.boxes_iou3d(d, d))
    got = iu.boxes_iou3d_gpu(dev(d[:1]), dev(d[10:11])).cpu().numpy()[0, 0]
    assert got == 0 and not np.isnan(got)           # zero-size box -> 0, not NaN


def test_ragged_and_empty(pkg, orc):
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
    a = cases.boxes(2, 37, extent=(3, 3, 1))
    b = cases.boxes(3, 101, extent=(3, 3, 1))
    assert close(iu.boxes_iou3d_gpu(dev(a), dev(b)).cpu().numpy(), orc.boxes_iou3d(a, b))
    e = torch.zeros(0, 7, device="cuda")
    assert tuple(iu.boxes_iou3d_gpu(e, dev(b)).shape) == (0, 101)
    assert tuple(iu.boxes_iou3d_gpu(dev(a), e).shape) == (37, 0)


def test_training_shape_and_block_diagonal(pkg, orc):
    """loss_helper_iou.py:95-111: (B*K) x (B*64) all pairs, of which only the diagonal blocks are consumed."""
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
    S, K, G = 4, 256, 64
    gt = np.stack([cases.boxes(10 + s, G, extent=(6, 6, 2)) for s in range(S)])
    gt[:, 40:, :3] = -1000                                    # padded GT slots
    pred = np.stack([cases.boxes(20 + s, K, extent=(6, 6, 2)) for s in range(S)])
    full = iu.boxes_iou3d_gpu(dev(pred.reshape(-1, 7)), dev(gt.reshape(-1, 7))).cpu().numpy()
    blk = iu.boxes_iou3d_batched(dev(pred), dev(gt)).cpu().numpy()
    for s in range(S):
        assert np.array_equal(blk[s], full[s * K:(s + 1) * K, s * G:(s + 1) * G])
        assert close(blk[s], orc.boxes_iou3d(pred[s], gt[s]))


@pytest.mark.parametrize("n", [1, 63, 64, 65, 256, 300, 1000])
def test_nms(pkg, orc, n):
    from pcdet.ops.iou3d_nms import iou3d_nms_cuda as native
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
    boxes = cases.boxes(30 + n, n, extent=(4, 4, 1))
    scores = np.random.default_rng(n).permutation(n).astype(np.float32)   # distinct scores
    for thr in (0.25, 0.05):
        kept, _ = iu.nms_gpu(dev(boxes), dev(scores), thr)
        assert kept.dtype == torch.int64 and kept.is_cuda
        assert np.array_equal(kept.cpu().numpy(), orc.nms_gpu(boxes, scores, thr))
        kept_n, _ = iu.nms_normal_gpu(dev(boxes), dev(scores), thr)
        assert np.array_equal(kept_n.cpu().numpy(), orc.nms_gpu(boxes, scores, thr, normal=True))
    # reference-native calling convention: sorted boxes, CPU int32 keep, returns the count
    order = np.argsort(-scores, kind="stable")
    keep = torch.zeros(n, dtype=torch.int32)
    num = native.nms_gpu(dev(boxes[order]), keep, 0.25)
    assert np.array_equal(keep[:num].numpy(), orc.nms(boxes[order], 0.25))
    kd, nd = native.nms_device(dev(boxes[order]), 0.25)
    assert np.array_equal(kd[:int(nd.item())].cpu().numpy(), orc.nms(boxes[order], 0.25))


def test_nms_pre_maxsize(pkg, orc):
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
    boxes = cases.boxes(77, 500, extent=(4, 4, 1))
    scores = np.random.default_rng(7).permutation(500).astype(np.float32)
    kept, _ = iu.nms_gpu(dev(boxes), dev(scores), 0.1, pre_maxsize=128)
    assert np.array_equal(kept.cpu().numpy(), orc.nms_gpu(boxes, scores, 0.1, pre_maxsize=128))
