"""GPU: the strongest parity pin -- the UNMODIFIED reference CUDA extensions (built from /root/reference into
oracle/_ref by oracle/build_ref.py) run on the same B200, compared with (a) the CPU oracle and (b) the product.
Indices bit-exact; IoU within 1e-5."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("kw,npoint,radius,nsample", [
    (dict(seed=0, B=1, N=2000, centre=False), 128, 0.2, 32),
    (dict(seed=1, B=3, N=1531, dup_frac=0.05, origin_frac=0.02), 200, 0.35, 16),
    (dict(seed=3, B=2, N=4096, dup_frac=0.5, origin_frac=0.01), 512, 0.15, 64),
    (dict(seed=2, B=2, N=300, dup_frac=0.3), 64, 0.5, 8),
    (dict(seed=8, B=2, N=40000, dup_frac=0.05), 2048, 0.2, 64),
])
def test_pointops_oracle_product_reference_agree(pkg, orc, ref_ext, kw, npoint, radius, nsample):
    if ref_ext is None:
        pytest.skip("oracle/_ref not present")
    import pointnet2._ext as ours
    ref = ref_ext["_ext"]
    xyz = cases.cloud(**kw)
    t = dev(xyz)
    r_fps = ref.furthest_point_sampling(t, npoint)
    assert torch.equal(ours.furthest_point_sampling(t, npoint), r_fps), "product FPS != reference CUDA"
    if xyz.shape[1] <= 5000:
        assert np.array_equal(orc.furthest_point_sampling(xyz, npoint), r_fps.cpu().numpy()), "oracle FPS != reference"
    new_xyz = ref.gather_points(t.transpose(1, 2).contiguous(), r_fps).transpose(1, 2).contiguous()
    r_bq = ref.ball_query(new_xyz, t, radius, nsample)
    assert torch.equal(ours.ball_query(new_xyz, t, radius, nsample), r_bq)
    if xyz.shape[1] <= 5000:
        assert np.array_equal(orc.ball_query(new_xyz.cpu().numpy(), xyz, radius, nsample), r_bq.cpu().numpy())
    r_d2, r_nn = ref.three_nn(t, new_xyz)
    o_d2, o_nn = ours.three_nn(t, new_xyz)
    assert torch.equal(o_nn, r_nn) and torch.equal(o_d2, r_d2)
    feats = torch.randn(xyz.shape[0], 6, xyz.shape[1], device="cuda")
    assert torch.equal(ours.group_points(feats, r_bq), ref.group_points(feats, r_bq))
    w = 1.0 / (torch.sqrt(r_d2) + 1e-8)
    w = (w / w.sum(2, keepdim=True)).contiguous()
    kf = ref.gather_points(feats, r_fps)
    assert torch.equal(ours.gather_points(feats, r_fps), kf)
    assert torch.equal(ours.three_interpolate(kf, r_nn, w), ref.three_interpolate(kf, r_nn, w))


def test_iou_nms_oracle_product_reference_agree(pkg, orc, ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref not present")
    from pcdet.ops.iou3d_nms import iou3d_nms_cuda as ours
    ref = ref_ext["iou3d_nms_cuda"]
    a = cases.boxes(0, 256)
    b = cases.boxes(1, 256, jitter_of=a)
    d = cases.degenerate_boxes()
    for (x, y) in ((a, b), (d, d)):
        r_ov = torch.zeros(x.shape[0], y.shape[0], device="cuda")
        o_ov = torch.zeros_like(r_ov)
        ref.boxes_overlap_bev_gpu(dev(x), dev(y), r_ov)
        ours.boxes_overlap_bev_gpu(dev(x), dev(y), o_ov)
        torch.cuda.synchronize()
        scale = max(1.0, float(r_ov.max()))
        assert float((o_ov - r_ov).abs().max()) <= 1e-5 * scale
        assert np.abs(orc.boxes_overlap_bev(x, y) - r_ov.cpu().numpy()).max() <= 1e-5 * scale
        r_i = torch.zeros_like(r_ov)
        o_i = torch.zeros_like(r_ov)
        ref.boxes_iou_bev_gpu(dev(x), dev(y), r_i)
        ours.boxes_iou_bev_gpu(dev(x), dev(y), o_i)
        torch.cuda.synchronize()
        ok = torch.isfinite(r_i)
        assert float((o_i[ok] - r_i[ok]).abs().max()) <= 1e-5
    boxes = cases.boxes(5, 300, extent=(4, 4, 1))
    order = np.argsort(-np.random.default_rng(6).random(300).astype(np.float32), kind="stable")
    sb = dev(boxes[order])
    for thr in (0.25, 0.05):
        k_r = torch.zeros(300, dtype=torch.int32)
        k_o = torch.zeros(300, dtype=torch.int32)
        n_r = ref.nms_gpu(sb, k_r, thr)
        n_o = ours.nms_gpu(sb, k_o, thr)
        assert n_r == n_o and torch.equal(k_r[:n_r], k_o[:n_o])
        assert np.array_equal(orc.nms(boxes[order], thr), k_r[:n_r].numpy())
        n_r = ref.nms_normal_gpu(sb, k_r, thr)
        n_o = ours.nms_normal_gpu(sb, k_o, thr)
        assert n_r == n_o and torch.equal(k_r[:n_r], k_o[:n_o])
