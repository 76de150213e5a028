"""CPU: the oracle pinned against the reference's own code.

(1) tests/golden/iou_bev_cpu.npz -- outputs of the reference's CPU entry boxes_iou_bev_cpu compiled unmodified from
    /root/reference (tests/golden/make_golden.py --cpu);
(2) tests/golden/ref_cuda_*.npz -- outputs of the reference's CUDA extension run on a B200 (make_golden.py --gpu);
(3) when oracle/_ref is present, the reference CPU entry is also called live.
"""
import os

import numpy as np
import pytest

import cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_iou_bev_matches_reference_cpu_golden(orc):
    g = np.load(os.path.join(GOLD, "iou_bev_cpu.npz"))
    for key in ("rand", "deg"):
        got = orc.boxes_iou_bev(g[key + "_a"], g[key + "_b"])
        ref = g[key + "_iou_bev"]
        ok = np.isfinite(ref)
        assert np.abs(got[ok] - ref[ok]).max() <= 1e-5, key
        assert np.array_equal(np.isnan(got), np.isnan(ref)), key


def test_iou_bev_matches_reference_cpu_live(orc, ref_ext):
    if ref_ext is None:
        pytest.skip("oracle/_ref not built here")
    import torch
    a = cases.boxes(10, 96)
    b = cases.boxes(11, 96, jitter_of=a, jitter=0.2)
    ans = torch.zeros((96, 96))
    ref_ext["iou3d_nms_cuda"].boxes_iou_bev_cpu(torch.from_numpy(a), torch.from_numpy(b), ans)
    got = orc.boxes_iou_bev(a, b)
    assert np.abs(got - ans.numpy()).max() <= 1e-5
    assert (ans.numpy() > 0.05).sum() > 50  # the case really exercises overlapping boxes


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "ref_cuda_pointops.npz")),
                    reason="reference-CUDA golden vectors not generated yet")
def test_pointops_match_reference_cuda_golden(orc):
    g = np.load(os.path.join(GOLD, "ref_cuda_pointops.npz"))
    names = sorted({k.rsplit("_xyz", 1)[0] for k in g.files if k.endswith("_xyz")})
    assert names
    for n in names:
        xyz = g[n + "_xyz"]
        npoint, nsample = [int(v) for v in g[n + "_cfg"]]
        radius = float(g[n + "_radius"][0])
        fps = orc.furthest_point_sampling(xyz, npoint)
        assert np.array_equal(fps, g[n + "_fps"]), "fps " + n
        new_xyz = np.take_along_axis(xyz, fps[:, :, None].astype(np.int64), 1)
        bq = orc.ball_query(new_xyz, xyz, radius, nsample)
        assert np.array_equal(bq, g[n + "_bq"]), "ball_query " + n
        d2, nn = orc.three_nn(xyz, new_xyz)
        assert np.array_equal(nn, g[n + "_nn_idx"]), "three_nn idx " + n
        assert np.array_equal(d2, g[n + "_nn_dist2"]), "three_nn dist2 " + n
        feats = g[n + "_feats"]
        known = orc.gather_points(feats, fps)
        interp = orc.three_interpolate(known, nn, g[n + "_w"])
        assert np.array_equal(interp, g[n + "_interp"]), "three_interpolate " + n
        grouped = orc.group_points(feats, bq)
        assert np.allclose(grouped.sum((2, 3)), g[n + "_grouped_sum"], rtol=1e-4, atol=1e-3), "group " + n


@pytest.mark.skipif(not os.path.exists(os.path.join(GOLD, "ref_cuda_iou.npz")),
                    reason="reference-CUDA golden vectors not generated yet")
def test_iou_nms_match_reference_cuda_golden(orc):
    g = np.load(os.path.join(GOLD, "ref_cuda_iou.npz"))
    for key in ("rand", "deg"):
        a, b = g[key + "_a"], g[key + "_b"]
        assert np.abs(orc.boxes_overlap_bev(a, b) - g[key + "_overlap"]).max() <= 1e-5 * max(1.0, g[key + "_overlap"].max())
        ref = g[key + "_iou_bev"]
        ok = np.isfinite(ref)
        assert np.abs(orc.boxes_iou_bev(a, b)[ok] - ref[ok]).max() <= 1e-5
    sb = g["nms_boxes_sorted"]
    for thr in (0.25, 0.05):
        assert np.array_equal(orc.nms(sb, thr), g["nms_keep_%g" % thr])
        assert np.array_equal(orc.nms(sb, thr, normal=True), g["nms_normal_keep_%g" % thr])
