"""GPU: device-side axis-aligned suppression (b200nms_aabb_suppress / b200nms_box_extents through the `utils.nms`
drop-in) against the golden vectors of the unmodified reference and against the numpy oracle.  Pick lists are index
work: bit-exact.  Corners: float64 math stored as float32, at most one float32 ulp from the reference."""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_aabb_nms.npz"))


@pytest.fixture(scope="module")
def nms(pkg):
    import importlib
    return importlib.import_module("utils.nms")


@pytest.mark.parametrize("name", ["k64", "k256", "k37", "k1"])
def test_reference_named_functions_match_golden(nms, name):
    b = G[name + "_boxes"]
    for thr in (0.25, 0.5):
        for old in (False, True):
            tag = "%s_t%g_o%d" % (name, thr, int(old))
            assert nms.nms_3d_faster(b[:, :7], thr, old) == G[tag + "_nms3d"].tolist(), tag
            assert nms.nms_3d_faster_samecls(b, thr, old) == G[tag + "_nms3d_cls"].tolist(), tag
            assert nms.lhs_3d_faster_samecls(b, thr, old) == G[tag + "_lhs_cls"].tolist(), tag
            assert nms.nms_2d_faster(b[:, [0, 2, 3, 5, 6]], thr, old) == G[tag + "_nms2d"].tolist(), tag


def _batch(seed, B, K, ncls, ties=False, nans=False):
    b = np.stack([cases.aabb_boxes(seed * 100 + i, K, ncls) for i in range(B)])
    rng = np.random.default_rng(seed)
    if ties:
        b[:, :, 6] = np.round(b[:, :, 6] * 8) / 8            # many equal scores
        q = K // 4
        b[:, :q] = b[:, q: 2 * q]                             # duplicate boxes (same score, same geometry)
    if nans and K >= 8:
        b[0, 3, 6] = np.nan
        b[0, 5, 0] = np.nan
        b[-1, 1, 3:6] = b[-1, 1, 0:3]                         # zero volume
        b[-1, 2, 3] = b[-1, 2, 0] - 1.0                       # negative extent
        b[-1, 6, 4] = np.inf
    valid = rng.random((B, K)) > 0.2
    valid[:, 0] = True
    return b, valid


@pytest.mark.parametrize("K,ncls,ties,nans", [(1, 1, False, False), (37, 2, False, True), (64, 18, True, False),
                                               (256, 18, False, True), (700, 5, True, True)])
def test_batch_matches_oracle(nms, orc, K, ncls, ties, nans):
    B = 5
    b, valid = _batch(K, B, K, ncls, ties, nans)
    t = torch.from_numpy(b).cuda()
    for use_cls, lhs in ((False, False), (True, False), (True, True), (False, True)):
        for old in (False, True):
            for v in (None, valid):
                pick, num, picked = nms.suppress_batch(t, 0.25, use_cls, lhs, old, None if v is None else torch.from_numpy(v).cuda())
                pick, num, picked = pick.cpu().numpy(), num.cpu().numpy(), picked.cpu().numpy()
                for i in range(B):
                    ref = orc.aabb_suppress(b[i], 0.25, use_cls, lhs, old, valid=None if v is None else v[i])
                    assert pick[i, : num[i]].tolist() == ref, (K, use_cls, lhs, old, v is not None, i)
                    assert (pick[i, num[i]:] == -1).all()
                    m = np.zeros(K, bool)
                    m[ref] = True
                    assert np.array_equal(picked[i], m)


def test_float32_input_and_thresholds(nms, orc):
    """float32 device tensors are widened exactly; a threshold that is not a float32 (0.3) is compared in float64."""
    b = cases.aabb_boxes(42, 200, 3)[None]
    t32 = torch.from_numpy(b.astype(np.float32)).cuda()
    for thr in (0.3, 0.05, 0.7):
        pick, num, _ = nms.suppress_batch(t32, thr, True, True)
        assert pick[0, : int(num[0])].cpu().tolist() == orc.aabb_suppress(b[0], thr, True, True)


def test_box_extents(nms, orc):
    c, s, h = G["corner_center"], G["corner_size"], G["corner_heading"]
    corners, ext = nms.box_extents_batch(torch.from_numpy(c).cuda()[None], torch.from_numpy(s).cuda()[None],
                                         torch.from_numpy(h).cuda()[None])
    corners, ext = corners[0].cpu().numpy(), ext[0].cpu().numpy()
    ref = G["corners"]
    assert np.all(np.abs(corners - ref) <= np.spacing(np.abs(ref)).astype(np.float32))
    assert np.array_equal(corners[:20], ref[:20])
    assert np.array_equal(ext[:, :3], corners.min(1)) and np.array_equal(ext[:, 3:], corners.max(1))
    oc, oe = orc.box_extents(c, s, h)
    assert np.all(np.abs(corners - oc) <= np.spacing(np.abs(oc)).astype(np.float32))
    _, ext_only = nms.box_extents_batch(torch.from_numpy(c).cuda()[None], torch.from_numpy(s).cuda()[None],
                                        torch.from_numpy(h).cuda()[None], return_corners=False)
    assert np.array_equal(ext_only[0].cpu().numpy(), ext)


def test_pseudo_label_filter_flow_on_device(nms, orc):
    """The reference's SSL filter (models/loss_helper_unlabeled.py:441-492) end to end: head outputs -> corners ->
    extents -> [extents, score, class] -> lower-half suppression -> pred_mask, without leaving the device."""
    rng = np.random.default_rng(3)
    B, K, ncls = 8, 64, 18
    center = (rng.random((B, K, 3)) * [6, 6, 2]).astype(np.float32)
    mean_size = rng.random((ncls, 3)) + 0.3
    size_cls = rng.integers(0, ncls, (B, K))
    size = mean_size[size_cls] + (rng.standard_normal((B, K, 3)) * 0.05).astype(np.float32)
    heading = np.zeros((B, K))                                    # ScanNet: class2angle == 0
    score = (rng.random((B, K)).astype(np.float32) * rng.random((B, K)).astype(np.float32))
    sem = rng.integers(0, ncls, (B, K))
    dc, ds, dh = (torch.from_numpy(a).cuda() for a in (center, size, heading))
    _, ext = nms.box_extents_batch(dc, ds, dh, return_corners=False)
    boxes8 = torch.cat([ext.double(), torch.from_numpy(score).cuda().double()[..., None],
                        torch.from_numpy(sem).cuda().double()[..., None]], -1)
    _, _, picked = nms.suppress_batch(boxes8, 0.25, use_cls=True, lhs=True)
    pred_mask = ~picked                                           # rows the reference zeroes in final_mask
    for i in range(B):
        _, e = orc.box_extents(center[i], size[i], heading[i])
        b8 = np.concatenate([e.astype(np.float64), score[i][:, None].astype(np.float64), sem[i][:, None].astype(np.float64)], 1)
        ref = np.ones(K, bool)
        ref[orc.aabb_suppress(b8, 0.25, True, True)] = False
        assert np.array_equal(pred_mask[i].cpu().numpy(), ref)


def test_full_size_properties_and_errors(nms):
    b, _ = _batch(9, 8, 1024, 18)
    t = torch.from_numpy(b).cuda()
    pick, num, picked = nms.suppress_batch(t, 0.25, True, False)
    pick, num, picked = pick.cpu().numpy(), num.cpu().numpy(), picked.cpu().numpy()
    for i in range(8):
        p = pick[i, : num[i]]
        assert len(np.unique(p)) == len(p) and picked[i].sum() == len(p)
        assert p[0] == np.argmax(b[i, :, 6])
        assert np.all(np.diff(b[i, p, 6]) < 0), "plain NMS picks come in descending score order"
    with pytest.raises(RuntimeError):
        nms.suppress_batch(torch.zeros((1, 1200, 8), device="cuda"), 0.25)
    with pytest.raises(RuntimeError):
        nms.suppress_batch(torch.zeros((1, 4, 7), device="cuda"), 0.25)
    assert nms.nms_3d_faster(np.zeros((0, 7)), 0.25) == []
