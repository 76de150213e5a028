"""CPU: the numpy restatement of the reference's host suppression loops (oracle.aabb_suppress / box_extents) against
golden vectors produced by the UNMODIFIED reference (tests/golden/make_golden.py --nms imports utils/nms.py,
utils/box_util.py from /root/reference; the reference cannot travel to the GPU box, the vectors do)."""
import os

import numpy as np
import pytest

import cases

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_aabb_nms.npz"))

MODES = {"nms3d": (False, False), "nms3d_cls": (True, False), "lhs_cls": (True, True)}


@pytest.mark.parametrize("name", ["k64", "k256", "k37", "k1"])
def test_suppress_matches_reference(orc, name):
    b = G[name + "_boxes"]
    assert np.array_equal(b, cases.aabb_boxes({"k64": 0, "k256": 1, "k37": 2, "k1": 3}[name], b.shape[0],
                                              {"k64": 3, "k256": 18, "k37": 1, "k1": 2}[name])), "fixture inputs drifted"
    for thr in (0.25, 0.5):
        for old in (0, 1):
            tag = "%s_t%g_o%d" % (name, thr, old)
            for mode, (use_cls, lhs) in MODES.items():
                got = orc.aabb_suppress(b, thr, use_cls, lhs, bool(old))
                assert got == G[tag + "_" + mode].tolist(), (tag, mode)
            b2 = b.copy()
            b2[:, 1], b2[:, 4] = b[:, 2], b[:, 5]   # the caller's 2-D boxes are (x, z) of the camera frame
            b2[:, 2], b2[:, 5] = 0.0, 1.0
            assert orc.aabb_suppress(b2, thr, False, False, bool(old)) == G[tag + "_nms2d"].tolist(), (tag, "nms2d")


def test_suppress_properties(orc):
    b = cases.aabb_boxes(7, 120, 4)
    plain = orc.aabb_suppress(b, 0.25, True, False)
    lhs = orc.aabb_suppress(b, 0.25, True, True)
    assert len(set(plain)) == len(plain) and len(set(lhs)) == len(lhs)
    assert set(plain) <= set(lhs), "LHS keeps every NMS pick plus half of the suppressed boxes"
    assert plain[0] == int(np.argmax(b[:, 6]))
    # threshold above 1: nothing is suppressed, picks = descending score order
    assert orc.aabb_suppress(b, 1.5, False, False) == np.argsort(-b[:, 6], kind="stable").tolist()
    # valid mask == dropping the rows first
    valid = np.arange(120) % 3 != 0
    sub = np.nonzero(valid)[0]
    assert orc.aabb_suppress(b, 0.25, True, True, valid=valid) == [int(sub[i]) for i in orc.aabb_suppress(b[sub], 0.25, True, True)]
    # equal scores: index order decides (lower index ranks lower, so the higher index is picked first)
    t = b.copy()
    t[:, 6] = 0.5
    assert orc.aabb_suppress(t, 1.5, False, False) == list(range(119, -1, -1))
    # degenerate boxes: zero volume -> 0/0 = NaN overlap never suppresses; with the LHS epsilon it is 0
    z = np.zeros((5, 8))
    z[:, 6] = np.arange(5)
    assert orc.aabb_suppress(z, 0.25, False, False) == [4, 3, 2, 1, 0]
    assert orc.aabb_suppress(z, 0.25, True, True) == [4, 3, 2, 1, 0]


def test_box_extents_matches_reference(orc):
    corners, ext = orc.box_extents(G["corner_center"], G["corner_size"], G["corner_heading"])
    ref = G["corners"]
    assert corners.dtype == np.float32 and corners.shape == ref.shape
    # float64 math rounded to float32: at most one float32 ulp apart (BLAS may fuse the 3-term dot product)
    assert np.all(np.abs(corners - ref) <= np.spacing(np.abs(ref)).astype(np.float32))
    assert np.array_equal(ext[:, :3], corners.min(1)) and np.array_equal(ext[:, 3:], corners.max(1))
    assert np.array_equal(corners[:20], ref[:20]), "heading 0 (ScanNet) involves no rounding at all"
