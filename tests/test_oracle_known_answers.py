"""CPU: the oracle against analytic / hand-derived known answers."""
import numpy as np

import cases


def box(x, y, z, dx, dy, dz, h):
    return np.asarray([[x, y, z, dx, dy, dz, h]], np.float32)


def test_iou_analytic(orc):
    sq = box(0, 0, 0, 2, 2, 1, 0)
    assert orc.boxes_iou_bev(sq, sq)[0, 0] == np.float32(1.0)
    assert abs(orc.boxes_iou_bev(sq, box(1, 0, 0, 2, 2, 1, 0))[0, 0] - 1 / 3) < 1e-6
    # 2x2 at 45 degrees vs 0: octagon area 8(sqrt2-1); IoU = A/(8-A) = 0.70710678
    a = 8 * (np.sqrt(2) - 1)
    assert abs(orc.boxes_iou_bev(sq, box(0, 0, 0, 2, 2, 1, np.pi / 4))[0, 0] - a / (8 - a)) < 1e-5
    assert orc.boxes_iou_bev(sq, box(5, 5, 0, 2, 2, 1, 0))[0, 0] == 0
    z = orc.boxes_iou3d(sq, box(0, 0, 0, 0, 0, 0, 0))[0, 0]
    assert z == 0 and not np.isnan(z)
    # 3D: identical footprint, half the height overlap -> (4*0.5)/(4+4-2)
    assert abs(orc.boxes_iou3d(sq, box(0, 0, 0.5, 2, 2, 1, 0))[0, 0] - 2 / 6) < 1e-6
    assert orc.boxes_iou3d(sq, box(0, 0, 2, 2, 2, 1, 0))[0, 0] == 0


def test_iou_polygon_literal(orc):
    """utils/box_util.py:553-556 demo: square (0,0)-(300,300) clipped by the diamond (150,150),(300,300),(150,450),
    (0,300) has area 22500 -- expressed as rotated rectangles: the diamond is a 45-degree square of side 150*sqrt2."""
    sq = box(150, 150, 0, 300, 300, 1, 0)
    side = 150 * np.sqrt(2)
    diamond = box(150, 300, 0, side, side, 1, np.pi / 4)
    assert abs(orc.boxes_overlap_bev(sq, diamond)[0, 0] - 22500) < 0.5


def test_iou_margin_quirk(orc):
    """Boxes separated by less than the 1e-2 corner margin still collect 4 corner vertices (iou3d_nms_kernel.cu:54):
    the reference reports a small positive overlap, and so must every implementation compared with it."""
    d = cases.degenerate_boxes()
    ov = orc.boxes_overlap_bev(d[0:1], d[4:5])[0, 0]
    assert 0 < ov < 0.05
    assert orc.boxes_overlap_bev(d[0:1], d[5:6])[0, 0] == 0


def test_opt_n_threads(orc):
    assert [orc.opt_n_threads(n) for n in (1, 2, 3, 9, 511, 512, 513, 40000)] == [1, 2, 2, 8, 256, 512, 512, 512]


def test_fps_basics(orc):
    xyz = cases.cloud(0, 2, 700, centre=False)
    idx = orc.furthest_point_sampling(xyz, 64)
    assert idx.shape == (2, 64) and idx.dtype == np.int32
    assert (idx[:, 0] == 0).all()
    for b in range(2):
        assert len(set(idx[b].tolist())) == 64  # distinct while points remain
    # greedy property: the second pick is the point furthest from point 0
    d = ((xyz[0] - xyz[0, 0]) ** 2).sum(1)
    assert idx[0, 1] == int(np.argmax(d))


def test_fps_tie_order_is_bit_reversed(orc):
    """Two exact duplicates of the furthest point at indices 130 and 300 (N >= 512 -> bs = 512): the reference's
    tree keeps the LEFT operand on ties, so the thread with the smaller bit-reversed id wins: rev9(130)=130>>... ->
    300 wins."""
    rng = np.random.default_rng(0)
    xyz = (rng.random((1, 600, 3), dtype=np.float32) * 0.5 + 1.0).astype(np.float32)
    far = np.asarray([50.0, 50.0, 50.0], np.float32)
    xyz[0, 130] = far
    xyz[0, 300] = far

    def rev9(t):
        return int(format(t, "09b")[::-1], 2)

    idx = orc.furthest_point_sampling(xyz, 2)
    expect = 130 if rev9(130) < rev9(300) else 300
    assert expect == 300
    assert idx[0, 1] == expect
    # same residue mod 512 (k and k+512): the lower index wins
    xyz2 = (rng.random((1, 700, 3), dtype=np.float32) * 0.5 + 1.0).astype(np.float32)
    xyz2[0, 40] = far
    xyz2[0, 552] = far
    assert orc.furthest_point_sampling(xyz2, 2)[0, 1] == 40


def test_fps_origin_skip(orc):
    """Points with x^2+y^2+z^2 <= 1e-3 never compete (sampling_gpu.cu:105-106); idx[0] is 0 regardless."""
    xyz = cases.cloud(5, 1, 600, centre=False) + 1.0
    xyz[0, 0] = 0.0          # start point inside the skip sphere: still selected first
    xyz[0, 77] = [0.01, 0.01, 0.01]
    xyz[0, 78] = [100, 100, 100]
    idx = orc.furthest_point_sampling(xyz.astype(np.float32), 600)
    assert idx[0, 0] == 0 and idx[0, 1] == 78
    assert 77 not in idx[0, 1:].tolist()
    # once every candidate is exhausted the selection degenerates to ties at 0 distance, never to a skipped point
    allskip = np.zeros((1, 40, 3), np.float32)
    assert (orc.furthest_point_sampling(allskip, 5) == 0).all()


def test_ball_query_semantics(orc):
    xyz = np.zeros((1, 10, 3), np.float32)
    xyz[0, :, 0] = np.arange(10) * 0.1          # points on a line, 0.1 apart
    centres = np.asarray([[[0.45, 0, 0], [5, 5, 5], [0.0, 0, 0]]], np.float32)
    idx = orc.ball_query(centres, xyz, 0.16, 4)
    assert idx[0, 0].tolist() == [3, 4, 5, 6]    # first nsample hits in index order (|dx| < 0.16)
    assert idx[0, 1].tolist() == [0, 0, 0, 0]    # empty ball keeps the zero-initialised row
    assert idx[0, 2].tolist() == [0, 1, 0, 0]    # two hits, padded with the first hit
    idx8 = orc.ball_query(centres, xyz, 10.0, 3)
    assert idx8[0, 0].tolist() == [0, 1, 2]      # more neighbours than nsample: stop early


def test_three_nn_and_interpolate(orc):
    known = np.asarray([[[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3], [1, 0, 0]]], np.float32)
    unknown = np.asarray([[[0.9, 0, 0], [0, 0, 0]]], np.float32)
    d2, idx = orc.three_nn(unknown, known)
    assert idx[0, 0].tolist() == [1, 4, 0]       # equal distances keep the earlier index in the better slot
    assert np.allclose(d2[0, 0], [0.01, 0.01, 0.81], atol=1e-6)
    assert idx[0, 1].tolist() == [0, 1, 4]
    # fewer than 3 known points: unused slots stay (+inf, 0)   (interpolate_gpu.cu:32-33,56-58)
    d2s, idxs = orc.three_nn(unknown, known[:, :2])
    assert np.isinf(d2s[0, 0, 2]) and idxs[0, 0, 2] == 0
    feats = np.arange(10, dtype=np.float32).reshape(1, 2, 5)
    w = np.asarray([[[0.5, 0.25, 0.25], [1, 0, 0]]], np.float32)
    out = orc.three_interpolate(feats, idx, w)
    assert np.allclose(out[0, 0], [0.5 * 1 + 0.25 * 4 + 0.25 * 0, 0.0])
    g = orc.three_interpolate_grad(np.ones((1, 2, 2), np.float32), idx, w, 5)
    assert np.allclose(g[0, 0], [0.25 + 1, 0.5, 0, 0, 0.25])


def test_group_gather_roundtrip(orc):
    rng = np.random.default_rng(0)
    pts = rng.standard_normal((2, 3, 50)).astype(np.float32)
    idx = rng.integers(0, 50, (2, 7, 4)).astype(np.int32)
    g = orc.group_points(pts, idx)
    assert g.shape == (2, 3, 7, 4) and g[1, 2, 3, 1] == pts[1, 2, idx[1, 3, 1]]
    gi = rng.integers(0, 50, (2, 9)).astype(np.int32)
    assert np.array_equal(orc.gather_points(pts, gi)[0, 1], pts[0, 1, gi[0]])
    # gradient of a gather/group is the scatter-add: <g(x), y> == <x, g^T(y)>
    y = rng.standard_normal(g.shape).astype(np.float32)
    lhs = float((g.astype(np.float64) * y).sum())
    rhs = float((pts.astype(np.float64) * orc.group_points_grad(y, idx, 50)).sum())
    assert abs(lhs - rhs) < 1e-3


def test_nms_greedy(orc):
    b = np.asarray([[0, 0, 0, 2, 2, 2, 0], [0.1, 0, 0, 2, 2, 2, 0], [5, 5, 0, 2, 2, 2, 0], [5.05, 5, 0, 2, 2, 2, 0.01],
                    [0, 0, 0, 2, 2, 2, 0.02]], np.float32)
    assert orc.nms(b, 0.5).tolist() == [0, 2]
    assert orc.nms(b, 0.99).tolist() == [0, 1, 2, 3, 4]
    assert orc.nms(b, 0.5, normal=True).tolist() == [0, 2]
    assert orc.nms(np.zeros((0, 7), np.float32), 0.5).tolist() == []


def test_module_restatement_float64_stack_agrees_with_the_literal_fp32_sequence(orc):
    """oracle/torch_ref.py evaluates conv1x1 -> BN(eval) -> ReLU in float64 (rounded once) so that the checker does not
    depend on the host's fp32 convolution.  It must be the same function as the literal fp32 torch sequence of the
    reference (pytorch_utils.py:14-39,70-123): both run here on the SA and FP shapes of the parity tests and agree within
    a fraction of the 1e-5 bar; max-pool / grouping / blending stay the fp32 ops they were."""
    import numpy as np
    import torch
    import torch_ref as tr

    import cases
    rng = np.random.default_rng(3)
    for spec, shape in (([7, 32, 64], (2, 7, 40, 16)), ([131, 128, 128, 256], (1, 131, 64, 32)), ([88, 48, 32], (2, 88, 512, 1))):
        layers = cases.mlp_params(5, spec)
        x = rng.standard_normal(shape).astype(np.float32)
        exact = tr.shared_mlp(torch.from_numpy(x), layers).numpy()
        literal = tr.shared_mlp(torch.from_numpy(x), layers, exact=False).numpy()
        assert exact.dtype == np.float32 and exact.shape == literal.shape == (shape[0], spec[-1], shape[2], shape[3])
        err = np.abs(exact - literal)
        assert (err <= 0.5 * (1e-5 + 1e-5 * np.abs(exact))).all(), float(err.max())
        assert np.array_equal(tr.shared_mlp(x, layers).numpy(), exact)          # arrays and tensors, deterministic
    # the SA / FP wrappers pass the switch through
    xyz = cases.cloud(4, 1, 300)
    feats = rng.standard_normal((1, 4, 300)).astype(np.float32)
    new_xyz = np.ascontiguousarray(xyz[:, :32])
    layers = cases.mlp_params(6, [7, 32, 32])
    a, ia = tr.sa_forward(xyz, feats, new_xyz, 0.3, 8, layers)
    b, ib = tr.sa_forward(xyz, feats, new_xyz, 0.3, 8, layers, exact=False)
    assert np.array_equal(ia, ib) and np.abs(a - b).max() <= 5e-6
