"""GPU: the sm_100a point operators (through the C ABI / pointnet2._ext drop-in) against the CPU oracle.
Indices (FPS, ball_query, three_nn) must be bit-exact; float outputs of the data-movement ops are exact too."""
import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu

CASES = {
    # name: (cloud kwargs, npoint, radius, nsample)
    "c1": (dict(seed=0, B=1, N=2000, centre=False), 128, 0.2, 32),
    "ragged": (dict(seed=1, B=3, N=1531, dup_frac=0.05, origin_frac=0.02), 200, 0.35, 16),
    "small": (dict(seed=2, B=2, N=300, dup_frac=0.3), 64, 0.5, 8),
    "dups": (dict(seed=3, B=2, N=4096, dup_frac=0.5, origin_frac=0.01), 512, 0.15, 64),
    "tiny": (dict(seed=4, B=2, N=9, extent=(1, 1, 1)), 9, 5.0, 6),
    "mid": (dict(seed=5, B=4, N=20000, dup_frac=0.02, origin_frac=0.001), 1024, 0.3, 32),
    "npoint_gt_n": (dict(seed=6, B=1, N=40), 64, 0.5, 4),
}


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", sorted(CASES))
def test_fps_ball_query_three_nn_bit_exact(pkg, orc, name):
    import pointnet2._ext as ext
    kw, npoint, radius, nsample = CASES[name]
    xyz = cases.cloud(**kw)
    t = dev(xyz)
    fps = ext.furthest_point_sampling(t, npoint)
    ref_fps = orc.furthest_point_sampling(xyz, npoint)
    assert fps.dtype == torch.int32 and tuple(fps.shape) == ref_fps.shape
    assert np.array_equal(fps.cpu().numpy(), ref_fps), "FPS indices differ"

    new_xyz_ref = np.take_along_axis(xyz, ref_fps[:, :, None].astype(np.int64), 1)
    new_xyz = ext.gather_points(t.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    assert np.array_equal(new_xyz.cpu().numpy(), new_xyz_ref)

    bq = ext.ball_query(new_xyz, t, radius, nsample)
    assert np.array_equal(bq.cpu().numpy(), orc.ball_query(new_xyz_ref, xyz, radius, nsample)), "ball_query differs"

    d2, nn = ext.three_nn(t, new_xyz)
    rd2, rnn = orc.three_nn(xyz, new_xyz_ref)
    assert np.array_equal(nn.cpu().numpy(), rnn), "three_nn idx differs"
    assert np.array_equal(d2.cpu().numpy(), rd2), "three_nn dist2 differs"


def test_fps_every_cluster_size_agrees(pkg, orc):
    """Every kernel generation (fps_owner_kernel with direct / redux record arg-max, fps_cluster_kernel) and every
    (cluster size, threads per CTA) shape it can pick must give the reference's indices: duplicated points (exact ties,
    also across warps and CTAs), points inside the origin-skip sphere, a cloud that leaves whole CTAs without candidates."""
    import importlib
    import pointnet2._ext as ext
    cabi = importlib.import_module("3dioumatch_b200._cabi")
    clouds = [(cases.cloud(7, 2, 9000, dup_frac=0.2, origin_frac=0.01), 300),
              (cases.cloud(8, 3, 700, dup_frac=0.5, origin_frac=0.3), 700),       # npoint == N: the tail is all ties at 0
              (np.concatenate([cases.cloud(9, 2, 40, dup_frac=0.3), np.zeros((2, 4000, 3), np.float32)], 1), 64)]
    tried = 0
    try:
        for xyz, m in clouds:
            ref = orc.furthest_point_sampling(xyz, m)
            t = dev(xyz)
            for kern in (-1, 0, 2):
                for cs in (1, 2, 4, 8, 16):
                    for th in (32, 64, 128, 256, 512):
                        cabi.force_fps_shape(kern, cs, th)
                        try:
                            got = ext.furthest_point_sampling(t, m)
                        except RuntimeError as e:
                            assert "needs a scratch buffer" in str(e), str(e)  # this shape cannot hold the cloud in registers
                            continue
                        tried += 1
                        assert np.array_equal(got.cpu().numpy(), ref), "kernel=%d cluster=%d threads=%d N=%d" % (kern, cs, th, xyz.shape[1])
    finally:
        cabi.force_fps_shape(-1, 0, 0)
    assert tried >= 100


def test_fps_prefix_speculation(pkg, orc):
    """Hierarchical levels: the input is the previous level's picks in pick order, for which FPS returns 0..m-1 unless a tie
    or the origin-skip rule intervenes.  The verify-then-skip path (fps_prefix_values_kernel + fps_prefix_check_kernel,
    b200pn2_fps_set_prefix_speculation) must give the oracle's indices when the speculation holds, when it fails for some
    scenes of the batch only, and when it fails everywhere -- and the same indices as with speculation off."""
    import importlib
    import pointnet2._ext as ext
    cabi = importlib.import_module("3dioumatch_b200._cabi")
    rng = np.random.default_rng(11)
    base = cases.scene_cloud(2, 4, 12000)[:, :, :3].copy()
    lvl1 = orc.furthest_point_sampling(base, 2048)
    ordered = np.take_along_axis(base, lvl1[:, :, None].astype(np.int64), 1)          # (4, 2048, 3) in furthest-point order
    inputs = {"ordered": (ordered.copy(), 1024)}
    mixed = ordered.copy()
    mixed[1] = mixed[1][rng.permutation(2048)]                                          # scene 1: arbitrary order
    mixed[2, 700] = mixed[2, 3]                                                         # scene 2: an exact duplicate -> a tie at the end of the chain
    mixed[3, 5] = np.float32([0.01, -0.01, 0.005])                                      # scene 3: pick 5 falls inside the origin-skip sphere
    inputs["mixed"] = (mixed, 1024)
    # exact ties between a speculated pick and a point that is never picked (k >= m): the tie key decides.  N = 2048 runs
    # with 512-thread blocks in the reference (L = 9): key(10) = bitrev9(10) = 160; point 1500 (1500 % 512 = 476 ->
    # bitrev9 = 119) wins the tie against pick 10 -> the speculation must be refuted; point 1535 (511 -> 511) loses it ->
    # the speculation holds; pick 0 is never contested (scene 2: a copy of point 0)
    ties = ordered.copy()
    ties[0, 1500] = ties[0, 10]
    ties[1, 1535] = ties[1, 10]
    ties[2, 1300] = ties[2, 0]
    ties[3, 1500] = ties[3, 1023]                                                       # a tie in the last column
    inputs["ties"] = (ties, 1024)
    inputs["random"] = (cases.cloud(12, 3, 1000, dup_frac=0.1, origin_frac=0.02), 500)
    inputs["ragged"] = (ordered[:3, :1000].copy(), 300)                                 # m, N not multiples of the 128-column tile
    inputs["whole"] = (ordered[:2, :512].copy(), 512)                                   # m == N
    lvl2 = orc.furthest_point_sampling(ordered, 1024)
    assert np.array_equal(lvl2[0], np.arange(1024))                                     # the nesting property itself (oracle = reference semantics)
    prev = cabi.set_fps_prefix_speculation(-1)
    try:
        for name, (xyz, m) in inputs.items():
            ref = orc.furthest_point_sampling(xyz, m)
            t = dev(xyz)
            cabi.set_fps_prefix_speculation(1)
            n0 = cabi.launch_count()
            on = ext.furthest_point_sampling(t, m)
            assert cabi.launch_count() - n0 == 4, "head + values + check + serial kernel"
            cabi.set_fps_prefix_speculation(0)
            n0 = cabi.launch_count()
            off = ext.furthest_point_sampling(t, m)
            assert cabi.launch_count() - n0 == 1
            assert np.array_equal(on.cpu().numpy(), ref), name
            assert torch.equal(on, off), name
        # a cloud larger than the speculation limit is never speculated on
        cabi.set_fps_prefix_speculation(1)
        big = dev(base[:1, :6000].copy())
        n0 = cabi.launch_count()
        ext.furthest_point_sampling(big, 64)
        assert cabi.launch_count() - n0 == 1
    finally:
        cabi.set_fps_prefix_speculation(prev)


def test_fps_scannet_shape_full_size(pkg, orc):
    """BASELINE config: (B,N)=(8,40000) -> 2048 samples; one scene checked against the oracle, all scenes for
    size-independent properties (distinct indices, idx[0]=0, greedy max-min property on a sample of steps)."""
    import pointnet2._ext as ext
    pc = cases.scene_cloud(0, 8, 40000)[:, :, :3].copy()
    t = dev(pc)
    fps = ext.furthest_point_sampling(t, 2048).cpu().numpy()
    assert (fps[:, 0] == 0).all()
    for b in range(8):
        assert len(np.unique(fps[b])) == 2048
    assert np.array_equal(fps[3:4], orc.furthest_point_sampling(pc[3:4], 2048))
    # greedy property at step j: the chosen point maximises the distance to the already chosen set
    b = 5
    for j in (1, 2, 17, 300, 2047):
        chosen = pc[b, fps[b, :j]]
        d = ((pc[b][:, None, :] - chosen[None, :, :]) ** 2).sum(-1).min(1)
        assert d[fps[b, j]] >= d.max() * (1 - 1e-5)


def test_fps_policy_is_result_neutral(pkg):
    """b200pn2_fps_set_policy: the throughput launch shape (fewer, fuller CTAs per scene) returns the same indices as the
    latency shape, on a ScanNet-sized and a SUN RGB-D-sized batch with duplicated points (exact ties)."""
    import importlib
    import pointnet2._ext as ext
    cabi = importlib.import_module("3dioumatch_b200._cabi")
    prev = cabi.set_fps_policy("latency")
    try:
        for (B, N, m, seed) in ((3, 40000, 2048, 4), (5, 20000, 2048, 5), (2, 2048, 1024, 6)):
            pc = cases.cloud(seed, B, N, dup_frac=0.05)
            t = dev(pc)
            cabi.set_fps_policy("latency")
            a = ext.furthest_point_sampling(t, m)
            cabi.set_fps_policy("throughput")
            b = ext.furthest_point_sampling(t, m)
            assert torch.equal(a, b)
    finally:
        cabi.set_fps_policy(prev)


def test_group_gather_interpolate_and_grads(pkg, orc):
    import pointnet2._ext as ext
    rng = np.random.default_rng(0)
    B, C, N, M, ns = 3, 37, 500, 64, 16
    pts = rng.standard_normal((B, C, N)).astype(np.float32)
    idx = rng.integers(0, N, (B, M, ns)).astype(np.int32)
    g = ext.group_points(dev(pts), dev(idx))
    assert np.array_equal(g.cpu().numpy(), orc.group_points(pts, idx))
    gi = rng.integers(0, N, (B, M)).astype(np.int32)
    assert np.array_equal(ext.gather_points(dev(pts), dev(gi)).cpu().numpy(), orc.gather_points(pts, gi))
    # gradients are atomic scatter-adds: compare with tolerance (summation order is not fixed, as in the reference)
    go = rng.standard_normal((B, C, M, ns)).astype(np.float32)
    got = ext.group_points_grad(dev(go), dev(idx), N).cpu().numpy()
    assert np.allclose(got, orc.group_points_grad(go, idx, N), rtol=1e-4, atol=1e-4)
    go2 = rng.standard_normal((B, C, M)).astype(np.float32)
    got = ext.gather_points_grad(dev(go2), dev(gi), N).cpu().numpy()
    assert np.allclose(got, orc.gather_points_grad(go2, gi, N), rtol=1e-4, atol=1e-4)
    # three_interpolate: bit-exact forward (same fma sequence), toleranced backward
    n, m = 333, 77
    known = rng.standard_normal((B, C, m)).astype(np.float32)
    tidx = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = rng.random((B, n, 3)).astype(np.float32)
    w /= w.sum(2, keepdims=True)
    out = ext.three_interpolate(dev(known), dev(tidx), dev(w)).cpu().numpy()
    assert np.array_equal(out, orc.three_interpolate(known, tidx, w))
    go3 = rng.standard_normal((B, C, n)).astype(np.float32)
    got = ext.three_interpolate_grad(dev(go3), dev(tidx), dev(w), m).cpu().numpy()
    assert np.allclose(got, orc.three_interpolate_grad(go3, tidx, w, m), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("shape", [(3, 37, 500, 64, 16), (2, 9, 4000, 700, 32), (1, 5, 64, 200, 8)])
def test_deterministic_gradients_bit_exact(pkg, orc, monkeypatch, shape):
    """B200_DETERMINISTIC=1 / torch.use_deterministic_algorithms: the sorted-segment gradient kernels sum every target's
    entries in ascending entry position -- the order of the oracle's sequential loops -- so the three gradients equal the
    CPU oracle bit for bit (many duplicates per target, empty targets, heavy collisions in the last shape) and repeat."""
    import pointnet2._ext as ext
    monkeypatch.setenv("B200_DETERMINISTIC", "1")
    B, C, N, M, ns = shape
    rng = np.random.default_rng(sum(shape))
    idx = rng.integers(0, N, (B, M, ns)).astype(np.int32)
    idx[:, : M // 2] = rng.integers(0, max(N // 50, 1), (B, M // 2, ns))          # pile-ups on a few targets
    go = rng.standard_normal((B, C, M, ns)).astype(np.float32)
    a = ext.group_points_grad(dev(go), dev(idx), N)
    assert np.array_equal(a.cpu().numpy(), orc.group_points_grad(go, idx, N))
    assert torch.equal(a, ext.group_points_grad(dev(go), dev(idx), N))
    gi = rng.integers(0, N, (B, M)).astype(np.int32)
    go2 = rng.standard_normal((B, C, M)).astype(np.float32)
    assert np.array_equal(ext.gather_points_grad(dev(go2), dev(gi), N).cpu().numpy(), orc.gather_points_grad(go2, gi, N))
    n, m = M * 3 + 1, max(N // 7, 3)
    tidx = rng.integers(0, m, (B, n, 3)).astype(np.int32)
    w = rng.random((B, n, 3)).astype(np.float32)
    w /= w.sum(2, keepdims=True)
    go3 = rng.standard_normal((B, C, n)).astype(np.float32)
    got = ext.three_interpolate_grad(dev(go3), dev(tidx), dev(w), m).cpu().numpy()
    assert np.array_equal(got, orc.three_interpolate_grad(go3, tidx, w, m))


def test_deterministic_flag_reaches_autograd(pkg):
    """torch.use_deterministic_algorithms(True) routes the autograd Functions' backward through the deterministic
    kernels: two backward passes of a grouping + interpolation graph give identical gradients."""
    import pointnet2.pointnet2_utils as U
    torch.manual_seed(0)
    feats = torch.randn(2, 16, 300, device="cuda", requires_grad=True)
    idx = torch.randint(0, 12, (2, 200, 16), device="cuda", dtype=torch.int32)     # heavy collisions
    known = torch.randn(2, 16, 40, device="cuda", requires_grad=True)
    tidx = torch.randint(0, 40, (2, 300, 3), device="cuda", dtype=torch.int32)
    w = torch.rand(2, 300, 3, device="cuda")
    prev = torch.are_deterministic_algorithms_enabled()
    torch.use_deterministic_algorithms(True, warn_only=True)
    try:
        grads = []
        for _ in range(2):
            feats.grad = known.grad = None
            (U.grouping_operation(feats, idx).square().sum() + U.three_interpolate(known, tidx, w).square().sum()).backward()
            grads.append((feats.grad.clone(), known.grad.clone()))
        assert torch.equal(grads[0][0], grads[1][0]) and torch.equal(grads[0][1], grads[1][1])
    finally:
        torch.use_deterministic_algorithms(prev)


def test_autograd_functions(pkg):
    """pointnet2_utils Functions: gradients flow through gather/group/interpolate (torch.autograd.gradcheck-style
    directional check in fp32)."""
    import pointnet2.pointnet2_utils as U
    torch.manual_seed(0)
    feats = torch.randn(2, 4, 50, device="cuda", requires_grad=True)
    idx = torch.randint(0, 50, (2, 6, 3), device="cuda", dtype=torch.int32)
    out = U.grouping_operation(feats, idx)
    out.sum().backward()
    counts = torch.zeros(2, 50, device="cuda")
    counts.scatter_add_(1, idx.view(2, -1).long(), torch.ones(2, 18, device="cuda"))
    assert torch.allclose(feats.grad, counts[:, None, :].expand(2, 4, 50))
    xyz = torch.rand(2, 50, 3, device="cuda")
    inds = U.furthest_point_sample(xyz, 8)
    assert inds.dtype == torch.int32 and not inds.requires_grad
    dist, nn = U.three_nn(xyz, xyz[:, :10].contiguous())
    assert dist.shape == (2, 50, 3) and (dist[:, :10, 0] == 0).all()


def test_three_nn_gridconv_shape(pkg, orc):
    """GridConv call shape (models/grid_conv_module.py:87): K*64 grid points vs 1024 seeds; one scene vs the oracle."""
    import pointnet2._ext as ext
    rng = np.random.default_rng(1)
    seeds = (rng.random((2, 1024, 3)) * [8, 8, 3]).astype(np.float32)
    grid = (rng.random((2, 256 * 64, 3)) * [8, 8, 3]).astype(np.float32)
    d2, idx = ext.three_nn(dev(grid), dev(seeds))
    rd2, ridx = orc.three_nn(grid[:1], seeds[:1])
    assert np.array_equal(idx[:1].cpu().numpy(), ridx) and np.array_equal(d2[:1].cpu().numpy(), rd2)
    d = d2.cpu().numpy()
    assert (d[..., 0] <= d[..., 1]).all() and (d[..., 1] <= d[..., 2]).all()


@pytest.mark.parametrize("name", ["c1", "ragged", "small", "dups", "tiny", "mid"])
def test_ball_query_grid_path_bit_exact(pkg, orc, name, monkeypatch):
    """The hashed-grid search (ball_query_grid.cu, default for N >= 8192) forced on for every case, including radii far
    larger than the cloud (every bucket overflows -> exact in-warp fallback) and duplicate-heavy clouds."""
    import pointnet2._ext as ext
    monkeypatch.setenv("B200_BQ_GRID", "1")
    kw, npoint, radius, nsample = CASES[name]
    xyz = cases.cloud(**kw)
    fps = orc.furthest_point_sampling(xyz, npoint)
    new_xyz = np.take_along_axis(xyz, fps[:, :, None].astype(np.int64), 1)
    for r, ns in ((radius, nsample), (radius * 0.3, 4), (radius * 6, nsample), (1e-4, 3)):
        got = ext.ball_query(dev(new_xyz), dev(xyz), r, ns).cpu().numpy()
        assert np.array_equal(got, orc.ball_query(new_xyz, xyz, r, ns)), (name, r, ns)
    monkeypatch.setenv("B200_BQ_GRID", "0")
    got = ext.ball_query(dev(new_xyz), dev(xyz), radius, nsample).cpu().numpy()
    assert np.array_equal(got, orc.ball_query(new_xyz, xyz, radius, nsample))


def test_ball_query_grid_far_coordinates(pkg, orc, monkeypatch):
    """Clouds far from the origin (cell indices beyond the fp32-safe range take the exact scan) and negative coordinates."""
    import pointnet2._ext as ext
    monkeypatch.setenv("B200_BQ_GRID", "1")
    for shift in (-37.5, 1.0e3, 3.0e5):
        xyz = cases.cloud(11, 2, 3000) + np.float32(shift)
        new_xyz = xyz[:, :200].copy()
        got = ext.ball_query(dev(new_xyz), dev(xyz), 0.25, 16).cpu().numpy()
        assert np.array_equal(got, orc.ball_query(new_xyz, xyz, 0.25, 16)), shift
