"""GPU: the fused set-abstraction forward (ball query -> group -> SharedMLP -> max on chip) against the restatement of
the reference module math (oracle/torch_ref.py: fp32 grouping with the reference's rounding, the conv/BN/ReLU stack
evaluated in float64 and rounded once), tolerance 1e-5 (abs + rel) on every element, and the module-level drop-in
against its own unfused path."""
import os

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-5, 1e-5


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def assert_close(got, ref):
    err = np.abs(got - ref)
    bound = ATOL + RTOL * np.abs(ref)
    assert (err <= bound).all(), "max err %.3g at ref %.3g" % (err.max(), np.abs(ref).flat[err.argmax()])


def run_case(orc, tr, B, N, M, C, radius, ns, spec, seed, use_xyz=True, normalize=True, dup=0.0):
    import pointnet2._ext as ext
    xyz = cases.cloud(seed, B, N, dup_frac=dup)
    rng = np.random.default_rng(seed + 1)
    feats = rng.standard_normal((B, C, N)).astype(np.float32) if C else None
    fps = orc.furthest_point_sampling(xyz, M)
    new_xyz = np.take_along_axis(xyz, fps[:, :, None].astype(np.int64), 1)
    full_spec = [(3 if use_xyz else 0) + C] + list(spec)
    layers = cases.mlp_params(seed + 2, full_spec)
    ref, ref_idx = tr.sa_forward(xyz, feats, new_xyz, radius, ns, layers, use_xyz=use_xyz, normalize_xyz=normalize)
    trip = [(dev(w), dev(s), dev(h)) for w, s, h in tr.fold(layers)]
    out, out_pm, idx = ext.sa_forward(dev(xyz), dev(feats), dev(new_xyz), radius, ns, trip, use_xyz=use_xyz,
                                      normalize_xyz=normalize, want_idx=True, want_pm=True)
    assert np.array_equal(idx.cpu().numpy(), ref_idx), "fused ball query differs"
    assert_close(out.cpu().numpy(), ref)
    assert torch.equal(out_pm, out.transpose(1, 2))
    return out


@pytest.fixture(scope="module")
def tr(orc):
    import torch_ref
    return torch_ref


def test_c1_config(orc, tr, pkg):
    """BASELINE configs[0]: one 2000-point cloud, npoint=128, r=0.2, ns=32, mlp=[4,32] (1 feature + xyz)."""
    import pointnet2._ext as ext
    xyz = cases.cloud(0, 1, 2000, centre=False)
    feats = np.random.default_rng(1).random((1, 1, 2000)).astype(np.float32)
    fps = orc.furthest_point_sampling(xyz, 128)
    new_xyz = np.take_along_axis(xyz, fps[:, :, None].astype(np.int64), 1)
    layers = cases.mlp_params(2, [4, 32])
    ref, ref_idx = tr.sa_forward(xyz, feats, new_xyz, 0.2, 32, layers, normalize_xyz=False)
    trip = [(dev(w), dev(s), dev(h)) for w, s, h in tr.fold(layers)]
    out, _, idx = ext.sa_forward(dev(xyz), dev(feats), dev(new_xyz), 0.2, 32, trip, want_idx=True)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert_close(out.cpu().numpy(), ref)


@pytest.mark.parametrize("shape", [
    # (B, N, M, C, radius, ns, mlp)   reduced-size versions of SA1..SA4 / vote aggregation (backbone_module.py:35-69)
    (2, 6000, 256, 1, 0.2, 64, [64, 64, 128]),
    (2, 2048, 200, 128, 0.4, 32, [128, 128, 256]),
    (2, 1024, 96, 256, 0.8, 16, [128, 128, 256]),
    (3, 512, 61, 256, 1.2, 16, [128, 128, 128]),
    (1, 700, 50, 5, 0.5, 24, [20, 33]),          # odd widths: padding paths, nsample not dividing the tile
    (1, 700, 50, 8, 0.5, 100, [16]),             # nsample close to the tile height
    (2, 300, 40, 0, 0.6, 8, [16, 16, 16, 24]),   # xyz only, 4 layers
])
def test_sa_shapes(orc, tr, pkg, shape):
    B, N, M, C, r, ns, spec = shape
    run_case(orc, tr, B, N, M, C, r, ns, spec, seed=B * 1000 + N, dup=0.05)


@pytest.mark.parametrize("shape", [
    # more tiles than SMs: every persistent CTA walks several tiles (tile queue, weight-ring wrap-around, both TMEM
    # accumulator sets, layer-1 stage parity across tiles)
    (2, 9000, 1024, 1, 0.25, 64, [64, 64, 128]),     # SA1-like: 1024 tiles, odd number of layer-1 k-blocks
    (2, 2048, 1024, 128, 0.4, 32, [128, 128, 256]),   # SA2-like: 512 tiles, 256-wide last layer (two halves)
    (3, 1024, 700, 256, 0.8, 16, [128, 128, 128]),    # vote-aggregation-like: 264 tiles, last tile of a scene partial
    (2, 1500, 900, 61, 0.5, 8, [32, 64]),             # nsample 8, unaligned features (scalar gather), 2 layers
])
def test_sa_many_tiles_per_cta(orc, tr, pkg, shape):
    B, N, M, C, r, ns, spec = shape
    run_case(orc, tr, B, N, M, C, r, ns, spec, seed=B * 77 + M, dup=0.02)


@pytest.mark.parametrize("factor", ["0", "1"])
@pytest.mark.parametrize("shape", [
    (2, 2048, 1024, 128, 0.4, 32, [128, 128, 256]),   # SA2-like (compacted tiles)
    (3, 1024, 333, 256, 0.8, 16, [128, 128, 128]),    # vote-aggregation-like, ragged last tile
    (2, 777, 128, 64, 0.5, 16, [96, 64, 32]),         # narrow widths, point count not a multiple of the 128-row GEMM tile
    (2, 6000, 700, 1, 0.2, 64, [64, 64, 128]),        # SA1-like: <= 4 raw channels -> layer 1 evaluated in the gather
    (2, 1500, 300, 3, 0.3, 16, [64, 64, 128]),        # three raw channels, uncompacted tiles
])
def test_sa_first_layer_factorised_and_not(orc, tr, pkg, monkeypatch, shape, factor):
    """B200_SA_TC_FACTOR: layer 1 as a per-point row GEMM + xyz FMAs in the gather (>= 32 feature channels) or entirely
    in the gather (<= 4 channels) vs one GEMM per grouped row; both meet the same 1e-5 bar against the fp32 reference,
    with and without relative-xyz channels."""
    monkeypatch.setenv("B200_SA_TC_FACTOR", factor)
    B, N, M, C, r, ns, spec = shape
    run_case(orc, tr, B, N, M, C, r, ns, spec, seed=B * 31 + M, dup=0.02)
    run_case(orc, tr, B, N, M, C, r, ns, spec, seed=B * 31 + M + 1, use_xyz=False)


@pytest.mark.parametrize("shape", [
    # (B, N, M, C, radius, ns, mlp): the levels whose ball query runs inside the fused kernel's producers
    (8, 1024, 512, 256, 0.8, 16, [128, 128, 256]),    # SA3 at the bench shape
    (8, 512, 256, 256, 1.2, 16, [128, 128, 256]),     # SA4
    (8, 1024, 256, 256, 0.3, 16, [128, 128, 128]),    # vote aggregation: sparse balls, many short / empty lists
    (3, 1024, 203, 256, 0.05, 16, [128, 128, 128]),   # radius so small that most balls hold only the centre itself
    (2, 2048, 77, 8, 0.4, 8, [32, 64]),               # nsample 8: four centres per producer warp, no factorised layer
    (2, 600, 150, 64, 0.5, 32, [64, 64]),             # nsample 32 (compaction off below): one centre per warp
])
def test_ball_query_fused_into_the_gather(orc, tr, pkg, monkeypatch, shape):
    """Small levels: the tensor-core kernel's producers stage the scene in shared memory (one cp.async.bulk) and run the
    radius search themselves (B200_SA_TC_QUERY, default on).  Neighbour lists bit-exact against the oracle when asked
    for, features within 1e-5 of the fp32 reference, and bit-identical to the two-kernel path (separate ball query)."""
    import pointnet2._ext as ext
    B, N, M, C, r, ns, spec = shape
    if ns == 32:
        monkeypatch.setenv("B200_SA_TC_UNITS", "0")
    run_case(orc, tr, B, N, M, C, r, ns, spec, seed=B * 13 + M, dup=0.03)
    xyz = cases.cloud(B * 13 + M, B, N, dup_frac=0.03)
    feats = np.random.default_rng(3).standard_normal((B, C, N)).astype(np.float32)
    fps = orc.furthest_point_sampling(xyz, M)
    new_xyz = np.take_along_axis(xyz, fps[:, :, None].astype(np.int64), 1)
    trip = [(dev(w), dev(s), dev(h)) for w, s, h in tr.fold(cases.mlp_params(4, [C + 3] + list(spec)))]
    args = (dev(xyz), dev(feats), dev(new_xyz), r, ns, trip)
    cabi = __import__("importlib").import_module("3dioumatch_b200._cabi")
    n0 = cabi.launch_count()
    fused, _, _ = ext.sa_forward(*args, normalize_xyz=True)                    # no idx requested: none is written
    n_fused = cabi.launch_count() - n0
    fused_i, _, idx_f = ext.sa_forward(*args, normalize_xyz=True, want_idx=True)
    monkeypatch.setenv("B200_SA_TC_QUERY", "0")
    n0 = cabi.launch_count()
    plain, _, idx_p = ext.sa_forward(*args, normalize_xyz=True, want_idx=True)
    n_plain = cabi.launch_count() - n0
    assert torch.equal(idx_f, idx_p) and torch.equal(fused, plain) and torch.equal(fused_i, plain)
    assert n_fused == n_plain - 1                                              # the ball-query launch is gone


def test_no_xyz_and_unnormalised(orc, tr, pkg):
    run_case(orc, tr, 2, 900, 64, 12, 0.5, 16, [32, 32], seed=5, use_xyz=False, normalize=False)
    run_case(orc, tr, 2, 900, 64, 12, 0.5, 16, [32, 32], seed=6, use_xyz=True, normalize=False)


def test_idx_in_and_point_major_inputs(orc, tr, pkg):
    """Caller-provided neighbour indices and a caller-provided point-major feature copy give the same result."""
    import pointnet2._ext as ext
    B, N, M, C, r, ns = 2, 1500, 128, 64, 0.4, 32
    xyz = cases.cloud(9, B, N)
    feats = np.random.default_rng(10).standard_normal((B, C, N)).astype(np.float32)
    fps = orc.furthest_point_sampling(xyz, M)
    new_xyz = np.take_along_axis(xyz, fps[:, :, None].astype(np.int64), 1)
    layers = cases.mlp_params(11, [C + 3, 64, 64])
    trip = [(dev(w), dev(s), dev(h)) for w, s, h in tr.fold(layers)]
    a, _, idx = ext.sa_forward(dev(xyz), dev(feats), dev(new_xyz), r, ns, trip, normalize_xyz=True, want_idx=True)
    b, _, _ = ext.sa_forward(dev(xyz), dev(feats), dev(new_xyz), r, ns, trip, normalize_xyz=True, idx=idx)
    c, _, _ = ext.sa_forward(dev(xyz), None, dev(new_xyz), r, ns, trip, normalize_xyz=True,
                             features_pm=dev(feats.transpose(0, 2, 1)))
    assert torch.equal(a, b) and torch.equal(a, c)


def test_module_dropin_fused_equals_unfused(pkg, monkeypatch):
    """PointnetSAModuleVotes in eval mode: fused single-kernel path vs the op-by-op path on the same kernels + torch
    (cudnn/cublas fp32, TF32 off) -- the reference's own dataflow."""
    import pointnet2_modules as M
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1)
    net = M.PointnetSAModuleVotes(npoint=256, radius=0.3, nsample=32, mlp=[16, 64, 64, 128], use_xyz=True,
                                  normalize_xyz=True).cuda()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
    net.eval()
    xyz = torch.from_numpy(cases.cloud(3, 2, 5000)).cuda()
    feats = torch.randn(2, 16, 5000, device="cuda")
    with torch.no_grad():
        nx1, f1, i1 = net(xyz, feats)
        monkeypatch.setenv("B200_SA_FUSED", "0")
        nx2, f2, i2 = net(xyz, feats)
    assert torch.equal(i1, i2) and torch.equal(nx1, nx2)
    assert torch.allclose(f1, f2, atol=1e-5, rtol=1e-5)
    # `inds` pass-through (pointnet2_modules.py:239-242)
    monkeypatch.delenv("B200_SA_FUSED")
    with torch.no_grad():
        nx3, f3, i3 = net(xyz, feats, i1)
    assert torch.equal(i3, i1) and torch.equal(f3, f1)


def test_module_training_path_backward(pkg):
    """Training mode (batch-statistics BN) takes the unfused, differentiable path."""
    import pointnet2_modules as M
    torch.manual_seed(0)
    net = M.PointnetSAModuleVotes(npoint=64, radius=0.4, nsample=16, mlp=[8, 32, 32], use_xyz=True,
                                  normalize_xyz=True).cuda().train()
    xyz = torch.from_numpy(cases.cloud(4, 2, 1000)).cuda()
    feats = torch.randn(2, 8, 1000, device="cuda", requires_grad=True)
    _, f, _ = net(xyz, feats)
    f.square().mean().backward()
    assert feats.grad is not None and torch.isfinite(feats.grad).all() and feats.grad.abs().sum() > 0
    assert net.mlp_module.layer0.conv.weight.grad is not None


def _load_mlp(mlp, layers):
    with torch.no_grad():
        for i, ly in enumerate(layers):
            blk = getattr(mlp, "layer%d" % i)
            blk.conv.weight.copy_(dev(ly["weight"]).view_as(blk.conv.weight))
            blk.bn.bn.weight.copy_(dev(ly["gamma"])); blk.bn.bn.bias.copy_(dev(ly["beta"]))
            blk.bn.bn.running_mean.copy_(dev(ly["mean"])); blk.bn.bn.running_var.copy_(dev(ly["var"]))


def _fp_diag(fp, orc, unknown, known, uf, kf, got, ref):
    """B200_TEST_DIAG=1: where a feature-propagation mismatch comes from (inputs of the fused kernel vs the kernel)."""
    import sys
    import pointnet2.pointnet2_utils as pu
    bad = np.abs(got - ref) > ATOL + RTOL * np.abs(ref)
    rows = np.unique(np.nonzero(bad)[2]); chans = np.unique(np.nonzero(bad)[1]); scenes = np.unique(np.nonzero(bad)[0])
    out = ["fp diag: %d bad elements; scenes %s; %d rows %s; %d channels %s" % (
        int(bad.sum()), scenes.tolist(), len(rows), rows[:24].tolist(), len(chans), chans[:24].tolist())]
    with torch.no_grad():
        again = fp(dev(unknown), dev(known), dev(uf), dev(kf)).cpu().numpy()
        out.append("second run in the same state: equal to the first %s, within bound %s" % (
            bool(np.array_equal(again, got)), bool((np.abs(again - ref) <= ATOL + RTOL * np.abs(ref)).all())))
        dist, idx = pu.three_nn(dev(unknown), dev(known))
        d2, i3 = orc.three_nn(unknown, known)
        out.append("three_nn idx equal %s, dist equal %s" % (bool(np.array_equal(idx.cpu().numpy(), i3)),
                                                          bool(np.array_equal(dist.cpu().numpy(), np.sqrt(d2)))))
        os.environ["B200_SA_FUSED"] = "0"
        try:
            unf = fp(dev(unknown), dev(known), dev(uf), dev(kf)).cpu().numpy()
        finally:
            del os.environ["B200_SA_FUSED"]
        out.append("op-by-op path max err %.3g; fused max err %.3g" % (float(np.abs(unf - ref).max()), float(np.abs(got - ref).max())))
        layers = fp.mlp.fold_affine()
        out.append("folded scale/shift finite: %s" % all(bool(torch.isfinite(t).all()) for tri in layers for t in tri))
        # the CPU side: is the fp32 restatement itself stable, and which of its stages differs from the device's?
        import torch_ref as tr2
        import cases as cs
        B, C1 = unknown.shape[0], (uf.shape[1] if uf is not None else 0)
        spec_layers = cs.mlp_params(3, [C1 + kf.shape[1]] + [w.size(0) for w, _, _ in layers])
        ref2 = tr2.fp_forward(unknown, known, uf, kf, spec_layers)
        out.append("CPU restatement recomputed: equal to the first %s (max diff %.3g); threads %d" % (
            bool(np.array_equal(ref2, ref)), float(np.abs(ref2 - ref).max()), torch.get_num_threads()))
        recip = 1.0 / (torch.from_numpy(np.sqrt(d2)) + 1e-8)
        w_cpu = (recip / recip.sum(2, keepdim=True)).numpy()
        rg = 1.0 / (dist + 1e-8)
        w_gpu = (rg / torch.sum(rg, dim=2, keepdim=True))
        out.append("blend weights: max |gpu - cpu| %.3g" % float(np.abs(w_gpu.cpu().numpy() - w_cpu).max()))
        interp_gpu = pu.three_interpolate(dev(kf), idx, w_gpu.contiguous()).cpu().numpy()
        interp_cpu = orc.three_interpolate(kf, i3, w_cpu)
        out.append("interpolated features: max |gpu - cpu| %.3g" % float(np.abs(interp_gpu - interp_cpu).max()))
        x = np.concatenate([interp_cpu, uf], 1) if uf is not None else interp_cpu
        xt = torch.from_numpy(x).unsqueeze(-1)
        y32 = tr2.shared_mlp_fp32(xt, spec_layers).squeeze(-1).numpy()
        y64 = xt.double()
        for ly in spec_layers:
            w64 = torch.from_numpy(ly["weight"]).double().view(ly["weight"].shape[0], -1, 1, 1)
            y64 = torch.nn.functional.conv2d(y64, w64)
            y64 = torch.nn.functional.batch_norm(y64, torch.from_numpy(ly["mean"]).double(), torch.from_numpy(ly["var"]).double(),
                                                 torch.from_numpy(ly["gamma"]).double(), torch.from_numpy(ly["beta"]).double(),
                                                 training=False, eps=1e-5)
            y64 = torch.relu(y64)
        y64 = y64.squeeze(-1).numpy()
        out.append("max |cpu fp32 MLP - cpu fp64 MLP| %.3g ; max |device - cpu fp64| %.3g ; max |first ref - cpu fp64| %.3g" % (
            float(np.abs(y32 - y64).max()), float(np.abs(got - y64).max()), float(np.abs(ref - y64).max())))
        torch.set_num_threads(1)
        y1 = tr2.shared_mlp_fp32(xt, spec_layers).squeeze(-1).numpy()
        out.append("cpu fp32 MLP with one thread: max |y - fp64| %.3g" % float(np.abs(y1 - y64).max()))
    print("\n".join(out), file=sys.stderr)


@pytest.mark.parametrize("shape", [
    (2, 512, 256, 24, 64, [48, 32], "small, hidden layer fused in one launch"),
    (2, 512, 256, 256, 256, [256, 256], "FP1 of the VoteNet backbone (backbone_module.py:71)"),
    (8, 1024, 512, 256, 256, [256, 256], "FP2 at the bench batch"),
    (1, 300, 7, 0, 32, [64], "no skip features, fewer known points than unknown, ragged rows"),
])
def test_fp_module(pkg, orc, tr, shape, monkeypatch):
    """PointnetFPModule (three_nn -> inverse-distance blend -> concat skip -> SharedMLP, pointnet2_modules.py:377-422):
    the fused path (rows built in the tensor-core kernel's producers, no interpolated / concatenated tensor in HBM)
    against the fp32 restatement, every element, 1e-5; and the op-by-op path (B200_SA_FUSED=0) on the same module."""
    import pointnet2_modules as M
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, n, m, C1, C2, spec, _ = shape
    rng = np.random.default_rng(0)
    unknown = cases.cloud(1, B, n)
    known = unknown[:, :m].copy() if m <= n else cases.cloud(2, B, m)
    uf = rng.standard_normal((B, C1, n)).astype(np.float32) if C1 else None
    kf = rng.standard_normal((B, C2, m)).astype(np.float32)
    layers = cases.mlp_params(3, [C1 + C2] + spec)
    fp = M.PointnetFPModule(mlp=[C1 + C2] + spec).cuda().eval()
    _load_mlp(fp.mlp, layers)
    launches0 = pkg.cabi().launch_count()
    with torch.no_grad():
        got = fp(dev(unknown), dev(known), dev(uf), dev(kf)).cpu().numpy()
    assert pkg.cabi().launch_count() - launches0 >= 3  # three_nn + transposes + the fused rows kernel(s): not torch convs
    ref = tr.fp_forward(unknown, known, uf, kf, layers)
    if os.environ.get("B200_TEST_DIAG") and not (np.abs(got - ref) <= ATOL + RTOL * np.abs(ref)).all():
        _fp_diag(fp, orc, unknown, known, uf, kf, got, ref)
    assert_close(got, ref)
    monkeypatch.setenv("B200_SA_FUSED", "0")
    with torch.no_grad():
        unfused = fp(dev(unknown), dev(known), dev(uf), dev(kf)).cpu().numpy()
    err = np.abs(unfused - ref)  # cuDNN's fp32 algorithm choice is not under our control -> 1e-4 for this path only
    assert (err <= 1e-4 + 1e-4 * np.abs(ref)).all(), float(err.max())


@pytest.mark.parametrize("spec", [[259, 128, 128, 128], [128, 128, 128, 97], [256, 256, 256, 3], [64, 32], [20, 256]])
def test_shared_mlp_forward_rows(pkg, tr, spec):
    """pt_utils.SharedMLP.forward in eval mode = fused row MLPs (1x1-conv heads, grid_conv_module.py:108-113) vs the fp32
    restatement, every element, 1e-5; widths that are no multiple of 32 / of 4, >128-wide layers chained."""
    import pointnet2.pytorch_utils as pt
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, spec[0], 50, 16)).astype(np.float32)
    layers = cases.mlp_params(7, spec)
    mlp = pt.SharedMLP(list(spec), bn=True).cuda().eval()
    _load_mlp(mlp, layers)
    launches0 = pkg.cabi().launch_count()
    with torch.no_grad():
        got = mlp(dev(x))
    assert pkg.cabi().launch_count() > launches0, "SharedMLP.forward did not reach libb200pc.so"
    ref = tr.shared_mlp(torch.from_numpy(x), layers).numpy()
    assert got.shape == ref.shape
    assert_close(got.cpu().numpy(), ref)


def test_frozen_plans_equal_per_call_packing(pkg, orc, tr):
    """freeze_inference(): packed weights cached per module (no tc_pack_weights_kernel in steady state) -- bit-identical
    outputs, fewer launches; unfreezing picks up a weight change made through .data (the reference's EMA idiom)."""
    import pointnet2_modules as M
    import pointnet2.pytorch_utils as pt
    torch.manual_seed(0)
    net = M.PointnetSAModuleVotes(npoint=256, radius=0.4, nsample=32, mlp=[128, 128, 128, 256], use_xyz=True,
                                  normalize_xyz=True).cuda().eval()
    xyz = dev(cases.cloud(3, 2, 2048))
    feats = torch.randn(2, 128, 2048, device="cuda")
    with torch.no_grad():
        _, base, _ = net(xyz, feats)
        c0 = pkg.cabi().launch_count()
        net(xyz, feats)
        per_call = pkg.cabi().launch_count() - c0
        assert pt.freeze_inference(net) == 1
        _, first, _ = net(xyz, feats)          # builds the plan
        c0 = pkg.cabi().launch_count()
        _, frozen, _ = net(xyz, feats)
        steady = pkg.cabi().launch_count() - c0
        assert torch.equal(base, first) and torch.equal(base, frozen)
        assert steady < per_call, (steady, per_call)
        net.mlp_module.layer0.conv.weight.data.mul_(1.5)   # no version bump
        _, stale, _ = net(xyz, feats)
        assert torch.equal(stale, frozen)                  # frozen = the caller's promise; documented
        pt.unfreeze(net)
        _, fresh, _ = net(xyz, feats)
        assert not torch.equal(fresh, frozen)


def test_grid_interp_mlp_max_fused_vs_oracle(pkg, orc, tr, monkeypatch):
    """IoU-branch sampler (grid_conv_module.py:87-113 shapes: K*64 grid points vs 1024 seeds, MLP [259,128,128,128]):
    fused tensor-core kernel vs the oracle stack (C three_nn / three_interpolate + fp32 torch MLP), every element 1e-5,
    and vs the op-by-op path on the same inputs."""
    import pointnet2.pointnet2_utils as U
    import pointnet2.pytorch_utils as pt
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(3)
    B, K, m, C = 2, 64, 1024, 256
    seeds = (rng.random((B, m, 3)) * 6).astype(np.float32)
    feats = rng.standard_normal((B, C, m)).astype(np.float32)
    grid = (rng.random((B, K * 64, 3)) * 6).astype(np.float32)
    rel = (rng.random((B, K * 64, 3)) - 0.5).astype(np.float32)
    layers = cases.mlp_params(9, [C + 3, 128, 128, 128])
    mlp = pt.SharedMLP([C + 3, 128, 128, 128], bn=True).cuda().eval()
    _load_mlp(mlp, layers)
    # oracle stack
    d2, oidx = orc.three_nn(grid, seeds)
    ow = 1.0 / (torch.sqrt(torch.from_numpy(d2)) + 1e-8)
    ow = (ow / ow.sum(2, keepdim=True)).numpy()
    interp = torch.from_numpy(orc.three_interpolate(feats, oidx, ow)).view(B, C, K, 64)
    x = torch.cat([torch.from_numpy(rel).transpose(1, 2).contiguous().view(B, 3, K, 64), interp], 1)
    ref = torch.nn.functional.max_pool2d(tr.shared_mlp(x, layers), kernel_size=[1, 64]).squeeze(-1).numpy()
    with torch.no_grad():
        dist, idx = U.three_nn(dev(grid), dev(seeds))
        assert np.array_equal(idx.cpu().numpy(), oidx)
        w = 1.0 / (dist + 1e-8)
        w = (w / w.sum(2, keepdim=True)).contiguous()
        fused = U.grid_interp_mlp_max(dev(feats), idx, w, dev(rel), 64, mlp)
        monkeypatch.setenv("B200_SA_FUSED", "0")
        generic = U.grid_interp_mlp_max(dev(feats), idx, w, dev(rel), 64, mlp)
    assert fused.shape == (B, 128, K)
    assert_close(fused.cpu().numpy(), ref)
    err = (generic.cpu().numpy() - ref)
    assert (np.abs(err) <= 1e-4 + 1e-4 * np.abs(ref)).all()


def test_msg_modules_forward_backward(pkg):
    """Multi-scale modules (pointnet2_modules.py:83-166,280-359 and the smoke demo :506-525): eval = fused per scale,
    train = differentiable op-by-op path; both produce (B, sum(mlp[-1]), npoint)."""
    import pointnet2_modules as M
    torch.manual_seed(1)
    xyz = torch.randn(2, 90, 3, device="cuda")
    feats = torch.randn(2, 6, 90, device="cuda", requires_grad=True)
    net = M.PointnetSAModuleMSG(npoint=8, radii=[1.0, 2.0], nsamples=[6, 3], mlps=[[6, 3], [6, 6]]).cuda()
    new_xyz, out = net(xyz, feats)                                   # train mode
    assert new_xyz.shape == (2, 8, 3) and out.shape == (2, 9, 8)
    out.sum().backward()
    assert feats.grad is not None and torch.isfinite(feats.grad).all()
    net.eval()
    with torch.no_grad():
        a = net(xyz, feats.detach())[1]                              # nsample 6/3 do not divide 128 -> fp32 FFMA kernel
        os.environ["B200_SA_FUSED"] = "0"
        try:
            b = net(xyz, feats.detach())[1]
        finally:
            del os.environ["B200_SA_FUSED"]
    assert torch.allclose(a, b, atol=1e-5, rtol=1e-5)
    votes = M.PointnetSAModuleMSGVotes(npoint=8, radii=[1.0], nsamples=[4], mlps=[[6, 5]]).cuda().eval()
    with torch.no_grad():
        nx, f, inds = votes(xyz, feats.detach())
    assert f.shape == (2, 5, 8) and inds.dtype == torch.int32


def test_training_step_with_flat_gradient_bucket(pkg):
    """One optimisation step through the drop-in (train-mode BN -> unfused differentiable path) + the flat-bucket
    gradient reduction helper (single rank here; the 2-rank reduction is covered on gloo in the CPU suite)."""
    import importlib
    import pointnet2_modules as M
    shard = importlib.import_module("3dioumatch_b200.shard")
    torch.manual_seed(0)
    net = M.PointnetSAModuleVotes(npoint=32, radius=0.5, nsample=8, mlp=[4, 16, 16], use_xyz=True, normalize_xyz=True).cuda()
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    xyz = torch.from_numpy(cases.cloud(12, 2, 400)).cuda()
    feats = torch.randn(2, 4, 400, device="cuda")
    before = [p.detach().clone() for p in net.parameters()]
    _, f, _ = net(xyz, feats)
    f.square().mean().backward()
    assert shard.allreduce_gradients(net) == sum(p.numel() for p in net.parameters() if p.grad is not None)
    opt.step()
    assert any(not torch.equal(a, b) for a, b in zip(before, net.parameters()))
