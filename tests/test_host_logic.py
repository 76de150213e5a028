"""CPU: host-side mirror of the reference interface -- module names, state_dict keys, argument checks, BN folding."""
import numpy as np
import pytest
import torch


def test_state_dict_keys_match_reference_names(pkg):
    import pointnet2_modules as M
    sa = M.PointnetSAModuleVotes(npoint=128, radius=0.2, nsample=32, mlp=[1, 32, 64], use_xyz=True, normalize_xyz=True)
    keys = list(sa.state_dict().keys())
    assert "mlp_module.layer0.conv.weight" in keys          # pytorch_utils.py:29-31,116
    assert "mlp_module.layer0.bn.bn.running_mean" in keys   # pytorch_utils.py:46,120
    assert "mlp_module.layer1.bn.bn.weight" in keys
    assert sa.mlp_module.layer0.conv.weight.shape == (32, 4, 1, 1)   # mlp[0] += 3 (pointnet2_modules.py:209-211)
    assert sa.mlp_module.layer0.conv.bias is None                     # bias = bias and not bn (pytorch_utils.py:90)
    fp = M.PointnetFPModule(mlp=[512, 256, 256])
    assert "mlp.layer1.conv.weight" in fp.state_dict()
    msg = M.PointnetSAModuleMSG(npoint=2, radii=[5.0, 10.0], nsamples=[6, 3], mlps=[[6, 3], [6, 6]])
    assert "mlps.1.layer0.bn.bn.bias" in msg.state_dict()


def test_reference_checkpoint_layout_loads(pkg):
    """A state_dict produced by the reference's SharedMLP (if oracle/_ref is installed) loads into ours."""
    import importlib.util
    import os
    ref_py = os.path.join(pkg.PKG_DIR, "..", "oracle", "_ref", "pointnet2", "pytorch_utils.py")
    if not os.path.exists(ref_py):
        pytest.skip("oracle/_ref not installed")
    spec = importlib.util.spec_from_file_location("ref_pytorch_utils", ref_py)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    import pointnet2.pytorch_utils as ours
    a = ref.SharedMLP([7, 16, 32], bn=True)
    b = ours.SharedMLP([7, 16, 32], bn=True)
    assert list(a.state_dict().keys()) == list(b.state_dict().keys())
    b.load_state_dict(a.state_dict())
    x = torch.randn(2, 7, 5, 3)
    a.eval(); b.eval()
    assert torch.equal(a(x), b(x))


def test_fold_affine_equals_eval_forward(pkg):
    import pointnet2.pytorch_utils as pt
    torch.manual_seed(0)
    mlp = pt.SharedMLP([5, 8, 6], bn=True)
    for m in mlp.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.3)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    mlp.eval()
    x = torch.randn(3, 5, 4, 2)
    y = x
    for w, sc, sh in mlp.fold_affine():
        y = torch.relu(torch.einsum("oc,bcmk->bomk", w, y) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
    assert torch.allclose(y, mlp(x), atol=1e-5)
    mlp.train()
    assert mlp.fold_affine() is None        # training-mode BN cannot be folded -> unfused path
    assert pt.SharedMLP([5, 8], bn=False).fold_affine() is not None


def test_ext_rejects_cpu_and_bad_dtypes(pkg):
    import pointnet2._ext as ext
    xyz = torch.rand(1, 16, 3)
    with pytest.raises(RuntimeError, match="CPU not supported"):     # sampling.cpp:85-87
        ext.furthest_point_sampling(xyz, 4)
    with pytest.raises(RuntimeError, match="float"):                  # utils.h CHECK_IS_FLOAT
        ext.furthest_point_sampling(xyz.double(), 4)
    with pytest.raises(RuntimeError, match="contiguous"):             # utils.h CHECK_CONTIGUOUS
        ext.ball_query(xyz.transpose(1, 2), xyz, 0.2, 4)
    with pytest.raises(RuntimeError, match="int"):                    # utils.h CHECK_IS_INT
        ext.gather_points(torch.rand(1, 3, 16), torch.zeros(1, 4, dtype=torch.int64))


def test_iou_cpu_entry_and_wrappers(pkg, orc):
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
    import cases
    a = cases.boxes(3, 40)
    b = cases.boxes(4, 40, jitter_of=a)
    got = iu.boxes_bev_iou_cpu(a, b)                                  # numpy in -> numpy out (iou3d_nms_utils.py:12-28)
    assert isinstance(got, np.ndarray)
    assert np.abs(got - orc.boxes_iou_bev(a, b)).max() <= 1e-5
    d = cases.degenerate_boxes()
    got = iu.boxes_bev_iou_cpu(torch.from_numpy(d), torch.from_numpy(d)).numpy()
    ref = orc.boxes_iou_bev(d, d)
    ok = np.isfinite(ref)
    assert np.abs(got[ok] - ref[ok]).max() <= 1e-5
    with pytest.raises(RuntimeError, match="CUDA"):                   # iou3d_nms.cpp:14-19, without exit(-1)
        from pcdet.ops.iou3d_nms import iou3d_nms_cuda
        iou3d_nms_cuda.boxes_overlap_bev_gpu(torch.zeros(1, 7), torch.zeros(1, 7), torch.zeros(1, 1))


def test_fps_policy_knob_roundtrip(pkg):
    """b200pn2_fps_set_policy is host state only (no CUDA call): returns the previous policy, clamps unknown values."""
    import importlib
    cabi = importlib.import_module("3dioumatch_b200._cabi")
    first = cabi.set_fps_policy("throughput")
    assert first in ("latency", "throughput")
    assert cabi.set_fps_policy("latency") == "throughput"
    assert cabi.lib().b200pn2_fps_set_policy(7) == 0          # 7 is not a policy: stored as latency (0)
    assert cabi.lib().b200pn2_fps_set_policy(0 if first == "latency" else 1) == 0
