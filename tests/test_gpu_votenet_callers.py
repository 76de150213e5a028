"""GPU: BASELINE configs[1] through the REFERENCE'S OWN CALLERS.

models/votenet_iou_branch.py (VoteNet: backbone_module.py, voting_module.py, proposal_module.py, grid_conv_module.py) and
models/loss_helper_iou.py run UNMODIFIED (installed verbatim into baseline/_ref by oracle/build_ref.py), once on this
package's drop-in operator stack and once on the reference's own operator stack (oracle/_ref: its python operator modules
on its CUDA extensions compiled for sm_100a, cuDNN convs, TF32 off), in ONE process, same seeded weights, at the bench
shape (B=8, N=40000, C=4, K=256):

  (a) every index tensor of the forward is bit-exact; float outputs are bounded end to end;
  (b) per layer, the reference's own layer INPUTS are fed to both stacks' modules and every output element must satisfy
      |err| <= 1e-5 + 1e-5 |ref| (SA1-SA4, FP1, FP2, vote aggregation, the GridConv sampler MLP, boxes_iou3d_gpu);
  (c) the execution mode bench.py times -- CUDA-graph replay, 5 lanes, frozen plans, throughput FPS policy -- reproduces
      the eager drop-in outputs bit for bit on rotating inputs.
"""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_OPS = os.path.join(ROOT, "oracle", "_ref")
B, N, K = 8, 40000, 256


def _stacks():
    ra = importlib.import_module("3dioumatch_b200.refapp")
    if not (ra.available() and os.path.exists(os.path.join(REF_OPS, "pointnet2", "_ext.so"))):
        pytest.skip("reference application / operator stack not installed (python oracle/build_ref.py)")
    ours = ra.load(ra.dropin_paths(), name="b200")
    ref = ra.load([os.path.join(REF_OPS, "pointnet2"), REF_OPS], name="reference")
    assert ours.ext.__file__.endswith("_ext.py") and ref.ext.__file__.endswith("_ext.so")
    assert ours.votenet.__file__ == ref.votenet.__file__ or open(ours.votenet.__file__).read() == open(ref.votenet.__file__).read()
    return ra, ours, ref


@pytest.fixture(scope="module")
def world():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ra, ours, ref = _stacks()
    net_o, cfg_o = ra.build_votenet(ours, "scannet", K, seed=1)
    net_r, cfg_r = ra.build_votenet(ref, "scannet", K, seed=1)
    sd_o, sd_r = net_o.state_dict(), net_r.state_dict()
    assert sd_o.keys() == sd_r.keys()                                   # same parameter / buffer names: checkpoints load
    assert all(torch.equal(sd_o[k], sd_r[k]) for k in sd_o)
    pc, labels = ra.make_inputs(B, N, seed=0, cfg=cfg_r)
    pc_d = torch.from_numpy(pc).cuda()
    lab_d = {k: torch.from_numpy(v).cuda() for k, v in labels.items()}
    # the reference stack, with every SA / FP module's inputs and outputs recorded
    rec = {}

    def hook(name):
        def fn(mod, args, kwargs, out):
            rec[name] = (args, kwargs, out)
        return fn
    handles = []
    for name in ("sa1", "sa2", "sa3", "sa4", "fp1", "fp2"):
        handles.append(getattr(net_r.backbone_net, name).register_forward_hook(hook(name), with_kwargs=True))
    handles.append(net_r.pnet.vote_aggregation.register_forward_hook(hook("vote_aggregation"), with_kwargs=True))
    handles.append(net_r.grid_conv.mlp_before_iou.register_forward_hook(hook("mlp_before_iou"), with_kwargs=True))
    with torch.no_grad():
        ep_r = ra.forward_with_iou_labels(ref, net_r, cfg_r, pc_d, lab_d)
        torch.cuda.synchronize()
        ep_o = ra.forward_with_iou_labels(ours, net_o, cfg_o, pc_d, lab_d)
        torch.cuda.synchronize()
    for h in handles:
        h.remove()
    return dict(ra=ra, ours=ours, ref=ref, net_o=net_o, net_r=net_r, cfg_o=cfg_o, cfg_r=cfg_r, pc=pc_d, labels=lab_d,
                ep_o=ep_o, ep_r=ep_r, rec=rec)


def _every_element(got, ref, what, atol=1e-5, rtol=1e-5):
    got, ref = got.float().cpu().numpy(), ref.float().cpu().numpy()
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    err = np.abs(got - ref)
    bad = err > atol + rtol * np.abs(ref)
    assert not bad.any(), "%s: %d of %d elements outside 1e-5 (max err %.3g at |ref| %.3g)" % (
        what, int(bad.sum()), bad.size, float(err.max()), float(np.abs(ref).flat[err.argmax()]))


def test_a_indices_bit_exact_and_outputs_bounded(world):
    ep_o, ep_r = world["ep_o"], world["ep_r"]
    for k in ("sa1_inds", "sa2_inds", "fp2_inds", "seed_inds", "aggregated_vote_inds"):
        assert ep_o[k].dtype == ep_r[k].dtype == torch.int32
        assert torch.equal(ep_o[k], ep_r[k]), k
    for k in ("sa1_xyz", "sa2_xyz", "sa3_xyz", "sa4_xyz", "seed_xyz", "fp2_xyz"):
        assert torch.equal(ep_o[k], ep_r[k]), k                         # gathers of exact indices are exact
    # backbone features: neighbour lists are identical (they depend on exact coordinates only), so the only difference
    # is fp32 rounding accumulated layer after layer -- every element bounded
    for k, tol in (("sa1_features", 2e-5), ("sa2_features", 5e-5), ("sa3_features", 1e-4), ("sa4_features", 1e-4),
                   ("fp2_features", 2e-4), ("vote_xyz", 2e-4), ("vote_features", 2e-4)):
        _every_element(ep_o[k], ep_r[k], k, atol=tol, rtol=tol)
    # past the votes, data-dependent DISCRETE choices sit between the two stacks' roundings (a vote crossing a ball
    # boundary, an arg-max over size classes, a three_nn neighbour swap): isolated proposals may differ, the bulk may not
    for k, tol in (("center", 1e-3), ("size", 1e-3), ("iou_scores", 2e-3), ("iou_labels", 2e-3), ("objectness_scores", 2e-3)):
        err = (ep_o[k].float() - ep_r[k].float()).abs()
        frac = float((err <= tol + tol * ep_r[k].float().abs()).float().mean())
        assert frac >= 0.995, (k, frac, float(err.max()))
    assert float((ep_r["iou_labels"] > 0).float().mean()) > 0.005         # the synthetic labels do overlap proposals
    assert torch.equal(ep_o["iou_assignment"][ep_r["iou_labels"] > 0.05], ep_r["iou_assignment"][ep_r["iou_labels"] > 0.05]) or \
        float((ep_o["iou_assignment"] == ep_r["iou_assignment"]).float().mean()) > 0.99


@pytest.mark.parametrize("name", ["sa1", "sa2", "sa3", "sa4", "fp1", "fp2", "vote_aggregation"])
def test_b_layer_parity_on_reference_inputs(world, name):
    """The reference stack's own layer inputs -> our module (same weights): every output element within 1e-5."""
    args, kwargs, out_r = world["rec"][name]
    net_o = world["net_o"]
    mod = net_o.pnet.vote_aggregation if name == "vote_aggregation" else getattr(net_o.backbone_net, name)
    launches0 = importlib.import_module("3dioumatch_b200._cabi").launch_count()
    with torch.no_grad():
        out_o = mod(*args, **kwargs)
    assert importlib.import_module("3dioumatch_b200._cabi").launch_count() > launches0
    if name.startswith("fp"):
        _every_element(out_o, out_r, name)
    else:
        assert torch.equal(out_o[0], out_r[0]), name + ": new_xyz"
        assert torch.equal(out_o[2], out_r[2]), name + ": inds"
        _every_element(out_o[1], out_r[1], name + ": new_features")


def test_b_gridconv_sampler_and_iou_on_reference_inputs(world):
    args, kwargs, out_r = world["rec"]["mlp_before_iou"]        # (B, 259, K, 64) interpolated grid features
    with torch.no_grad():
        out_o = world["net_o"].grid_conv.mlp_before_iou(*args, **kwargs)
    _every_element(out_o, out_r, "GridConv mlp_before_iou (B,128,K,64)")
    # the IoU label op on the reference's predicted boxes: all (B*K) x (B*64) pairs (loss_helper_iou.py:95-96)
    pred = world["ep_r"]["pred_bbox"].view(-1, 7).contiguous()
    lab = world["labels"]
    cfg = world["cfg_r"]
    center = lab["center_label"].clone()
    center[(1 - lab["box_label_mask"]).unsqueeze(-1).expand(-1, -1, 3).bool()] = -1000
    gt = torch.cat([center, cfg.class2size_gpu(lab["size_class_label"], lab["size_residual_label"]),
                    -cfg.class2angle_gpu(lab["heading_class_label"], lab["heading_residual_label"])[:, :, None]], 2)
    gt = gt.view(-1, 7).contiguous()
    iou_r = world["ref"].iou.boxes_iou3d_gpu(pred, gt)
    torch.cuda.synchronize()
    iou_o = world["ours"].iou.boxes_iou3d_gpu(pred, gt)
    assert float(iou_r.max()) > 0.1
    _every_element(iou_o, iou_r, "boxes_iou3d_gpu (%d x %d)" % (pred.shape[0], gt.shape[0]), atol=1e-5, rtol=0)


def test_c_graph_replay_lanes_frozen_throughput_policy(world):
    """bench.py's execution mode against the eager drop-in outputs: bit-identical on rotating inputs."""
    ra, ours, net_o, cfg_o = world["ra"], world["ours"], world["net_o"], world["cfg_o"]
    runner_mod = importlib.import_module("3dioumatch_b200.runner")
    cabi = importlib.import_module("3dioumatch_b200._cabi")
    keys = ra.LABEL_KEYS_IOU
    out_keys = ("sa1_inds", "aggregated_vote_inds", "fp2_features", "center", "iou_scores", "iou_labels")
    rng = np.random.default_rng(5)
    pcs = []
    base = world["pc"].cpu().numpy()
    for i in range(7):
        pc = base.copy()
        pc[:, :, :3] += rng.normal(0, 0.01, (B, 1, 3)).astype(np.float32)
        pcs.append(torch.from_numpy(pc[:, rng.permutation(N)] if i else pc).pin_memory())
    labels = [world["labels"][k] for k in keys]

    def step(pc, *lab):
        ep = ra.forward_with_iou_labels(ours, net_o, cfg_o, pc, dict(zip(keys, lab)))
        return {k: ep[k] for k in out_keys}
    with torch.no_grad():
        eager = []
        for pc in pcs:
            o = step(pc.cuda(), *labels)
            eager.append({k: v.clone() for k, v in o.items()})
        torch.cuda.synchronize()
        assert ours.pt.freeze_inference(net_o) >= 6
        prev = cabi.set_fps_policy("throughput")
        try:
            runner = runner_mod.LaneRunner(step, [pcs[0].cuda()] + labels, lanes=5, graphs=True)
            assert runner.use_graphs, runner.capture_error
            outs = []
            for i, pc in enumerate(pcs):
                o = runner.run(i, [pc] + labels)
                lane = runner.lanes[i % 5]
                with torch.cuda.stream(lane):
                    outs.append({k: v.clone() for k, v in o.items()})
            torch.cuda.synchronize()
        finally:
            cabi.set_fps_policy(prev)
            ours.pt.unfreeze(net_o)
    for i, (e, g) in enumerate(zip(eager, outs)):
        for k in out_keys:
            assert torch.equal(e[k], g[k]), "step %d: %s differs between eager and graph replay" % (i, k)


def test_d_fast_callers_mirrors_match_the_reference(world):
    """--callers fast: this package's mirrors of backbone / voting / proposal / GridConv modules and compute_iou_labels
    (dropin_callers/, SURVEY 8f n1 + n2) under the reference's VoteNet: same state dict, same indices, outputs within the
    same bounds as the unchanged callers, and no cuDNN / cuBLAS convolution left in the step."""
    ra = world["ra"]
    fast = ra.load(ra.dropin_paths(fast_callers=True), name="b200-fast")
    assert "dropin_callers" in fast.files["models.grid_conv_module"] and "baseline" in fast.files["models.votenet_iou_branch"]
    net_f, cfg_f = ra.build_votenet(fast, "scannet", K, seed=1)
    sd_f, sd_r = net_f.state_dict(), world["net_r"].state_dict()
    assert sd_f.keys() == sd_r.keys() and all(torch.equal(sd_f[k], sd_r[k]) for k in sd_f)
    from torch.profiler import ProfilerActivity, profile
    fast.pt.freeze_inference(net_f)
    with torch.no_grad():
        ra.forward_with_iou_labels(fast, net_f, cfg_f, world["pc"], world["labels"])
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            ep_f = ra.forward_with_iou_labels(fast, net_f, cfg_f, world["pc"], world["labels"])
            torch.cuda.synchronize()
    names = [e.name for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    # the one library GEMM left is torch.bmm of the (B*K, 64, 3) grids with their 3x3 rotation matrices (grid_conv_module.py:79)
    libs = [n for n in names if any(t in n.lower() for t in ("cudnn", "implicit_convolve", "tensorop", "xmma", "bn_fw"))]
    assert not libs, "library convolution / tensor-core GEMM kernels in the fast-caller step: %s" % sorted(set(libs))[:5]
    assert len(names) < 260
    ep_r = world["ep_r"]
    for k in ("sa1_inds", "sa2_inds", "fp2_inds", "seed_inds", "aggregated_vote_inds"):
        assert torch.equal(ep_f[k], ep_r[k]), k
    for k, tol in (("sa4_features", 1e-4), ("fp2_features", 2e-4), ("vote_xyz", 2e-4), ("vote_features", 2e-4)):
        _every_element(ep_f[k], ep_r[k], k, atol=tol, rtol=tol)
    for k, tol in (("center", 1e-3), ("size", 1e-3), ("iou_scores", 2e-3), ("iou_labels", 2e-3), ("objectness_scores", 2e-3)):
        err = (ep_f[k].float() - ep_r[k].float()).abs()
        frac = float((err <= tol + tol * ep_r[k].float().abs()).float().mean())
        assert frac >= 0.995, (k, frac, float(err.max()))
    assert torch.equal(ep_f["iou_objectness_label"], ep_r["iou_objectness_label"]) or \
        float((ep_f["iou_objectness_label"] == ep_r["iou_objectness_label"]).float().mean()) > 0.995
    # block-diagonal IoU labels on the REFERENCE's boxes == the diagonal blocks of its all-pairs matrix
    pred, lab, cfg = ep_r["pred_bbox"], world["labels"], world["cfg_r"]
    center = lab["center_label"].clone()
    center[(1 - lab["box_label_mask"]).unsqueeze(-1).expand(-1, -1, 3).bool()] = -1000
    gt = torch.cat([center, cfg.class2size_gpu(lab["size_class_label"], lab["size_residual_label"]),
                    -cfg.class2angle_gpu(lab["heading_class_label"], lab["heading_residual_label"])[:, :, None]], 2)
    blk = fast.iou.boxes_iou3d_batched(pred.contiguous(), gt.contiguous()).max(dim=2)[0]
    _every_element(blk, ep_r["iou_labels"], "block-diagonal iou_labels", atol=1e-5, rtol=0)
