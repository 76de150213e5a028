"""Run the reference APPLICATION unmodified on an operator stack.

The reference's callers -- models/votenet_iou_branch.py (VoteNet), backbone_module.py, voting_module.py,
proposal_module.py, grid_conv_module.py, the loss helpers, utils/box_util.py and the dataset configs -- are pure python
ABOVE the drop-in boundary.  oracle/build_ref.py installs them (copied verbatim from /root/reference, git-ignored)
into baseline/_ref/.  `load()` imports them with a chosen operator stack in front of sys.path:

    the drop-in stack   3dioumatch_b200/dropin/            -> every native op resolves to libb200pc.so (this package)
    any other stack     e.g. the reference's own operator modules + extensions (tests / bench.py --impl reference
                        pass that path; the package itself never names it)

so the SAME caller source runs on either stack (BASELINE.json north_star: "votenet_iou_branch.py and the loss helpers
call it unchanged").  Several stacks can be loaded in one process: every load gets private copies of the colliding
module names (pointnet2, pointnet2_modules, pcdet, models, utils ...), removed from sys.modules again afterwards.
"""
import importlib
import os
import sys
import types

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
APP_ROOT = os.path.join(ROOT, "baseline", "_ref")
DROPIN_DIR = os.path.join(PKG_DIR, "dropin")
FAST_CALLERS_DIR = os.path.join(PKG_DIR, "dropin_callers")

# top-level module names both stacks (and the application) define
_OWNED = ("pointnet2", "pointnet2_modules", "pointnet2_utils", "pytorch_utils", "pcdet", "models", "utils", "scannet",
          "sunrgbd", "_ext", "iou3d_nms_cuda", "nn_distance", "box_util", "pc_util", "nms", "model_util_scannet",
          "model_util_sunrgbd", "_b200_rows")
# optional third-party imports of reference files that are not on the measured path (plotting / mesh IO)
_STUBS = ("trimesh", "matplotlib", "matplotlib.pyplot", "plyfile", "cv2", "mayavi", "scipy.io")


class _Missing:
    """Placeholder for a name of an optional third-party module that is absent here: attribute access works (default
    arguments like `pyplot.cm.jet` are evaluated at import), calling it raises."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Missing(self._name + "." + k)

    def __call__(self, *a, **k):
        raise RuntimeError("optional dependency of the reference is not installed in this image: %s" % self._name)


def _stub_attr(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Missing(name)


def dropin_paths(fast_callers=False):
    """sys.path entries of this package's operator stack.  fast_callers=True additionally shadows the reference's
    voting / proposal / grid-conv modules and compute_iou_labels with the drop-in mirrors of dropin_callers/ (SURVEY 8f
    rows n1, n2: same class names, constructor arguments and state-dict keys, fused kernels underneath)."""
    paths = [os.path.join(DROPIN_DIR, "pointnet2"), DROPIN_DIR]
    if fast_callers:
        paths.insert(0, FAST_CALLERS_DIR)
    return paths


def available(app_root=APP_ROOT):
    return os.path.exists(os.path.join(app_root, "models", "votenet_iou_branch.py"))


def load(ops_paths, app_root=APP_ROOT, name="stack", with_losses=False):
    """Import the reference application on the operator stack found under `ops_paths` (front of sys.path)."""
    if not available(app_root):
        raise RuntimeError("reference application not installed under %s (python oracle/build_ref.py)" % app_root)
    owned = lambda k: k.split(".")[0] in _OWNED  # noqa: E731
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if owned(k)}
    saved_path = list(sys.path)
    stubbed = []
    sys.path[:0] = list(ops_paths) + [app_root]
    try:
        for s in _STUBS:
            if s not in sys.modules:
                try:
                    importlib.import_module(s)
                except Exception:  # noqa: BLE001 -- absent in this image; never called on the measured path
                    m = types.ModuleType(s)
                    m.__path__ = []
                    m.__dict__["__getattr__"] = _stub_attr  # `from plyfile import PlyData` etc.
                    sys.modules[s] = m
                    stubbed.append(s)
        ns = types.SimpleNamespace(name=name)
        ns.utils = importlib.import_module("pointnet2.pointnet2_utils")
        ns.pt = importlib.import_module("pointnet2.pytorch_utils")
        ns.modules = importlib.import_module("pointnet2_modules")
        ns.ext = importlib.import_module("pointnet2._ext")
        ns.iou = importlib.import_module("pcdet.ops.iou3d_nms.iou3d_nms_utils")
        ns.votenet = importlib.import_module("models.votenet_iou_branch")
        ns.loss_iou = importlib.import_module("models.loss_helper_iou")
        ns.box_util = importlib.import_module("utils.box_util")
        ns.scannet = importlib.import_module("scannet.model_util_scannet")
        ns.sunrgbd = importlib.import_module("sunrgbd.model_util_sunrgbd")
        if with_losses:
            ns.loss_labeled = importlib.import_module("models.loss_helper_labeled")
            ns.loss_unlabeled = importlib.import_module("models.loss_helper_unlabeled")
        ns.files = {k: getattr(v, "__file__", None) for k, v in sys.modules.items() if owned(k)}
    finally:
        for k in [k for k in sys.modules if owned(k)]:
            del sys.modules[k]
        for s in stubbed:
            sys.modules.pop(s, None)
        sys.modules.update(saved)
        sys.path[:] = saved_path
    return ns


def dataset_config(ns, dataset="scannet"):
    """The reference's own dataset config object (ScanNet: 18 classes / 1 heading bin / 18 size clusters,
    scannet/model_util_scannet.py:19-35; SUN RGB-D: 10 / 12 / 10, sunrgbd/model_util_sunrgbd.py:19-45)."""
    return ns.scannet.ScannetDatasetConfig() if dataset == "scannet" else ns.sunrgbd.SunrgbdDatasetConfig()


def build_votenet(ns, dataset="scannet", num_proposal=256, seed=1, device="cuda", train=False):
    """VoteNet exactly as train.py:176-185 builds it (input_feature_dim=1: height), random-initialised from `seed`,
    with non-trivial BatchNorm statistics so that eval-mode parity exercises the folded affine."""
    import torch
    import torch.nn as nn
    cfg = dataset_config(ns, dataset)
    torch.manual_seed(seed)
    net = ns.votenet.VoteNet(num_class=cfg.num_class, num_heading_bin=cfg.num_heading_bin,
                             num_size_cluster=cfg.num_size_cluster, mean_size_arr=cfg.mean_size_arr,
                             dataset_config=cfg, num_proposal=num_proposal, input_feature_dim=1,
                             sampling="seed_fps", query_feats="seed")
    gen = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) * 0.5 + 0.75)
    if hasattr(net.backbone_net, "prefetch_proposals") and net.pnet.sampling == "seed_fps":
        net.backbone_net.prefetch_proposals = num_proposal   # drop-in backbone mirror: seed FPS joins the prefetched chain
    net = net.to(device)
    return (net.train() if train else net.eval()), cfg


def make_inputs(B=8, N=40000, seed=0, room=(8.0, 8.0, 3.0), dataset="scannet", cfg=None, max_gt=64):
    """Synthetic scenes + labels in the reference's dataset format (numpy, host): ScanNet-shaped 8x8x3 m rooms or SUN
    RGB-D-shaped 5x5x2.5 m."""
    synth = importlib.import_module(__name__.rsplit(".", 1)[0] + ".synth")
    pc = synth.scene_cloud(seed, B, N, room=room)
    if cfg is None:
        nc, nh, ns_, msa = (18, 1, 18, np.full((18, 3), 0.8, np.float32)) if dataset == "scannet" else \
            (10, 12, 10, np.full((10, 3), 0.8, np.float32))
    else:
        nc, nh, ns_, msa = cfg.num_class, cfg.num_heading_bin, cfg.num_size_cluster, cfg.mean_size_arr
    labels = synth.scene_labels(seed, pc, nc, nh, ns_, msa, max_gt=max_gt,
                                extent=(room[0] * 0.75, room[1] * 0.75, room[2] * 0.66))
    return pc, labels


LABEL_KEYS_IOU = ("center_label", "box_label_mask", "heading_class_label", "heading_residual_label",
                  "size_class_label", "size_residual_label")


def forward_with_iou_labels(ns, net, cfg, point_clouds, labels):
    """One step of BASELINE configs[1]: VoteNet.forward (votenet_iou_branch.py:139-151) + the IoU labels of the proposals
    against the padded GT boxes (loss_helper_iou.py:52-112, called as loss_helper_labeled.py:219-226 does)."""
    import torch
    end_points = net({"point_clouds": point_clouds})
    for k in LABEL_KEYS_IOU:
        end_points[k] = labels[k].clone() if k == "center_label" else labels[k]  # center_label is modified in place (:56-58)
    inds = torch.arange(point_clouds.shape[0], device=point_clouds.device)
    iou_labels, objectness_label, assignment = ns.loss_iou.compute_iou_labels(
        end_points, inds, end_points["aggregated_vote_xyz"], end_points["center"], None, None,
        end_points["heading_scores"], end_points["heading_residuals"], end_points["size_scores"],
        end_points["size_residuals"], config_dict={"dataset_config": cfg})
    end_points["iou_labels"] = iou_labels
    end_points["iou_objectness_label"] = objectness_label
    end_points["iou_assignment"] = assignment
    return end_points
