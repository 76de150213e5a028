"""GPU: the device pseudo-label filter (SURVEY 8f row n3 wired into a caller mirror, dropin_callers/models/
loss_helper_unlabeled.py) against the UNMODIFIED reference function models/loss_helper_unlabeled.py:get_pseudo_labels
running its host loops and numpy lhs_3d_faster_samecls (utils/nms.py:168-215), ScanNet and SUN RGB-D head shapes."""
import importlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_OPS = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def stacks():
    ra = importlib.import_module("3dioumatch_b200.refapp")
    if not (ra.available() and os.path.exists(os.path.join(REF_OPS, "pointnet2", "_ext.so"))):
        pytest.skip("reference application / operator stack not installed (python oracle/build_ref.py)")
    ref = ra.load([os.path.join(REF_OPS, "pointnet2"), REF_OPS], name="reference", with_losses=True)
    fast = ra.load(ra.dropin_paths(fast_callers=True), name="b200-fast", with_losses=True)
    assert "baseline" in ref.files["models.loss_helper_unlabeled"] and "baseline" in ref.files["utils.nms"]
    assert "dropin_callers" in fast.files["models.loss_helper_unlabeled"]
    return ra, ref, fast


@pytest.mark.parametrize("dataset", ["scannet", "sunrgbd"])
def test_pseudo_label_filter_mirror(stacks, dataset):
    ra, ref, fast = stacks
    cfg_r, cfg_f = ra.dataset_config(ref, dataset), ra.dataset_config(fast, dataset)
    n_lab, n_unl, K = 2, 5, 128
    g = torch.Generator().manual_seed(3 if dataset == "scannet" else 4)
    rnd = lambda *s: torch.randn(*s, generator=g)  # noqa: E731
    B = n_lab + n_unl
    nh, nsz, nc = cfg_r.num_heading_bin, cfg_r.num_size_cluster, cfg_r.num_class
    # boxes clustered around a few objects so that the suppression has work to do; confident heads so that many pass
    obj_centres = torch.rand(n_unl, 6, 3, generator=g) * torch.tensor([6.0, 6.0, 1.5]) - torch.tensor([3.0, 3.0, 0.0])
    which = torch.randint(0, 6, (n_unl, K), generator=g)
    center = torch.gather(obj_centres, 1, which.unsqueeze(-1).expand(-1, -1, 3)) + 0.15 * rnd(n_unl, K, 3)
    sem = 0.5 * rnd(n_unl, K, nc)
    sem.scatter_(2, (which % nc).unsqueeze(-1), 9.0)                               # one confident class per object
    objn = torch.stack([rnd(n_unl, K), rnd(n_unl, K) + 5.0], dim=2)               # mostly positive
    args = dict(pred_center=center, pred_sem_cls=sem, pred_objectness=objn, pred_heading_scores=rnd(n_unl, K, nh),
                pred_heading_residuals=0.1 * rnd(n_unl, K, nh), pred_size_scores=rnd(n_unl, K, nsz),
                pred_size_residuals=0.2 * rnd(n_unl, K, nsz, 3), pred_vote_xyz=center + 0.05 * rnd(n_unl, K, 3))
    args = {k: v.cuda() for k, v in args.items()}
    ema = {"iou_scores": (2.0 * rnd(B, K, nc) + 1.0).cuda()}
    ep = {"supervised_mask": torch.cat([torch.ones(n_lab), torch.zeros(n_unl)]).long().cuda()}
    conf = {"use_lhs": True, "nms_iou": 0.25, "use_old_type_nms": False, "obj_threshold": 0.9, "cls_threshold": 0.9,
            "iou_threshold": 0.25, "view_stats": False, "dataset": dataset}

    def run(ns, cfg):
        e = dict(ep)
        out = ns.loss_unlabeled.get_pseudo_labels(e, ema, **{k: v.clone() for k, v in args.items()},
                                                  config_dict=dict(conf, dataset_config=cfg))
        return [o.clone() for o in out], e
    out_r, ep_r = run(ref, cfg_r)
    cabi = importlib.import_module("3dioumatch_b200._cabi")
    n0 = cabi.launch_count()
    out_f, ep_f = run(fast, cfg_f)
    assert cabi.launch_count() - n0 == 2                                            # box extents + suppression, whole batch
    names = ("label_mask", "center_label", "sem_cls_label", "heading_label", "heading_residual_label", "size_label",
             "size_residual_label", "false_center_label", "iou_label")
    kept = int(out_r[0].sum())
    assert 0 < kept < n_unl * 64 and kept < int((out_r[1][:, :, 0] > -999).sum()) + 1
    for n, a, b in zip(names, out_r, out_f):
        assert a.dtype == b.dtype and torch.equal(a, b), n
    assert torch.equal(ep_r["pseudo_gt_ratio"], ep_f["pseudo_gt_ratio"])
    # the suppression did remove candidates (otherwise the test would not exercise it)
    no_lhs = ref.loss_unlabeled.get_pseudo_labels(dict(ep), ema, **{k: v.clone() for k, v in args.items()},
                                                  config_dict=dict(conf, dataset_config=cfg_r, use_lhs=False))
    assert int(no_lhs[0].sum()) > kept
