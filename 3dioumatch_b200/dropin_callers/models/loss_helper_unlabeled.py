"""Drop-in mirror of models/loss_helper_unlabeled.py: every name of the reference module is
re-exported unchanged; `get_pseudo_labels` keeps the reference's code for everything except the pseudo-label
suppression, which runs on the device.

The reference filters the teacher's boxes on the HOST (loss_helper_unlabeled.py:441-492): per unlabeled scene and per box
it copies heading / size values back one scalar at a time (four device->host copies per box: 2048 blocking copies per
SSL step at 8 unlabeled scenes x 64 boxes), builds the 8 corners in numpy, takes their min / max in a Python loop and
runs lhs_3d_faster_samecls.  Here the same boxes go through two launches for the whole batch -- b200nms_box_extents
(corners + extents, float64 math / float32 storage like the numpy code) and b200nms_aabb_suppress (lower-half
suppression, float64 overlap decisions in numpy's operation order) -- and nothing is copied to the host.

How: the reference function is called with use_lhs=False, which runs all of its code but the host loop; the suppression
result is then applied the way :492 applies it (slots that are not picked lose their label, their centre goes to
-1000).  The statistics path (view_stats) and use_lhs=False calls go to the reference function untouched.  Pick lists
are identical to the reference's for distinct scores (csrc/aabb_nms.cu header; tests/test_gpu_aabb_nms.py,
tests/test_gpu_sa_train.py::test_pseudo_label_filter_mirror)."""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn

from utils import nms as _nms


def _load_reference_module():
    here = os.path.abspath(__file__)
    for p in sys.path:
        cand = os.path.join(p, "models", "loss_helper_unlabeled.py")
        if os.path.isfile(cand) and os.path.abspath(cand) != here:
            spec = importlib.util.spec_from_file_location("models._reference_loss_helper_unlabeled", cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("the reference's models/loss_helper_unlabeled.py is not on sys.path")


_ref = _load_reference_module()
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
_reference_get_pseudo_labels = _ref.get_pseudo_labels
MAX_NUM_OBJ = _ref.MAX_NUM_OBJ


def _top_slots(ema_end_points, end_points, pred_sem_cls, pred_objectness, config_dict):
    """The slot order of :371-425 (objectness / class / IoU thresholds, the MAX_NUM_OBJ best by pos_obj * max_cls among
    the boxes that pass), recomputed from the same tensors with the same torch calls -> the same `inds`."""
    pos_obj = nn.Softmax(dim=2)(pred_objectness)[:, :, 1]
    max_cls, argmax_cls = torch.max(nn.Softmax(dim=2)(pred_sem_cls), dim=2)
    unsup = torch.nonzero(1 - end_points['supervised_mask']).squeeze(1).long()
    iou_pred = nn.Sigmoid()(ema_end_points['iou_scores'][unsup, ...])
    iou_pred = torch.gather(iou_pred, 2, argmax_cls.unsqueeze(-1)).squeeze(-1) if iou_pred.shape[2] > 1 else iou_pred.squeeze(-1)
    keep = torch.logical_and(torch.logical_and(max_cls > config_dict['cls_threshold'], pos_obj > config_dict['obj_threshold']),
                             iou_pred > config_dict['iou_threshold'])
    inds = torch.argsort(pos_obj * max_cls * keep, dim=1, descending=True)[:, :MAX_NUM_OBJ].long()
    return inds, pos_obj, argmax_cls, iou_pred


def get_pseudo_labels(end_points, ema_end_points, pred_center, pred_sem_cls, pred_objectness, pred_heading_scores,
                      pred_heading_residuals, pred_size_scores, pred_size_residuals, pred_vote_xyz, config_dict):
    if not config_dict['use_lhs'] or config_dict['view_stats']:
        return _reference_get_pseudo_labels(end_points, ema_end_points, pred_center, pred_sem_cls, pred_objectness,
                                            pred_heading_scores, pred_heading_residuals, pred_size_scores,
                                            pred_size_residuals, pred_vote_xyz, config_dict)
    no_lhs = dict(config_dict)
    no_lhs['use_lhs'] = False
    (label_mask, center_label, sem_cls_label, heading_label, heading_residual_label, size_label, size_residual_label,
     false_center_label, iou_label) = _reference_get_pseudo_labels(
        end_points, ema_end_points, pred_center, pred_sem_cls, pred_objectness, pred_heading_scores,
        pred_heading_residuals, pred_size_scores, pred_size_residuals, pred_vote_xyz, no_lhs)

    # ---- the boxes of :441-483 for every slot, on the device ---------------------------------------------------------
    cfg = config_dict['dataset_config']
    inds, pos_obj, _, _ = _top_slots(ema_end_points, end_points, pred_sem_cls, pred_objectness, config_dict)
    centre = torch.gather(pred_center, 1, inds.unsqueeze(-1).expand(-1, -1, 3))
    mean_size = torch.from_numpy(np.asarray(cfg.mean_size_arr, np.float64)).to(centre.device)
    size = mean_size[size_label] + size_residual_label.double()            # class2size in numpy: float64 + float32
    if cfg.num_heading_bin > 1:                                             # class2angle in numpy (sunrgbd :110-120)
        heading = heading_label.double() * (2 * np.pi / float(cfg.num_heading_bin)) + heading_residual_label.double()
        heading = heading - 2 * np.pi * (heading > np.pi)
    else:                                                                   # ScanNet boxes are axis aligned (:60-64)
        heading = torch.zeros(heading_label.shape, dtype=torch.float64, device=centre.device)
    _, extents = _nms.box_extents_batch(centre, size, heading, return_corners=False)
    score = (torch.gather(pos_obj, 1, inds).float() * iou_label.float()).double()       # float32 product, like :480
    boxes = torch.cat([extents.double(), score.unsqueeze(-1), sem_cls_label.double().unsqueeze(-1)], dim=-1)
    _, _, picked = _nms.suppress_batch(boxes, config_dict['nms_iou'], use_cls=True, lhs=True,
                                       old_type=config_dict['use_old_type_nms'])
    # ---- :492 + :528-533: slots the suppression drops lose their label -----------------------------------------------
    label_mask = label_mask * picked.long()
    center_label[(1 - label_mask).unsqueeze(-1).expand(-1, -1, 3).bool()] = -1000
    return (label_mask, center_label, sem_cls_label, heading_label, heading_residual_label, size_label,
            size_residual_label, false_center_label, iou_label)


# get_unlabeled_loss (reference code, :541-600) looks the function up in the reference module's globals
_ref.get_pseudo_labels = get_pseudo_labels
