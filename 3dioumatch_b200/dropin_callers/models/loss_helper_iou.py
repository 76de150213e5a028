"""Drop-in mirror of models/loss_helper_iou.py: every name of the reference module is re-exported
unchanged, except that the two functions which compute IoU labels against the padded GT boxes evaluate ONLY the pairs the
reference ever consumes.

The reference flattens predictions and GT boxes of the whole batch, computes all (B*K) x (B*G) rotated IoUs, views the
result as (B*K, B, G), takes the max over G and then gathers the diagonal batch block (loss_helper_iou.py:40-49,
:95-111): (B-1)/B of the pairs are thrown away.  Here the (B, K, G) block-diagonal is one launch of
b200iou_boxes_iou3d_batched; max / argmax over G give the same (iou_labels, object_assignment) -- the diagonal blocks of
the all-pairs matrix are bit-identical to the batched result (tests/test_gpu_iou.py)."""
import importlib.util
import os
import sys

import torch

from pcdet.ops.iou3d_nms import iou3d_nms_utils as _iou_utils


def _load_reference_module():
    here = os.path.abspath(__file__)
    for p in sys.path:
        cand = os.path.join(p, "models", "loss_helper_iou.py")
        if os.path.isfile(cand) and os.path.abspath(cand) != here:
            spec = importlib.util.spec_from_file_location("models._reference_loss_helper_iou", cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    raise ImportError("the reference's models/loss_helper_iou.py is not on sys.path")


_ref = _load_reference_module()
globals().update({k: v for k, v in vars(_ref).items() if not k.startswith("__")})
nn_distance = _ref.nn_distance
NEAR_THRESHOLD = _ref.NEAR_THRESHOLD


def _gt_boxes(end_points, inds, dataset_config):
    center_label = end_points['center_label'][inds, ...]
    pad = (1 - end_points['box_label_mask'][inds, ...]).unsqueeze(-1).expand(-1, -1, 3).bool()
    center_label[pad] = -1000                                        # padded slots far away (:56-58), in place like the reference
    gt_size = dataset_config.class2size_gpu(end_points['size_class_label'][inds, ...],
                                            end_points['size_residual_label'][inds, ...])
    gt_angle = dataset_config.class2angle_gpu(end_points['heading_class_label'][inds, ...],
                                              end_points['heading_residual_label'][inds, ...])
    return center_label, torch.cat([center_label, gt_size, -gt_angle[:, :, None]], dim=2)


def _block_diagonal(pred_bbox, gt_bbox):
    iou = _iou_utils.boxes_iou3d_batched(pred_bbox.contiguous(), gt_bbox.contiguous())     # (B, K, G)
    iou_labels, object_assignment = iou.max(dim=2)
    return iou_labels.detach(), object_assignment


def compute_iou_from_given_size(end_points, unsupervised_inds, pred_center, pred_size, pred_heading, config_dict):
    _, gt_bbox = _gt_boxes(end_points, unsupervised_inds, config_dict['dataset_config'])
    pred_size[pred_size <= 0] = 1e-6
    pred_bbox = torch.cat([pred_center, pred_size, -pred_heading[:, :, None]], axis=2)
    end_points['pred_bbox'] = pred_bbox
    iou_labels, object_assignment = _block_diagonal(pred_bbox, gt_bbox)
    return iou_labels, None, object_assignment


def compute_iou_labels(end_points, unsupervised_inds, pred_votes, pred_center, pred_sem_cls, pred_objectness,
                       pred_heading_scores, pred_heading_residuals, pred_size_scores, pred_size_residuals, config_dict,
                       reverse=False):
    if reverse:  # GT-major layout: rarely used, keep the reference's formulation
        return _ref.compute_iou_labels(end_points, unsupervised_inds, pred_votes, pred_center, pred_sem_cls, pred_objectness,
                                       pred_heading_scores, pred_heading_residuals, pred_size_scores, pred_size_residuals,
                                       config_dict, reverse=True)
    cfg = config_dict['dataset_config']
    center_label, gt_bbox = _gt_boxes(end_points, unsupervised_inds, cfg)
    pred_heading_class = torch.argmax(pred_heading_scores, -1)
    pred_heading_residual = torch.gather(pred_heading_residuals, 2, pred_heading_class.unsqueeze(-1)).squeeze(2)
    pred_size_class = torch.argmax(pred_size_scores, -1)
    pred_size_residual = torch.gather(pred_size_residuals, 2,
                                      pred_size_class.unsqueeze(-1).unsqueeze(-1).repeat(1, 1, 1, 3)).squeeze(2)
    # objectness labels: proposals whose vote cluster centre is within NEAR_THRESHOLD of a GT centre (:70-74)
    dist1, _, _, _ = nn_distance(pred_votes, center_label)
    near = torch.sqrt(dist1 + 1e-6) < NEAR_THRESHOLD
    objectness_label = near.long()
    pred_size = cfg.class2size_gpu(pred_size_class.detach(), pred_size_residual)
    pred_size[pred_size <= 0] = 1e-6
    if cfg.num_heading_bin == 1:
        pred_angle = torch.zeros(pred_size.shape[:2], device=pred_size.device)
    else:
        pred_angle = cfg.class2angle_gpu(pred_heading_class.detach(), pred_heading_residual)
    pred_bbox = torch.cat([pred_center, pred_size, -pred_angle[:, :, None]], axis=2)
    end_points['pred_bbox'] = pred_bbox
    iou_labels, object_assignment = _block_diagonal(pred_bbox, gt_bbox)
    return iou_labels, objectness_label, object_assignment
