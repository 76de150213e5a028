"""Drop-in mirror of models/grid_conv_module.py (GridConv, :22-116): same constructor, child modules
and state-dict keys (mlp_before_iou, conv1_iou-3, bn1_iou-2), same end_points['iou_scores'].

The reference builds the 4x4x4 grid of every box, finds the three nearest seeds of every grid point, then MATERIALISES the
gathered neighbour features (B, K*64*3, 256) with a python list comprehension over the batch, blends them, concatenates the
relative grid coordinates, runs SharedMLP[259,128,128,128] and max-pools over the 64 grid points.  Here the grid and the
weights are computed with the same torch expressions (so the same floats), and everything from the gather to the max-pool
is ONE fused tensor-core kernel (pointnet2_utils.grid_interp_mlp_max -> b200pn2_interp_mlp_forward); the three-layer IoU
head is one fused row-MLP launch."""
import torch
import torch.nn as nn

from utils.box_util import rot_gpu

import pointnet2._ext as _ext
import pointnet2.pointnet2_utils as pointnet2_utils
import pointnet2.pytorch_utils as pt_utils
from _b200_rows import Head, fusable, torch_head


class GridConv(nn.Module):
    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal, sampling,
                 seed_feat_dim=256, query_feats='seed', iou_class_depend=True):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.query_feats = query_feats
        self.iou_class_depend = iou_class_depend
        self.iou_size = num_class if self.iou_class_depend else 1
        self.mlp_before_iou = pt_utils.SharedMLP([self.seed_feat_dim + 3, 128, 128, 128], bn=True)
        self.conv1_iou = torch.nn.Conv1d(128, 128, 1)
        self.conv2_iou = torch.nn.Conv1d(128, 128, 1)
        self.conv3_iou = torch.nn.Conv1d(128, 3 + num_heading_bin * 2 + num_size_cluster * 3 + self.iou_size, 1)
        self.bn1_iou = torch.nn.BatchNorm1d(128)
        self.bn2_iou = torch.nn.BatchNorm1d(128)
        object.__setattr__(self, "_b200_head", Head([(self.conv1_iou, self.bn1_iou), (self.conv2_iou, self.bn2_iou),
                                                     (self.conv3_iou, None)]))

    def forward(self, center, size, heading, end_points):
        if self.query_feats == 'vote':
            origin_xyz, origin_features = end_points['vote_xyz'], end_points['vote_features']
        elif self.query_feats == 'seed':
            origin_xyz, origin_features = end_points['seed_xyz'], end_points['seed_features']
        elif self.query_feats == 'seed+vote':
            origin_xyz, origin_features = end_points['seed_xyz'], end_points['vote_features']
        else:
            raise NotImplementedError()
        origin_features = origin_features.detach().contiguous()
        origin_xyz = origin_xyz.detach().contiguous()
        B, K = size.shape[:2]
        G = 4
        # the 4x4x4 lattice of every box, x slowest (grid_conv_module.py:65-76): same products, same bmm, same sums
        step = torch.linspace(-1, 1, G).cuda()
        gx = step.view(G, 1, 1).repeat(1, G, G).view(1, 1, -1)
        gy = step.view(1, G, 1).repeat(G, 1, G).view(1, 1, -1)
        gz = step.view(1, 1, G).repeat(G, G, 1).view(1, 1, -1)
        whole_grid = torch.stack([gx * size[:, :, 0:1], gy * size[:, :, 1:2], gz * size[:, :, 2:3]], dim=-1)  # (B,K,64,3)
        rot_mat = rot_gpu(heading).view(-1, 3, 3)
        whole_grid = torch.bmm(whole_grid.view(B * K, -1, 3), rot_mat.transpose(1, 2)).view(B, K, -1, 3)
        whole_grid = whole_grid + center.unsqueeze(2).expand(-1, -1, G * G * G, -1)
        whole_grid = whole_grid.view(B, -1, 3).contiguous()

        _, idx = pointnet2_utils.three_nn(whole_grid, origin_xyz)                       # (B, K*64, 3)
        # inverse-distance weights from the gathered coordinates, as the reference computes them (:89-101)
        nbr = torch.gather(origin_xyz, dim=1, index=idx.view(B, -1, 1).expand(-1, -1, 3).long())
        diff = nbr - whole_grid.unsqueeze(2).expand(-1, -1, 3, -1).contiguous().view(B, -1, 3)
        dist = torch.sqrt(torch.sum(diff * diff, dim=2))
        weight = (1 / (dist + 1e-8)).view(B, -1, 3)
        weight = (weight / torch.sum(weight, dim=2, keepdim=True)).contiguous()
        relative_grid = whole_grid - center.unsqueeze(2).expand(-1, -1, G * G * G, -1).contiguous().view(B, -1, 3)

        # gather -> blend -> concat(relative grid) -> SharedMLP -> max over the 64 grid points: one fused kernel
        iou_features = pointnet2_utils.grid_interp_mlp_max(origin_features, idx, weight, relative_grid.contiguous(),
                                                           G * G * G, self.mlp_before_iou)   # (B, 128, K)
        if fusable(iou_features, self):
            net_iou, _ = self._b200_head(_ext.transpose_cn(iou_features.contiguous()), want_cm=True)
        else:
            net_iou = torch_head(iou_features, self._b200_head.pairs)
        end_points['iou_scores'] = net_iou.transpose(2, 1)[:, :, -self.iou_size:]
        return end_points
