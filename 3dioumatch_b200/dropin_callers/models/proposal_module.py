"""Drop-in mirror of models/proposal_module.py (decode_scores :24-54, ProposalModule :57-125): same constructor, child
modules and state-dict keys (vote_aggregation, conv1-3, bn1-2), same end_points.  The vote-aggregation SA layer hands its
output over point-major and the three-layer proposal head runs as ONE fused row-MLP launch in eval mode."""
import numpy as np
import torch
import torch.nn as nn

import pointnet2._ext as _ext
from pointnet2 import pointnet2_utils
from pointnet2.pointnet2_modules import PointnetSAModuleVotes
from _b200_rows import Head, fusable, torch_head


def decode_scores(net, end_points, num_class, num_heading_bin, num_size_cluster, mean_size_arr):
    """net (B, 2+3+NH*2+NS*4+NC, K) -> objectness / center / heading / size / class entries of end_points
    (channel layout of proposal_module.py:24-54)."""
    t = net.transpose(2, 1)
    B, K = t.shape[0], t.shape[1]
    NH, NS = num_heading_bin, num_size_cluster
    end_points['objectness_scores'] = t[:, :, 0:2]
    end_points['center'] = end_points['aggregated_vote_xyz'] + t[:, :, 2:5]
    end_points['heading_scores'] = t[:, :, 5:5 + NH]
    hres = t[:, :, 5 + NH:5 + NH * 2]
    end_points['heading_residuals_normalized'] = hres                    # -1 .. 1
    end_points['heading_residuals'] = hres * (np.pi / NH)
    end_points['size_scores'] = t[:, :, 5 + NH * 2:5 + NH * 2 + NS]
    sres = t[:, :, 5 + NH * 2 + NS:5 + NH * 2 + NS * 4].view([B, K, NS, 3])
    sres = torch.nn.functional.softplus(sres) - 1
    end_points['size_residuals_normalized'] = sres
    end_points['size_residuals'] = sres * torch.from_numpy(mean_size_arr.astype(np.float32)).cuda().unsqueeze(0).unsqueeze(0)
    end_points['sem_cls_scores'] = t[:, :, 5 + NH * 2 + NS * 4:]
    return end_points


class ProposalModule(nn.Module):
    def __init__(self, num_class, num_heading_bin, num_size_cluster, mean_size_arr, num_proposal, sampling,
                 seed_feat_dim=256, query_feats='seed'):
        super().__init__()
        self.num_class = num_class
        self.num_heading_bin = num_heading_bin
        self.num_size_cluster = num_size_cluster
        self.mean_size_arr = mean_size_arr
        self.num_proposal = num_proposal
        self.sampling = sampling
        self.seed_feat_dim = seed_feat_dim
        self.query_feats = query_feats
        self.vote_aggregation = PointnetSAModuleVotes(npoint=self.num_proposal, radius=0.3, nsample=16,
                                                      mlp=[self.seed_feat_dim, 128, 128, 128], use_xyz=True,
                                                      normalize_xyz=True)
        self.conv1 = torch.nn.Conv1d(128, 128, 1)
        self.conv2 = torch.nn.Conv1d(128, 128, 1)
        self.conv3 = torch.nn.Conv1d(128, 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + self.num_class, 1)
        self.bn1 = torch.nn.BatchNorm1d(128)
        self.bn2 = torch.nn.BatchNorm1d(128)
        object.__setattr__(self, "_b200_head", Head([(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, None)]))

    def forward(self, xyz, features, end_points):
        """xyz (B,n,3) votes, features (B,C,n) -> proposals decoded into end_points."""
        if self.sampling == 'vote_fps':
            xyz, features, sample_inds = self.vote_aggregation(xyz, features)
        elif self.sampling == 'seed_fps':
            sample_inds = end_points.pop('_b200_proposal_inds', None)   # prefetched by the backbone mirror
            if sample_inds is None:
                sample_inds = pointnet2_utils.furthest_point_sample(end_points['seed_xyz'], self.num_proposal)
            xyz, features, _ = self.vote_aggregation(xyz, features, sample_inds)
        elif self.sampling == 'random':
            B, num_seed = end_points['seed_xyz'].shape[:2]
            sample_inds = torch.randint(0, num_seed, (B, self.num_proposal), dtype=torch.int).cuda()
            xyz, features, _ = self.vote_aggregation(xyz, features, sample_inds)
        else:
            raise ValueError('Unknown sampling strategy: %s' % (self.sampling,))
        end_points['aggregated_vote_xyz'] = xyz
        end_points['aggregated_vote_inds'] = sample_inds
        if fusable(features, self):
            net, _ = self._b200_head(_ext.transpose_cn(features.contiguous()), want_cm=True)
        else:
            net = torch_head(features, self._b200_head.pairs)
        return decode_scores(net, end_points, self.num_class, self.num_heading_bin, self.num_size_cluster, self.mean_size_arr)
