"""Drop-in mirror of models/voting_module.py (VotingModule, :16-65): same constructor, same child modules and state-dict
keys (conv1, conv2, conv3, bn1, bn2), same outputs.  In eval mode without gradients the three 1x1 convolutions run as
fused row MLPs on the tensor-core kernel (no cuDNN / cuBLAS launch); otherwise the torch formulation of the reference."""
import torch
import torch.nn as nn

import pointnet2._ext as _ext
from _b200_rows import Head, fusable, torch_head


class VotingModule(nn.Module):
    def __init__(self, vote_factor, seed_feature_dim):
        super().__init__()
        self.vote_factor = vote_factor
        self.in_dim = seed_feature_dim
        self.out_dim = self.in_dim  # residual features: in_dim == out_dim
        self.conv1 = torch.nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv2 = torch.nn.Conv1d(self.in_dim, self.in_dim, 1)
        self.conv3 = torch.nn.Conv1d(self.in_dim, (3 + self.out_dim) * self.vote_factor, 1)
        self.bn1 = torch.nn.BatchNorm1d(self.in_dim)
        self.bn2 = torch.nn.BatchNorm1d(self.in_dim)
        object.__setattr__(self, "_b200_head", Head([(self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, None)]))

    def forward(self, seed_xyz, seed_features):
        """seed_xyz (B,n,3), seed_features (B,C,n) -> vote_xyz (B,n*vf,3), vote_features (B,C,n*vf)."""
        B, n = seed_xyz.shape[0], seed_xyz.shape[1]
        if self.vote_factor == 1 and fusable(seed_features, self) and self.in_dim % 4 == 0 and self.in_dim <= 256:
            rows = _ext.transpose_cn(seed_features.contiguous())                        # (B, n, C)
            (_, offset), (residual, _) = self._b200_head(rows, want_cm=True, want_pm=True, split_last=[3])
            vote_xyz = seed_xyz + offset                                                 # (B, n, 3)
            vote_features = seed_features + residual                                     # (B, C, n)
            return vote_xyz.contiguous(), vote_features.contiguous()
        net = torch_head(seed_features, self._b200_head.pairs)
        net = net.transpose(2, 1).view(B, n, self.vote_factor, 3 + self.out_dim)
        vote_xyz = (seed_xyz.unsqueeze(2) + net[:, :, :, 0:3].contiguous()).contiguous().view(B, n * self.vote_factor, 3)
        vote_features = seed_features.transpose(2, 1).unsqueeze(2) + net[:, :, :, 3:]
        vote_features = vote_features.contiguous().view(B, n * self.vote_factor, self.out_dim).transpose(2, 1).contiguous()
        return vote_xyz, vote_features
