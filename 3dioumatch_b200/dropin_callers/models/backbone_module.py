"""Drop-in mirror of models/backbone_module.py (Pointnet2Backbone, :21-133): same constructor, child modules (sa1-sa4, fp1,
fp2) and state-dict keys, same end_points.

One scheduling change: the sampling indices of every level depend on COORDINATES only (FPS of FPS of ...), never on
features, so the whole index chain -- and the proposal module's seed FPS -- is issued ahead of the feature path on a
side stream and handed to the SA modules through the reference's own `inds` argument (pointnet2_modules.py:239-242).  The
latency-bound FPS kernels (a few SMs each) then overlap the throughput-bound MLP kernels instead of serialising with
them.  Indices and features are identical to the serial order (same kernels, same inputs)."""
import torch
import torch.nn as nn

import pointnet2.pointnet2_utils as pointnet2_utils
from pointnet2.pointnet2_modules import PointnetFPModule, PointnetSAModuleVotes


class Pointnet2Backbone(nn.Module):
    def __init__(self, input_feature_dim=0):
        super().__init__()
        self.sa1 = PointnetSAModuleVotes(npoint=2048, radius=0.2, nsample=64, mlp=[input_feature_dim, 64, 64, 128],
                                         use_xyz=True, normalize_xyz=True)
        self.sa2 = PointnetSAModuleVotes(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.sa3 = PointnetSAModuleVotes(npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.sa4 = PointnetSAModuleVotes(npoint=256, radius=1.2, nsample=16, mlp=[256, 128, 128, 256],
                                         use_xyz=True, normalize_xyz=True)
        self.fp1 = PointnetFPModule(mlp=[256 + 256, 256, 256])
        self.fp2 = PointnetFPModule(mlp=[256 + 256, 256, 256])
        self.prefetch_proposals = 0        # set by VoteNet users that sample proposals with 'seed_fps' (see refapp)
        object.__setattr__(self, "_side", {})

    def _break_up_pc(self, pc):
        xyz = pc[..., 0:3].contiguous()
        features = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        return xyz, features

    def _sample_chain(self, xyz):
        """[(inds, ready event)] for sa1..sa4 (+ the proposal FPS on the sa2 level), issued on a side stream."""
        fps, gather = pointnet2_utils.furthest_point_sample, pointnet2_utils.gather_operation
        main = torch.cuda.current_stream()
        side = self._side.setdefault(main.cuda_stream, torch.cuda.Stream())
        side.wait_stream(main)
        out = []
        with torch.cuda.stream(side):
            cur, seeds = xyz, None
            for level, m in enumerate((self.sa1.npoint, self.sa2.npoint, self.sa3.npoint, self.sa4.npoint)):
                inds = fps(cur, m)
                ev = torch.cuda.Event()
                ev.record(side)
                out.append((inds, ev))
                cur = gather(cur.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
                if level == 1:
                    seeds = cur
            if self.prefetch_proposals:
                pinds = fps(seeds, self.prefetch_proposals)
                ev = torch.cuda.Event()
                ev.record(side)
                out.append((pinds, ev))
        for inds, _ in out:
            inds.record_stream(main)
        return out

    def forward(self, pointcloud, end_points=None):
        if not end_points:
            end_points = {}
        xyz, features = self._break_up_pc(pointcloud)
        chain = self._sample_chain(xyz) if (xyz.is_cuda and not torch.is_grad_enabled()) else None
        main = torch.cuda.current_stream() if chain is not None else None

        def inds_of(level):
            if chain is None:
                return None
            inds, ev = chain[level]
            main.wait_event(ev)
            return inds
        xyz, features, fps_inds = self.sa1(xyz, features, inds_of(0))
        end_points['sa1_inds'], end_points['sa1_xyz'], end_points['sa1_features'] = fps_inds, xyz, features
        xyz, features, fps_inds = self.sa2(xyz, features, inds_of(1))
        end_points['sa2_inds'], end_points['sa2_xyz'], end_points['sa2_features'] = fps_inds, xyz, features
        xyz, features, fps_inds = self.sa3(xyz, features, inds_of(2))
        end_points['sa3_xyz'], end_points['sa3_features'] = xyz, features
        xyz, features, fps_inds = self.sa4(xyz, features, inds_of(3))
        end_points['sa4_xyz'], end_points['sa4_features'] = xyz, features
        features = self.fp1(end_points['sa3_xyz'], end_points['sa4_xyz'], end_points['sa3_features'], end_points['sa4_features'])
        features = self.fp2(end_points['sa2_xyz'], end_points['sa3_xyz'], end_points['sa2_features'], features)
        end_points['fp2_features'] = features
        end_points['fp2_xyz'] = end_points['sa2_xyz']
        num_seed = end_points['fp2_xyz'].shape[1]
        end_points['fp2_inds'] = end_points['sa1_inds'][:, 0:num_seed]
        if chain is not None and self.prefetch_proposals:
            end_points['_b200_proposal_inds'] = inds_of(4)
        return end_points
