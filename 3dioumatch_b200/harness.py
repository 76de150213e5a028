"""Benchmark / integration harness: the VoteNet-with-IoU-branch forward DATAFLOW on a pluggable operator stack.

This is not a re-implementation of the reference's training code.  It strings the hot-path operator modules
together in the order and shapes of models/votenet_iou_branch.py:75-151 (backbone_module.py:83-133 -> voting_module.py:38-65
-> proposal_module.py:90-123 -> calculate_bbox :111-137 -> grid_conv_module.py:48-116) and
models/loss_helper_iou.py:95-111 (IoU labels against 64 padded GT boxes), with random-initialised weights, so that the
SAME module graph can be timed on
  * this package's drop-in operator stack (`stack_b200()`), and
  * the unmodified reference operator stack installed in oracle/_ref (bench.py --impl reference).
`ops` is a namespace with: modules (pointnet2_modules), utils (pointnet2_utils), pt (pytorch_utils), iou (iou3d_nms_utils).
"""
import importlib
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def stack_b200():
    pkg = importlib.import_module(__name__.rsplit(".", 1)[0])
    pkg.install_dropin()
    import pointnet2.pointnet2_utils as utils
    import pointnet2.pytorch_utils as pt
    import pointnet2_modules as modules
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iou
    return types.SimpleNamespace(name="b200", modules=modules, utils=utils, pt=pt, iou=iou)


def stack_from_path(root):
    """Operator stack found under `root` (e.g. oracle/_ref: the reference's own python + extensions)."""
    for p in (os.path.join(root, "pointnet2"), root):
        sys.path.insert(0, p)
    import pointnet2.pointnet2_utils as utils
    import pointnet2.pytorch_utils as pt
    import pointnet2_modules as modules
    from pcdet.ops.iou3d_nms import iou3d_nms_utils as iou
    return types.SimpleNamespace(name="reference", modules=modules, utils=utils, pt=pt, iou=iou)


class Backbone(nn.Module):
    """SA1..SA4 + FP1, FP2 with the hyper-parameters of models/backbone_module.py:35-72."""

    def __init__(self, ops, input_feature_dim=1):
        super().__init__()
        SA, FP = ops.modules.PointnetSAModuleVotes, ops.modules.PointnetFPModule
        self.sa1 = SA(npoint=2048, radius=0.2, nsample=64, mlp=[input_feature_dim, 64, 64, 128], use_xyz=True,
                      normalize_xyz=True)
        self.sa2 = SA(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.sa3 = SA(npoint=512, radius=0.8, nsample=16, mlp=[256, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.sa4 = SA(npoint=256, radius=1.2, nsample=16, mlp=[256, 128, 128, 256], use_xyz=True, normalize_xyz=True)
        self.fp1 = FP(mlp=[256 + 256, 256, 256])
        self.fp2 = FP(mlp=[256 + 256, 256, 256])
        self.ops = ops
        self.prefetch = True      # issue the FPS index chain ahead of the feature path on a side stream
        self._side = {}
        self.proposal_inds = None

    def sample_chain(self, xyz, num_proposal):
        """The sampling indices of every level depend on coordinates only (FPS of FPS of ...), never on features, so
        the whole chain can be issued ahead of the feature path on a side stream and handed to the SA modules through
        their `inds` argument (pointnet2_modules.py:239-242).  Returns [(inds, ready_event)] per level + proposals."""
        fps, gather = self.ops.utils.furthest_point_sample, self.ops.utils.gather_operation
        main = torch.cuda.current_stream()
        side = self._side.setdefault(main.cuda_stream, torch.cuda.Stream())
        side.wait_stream(main)
        out = []
        with torch.cuda.stream(side):
            cur = xyz
            for level, m in enumerate((2048, 1024, 512, 256)):
                inds = fps(cur, m)
                ev = torch.cuda.Event()
                ev.record(side)
                out.append((inds, ev))
                cur = gather(cur.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
                if level == 1:
                    seeds = cur
            pinds = fps(seeds, num_proposal)
            ev = torch.cuda.Event()
            ev.record(side)
            out.append((pinds, ev))
        for inds, _ in out:
            inds.record_stream(main)
        return out

    def forward(self, pc, num_proposal=None):
        xyz = pc[..., 0:3].contiguous()
        feats = pc[..., 3:].transpose(1, 2).contiguous() if pc.size(-1) > 3 else None
        chain = self.sample_chain(xyz, num_proposal) if (self.prefetch and num_proposal) else None
        main = torch.cuda.current_stream()

        def inds_of(level):
            if chain is None:
                return None
            inds, ev = chain[level]
            main.wait_event(ev)
            return inds
        x1, f1, i1 = self.sa1(xyz, feats, inds_of(0))
        x2, f2, _ = self.sa2(x1, f1, inds_of(1))
        x3, f3, _ = self.sa3(x2, f2, inds_of(2))
        x4, f4, _ = self.sa4(x3, f3, inds_of(3))
        f = self.fp1(x3, x4, f3, f4)
        f = self.fp2(x2, x3, f2, f)
        self.proposal_inds = inds_of(4)
        return x2, f, i1[:, :x2.shape[1]]  # seeds, seed features, seed indices into the input cloud


class Voting(nn.Module):
    """models/voting_module.py:16-65 (vote_factor 1): three 1x1 conv1d, xyz offset + residual features."""

    def __init__(self, dim=256):
        super().__init__()
        self.conv1, self.conv2 = nn.Conv1d(dim, dim, 1), nn.Conv1d(dim, dim, 1)
        self.conv3 = nn.Conv1d(dim, 3 + dim, 1)
        self.bn1, self.bn2 = nn.BatchNorm1d(dim), nn.BatchNorm1d(dim)

    def forward(self, seed_xyz, seed_features):
        net = F.relu(self.bn1(self.conv1(seed_features)))
        net = F.relu(self.bn2(self.conv2(net)))
        net = self.conv3(net).transpose(2, 1)
        vote_xyz = seed_xyz + net[:, :, 0:3]
        vote_features = (seed_features.transpose(2, 1) + net[:, :, 3:]).transpose(2, 1).contiguous()
        return vote_xyz.contiguous(), vote_features


class VoteNetPath(nn.Module):
    def __init__(self, ops, num_class=18, num_heading_bin=1, num_size_cluster=18, input_feature_dim=1,
                 num_proposal=256, mean_size_seed=0):
        super().__init__()
        self.ops = ops
        self.block_diagonal_iou = True   # use the batched block-diagonal IoU entry when the stack offers it
        self.K, self.NH, self.NS, self.NC = num_proposal, num_heading_bin, num_size_cluster, num_class
        rng = np.random.default_rng(mean_size_seed)
        self.register_buffer("mean_size", torch.from_numpy((rng.random((num_size_cluster, 3)) + 0.3).astype(np.float32)))
        self.backbone = Backbone(ops, input_feature_dim)
        self.vgen = Voting(256)
        self.vote_aggregation = ops.modules.PointnetSAModuleVotes(npoint=num_proposal, radius=0.3, nsample=16,
                                                                  mlp=[256, 128, 128, 128], use_xyz=True,
                                                                  normalize_xyz=True)
        out = 2 + 3 + num_heading_bin * 2 + num_size_cluster * 4 + num_class
        self.conv1, self.conv2, self.conv3 = nn.Conv1d(128, 128, 1), nn.Conv1d(128, 128, 1), nn.Conv1d(128, out, 1)
        self.bn1, self.bn2 = nn.BatchNorm1d(128), nn.BatchNorm1d(128)
        # IoU branch (grid_conv_module.py:38-44)
        self.mlp_before_iou = ops.pt.SharedMLP([256 + 3, 128, 128, 128], bn=True)
        self.conv1_iou, self.conv2_iou = nn.Conv1d(128, 128, 1), nn.Conv1d(128, 128, 1)
        self.conv3_iou = nn.Conv1d(128, 3 + num_heading_bin * 2 + num_size_cluster * 3 + num_class, 1)
        self.bn1_iou, self.bn2_iou = nn.BatchNorm1d(128), nn.BatchNorm1d(128)

    # ---- proposal decode (proposal_module.py:24-54, votenet_iou_branch.py:111-137) ----------------------------
    def decode(self, net, base_xyz):
        t = net.transpose(2, 1)
        NH, NS = self.NH, self.NS
        center = base_xyz + t[:, :, 2:5]
        heading_scores = t[:, :, 5:5 + NH]
        heading_res = t[:, :, 5 + NH:5 + 2 * NH] * (math.pi / NH)
        size_scores = t[:, :, 5 + 2 * NH:5 + 2 * NH + NS]
        size_res = (F.softplus(t[:, :, 5 + 2 * NH + NS:5 + 2 * NH + 4 * NS].reshape(t.shape[0], t.shape[1], NS, 3)) - 1)
        size_res = size_res * self.mean_size[None, None]
        size_cls = size_scores.argmax(-1)
        size = (self.mean_size[size_cls] + torch.gather(size_res, 2, size_cls[..., None, None].expand(-1, -1, -1, 3))
                .squeeze(2)) / 2                                   # half sizes
        size = torch.where(size < 0, torch.full_like(size, 1e-6), size)
        hcls = heading_scores.argmax(-1)
        heading = hcls.float() * (2 * math.pi / NH) + torch.gather(heading_res, 2, hcls[..., None]).squeeze(2)
        return center, size, heading, t[:, :, 0:2]

    # ---- IoU branch (grid_conv_module.py:48-116) -------------------------------------------------------------
    def grid_conv(self, center, size, heading, seed_xyz, seed_features):
        ops = self.ops
        B, K = size.shape[:2]
        g = torch.linspace(-1, 1, 4, device=size.device)
        gx, gy, gz = torch.meshgrid(g, g, g, indexing="ij")
        unit = torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], -1)            # (64,3)
        grid = unit[None, None] * size[:, :, None, :]                                          # (B,K,64,3)
        c, s = torch.cos(heading), torch.sin(heading)
        zeros, ones = torch.zeros_like(c), torch.ones_like(c)
        rot = torch.stack([c, s, zeros, -s, c, zeros, zeros, zeros, ones], -1).view(B * K, 3, 3)  # rot_gpu, box_util.py:292-306
        grid = torch.bmm(grid.view(B * K, 64, 3), rot.transpose(1, 2)).view(B, K, 64, 3) + center[:, :, None, :]
        whole = grid.view(B, K * 64, 3).contiguous()
        feat_dim = seed_features.shape[1]
        _, idx = ops.utils.three_nn(whole, seed_xyz)                                           # (B,K*64,3)
        nbr = torch.gather(seed_xyz, 1, idx.view(B, -1, 1).expand(-1, -1, 3).long())           # (B,K*64*3,3)
        d = nbr - whole[:, :, None, :].expand(-1, -1, 3, -1).reshape(B, -1, 3)
        dist = torch.sqrt((d * d).sum(2))
        w = (1 / (dist + 1e-8)).view(B, -1, 3)
        w = (w / w.sum(2, keepdim=True)).contiguous()
        rel = whole - center[:, :, None, :].expand(-1, -1, 64, -1).reshape(B, -1, 3)
        if hasattr(ops.utils, "grid_interp_mlp_max"):     # fused sampler + MLP + max
            x = ops.utils.grid_interp_mlp_max(seed_features, idx, w, rel.contiguous(), 64, self.mlp_before_iou)
        else:
            interp = ops.utils.three_interpolate(seed_features, idx, w)                         # (B,C,K*64)
            x = torch.cat([rel.transpose(1, 2).contiguous().view(B, 3, K, 64), interp.view(B, feat_dim, K, 64)], 1)
            x = self.mlp_before_iou(x)
            x = F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)
        net = F.relu(self.bn1_iou(self.conv1_iou(x)))
        net = F.relu(self.bn2_iou(self.conv2_iou(net)))
        return self.conv3_iou(net).transpose(2, 1)[:, :, -self.NC:]

    def forward(self, point_clouds, gt_boxes):
        """point_clouds (B,N,3+C); gt_boxes (B,G,7) [x,y,z,dx,dy,dz,heading] -> dict of proposal / IoU tensors."""
        ops = self.ops
        seed_xyz, seed_features, seed_inds = self.backbone(point_clouds, self.K)
        vote_xyz, vote_features = self.vgen(seed_xyz, seed_features)
        vote_features = vote_features / torch.norm(vote_features, p=2, dim=1, keepdim=True)
        sample_inds = self.backbone.proposal_inds                                                # 'seed_fps'
        if sample_inds is None:
            sample_inds = ops.utils.furthest_point_sample(seed_xyz, self.K)
        agg_xyz, agg_feat, _ = self.vote_aggregation(vote_xyz, vote_features, sample_inds)
        net = F.relu(self.bn1(self.conv1(agg_feat)))
        net = F.relu(self.bn2(self.conv2(net)))
        net = self.conv3(net)
        center, size, heading, objectness = self.decode(net, agg_xyz)
        iou_scores = self.grid_conv(center.detach(), size.detach(), heading.detach(), seed_xyz.detach(),
                                    seed_features.detach())
        # IoU labels: all (B*K) x (B*G) pairs, then the per-scene diagonal blocks (loss_helper_iou.py:95-111)
        B, K, G = center.shape[0], self.K, gt_boxes.shape[1]
        pred = torch.cat([center, size * 2, -heading[..., None]], 2)
        pa, pb = pred.view(-1, 7).contiguous(), gt_boxes.reshape(-1, 7).contiguous()
        if hasattr(ops.iou, "boxes_iou3d_batched") and self.block_diagonal_iou:
            # only the per-scene diagonal blocks of the all-pairs matrix are ever consumed
            # (loss_helper_iou.py:107-109) -> evaluate exactly those (B, K, G) pairs in one launch
            blk = ops.iou.boxes_iou3d_batched(pred.contiguous(), gt_boxes.contiguous())
            iou_labels, assignment = blk.max(dim=2)
            return dict(seed_xyz=seed_xyz, seed_inds=seed_inds, vote_xyz=vote_xyz, aggregated_vote_inds=sample_inds,
                        center=center, size=size, heading=heading, objectness=objectness, iou_scores=iou_scores,
                        iou_labels=iou_labels, object_assignment=assignment, pred_bbox=pred)
        if ops.name == "reference":
            # the reference launches its IoU kernel on the LEGACY default stream (iou3d_nms_kernel.cu:396); when the
            # harness runs on a non-blocking stream that launch must be ordered by hand
            cur, legacy = torch.cuda.current_stream(), torch.cuda.default_stream()
            if cur != legacy:
                legacy.wait_stream(cur)
                with torch.cuda.stream(legacy):
                    iou = ops.iou.boxes_iou3d_gpu(pa, pb)
                cur.wait_stream(legacy)
                iou.record_stream(cur)
            else:
                iou = ops.iou.boxes_iou3d_gpu(pa, pb)
        else:
            iou = ops.iou.boxes_iou3d_gpu(pa, pb)
        iou_labels, assignment = iou.view(B * K, B, G).max(dim=2)
        sel = torch.arange(B, device=iou.device).unsqueeze(1).expand(-1, K).reshape(-1, 1)
        iou_labels = iou_labels.gather(1, sel).view(B, K)
        assignment = assignment.gather(1, sel).view(B, K)
        return dict(seed_xyz=seed_xyz, seed_inds=seed_inds, vote_xyz=vote_xyz, aggregated_vote_inds=sample_inds,
                    center=center, size=size, heading=heading, objectness=objectness, iou_scores=iou_scores,
                    iou_labels=iou_labels, object_assignment=assignment, pred_bbox=pred)


def make_model(ops, seed=1, num_proposal=256, device="cuda"):
    """Random-initialised weights (torch.manual_seed(seed)) + non-trivial BN statistics, eval mode."""
    torch.manual_seed(seed)
    net = VoteNetPath(ops, num_proposal=num_proposal)
    gen = torch.Generator().manual_seed(seed + 1)
    for m in net.modules():
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d)):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=gen) * 0.5 + 0.75)
    return net.to(device).eval()


def make_inputs(B=8, N=40000, G=64, seed=0, room=(8.0, 8.0, 3.0)):
    """ScanNet-shaped synthetic scenes (tests/cases.py:scene_cloud) + 64-slot padded GT boxes (numpy, host)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tests"))
    import cases
    pc = cases.scene_cloud(seed, B, N, room=room)
    gt = np.zeros((B, G, 7), np.float32)
    rng = np.random.default_rng(seed + 100)
    for b in range(B):
        n = int(rng.integers(3, 13))
        bx = cases.boxes(seed * 1000 + b, n, extent=(6.0, 6.0, 2.0))
        bx[:, 0:2] -= 3.0
        gt[b, :n] = bx
        gt[b, n:, 0:3] = -1000.0       # padded slots (loss_helper_iou.py:56-58)
        gt[b, n:, 3:6] = 1.0
    return pc, gt
