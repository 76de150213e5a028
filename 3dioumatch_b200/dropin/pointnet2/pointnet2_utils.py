"""Autograd operator surface of the reference's pointnet2/pointnet2_utils.py (:52-426) on the sm_100a kernels.

Same public names, argument orders, dtypes and return shapes:
    furthest_point_sample, gather_operation, three_nn, three_interpolate, grouping_operation, ball_query,
    QueryAndGroup, GroupAll, RandomDropout
The native module is `pointnet2._ext` of this drop-in tree (C ABI of libb200pc.so); nothing here runs on CPU.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

import pointnet2._ext as _ext


class RandomDropout(nn.Module):
    """Feature dropout with a random rate in [0, p) (reference :42-50)."""

    def __init__(self, p=0.5, inplace=False):
        super().__init__()
        self.p, self.inplace = p, inplace

    def forward(self, X):
        theta = float(torch.empty(1).uniform_(0, self.p)[0])
        return nn.functional.dropout(X, theta, self.training, self.inplace) * (1.0 - theta) if self.training else X


class FurthestPointSampling(Function):
    """xyz (B,N,3) f32, npoint -> (B,npoint) int32; not differentiable (reference :52-78)."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint) (reference :84-115)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_points = features.size(2)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n_points), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) = sqrt of squared distances, idx (B,n,3) int32)
    (reference :121-147)."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B,c,m), idx (B,n,3) int32, weight (B,n,3) -> (B,c,n) (reference :153-204).
    The backward is the true scatter-add gradient (see _ext.three_interpolate_grad)."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.m = features.size(2)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        return _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m), None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """features (B,C,N), idx (B,npoint,nsample) int32 -> (B,C,npoint,nsample) (reference :210-255)."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_points = features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n_points), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """(radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> (B,npoint,nsample) int32 (reference :261-289;
    note the native call swaps the two point arguments, :283)."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


def _resample_uniformly(idx, nsample):
    """`sample_uniformly` of the reference (:337-346): keep the unique neighbours of each ball and refill the
    remaining slots by sampling them with replacement.  Host-side, like the reference."""
    unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
    for b in range(idx.shape[0]):
        for r in range(idx.shape[1]):
            uniq = torch.unique(idx[b, r, :])
            k = uniq.shape[0]
            unique_cnt[b, r] = k
            pick = torch.randint(0, k, (nsample - k,), dtype=torch.long)
            idx[b, r, :] = torch.cat((uniq, uniq[pick]))
    return unique_cnt


class QueryAndGroup(nn.Module):
    """Ball query + grouping: returns (B, 3+C, npoint, nsample) [, grouped_xyz][, unique_cnt] (reference :295-377)."""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if ret_unique_cnt:
            assert sample_uniformly

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        unique_cnt = _resample_uniformly(idx, self.nsample) if self.sample_uniformly else None

        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,nsample)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz /= self.radius

        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        else:
            grouped = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped

        extra = ([grouped_xyz] if self.ret_grouped_xyz else []) + ([unique_cnt] if self.ret_unique_cnt else [])
        return (new_features, *extra) if extra else new_features


class GroupAll(nn.Module):
    """One group holding every point: (B, 3+C, 1, N) (reference :380-426)."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz  # the reference accepts but drops this flag (:389-392) ...

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            new_features = grouped_xyz
        else:
            grouped = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, grouped], dim=1) if self.use_xyz else grouped
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features


def grid_interp_mlp_max(known_feats, idx, weight, rel_xyz, nsample, shared_mlp):
    """IoU-branch feature sampler (reference models/grid_conv_module.py:87-113): blend the three nearest seeds with
    `weight`, prepend the relative grid coordinates, run `shared_mlp` (a pytorch_utils.SharedMLP) and max-pool over the
    `nsample` grid points of each box.  known_feats (B,C,m), idx/weight/rel_xyz (B, K*nsample, 3) -> (B, C_out, K).

    One fused tensor-core kernel when the MLP is in eval mode and no gradient is required; otherwise the same math
    op by op (three_interpolate -> cat -> SharedMLP -> max_pool2d)."""
    import os
    B, C, _ = known_feats.shape
    K = idx.shape[1] // nsample
    need_grad = torch.is_grad_enabled() and (known_feats.requires_grad or weight.requires_grad or rel_xyz.requires_grad or
                                             any(p.requires_grad for p in shared_mlp.parameters()))
    layers = None
    if not need_grad and os.environ.get("B200_SA_FUSED", "1") != "0" and hasattr(shared_mlp, "fold_affine"):
        layers = shared_mlp.fold_affine()
    if layers is not None and C % 4 == 0 and 128 % nsample == 0 and nsample >= 8:
        try:
            plan = shared_mlp.b200_plan(layers, C, True) if hasattr(shared_mlp, "b200_plan") else None
            return _ext.interp_mlp_forward(known_feats.contiguous(), idx.contiguous(), weight.contiguous(),
                                           rel_xyz.contiguous(), nsample, layers, plan=plan)
        except RuntimeError as e:  # widths outside the tensor-core kernel's range -> generic path
            if "not supported" not in str(e):
                raise
    interp = three_interpolate(known_feats, idx, weight)                                  # (B, C, K*nsample)
    x = torch.cat([rel_xyz.transpose(1, 2).contiguous().view(B, 3, K, nsample), interp.view(B, C, K, nsample)], 1)
    x = shared_mlp(x)
    return torch.nn.functional.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1)
