"""Set-abstraction / feature-propagation modules with the constructor and forward signatures of the reference's
pointnet2/pointnet2_modules.py (:31-503) -- PointnetSAModuleVotes, PointnetSAModuleMSGVotes, PointnetSAModule(MSG),
PointnetFPModule, PointnetLFPModuleMSG -- so models/backbone_module.py:35-72, proposal_module.py:72-79 and
grid_conv_module.py run unchanged, and reference checkpoints load (same child-module names).

Execution differs:  when a grouper+SharedMLP+max-pool stage is in the shape the fused sm_100a kernel supports
(eval-mode BatchNorm, max pooling, no gradient needed, nsample <= 128, <= 4 layers), the whole stage is ONE call into
b200pn2_sa_forward and grouped tensors never reach HBM.  Otherwise (training, avg/rbf pooling, sample_uniformly,
GroupAll) the stage runs op by op on the same kernels through the autograd Functions of pointnet2_utils.
Set B200_SA_FUSED=0 to force the unfused path.
"""
import os
import sys
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F

_HERE = os.path.dirname(os.path.abspath(__file__))
if _HERE not in sys.path:
    sys.path.append(_HERE)
if os.path.dirname(_HERE) not in sys.path:
    sys.path.append(os.path.dirname(_HERE))

import pointnet2.pointnet2_utils as pointnet2_utils  # noqa: E402
import pointnet2.pytorch_utils as pt_utils  # noqa: E402
import pointnet2._ext as _ext  # noqa: E402


def _fused_enabled():
    return os.environ.get("B200_SA_FUSED", "1") != "0"


def _sample_centres(xyz, npoint, inds=None):
    """FPS (unless indices are given) + gather of the centre coordinates -> (new_xyz (B,npoint,3), inds)."""
    if inds is None:
        inds = pointnet2_utils.furthest_point_sample(xyz, npoint)
    flipped = xyz.transpose(1, 2).contiguous()
    new_xyz = pointnet2_utils.gather_operation(flipped, inds).transpose(1, 2).contiguous()
    return new_xyz, inds


def _can_fuse(grouper, mlp, xyz, features, pooling="max"):
    if not _fused_enabled() or pooling != "max" or not isinstance(grouper, pointnet2_utils.QueryAndGroup):
        return None
    if grouper.sample_uniformly or grouper.nsample > 128 or not xyz.is_cuda:
        return None
    if torch.is_grad_enabled() and (xyz.requires_grad or (features is not None and features.requires_grad) or
                                    any(p.requires_grad for p in mlp.parameters())):
        return None
    layers = mlp.fold_affine() if hasattr(mlp, "fold_affine") else None
    if not layers or len(layers) > 4:
        return None
    return layers


def _fused_stage(grouper, layers, xyz, new_xyz, features, features_pm=None, want_pm=False, mlp=None):
    C = features.size(1) if features is not None else (features_pm.size(2) if features_pm is not None else 0)
    plan = mlp.b200_plan(layers, C, grouper.use_xyz) if (mlp is not None and hasattr(mlp, "b200_plan")) else None
    out, out_pm, _ = _ext.sa_forward(xyz, features, new_xyz, grouper.radius, grouper.nsample, layers,
                                     use_xyz=grouper.use_xyz, normalize_xyz=grouper.normalize_xyz,
                                     features_pm=features_pm, want_pm=want_pm, plan=plan)
    return out, out_pm


class _FusedTrainSA(torch.autograd.Function):
    """group -> [conv1x1 -> BatchNorm2d(batch statistics) -> ReLU] x L -> max over nsample, forward and backward on the fused
    training kernels (csrc/sa_train.cu).  Differentiable w.r.t. features, conv weights, BN gamma / beta;
    coordinates are constants.  Running statistics are updated in place by the forward (torch's momentum rule)."""

    @staticmethod
    def forward(ctx, features, xyz, new_xyz, idx, cfg, *params):
        radius, nsample, use_xyz, normalize, eps, momentum, running = cfg
        L = len(params) // 3
        layers = [(params[3 * i].detach(), params[3 * i + 1].detach(), params[3 * i + 2].detach(), running[i][0], running[i][1])
                  for i in range(L)]
        fpm = _ext.transpose_cn(features.detach().contiguous()) if features is not None else None
        out, saved = _ext.sa_train_forward(xyz, fpm, new_xyz, idx, radius, nsample, layers, eps, momentum, use_xyz=use_xyz,
                                           normalize_xyz=normalize)
        ctx.save_for_backward(saved, idx, *[p.detach() for p in params])
        ctx.meta = (xyz.size(0), xyz.size(1), new_xyz.size(1), features.size(1) if features is not None else 0, nsample, use_xyz,
                    features is not None and features.requires_grad, [tuple(p.shape) for p in params])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        saved, idx = ctx.saved_tensors[0], ctx.saved_tensors[1]
        params = ctx.saved_tensors[2:]
        B, N, M, C, nsample, use_xyz, want_feat, shapes = ctx.meta
        L = len(params) // 3
        layers = [(params[3 * i], params[3 * i + 1], params[3 * i + 2], None, None) for i in range(L)]
        gin, gw, gg, gb = _ext.sa_train_backward(grad_out.contiguous(), saved, idx, nsample, layers, B, N, M, C, use_xyz=use_xyz,
                                                 want_input_grad=want_feat)
        grads = []
        for i in range(L):
            grads += [gw[i].view(shapes[3 * i]), gg[i], gb[i]]
        return (gin if want_feat else None, None, None, None, None, *grads)


def _train_fusable(grouper, mlp, xyz, features, pooling="max"):
    """[(conv, bn)] when the stage can run on the fused training kernels: every block conv1x1(bias=False) -> BatchNorm2d in
    training mode -> ReLU, max pooling, constant coordinates, one (eps, momentum) for the stack."""
    if os.environ.get("B200_SA_TRAIN_FUSED", "1") == "0" or pooling != "max" or not xyz.is_cuda or xyz.requires_grad:
        return None
    if not isinstance(grouper, pointnet2_utils.QueryAndGroup) or grouper.sample_uniformly:
        return None
    pairs = []
    for block in mlp.children():
        conv = bn = act = None
        for key, mod in block.named_children():
            if key.endswith("conv"):
                conv = mod
            elif key.endswith("bn"):
                bn = next(iter(mod.children()))
            elif key.endswith("activation"):
                act = mod
        if conv is None or bn is None or not isinstance(act, nn.ReLU) or conv.bias is not None or not bn.training:
            return None
        if tuple(conv.kernel_size) != (1, 1) or not bn.affine or not bn.track_running_stats or bn.momentum is None:
            return None
        pairs.append((conv, bn))
    if not pairs or len({(bn.eps, bn.momentum) for _, bn in pairs}) != 1:
        return None
    C = features.size(1) if features is not None else 0
    probe = [(c.weight, b.weight, b.bias, None, None) for c, b in pairs]
    if not _ext.sa_train_supported(C, grouper.use_xyz, probe):
        return None
    return pairs


def _fused_train_stage(grouper, pairs, xyz, new_xyz, features):
    idx = pointnet2_utils.ball_query(grouper.radius, grouper.nsample, xyz, new_xyz)
    bn0 = pairs[0][1]
    cfg = (grouper.radius, grouper.nsample, grouper.use_xyz, grouper.normalize_xyz, bn0.eps, bn0.momentum,
           [(bn.running_mean, bn.running_var) for _, bn in pairs])
    params = []
    for conv, bn in pairs:
        params += [conv.weight, bn.weight, bn.bias]
    out = _FusedTrainSA.apply(features, xyz, new_xyz, idx, cfg, *params)
    with torch.no_grad():
        for _, bn in pairs:
            bn.num_batches_tracked += 1
    return out


def _pool(new_features, pooling, grouped_xyz=None, sigma=None, nsample=None):
    if pooling == "max":
        out = F.max_pool2d(new_features, kernel_size=[1, new_features.size(3)])
    elif pooling == "avg":
        out = F.avg_pool2d(new_features, kernel_size=[1, new_features.size(3)])
    elif pooling == "rbf":
        # radial-basis weighting of the neighbours, normalised by nsample (reference :267-271)
        rbf = torch.exp(-1 * grouped_xyz.pow(2).sum(1, keepdim=False) / (sigma ** 2) / 2)
        out = torch.sum(new_features * rbf.unsqueeze(1), -1, keepdim=True) / float(nsample)
    else:
        raise ValueError("unknown pooling %r" % (pooling,))
    return out.squeeze(-1)


class _PointnetSAModuleBase(nn.Module):
    """FPS -> per-scale (group -> SharedMLP -> max) -> concat (reference :31-80)."""

    def __init__(self):
        super().__init__()
        self.npoint = None
        self.groupers = None
        self.mlps = None

    def _scales(self, xyz, new_xyz, features):
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            layers = _can_fuse(grouper, mlp, xyz, features) if new_xyz is not None else None
            if layers is not None:
                outs.append(_fused_stage(grouper, layers, xyz, new_xyz, features, mlp=mlp)[0])
            else:
                outs.append(_pool(mlp(grouper(xyz, new_xyz, features)), "max"))
        return torch.cat(outs, dim=1)

    def forward(self, xyz, features=None):
        new_xyz = _sample_centres(xyz, self.npoint)[0] if self.npoint is not None else None
        return new_xyz, self._scales(xyz, new_xyz, features)


def _build_scales(module, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly):
    assert len(radii) == len(nsamples) == len(mlps)
    module.groupers = nn.ModuleList()
    module.mlps = nn.ModuleList()
    for radius, nsample, spec in zip(radii, nsamples, mlps):
        module.groupers.append(
            pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, sample_uniformly=sample_uniformly)
            if npoint is not None else pointnet2_utils.GroupAll(use_xyz))
        if use_xyz:
            spec[0] += 3  # in place, like the reference (:125-126)
        module.mlps.append(pt_utils.SharedMLP(spec, bn=bn))


class PointnetSAModuleMSG(_PointnetSAModuleBase):
    """Multi-scale grouping SA layer (reference :83-129)."""

    def __init__(self, *, npoint: int, radii: List[float], nsamples: List[int], mlps: List[List[int]],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        self.npoint = npoint
        _build_scales(self, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)


class PointnetSAModule(PointnetSAModuleMSG):
    """Single-scale SA layer (reference :132-166)."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True):
        super().__init__(mlps=[mlp], npoint=npoint, radii=[radius], nsamples=[nsample], bn=bn, use_xyz=use_xyz)


class PointnetSAModuleVotes(nn.Module):
    """SA layer that also returns the sampled indices (reference :169-277).

    forward(xyz (B,N,3), features (B,C,N), inds=None) -> (new_xyz (B,npoint,3), new_features (B,C_out,npoint),
    inds (B,npoint) int32 [, unique_cnt])."""

    def __init__(self, *, mlp: List[int], npoint: int = None, radius: float = None, nsample: int = None,
                 bn: bool = True, use_xyz: bool = True, pooling: str = 'max', sigma: float = None,
                 normalize_xyz: bool = False, sample_uniformly: bool = False, ret_unique_cnt: bool = False):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.pooling = pooling
        self.use_xyz = use_xyz
        self.sigma = sigma if sigma is not None else (self.radius / 2 if self.radius is not None else None)
        self.normalize_xyz = normalize_xyz
        self.ret_unique_cnt = ret_unique_cnt
        if npoint is not None:
            self.grouper = pointnet2_utils.QueryAndGroup(radius, nsample, use_xyz=use_xyz, ret_grouped_xyz=True,
                                                         normalize_xyz=normalize_xyz,
                                                         sample_uniformly=sample_uniformly,
                                                         ret_unique_cnt=ret_unique_cnt)
        else:
            self.grouper = pointnet2_utils.GroupAll(use_xyz, ret_grouped_xyz=True)
        mlp_spec = mlp
        if use_xyz and len(mlp_spec) > 0:
            mlp_spec[0] += 3
        self.mlp_module = pt_utils.SharedMLP(mlp_spec, bn=bn)

    def forward(self, xyz, features=None, inds=None):
        if inds is not None:
            assert inds.shape[1] == self.npoint
        if self.npoint is not None:
            new_xyz, inds = _sample_centres(xyz, self.npoint, inds)
        else:
            new_xyz = None
            if inds is None:  # the reference samples even when npoint is None (:239-240); keep `inds` defined
                inds = torch.zeros((xyz.size(0), 0), dtype=torch.int32, device=xyz.device)

        layers = None
        if new_xyz is not None and not self.ret_unique_cnt:
            layers = _can_fuse(self.grouper, self.mlp_module, xyz, features, self.pooling)
        if layers is not None:
            new_features, _ = _fused_stage(self.grouper, layers, xyz, new_xyz, features, mlp=self.mlp_module)
            return new_xyz, new_features, inds
        if new_xyz is not None and not self.ret_unique_cnt:
            pairs = _train_fusable(self.grouper, self.mlp_module, xyz, features, self.pooling)
            if pairs is not None:
                return new_xyz, _fused_train_stage(self.grouper, pairs, xyz, new_xyz, features), inds

        grouped = self.grouper(xyz, new_xyz, features)
        if self.ret_unique_cnt:
            grouped_features, grouped_xyz, unique_cnt = grouped
        else:
            grouped_features, grouped_xyz = grouped
        new_features = _pool(self.mlp_module(grouped_features), self.pooling, grouped_xyz, self.sigma, self.nsample)
        if self.ret_unique_cnt:
            return new_xyz, new_features, inds, unique_cnt
        return new_xyz, new_features, inds


class PointnetSAModuleMSGVotes(nn.Module):
    """Multi-scale SA layer that also returns the sampled indices (reference :280-359)."""

    def __init__(self, *, mlps: List[List[int]], npoint: int, radii: List[float], nsamples: List[int],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        self.npoint = npoint
        _build_scales(self, npoint, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)

    def forward(self, xyz, features=None, inds=None):
        if self.npoint is not None:
            new_xyz, inds = _sample_centres(xyz, self.npoint, inds)
        else:
            new_xyz = None
        return new_xyz, _PointnetSAModuleBase._scales(self, xyz, new_xyz, features), inds


class PointnetFPModule(nn.Module):
    """Feature propagation: three_nn -> inverse-distance weights -> three_interpolate -> concat skip -> SharedMLP
    (reference :362-422)."""

    def __init__(self, *, mlp: List[int], bn: bool = True):
        super().__init__()
        self.mlp = pt_utils.SharedMLP(mlp, bn=bn)

    def forward(self, unknown, known, unknow_feats, known_feats):
        if known is not None:
            dist, idx = pointnet2_utils.three_nn(unknown, known)
            dist_recip = 1.0 / (dist + 1e-8)
            weight = dist_recip / torch.sum(dist_recip, dim=2, keepdim=True)
            fused = self._fused(unknown, unknow_feats, known_feats, idx, weight)
            if fused is not None:
                return fused
            interpolated = pointnet2_utils.three_interpolate(known_feats, idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)

    def _fused(self, unknown, unknow_feats, known_feats, idx, weight):
        """Eval mode, no gradient: blend -> concat -> SharedMLP without the interpolated / concatenated tensors ever
        reaching HBM -- the producers of the tensor-core kernel build each row [sum_t w_t * known[idx_t] | skip] on the
        fly (b200pn2_fp_rows_forward); layers wider than 128 chain through the row-MLP entry."""
        if not _fused_enabled() or not known_feats.is_cuda or known_feats.dtype != torch.float32:
            return None
        if torch.is_grad_enabled() and (known_feats.requires_grad or (unknow_feats is not None and unknow_feats.requires_grad)
                                        or unknown.requires_grad or any(p.requires_grad for p in self.mlp.parameters())):
            return None
        layers = self.mlp.fold_affine() if hasattr(self.mlp, "fold_affine") else None
        C2 = known_feats.size(1)
        C1 = unknow_feats.size(1) if unknow_feats is not None else 0
        if (not layers or C2 % 4 != 0 or C1 % 4 != 0 or layers[0][0].size(1) != C1 + C2 or
                any(w.size(0) > 256 for w, _, _ in layers)):
            return None
        groups = _ext.split_row_groups(layers)
        known_pm = _ext.transpose_cn(known_feats.contiguous())
        skip_pm = _ext.transpose_cn(unknow_feats.contiguous()) if unknow_feats is not None else None
        out_cm = rows = None
        chan = C1 + C2
        for gi, grp in enumerate(groups):
            last = gi == len(groups) - 1
            plan = self.mlp.b200_plan(grp, chan, False, row_output=True, plain_rows=gi > 0)
            if gi == 0:
                out_cm, rows = _ext.fp_rows_forward(known_pm, skip_pm, idx.contiguous(), weight.contiguous(), grp,
                                                    relu_last=True, want_cm=last, want_pm=not last, plan=plan)
            else:
                out_cm, rows = _ext.row_mlp_forward(rows, grp, relu_last=True, want_cm=last, want_pm=not last, plan=plan)
            chan = grp[-1][0].size(0)
        return out_cm


class PointnetLFPModuleMSG(nn.Module):
    """Learnable feature propagation (reference :425-503): group features1 around xyz2, SharedMLP, max, concat
    features2, post-MLP."""

    def __init__(self, *, mlps: List[List[int]], radii: List[float], nsamples: List[int], post_mlp: List[int],
                 bn: bool = True, use_xyz: bool = True, sample_uniformly: bool = False):
        super().__init__()
        self.post_mlp = pt_utils.SharedMLP(post_mlp, bn=bn)
        _build_scales(self, 1, radii, nsamples, mlps, bn, use_xyz, sample_uniformly)

    def forward(self, xyz2, xyz1, features2, features1):
        outs = []
        for grouper, mlp in zip(self.groupers, self.mlps):
            layers = _can_fuse(grouper, mlp, xyz1, features1)
            if layers is not None:
                pooled = _fused_stage(grouper, layers, xyz1, xyz2, features1, mlp=mlp)[0]
            else:
                pooled = _pool(mlp(grouper(xyz1, xyz2, features1)), "max")
            if features2 is not None:
                pooled = torch.cat([pooled, features2], dim=1)
            outs.append(self.post_mlp(pooled.unsqueeze(-1)))
        return torch.cat(outs, dim=1).squeeze(-1)


if __name__ == "__main__":
    # the reference's smoke demo (:506-525), on the fused path
    torch.manual_seed(1)
    xyz = torch.randn(2, 9, 3).cuda()
    feats = torch.randn(2, 9, 6).cuda().transpose(1, 2).contiguous()
    net = PointnetSAModuleMSG(npoint=2, radii=[5.0, 10.0], nsamples=[6, 3], mlps=[[6, 3], [6, 6]]).cuda().eval()
    with torch.no_grad():
        print(net(xyz, feats))
