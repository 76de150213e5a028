"""fp32 PyTorch restatement of the set-abstraction / feature-propagation MODULE math -- TEST INFRASTRUCTURE ONLY.

The native index ops come from the C oracle (oracle/oracle.py); the dense half follows the reference's torch code:
  QueryAndGroup.forward      pointnet2/pointnet2_utils.py:318-377
  SharedMLP (conv1x1+BN+ReLU) pointnet2/pytorch_utils.py:14-39,70-123   (eval-mode BN: running statistics)
  max_pool2d over nsample    pointnet2/pointnet2_modules.py:256-262
  PointnetFPModule.forward   pointnet2/pointnet2_modules.py:377-422
Runs on CPU in fp32 (no TF32 anywhere).
"""
import numpy as np
import torch
import torch.nn.functional as F

import oracle as orc


def shared_mlp(x, layers, eps=1e-5):
    """x (B,C,M,K); layers = list of dict(weight (cout,cin), gamma, beta, mean, var) -> relu(bn(conv(x)))..."""
    for ly in layers:
        w = torch.from_numpy(ly["weight"]).view(ly["weight"].shape[0], -1, 1, 1)
        x = F.conv2d(x, w)
        x = F.batch_norm(x, torch.from_numpy(ly["mean"]), torch.from_numpy(ly["var"]), torch.from_numpy(ly["gamma"]),
                         torch.from_numpy(ly["beta"]), training=False, eps=eps)
        x = F.relu(x)
    return x


def sa_forward(xyz, features, new_xyz, radius, nsample, layers, use_xyz=True, normalize_xyz=False, idx=None):
    """numpy in/out.  Returns (new_features (B,Cout,M), idx (B,M,nsample))."""
    if idx is None:
        idx = orc.ball_query(new_xyz, xyz, radius, nsample)
    grouped_xyz = torch.from_numpy(orc.group_points(np.ascontiguousarray(xyz.transpose(0, 2, 1)), idx))
    grouped_xyz = grouped_xyz - torch.from_numpy(new_xyz).transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        grouped_xyz = grouped_xyz * np.float32(1.0 / radius)  # CUDA torch: x * fp32(1/r) for `x /= python_float`
    if features is not None:
        grouped = torch.from_numpy(orc.group_points(features, idx))
        x = torch.cat([grouped_xyz, grouped], 1) if use_xyz else grouped
    else:
        x = grouped_xyz
    x = shared_mlp(x, layers)
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1).numpy(), idx


def fold(layers, eps=1e-5):
    """(weight, scale, shift) triples of the eval-mode affine, as SharedMLP.fold_affine produces."""
    out = []
    for ly in layers:
        scale = ly["gamma"] / np.sqrt(ly["var"] + np.float32(eps))
        out.append((ly["weight"], scale.astype(np.float32), (ly["beta"] - ly["mean"] * scale).astype(np.float32)))
    return out


def fp_forward(unknown, known, unknow_feats, known_feats, layers):
    """PointnetFPModule.forward in fp32 (numpy in/out)."""
    dist2, idx = orc.three_nn(unknown, known)
    dist = torch.sqrt(torch.from_numpy(dist2))
    recip = 1.0 / (dist + 1e-8)
    weight = (recip / recip.sum(2, keepdim=True)).numpy()
    interp = torch.from_numpy(orc.three_interpolate(known_feats, idx, weight))
    x = torch.cat([interp, torch.from_numpy(unknow_feats)], 1) if unknow_feats is not None else interp
    return shared_mlp(x.unsqueeze(-1), layers).squeeze(-1).numpy()
