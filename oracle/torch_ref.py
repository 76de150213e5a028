"""PyTorch / numpy restatement of the set-abstraction / feature-propagation MODULE math -- TEST INFRASTRUCTURE ONLY.

The native index ops come from the C oracle (oracle/oracle.py); the dense half follows the reference's torch code:
  QueryAndGroup.forward      pointnet2/pointnet2_utils.py:318-377
  SharedMLP (conv1x1+BN+ReLU) pointnet2/pytorch_utils.py:14-39,70-123   (eval-mode BN: running statistics)
  max_pool2d over nsample    pointnet2/pointnet2_modules.py:256-262
  PointnetFPModule.forward   pointnet2/pointnet2_modules.py:377-422
Everything up to the grouped / interpolated tensor is fp32 with the reference's rounding sequence.  The conv1x1 -> BN ->
ReLU stack is evaluated in float64 (numpy matmul) and rounded to fp32 once (`exact=True`, the default): the checker must
not depend on the host's fp32 convolution.  Measured on this pool's hosts (Xeon with AMX): the FIRST `F.conv2d` of a shape
in a process occasionally returned one thread's chunk of the output with ~6e-4 error while the recomputation, the fp64
result and both device paths agreed to 4e-6 -- a parity test against that would blame the kernel for the checker's
error.  `exact=False` keeps the literal fp32 torch sequence (conv2d -> batch_norm -> relu) for the CPU baseline timing.
"""
import numpy as np
import torch
import torch.nn.functional as F

import oracle as orc


def shared_mlp_fp32(x, layers, eps=1e-5):
    """The literal fp32 torch sequence: x (B,C,M,K); layers = list of dict(weight (cout,cin), gamma, beta, mean, var)."""
    for ly in layers:
        w = torch.from_numpy(ly["weight"]).view(ly["weight"].shape[0], -1, 1, 1)
        x = F.conv2d(x, w)
        x = F.batch_norm(x, torch.from_numpy(ly["mean"]), torch.from_numpy(ly["var"]), torch.from_numpy(ly["gamma"]),
                         torch.from_numpy(ly["beta"]), training=False, eps=eps)
        x = F.relu(x)
    return x


def shared_mlp(x, layers, eps=1e-5, exact=True):
    """relu(bn(conv1x1(x))) per layer on x (B,C,M,K) (torch tensor or array) -> torch fp32 tensor (B,Cout,M,K).
    exact: float64 arithmetic, rounded to fp32 at the end (see the module docstring); else the fp32 torch ops."""
    if not exact:
        return shared_mlp_fp32(x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(x)), layers, eps)
    a = (x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x)).astype(np.float64)
    B, C, M, K = a.shape
    a = np.ascontiguousarray(a.transpose(1, 0, 2, 3)).reshape(C, B * M * K)
    for ly in layers:
        w = ly["weight"].astype(np.float64).reshape(ly["weight"].shape[0], -1)
        scale = ly["gamma"].astype(np.float64) / np.sqrt(ly["var"].astype(np.float64) + eps)
        z = (w @ a - ly["mean"].astype(np.float64)[:, None]) * scale[:, None] + ly["beta"].astype(np.float64)[:, None]
        a = np.maximum(z, 0.0)
    out = a.reshape(-1, B, M, K).transpose(1, 0, 2, 3)
    return torch.from_numpy(np.ascontiguousarray(out).astype(np.float32))


def sa_forward(xyz, features, new_xyz, radius, nsample, layers, use_xyz=True, normalize_xyz=False, idx=None, exact=True):
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
    x = shared_mlp(x, layers, exact=exact)
    return F.max_pool2d(x, kernel_size=[1, x.size(3)]).squeeze(-1).numpy(), idx


def fold(layers, eps=1e-5):
    """(weight, scale, shift) triples of the eval-mode affine, as SharedMLP.fold_affine produces."""
    out = []
    for ly in layers:
        scale = ly["gamma"] / np.sqrt(ly["var"] + np.float32(eps))
        out.append((ly["weight"], scale.astype(np.float32), (ly["beta"] - ly["mean"] * scale).astype(np.float32)))
    return out


def fp_forward(unknown, known, unknow_feats, known_feats, layers, exact=True):
    """PointnetFPModule.forward (numpy in/out): fp32 three_nn / weights / blend, then the MLP stack (see shared_mlp)."""
    dist2, idx = orc.three_nn(unknown, known)
    dist = torch.sqrt(torch.from_numpy(dist2))
    recip = 1.0 / (dist + 1e-8)
    weight = (recip / recip.sum(2, keepdim=True)).numpy()
    interp = torch.from_numpy(orc.three_interpolate(known_feats, idx, weight))
    x = torch.cat([interp, torch.from_numpy(unknow_feats)], 1) if unknow_feats is not None else interp
    return shared_mlp(x.unsqueeze(-1), layers, exact=exact).squeeze(-1).numpy()
