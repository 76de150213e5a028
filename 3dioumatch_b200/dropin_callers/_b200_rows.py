"""1x1-conv heads (torch.nn.Conv1d [+ BatchNorm1d] [+ ReLU] chains) as fused row MLPs on the tensor-core kernel.

The reference's voting / proposal / IoU heads are plain torch modules (voting_module.py:27-31, proposal_module.py:84-88,
grid_conv_module.py:42-46); the drop-in caller mirrors keep those modules -- same names, same state-dict keys -- and in
eval mode without gradients run them through b200pn2_row_mlp_forward: one launch per run of layers, hidden activations
never leave the SM.  Anything else (training, gradients) takes the torch path of the reference."""
import os

import torch
import torch.nn.functional as F

import pointnet2._ext as _ext


def fold(conv, bn=None):
    """(weight (cout,cin), scale, shift) with y = scale * (W x) + shift  ==  bn(conv(x)) in eval mode."""
    w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).contiguous().float()
    bias = conv.bias.detach().float() if conv.bias is not None else None
    if bn is not None:
        inv = torch.rsqrt(bn.running_var.detach() + bn.eps)
        scale = (bn.weight.detach() * inv) if bn.affine else inv
        shift = (bn.bias.detach() if bn.affine else torch.zeros_like(inv)) - bn.running_mean.detach() * scale
        if bias is not None:
            shift = shift + bias * scale
    else:
        scale = torch.ones(conv.out_channels, device=w.device)
        shift = bias if bias is not None else torch.zeros_like(scale)
    return w, scale.float().contiguous(), shift.float().contiguous()


def fusable(x, *modules):
    if os.environ.get("B200_SA_FUSED", "1") == "0" or not x.is_cuda or x.dtype != torch.float32:
        return False
    if any(m.training for m in modules):
        return False
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for m in modules for p in m.parameters())):
        return False
    return True


class Head:
    """Cache of folded layers + packed plans for one chain of (conv, bn) pairs; frozen with the model's SharedMLPs
    (pointnet2.pytorch_utils.freeze_inference sets `_b200_frozen_heads` through freeze())."""

    def __init__(self, pairs):
        self.pairs = pairs      # [(conv, bn or None)], ReLU after every layer but the last
        self.frozen = None

    def layers(self):
        if self.frozen is not None:
            return self.frozen["layers"]
        return [fold(c, b) for c, b in self.pairs]

    def freeze(self):
        self.frozen = {"layers": [fold(c, b) for c, b in self.pairs], "plans": {}}

    def unfreeze(self):
        self.frozen = None

    def plan(self, key, grp, chan):
        if self.frozen is None:
            return None
        if key not in self.frozen["plans"]:
            self.frozen["plans"][key] = _ext.mlp_plan(grp, chan, False, row_output=True, plain_rows=True)
        return self.frozen["plans"][key]

    def __call__(self, rows, want_cm=True, want_pm=False, split_last=None):
        """rows (B, n, C) point-major -> the chain's output, channel-major (B, cout, n) and/or point-major.
        split_last = [c0, c1, ...]: the last layer's output rows are produced as separate row ranges (e.g. the voting
        module's 3 offset channels + 256 residual channels, together wider than one launch takes)."""
        layers = self.layers()
        body, last = layers[:-1], layers[-1]
        chan = rows.size(2)
        groups = _ext.split_row_groups(body) if body else []
        for gi, grp in enumerate(groups):
            _, rows = _ext.row_mlp_forward(rows, grp, relu_last=True, want_cm=False, want_pm=True,
                                           plan=self.plan(("body", gi), grp, chan))
            chan = grp[-1][0].size(0)
        w, sc, sh = last
        bounds = [0] + list(split_last or []) + [w.size(0)]
        outs = []
        for pi in range(len(bounds) - 1):
            a, b = bounds[pi], bounds[pi + 1]
            grp = [(w[a:b].contiguous(), sc[a:b].contiguous(), sh[a:b].contiguous())] if split_last else [last]
            if self.frozen is not None and split_last:
                cache = self.frozen.setdefault("split", {})
                grp = cache.setdefault(pi, grp)
            outs.append(_ext.row_mlp_forward(rows, grp, relu_last=False, want_cm=want_cm, want_pm=want_pm,
                                             plan=self.plan(("last", pi), grp, chan)))
        return outs if split_last else outs[0]


def torch_head(x, pairs):
    """The reference's own formulation: relu(bn(conv(x))) ... conv_last(x), channel-major (B,C,n)."""
    for conv, bn in pairs[:-1]:
        x = F.relu(bn(conv(x)) if bn is not None else conv(x))
    conv, bn = pairs[-1]
    return bn(conv(x)) if bn is not None else conv(x)
