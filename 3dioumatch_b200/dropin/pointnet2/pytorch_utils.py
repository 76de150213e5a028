"""Shared-MLP building blocks with the module/parameter names of the reference's pointnet2/pytorch_utils.py
(so its checkpoints load unchanged: `layer{i}.conv.weight`, `layer{i}.bn.bn.{weight,bias,running_mean,...}`;
reference pytorch_utils.py:14-39,42-67,70-123,226-263,265-299).

Written for the fused sm_100a path: every block can export itself as an (weight, scale, shift) affine triple
(`fold_affine`), which is what b200pn2_sa_forward consumes in eval mode; `freeze_inference(model)` additionally keeps
the packed tensor-core weight images per module, and SharedMLP.forward itself runs as fused row MLPs in eval mode.
"""
import os

import torch
import torch.nn as nn

try:  # bound at import time: callers may load several operator stacks into one process (3dioumatch_b200/refapp.py)
    from . import _ext
except ImportError:  # imported flat (`import pytorch_utils`), as the reference's pointnet2_modules.py does
    import _ext

_DEFAULT_ACT = nn.ReLU(inplace=True)


class _NormWrap(nn.Sequential):
    """`<name>bn` child holding the torch BatchNorm; gamma=1, beta=0 (reference _BNBase, :42-50)."""

    def __init__(self, width, norm_cls, name=""):
        super().__init__()
        norm = norm_cls(width)
        nn.init.constant_(norm.weight, 1.0)
        nn.init.constant_(norm.bias, 0)
        self.add_module(name + "bn", norm)


class BatchNorm1d(_NormWrap):
    def __init__(self, in_size, *, name=""):
        super().__init__(in_size, nn.BatchNorm1d, name)


class BatchNorm2d(_NormWrap):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, nn.BatchNorm2d, name)


class BatchNorm3d(_NormWrap):
    def __init__(self, in_size, name=""):
        super().__init__(in_size, nn.BatchNorm3d, name)


def _assemble(seq, name, core_key, core, norm, activation, preact):
    """Child order of the reference blocks: [bn, act,] core [, bn, act] depending on `preact` (:107-123)."""
    tail = []
    if norm is not None:
        tail.append((name + "bn", norm))
    if activation is not None:
        tail.append((name + "activation", activation))
    order = tail + [(name + core_key, core)] if preact else [(name + core_key, core)] + tail
    for key, mod in order:
        seq.add_module(key, mod)


class _ConvBase(nn.Sequential):
    """conv (bias only without BN, :90) -> BN -> activation."""

    def __init__(self, in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=None,
                 batch_norm=None, bias=True, preact=False, name=""):
        super().__init__()
        use_bias = bias and not bn
        unit = conv(in_size, out_size, kernel_size=kernel_size, stride=stride, padding=padding, bias=use_bias)
        init(unit.weight)
        if use_bias:
            nn.init.constant_(unit.bias, 0)
        norm = batch_norm(in_size if preact else out_size) if bn else None
        _assemble(self, name, "conv", unit, norm, activation, preact)


class Conv1d(_ConvBase):
    def __init__(self, in_size, out_size, *, kernel_size=1, stride=1, padding=0, activation=_DEFAULT_ACT, bn=False,
                 init=nn.init.kaiming_normal_, bias=True, preact=False, name=""):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv1d,
                         batch_norm=BatchNorm1d, bias=bias, preact=preact, name=name)


class Conv2d(_ConvBase):
    def __init__(self, in_size, out_size, *, kernel_size=(1, 1), stride=(1, 1), padding=(0, 0),
                 activation=_DEFAULT_ACT, bn=False, init=nn.init.kaiming_normal_, bias=True, preact=False, name=""):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv2d,
                         batch_norm=BatchNorm2d, bias=bias, preact=preact, name=name)


class Conv3d(_ConvBase):
    def __init__(self, in_size, out_size, *, kernel_size=(1, 1, 1), stride=(1, 1, 1), padding=(0, 0, 0),
                 activation=_DEFAULT_ACT, bn=False, init=nn.init.kaiming_normal_, bias=True, preact=False, name=""):
        super().__init__(in_size, out_size, kernel_size, stride, padding, activation, bn, init, conv=nn.Conv3d,
                         batch_norm=BatchNorm3d, bias=bias, preact=preact, name=name)


class FC(nn.Sequential):
    def __init__(self, in_size, out_size, *, activation=_DEFAULT_ACT, bn=False, init=None, preact=False, name=""):
        super().__init__()
        unit = nn.Linear(in_size, out_size, bias=not bn)
        if init is not None:
            init(unit.weight)
        if not bn:
            nn.init.constant_(unit.bias, 0)
        norm = BatchNorm1d(in_size if preact else out_size) if bn else None
        _assemble(self, name, "fc", unit, norm, activation, preact)


class SharedMLP(nn.Sequential):
    """Stack of 1x1 Conv2d blocks named `layer{i}` (reference :14-39)."""

    def __init__(self, args, *, bn=False, activation=_DEFAULT_ACT, preact=False, first=False, name=""):
        super().__init__()
        for i in range(len(args) - 1):
            plain_first = first and preact and i == 0  # the very first pre-activation block has no bn/act
            self.add_module(name + "layer{}".format(i),
                            Conv2d(args[i], args[i + 1], bn=bn and not plain_first,
                                   activation=None if plain_first else activation, preact=preact))

    # ---- export for the fused kernels --------------------------------------------------------------------
    def fold_affine(self):
        """[(weight (cout,cin), scale (cout,), shift (cout,))] such that each block is relu(scale*(W x)+shift),
        or None when a block is not conv(1x1) -> [eval BN] -> ReLU.

        Recomputed from the live parameters on EVERY call (a handful of tiny elementwise kernels): weights may be
        rewritten behind autograd's back -- the reference's EMA teacher does `ema_param.data.mul_(alpha).add_(...)`
        (train.py:285-289), which bumps no version counter -- and a captured CUDA graph must read them where they
        live.  `b200_freeze()` trades that for zero per-call work once the weights are final."""
        frozen = getattr(self, "_b200_frozen", None)
        if frozen is not None:
            return frozen["layers"]
        return self._fold_affine_uncached()

    def b200_freeze(self):
        """Inference with final weights: fold the BN affine once and keep the packed tensor-core weight images
        ("plans", include/b200_pointnet2.h) per call shape.  The caller promises not to modify parameters or BN
        statistics until b200_unfreeze() / train()."""
        layers = self._fold_affine_uncached() if not any(m.training for m in self.modules()) else None
        object.__setattr__(self, "_b200_frozen", {"layers": layers, "plans": {}} if layers is not None else None)
        return layers is not None

    def b200_unfreeze(self):
        object.__setattr__(self, "_b200_frozen", None)

    def train(self, mode=True):
        if mode:
            self.b200_unfreeze()
        return super().train(mode)

    def b200_plan(self, layers, C_feat, use_xyz, row_output=False, plain_rows=False):
        """Cached packed weights for `layers` (a slice of the frozen stack) or None when not frozen."""
        frozen = getattr(self, "_b200_frozen", None)
        if frozen is None or frozen["layers"] is None:
            return None
        key = (layers[0][0].data_ptr(), len(layers), int(C_feat), bool(use_xyz), bool(row_output), bool(plain_rows),
               os.environ.get("B200_SA_TC_FACTOR", ""), str(layers[0][0].device))
        if key not in frozen["plans"]:
            frozen["plans"][key] = _ext.mlp_plan(layers, C_feat, use_xyz, row_output=row_output, plain_rows=plain_rows)
        return frozen["plans"][key]

    def forward(self, x):
        """Eval mode, no gradient, CUDA: the whole stack as row MLPs on the tensor-core kernel (one launch per run of
        layers, activations between the layers of a run never leave the SM); otherwise nn.Sequential."""
        if (x.is_cuda and x.dim() == 4 and x.dtype == torch.float32 and os.environ.get("B200_SA_FUSED", "1") != "0" and
                not (torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())))):
            layers = self.fold_affine()
            if layers:
                out = self._rows_forward(x, layers)
                if out is not None:
                    return out
        return super().forward(x)

    def _rows_forward(self, x, layers):
        B, C, H, W = x.shape
        if B * H * W == 0 or any(w.size(0) > 256 for w, _, _ in layers) or layers[0][0].size(1) != C:
            return None
        groups = _ext.split_row_groups(layers)
        x_cm = x.contiguous().view(B, C, H * W)
        rows, chan = None, C
        for gi, grp in enumerate(groups):
            last = gi == len(groups) - 1
            plan = self.b200_plan(grp, chan, False, row_output=True, plain_rows=True)
            if gi == 0 and os.environ.get("B200_ROWS_CM", "1") != "0":
                # the first run reads the channel-major conv input in place (no transpose pass)
                out_cm, out_pm = _ext.row_mlp_forward_cm(x_cm, grp, relu_last=True, want_cm=last, want_pm=not last, plan=plan)
            else:
                if rows is None:
                    rows = _ext.transpose_cn(x_cm, ld=(C + 3) // 4 * 4)          # (B, H*W, ld)
                out_cm, out_pm = _ext.row_mlp_forward(rows, grp, relu_last=True, want_cm=last, want_pm=not last, plan=plan,
                                                      channels=chan)
            rows, chan = out_pm, grp[-1][0].size(0)
        return out_cm.view(B, chan, H, W)

    def _fold_affine_uncached(self):
        triples = []
        for block in self.children():
            conv = bnorm = act = None
            seen = []
            for key, mod in block.named_children():
                seen.append(key)
                if key.endswith("conv"):
                    conv = mod
                elif key.endswith("bn"):
                    bnorm = next(iter(mod.children()))
                elif key.endswith("activation"):
                    act = mod
            if conv is None or not isinstance(act, nn.ReLU) or not seen or not seen[0].endswith("conv"):
                return None
            if tuple(conv.kernel_size) != (1, 1) or tuple(conv.stride) != (1, 1) or tuple(conv.padding) != (0, 0):
                return None
            w = conv.weight.detach().reshape(conv.out_channels, conv.in_channels).contiguous()
            cb = conv.bias.detach() if conv.bias is not None else None
            if bnorm is not None:
                if bnorm.training or not bnorm.track_running_stats or bnorm.running_var is None:
                    return None
                inv = torch.rsqrt(bnorm.running_var.detach() + bnorm.eps)
                gamma = bnorm.weight.detach() if bnorm.affine else torch.ones_like(inv)
                beta = bnorm.bias.detach() if bnorm.affine else torch.zeros_like(inv)
                scale = gamma * inv
                shift = beta - bnorm.running_mean.detach() * scale
                if cb is not None:
                    shift = shift + cb * scale
            else:
                scale = torch.ones(conv.out_channels, device=w.device, dtype=w.dtype)
                shift = cb.clone() if cb is not None else torch.zeros_like(scale)
            triples.append((w.float(), scale.float().contiguous(), shift.float().contiguous()))
        return triples


def freeze_inference(model):
    """Freeze every SharedMLP of `model` for inference with final weights (see SharedMLP.b200_freeze): call after
    model.eval() and after the checkpoint is loaded.  Returns the number of stacks frozen."""
    n = 0
    for m in model.modules():
        if isinstance(m, SharedMLP) and m.b200_freeze():
            n += 1
        head = getattr(m, "_b200_head", None)   # 1x1-conv heads of the drop-in caller mirrors (dropin_callers/)
        if head is not None and not m.training:
            head.freeze()
    return n


def unfreeze(model):
    for m in model.modules():
        if isinstance(m, SharedMLP):
            m.b200_unfreeze()
        head = getattr(m, "_b200_head", None)
        if head is not None:
            head.unfreeze()


def set_bn_momentum_default(bn_momentum):
    def fn(m):
        if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            m.momentum = bn_momentum
    return fn


class BNMomentumScheduler(object):
    """Sets BatchNorm momentum from `bn_lambda(epoch)` (reference :274-299)."""

    def __init__(self, model, bn_lambda, last_epoch=-1, setter=set_bn_momentum_default):
        if not isinstance(model, nn.Module):
            raise RuntimeError("Class '{}' is not a PyTorch nn Module".format(type(model).__name__))
        self.model, self.setter, self.lmbd = model, setter, bn_lambda
        self.step(last_epoch + 1)
        self.last_epoch = last_epoch

    def step(self, epoch=None):
        if epoch is None:
            epoch = self.last_epoch + 1
        self.last_epoch = epoch
        self.model.apply(self.setter(self.lmbd(epoch)))
