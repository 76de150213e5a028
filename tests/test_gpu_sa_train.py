"""GPU: the fused TRAINING-mode set-abstraction MLP (csrc/sa_train.cu) -- conv1x1 -> BatchNorm2d on batch
statistics -> ReLU per layer, max over nsample, forward AND backward -- against
  * oracle/sa_train_ref.reference(): the reference semantics run literally through torch on the CPU (conv2d -> batch_norm(
    training=True) -> relu -> max_pool2d, autograd, running statistics), on given rows;
  * the op-by-op training path of the same drop-in module (group_points + cuDNN + autograd), through the module API.
Outputs, input / weight / gamma / beta gradients and running statistics; bit-reproducibility of the reductions."""
import os
import sys

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def close(got, ref, tol, what):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref).max()
    assert err <= tol * (1 + np.abs(ref).max()), "%s: max err %.3g (scale %.3g)" % (what, err, np.abs(ref).max())


def _case(seed, G, ns, spec):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((G * ns, spec[0])).astype(np.float32)
    ws = [(rng.standard_normal((spec[i + 1], spec[i])) / np.sqrt(spec[i])).astype(np.float32) for i in range(len(spec) - 1)]
    gs = [(1 + 0.2 * rng.standard_normal(c)).astype(np.float32) for c in spec[1:]]
    bs = [(0.1 * rng.standard_normal(c)).astype(np.float32) for c in spec[1:]]
    go = rng.standard_normal((G, spec[-1])).astype(np.float32)
    run = ([rng.standard_normal(c).astype(np.float32) for c in spec[1:]], [(0.5 + rng.random(c)).astype(np.float32) for c in spec[1:]])
    return x, ws, gs, bs, go, run


@pytest.mark.parametrize("shape", [(40, 16, [8, 32, 32, 64]), (24, 32, [36, 64, 48]), (9, 8, [4, 16]), (130, 4, [12, 24, 24, 24]),
                                   (300, 16, [260, 128, 128, 256]), (64, 64, [4, 64, 64, 128])])
def test_rows_forward_backward_vs_reference_semantics(pkg, shape):
    import pointnet2._ext as ext
    import sa_train_ref as T
    G, ns, spec = shape
    x, ws, gs, bs, go, run = _case(G + ns, G, ns, spec)
    ref = T.reference(x, ns, ws, gs, bs, go, running=run, momentum=0.1, eps=1e-5)
    rm = [dev(m.copy()) for m in run[0]]
    rv = [dev(v.copy()) for v in run[1]]
    layers = [(dev(w), dev(g), dev(b), rm[l], rv[l]) for l, (w, g, b) in enumerate(zip(ws, gs, bs))]
    xr = dev(x)
    out, saved = ext.sa_train_forward(None, None, None, None, 1.0, ns, layers, 1e-5, 0.1, use_xyz=False, x_rows=xr, groups=(1, G))
    close(out[0].t().cpu().numpy(), ref["out"], 1e-5, "out")
    for l in range(len(ws)):
        close(rm[l].cpu().numpy(), ref["running_mean"][l], 1e-5, "running_mean[%d]" % l)
        close(rv[l].cpu().numpy(), ref["running_var"][l], 1e-5, "running_var[%d]" % l)
    gout = dev(go.T.copy()).unsqueeze(0).contiguous()                     # (1, C_L, G)
    gin, gw, gg, gb = ext.sa_train_backward(gout, saved, None, ns, layers, 1, 0, G, spec[0], use_xyz=False, x_rows=xr)
    close(gin.cpu().numpy(), ref["grad_x"], 2e-4, "grad_x")
    for l in range(len(ws)):
        close(gw[l].cpu().numpy(), ref["grad_w"][l], 2e-4, "grad_w[%d]" % l)
        close(gg[l].cpu().numpy(), ref["grad_gamma"][l], 2e-4, "grad_gamma[%d]" % l)
        close(gb[l].cpu().numpy(), ref["grad_beta"][l], 2e-4, "grad_beta[%d]" % l)
    # fixed-order reductions: a second run reproduces every parameter gradient bit for bit
    gin2, gw2, gg2, gb2 = ext.sa_train_backward(gout, saved, None, ns, layers, 1, 0, G, spec[0], use_xyz=False, x_rows=xr)
    assert all(torch.equal(a, b) for a, b in zip(gw + gg + gb + [gin], gw2 + gg2 + gb2 + [gin2]))


def _module_pair(spec, npoint, radius, nsample, seed=0):
    import pointnet2_modules as M
    torch.manual_seed(seed)
    a = M.PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=nsample, mlp=list(spec), use_xyz=True, normalize_xyz=True).cuda()
    torch.manual_seed(seed)
    b = M.PointnetSAModuleVotes(npoint=npoint, radius=radius, nsample=nsample, mlp=list(spec), use_xyz=True, normalize_xyz=True).cuda()
    with torch.no_grad():
        for m in list(a.modules()) + list(b.modules()):
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.uniform_(0.7, 1.3)
                m.bias.normal_(0, 0.1)
    b.load_state_dict(a.state_dict())
    return a.train(), b.train()


@pytest.mark.parametrize("shape", [
    (2, 5000, 16, 256, 0.3, 32, [16, 64, 64, 128], "generic"),
    (4, 2048, 128, 1024, 0.4, 32, [128, 128, 128, 256], "SA2 of the backbone at the pretrain batch (backbone_module.py:44-51)"),
    (2, 1024, 256, 512, 0.8, 16, [256, 128, 128, 256], "SA3"),
    (2, 20000, 1, 2048, 0.2, 64, [1, 64, 64, 128], "SA1: one feature channel + xyz"),
])
def test_module_training_fused_vs_op_by_op(pkg, monkeypatch, shape):
    """PointnetSAModuleVotes in train() mode: fused training kernels vs the op-by-op autograd path (B200_SA_TRAIN_FUSED=0)."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B, N, C, npoint, radius, ns, spec, _ = shape
    fused, plain = _module_pair(spec, npoint, radius, ns)
    xyz = dev(cases.scene_cloud(3, B, N)[:, :, :3].copy())
    torch.manual_seed(5)
    f1 = torch.randn(B, C, N, device="cuda", requires_grad=True)
    f2 = f1.detach().clone().requires_grad_(True)
    launches0 = pkg.cabi().launch_count()
    nx1, o1, i1 = fused(xyz, f1)
    assert pkg.cabi().launch_count() - launches0 >= 8
    monkeypatch.setenv("B200_SA_TRAIN_FUSED", "0")
    nx2, o2, i2 = plain(xyz, f2)
    monkeypatch.delenv("B200_SA_TRAIN_FUSED")
    assert torch.equal(i1, i2) and torch.equal(nx1, nx2)
    close(o1.detach().cpu().numpy(), o2.detach().cpu().numpy(), 1e-5, "out")
    g = torch.randn_like(o1)
    o1.backward(g)
    o2.backward(g)
    close(f1.grad.cpu().numpy(), f2.grad.cpu().numpy(), 2e-4, "grad features")
    for (n1, p1), (n2, p2) in zip(fused.named_parameters(), plain.named_parameters()):
        assert n1 == n2 and p1.grad is not None, n1
        close(p1.grad.cpu().numpy(), p2.grad.cpu().numpy(), 2e-4, "grad " + n1)
    for (n1, b1), (n2, b2) in zip(fused.named_buffers(), plain.named_buffers()):
        close(b1.float().cpu().numpy(), b2.float().cpu().numpy(), 1e-5, "buffer " + n1)


def test_teacher_forward_under_no_grad_uses_the_fused_training_forward(pkg):
    """train.py:309-335: the EMA teacher runs .train() under torch.no_grad(): batch statistics, running-stat updates, no graph."""
    fused, plain = _module_pair([16, 64, 64, 128], 256, 0.3, 32)
    xyz = dev(cases.scene_cloud(4, 2, 4000)[:, :, :3].copy())
    f = torch.randn(2, 16, 4000, device="cuda")
    with torch.no_grad():
        c0 = pkg.cabi().launch_count()
        _, o1, _ = fused(xyz, f)
        fused_launches = pkg.cabi().launch_count() - c0
        os.environ["B200_SA_TRAIN_FUSED"] = "0"
        try:
            _, o2, _ = plain(xyz, f)
        finally:
            del os.environ["B200_SA_TRAIN_FUSED"]
    assert fused_launches >= 8
    close(o1.cpu().numpy(), o2.cpu().numpy(), 1e-5, "out")
    for (n1, b1), (_, b2) in zip(fused.named_buffers(), plain.named_buffers()):
        close(b1.float().cpu().numpy(), b2.float().cpu().numpy(), 1e-5, "buffer " + n1)
