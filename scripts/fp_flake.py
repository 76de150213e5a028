"""Diagnostic: is the fused feature-propagation path (b200pn2_fp_rows_forward) run-to-run deterministic, and does it depend
on stale device memory?  Runs test_fp_module's shape 0 (and the FP1 shape) many times in one process, poisoning torch's
free blocks with NaN between runs, and compares every output bit for bit with the first and with the fp32 / fp64 oracle."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
pkg = importlib.import_module("3dioumatch_b200")
pkg.install_dropin()
import cases  # noqa: E402
import torch_ref as tr  # noqa: E402
import pointnet2_modules as M  # noqa: E402
import pointnet2._ext as ext  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def dev(a):
    return None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()


def load(mlp, layers):
    with torch.no_grad():
        for i, ly in enumerate(layers):
            blk = getattr(mlp, "layer%d" % i)
            blk.conv.weight.copy_(dev(ly["weight"]).view_as(blk.conv.weight))
            blk.bn.bn.weight.copy_(dev(ly["gamma"])); blk.bn.bn.bias.copy_(dev(ly["beta"]))
            blk.bn.bn.running_mean.copy_(dev(ly["mean"])); blk.bn.bn.running_var.copy_(dev(ly["var"]))


def poison(rng):
    for _ in range(4):
        n = int(rng.integers(1 << 10, 1 << 22))
        t = torch.full((n,), float("nan"), device="cuda")
        del t
    # other kernels of the library in between (scratch pool reuse)
    pts = torch.rand(2, 1024, 3, device="cuda")
    ext.furthest_point_sampling(pts, 512)


for shape in [(2, 512, 256, 24, 64, [48, 32]), (2, 512, 256, 256, 256, [256, 256]), (1, 300, 7, 0, 32, [64])]:
    B, n, m, C1, C2, spec = shape
    rng = np.random.default_rng(0)
    unknown = cases.cloud(1, B, n)
    known = unknown[:, :m].copy() if m <= n else cases.cloud(2, B, m)
    uf = rng.standard_normal((B, C1, n)).astype(np.float32) if C1 else None
    kf = rng.standard_normal((B, C2, m)).astype(np.float32)
    layers = cases.mlp_params(3, [C1 + C2] + spec)
    ref = tr.fp_forward(unknown, known, uf, kf, layers)
    bound = 1e-5 + 1e-5 * np.abs(ref)
    prng = np.random.default_rng(5)
    first = None
    n_diff = 0
    worst = 0.0
    for it in range(60):
        fp = M.PointnetFPModule(mlp=[C1 + C2] + spec).cuda().eval()
        load(fp.mlp, layers)
        if it % 2:
            poison(prng)
        with torch.no_grad():
            got = fp(dev(unknown), dev(known), dev(uf), dev(kf)).cpu().numpy()
        err = np.abs(got - ref)
        ratio = float(np.nanmax(err / bound)) if np.isfinite(got).all() else float("inf")
        worst = max(worst, ratio)
        if first is None:
            first = got
        elif not np.array_equal(first, got):
            n_diff += 1
            d = np.abs(got - first)
            w = np.unravel_index(np.nanargmax(np.where(np.isfinite(d), d, np.inf)), d.shape)
            print("  iter %d differs from iter 0: %d elements, max |diff| %.3g at %s (ref %.4g), nonfinite %d; err/bound %.3f"
                  % (it, int((d != 0).sum()), float(d[w]), w, float(ref[w]), int((~np.isfinite(got)).sum()), ratio))
    print("shape %s: %d of 59 repeats differ from the first run; worst err/bound vs fp32 oracle %.3f" % (shape[:5], n_diff, worst))
