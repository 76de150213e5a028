"""Development: repeat the FP-module parity case and report where the worst elements are."""
import importlib, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
pkg = importlib.import_module("3dioumatch_b200"); pkg.install_dropin()
import cases, oracle as orc, torch_ref as tr
orc.build()
import pointnet2_modules as M
dev = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).cuda()
B, n, m, C1, C2, spec = 2, 512, 256, 24, 64, [48, 32]
rng = np.random.default_rng(0)
unknown = cases.cloud(1, B, n); known = unknown[:, :m].copy()
uf = rng.standard_normal((B, C1, n)).astype(np.float32); kf = rng.standard_normal((B, C2, m)).astype(np.float32)
layers = cases.mlp_params(3, [C1 + C2] + spec)
fp = M.PointnetFPModule(mlp=[C1 + C2] + spec).cuda().eval()
with torch.no_grad():
    for i, ly in enumerate(layers):
        blk = getattr(fp.mlp, "layer%d" % i)
        blk.conv.weight.copy_(dev(ly["weight"]).view_as(blk.conv.weight)); blk.bn.bn.weight.copy_(dev(ly["gamma"])); blk.bn.bn.bias.copy_(dev(ly["beta"]))
        blk.bn.bn.running_mean.copy_(dev(ly["mean"])); blk.bn.bn.running_var.copy_(dev(ly["var"]))
ref = tr.fp_forward(unknown, known, uf, kf, layers)
outs = []
with torch.no_grad():
    for it in range(6):
        got = fp(dev(unknown), dev(known), dev(uf), dev(kf)).cpu().numpy()
        outs.append(got)
        err = np.abs(got - ref); bad = err > 1e-5 + 1e-5 * np.abs(ref)
        print("iter", it, "bad", int(bad.sum()), "max", float(err.max()), "where", np.argwhere(bad)[:6].tolist())
print("deterministic:", all(np.array_equal(outs[0], o) for o in outs))
# is it the weights?  points whose nearest known point is (almost) themselves
d2, idx = orc.three_nn(unknown, known)
bad = np.abs(outs[0] - ref) > 1e-5 + 1e-5 * np.abs(ref)
pts = sorted(set((b, nn) for b, c, nn in np.argwhere(bad).tolist()))
for b, nn in pts[:8]:
    print("point", b, nn, "dist2", d2[b, nn], "idx", idx[b, nn])
