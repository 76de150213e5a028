"""Training-mode set-abstraction MLP (conv1x1 -> BatchNorm2d with BATCH statistics -> ReLU, ..., max over nsample) --
TEST INFRASTRUCTURE ONLY: the oracle for SURVEY.md section 8f row n4 (second half), written before the kernel.

Two functions over the same inputs (rows = every (scene, centre, sample) of the grouped tensor, row-major):

* `reference(...)`   the reference semantics, literally: `nn.Conv2d(bias=False)` -> `nn.BatchNorm2d` in training mode ->
  `ReLU` per layer (pointnet2/pytorch_utils.py:14-39,70-123), `F.max_pool2d` over nsample
  (pointnet2/pointnet2_modules.py:256-262), gradients by torch autograd, running statistics updated with torch's
  momentum rule (unbiased variance).
* `tiled(...)`       the algorithm a fused kernel can run WITHOUT materialising the (rows, C) activations in HBM: every
  pass walks the rows in tiles of 128 and recomputes the activations it needs from the gathered input rows.
    forward : one statistics pass per layer l (recompute layers < l with their final affine, accumulate the
              per-channel sum and sum of squares of layer l's conv output as per-tile partials, combined in tile order),
              then one output pass (max + arg-max row per centre and channel);
    backward: per layer, from the last to the first, two passes -- (1) the per-channel sums BatchNorm's gradient needs
              (sum dy, sum dy * xhat), (2) dz = gamma / sigma * (dy - mean(dy) - xhat * mean(dy * xhat)), dW += dz^T a,
              da = dz W -- each recomputing the forward of its tile.  L statistics passes + 1 forward, 2L backward.
  Deterministic by construction (fixed tile order, no atomics).  `tests/test_oracle_sa_train.py` checks it against
  `reference` (outputs, input gradients, weight / gamma / beta gradients, running statistics).
"""
import numpy as np
import torch
import torch.nn.functional as F

TILE = 128


def reference(x_rows, nsample, weights, gammas, betas, grad_out, running=None, momentum=0.1, eps=1e-5):
    """x_rows (R, C0) float32, R = centres * nsample; weights[l] (C_{l+1}, C_l); grad_out (centres, C_L).
    Returns dict(out, grad_x, grad_w, grad_gamma, grad_beta, running_mean, running_var)."""
    R, C0 = x_rows.shape
    G = R // nsample
    x = torch.from_numpy(x_rows).clone().requires_grad_(True)
    ws = [torch.from_numpy(w).clone().requires_grad_(True) for w in weights]
    gs = [torch.from_numpy(g).clone().requires_grad_(True) for g in gammas]
    bs = [torch.from_numpy(b).clone().requires_grad_(True) for b in betas]
    rm = [torch.zeros(w.shape[0]) if running is None else torch.from_numpy(running[0][l]).clone() for l, w in enumerate(weights)]
    rv = [torch.ones(w.shape[0]) if running is None else torch.from_numpy(running[1][l]).clone() for l, w in enumerate(weights)]
    a = x.t().reshape(1, C0, G, nsample)                                # (B=1, C, npoint, nsample): BN sees all rows
    for l, w in enumerate(ws):
        # the 1x1 convolution as a plain fp32 matrix product (same math as F.conv2d with a (Cout, Cin, 1, 1) kernel): the
        # host's oneDNN convolution is kept out of the checker (oracle/torch_ref.py explains why)
        a = torch.matmul(w, a.reshape(1, a.shape[1], G * nsample)).reshape(1, w.shape[0], G, nsample)
        a = F.batch_norm(a, rm[l], rv[l], gs[l], bs[l], training=True, momentum=momentum, eps=eps)
        a = F.relu(a)
    out = F.max_pool2d(a, kernel_size=[1, nsample]).squeeze(-1).squeeze(0).t()   # (G, C_L)
    out.backward(torch.from_numpy(grad_out))
    return dict(out=out.detach().numpy(), grad_x=x.grad.numpy(), grad_w=[w.grad.numpy() for w in ws],
                grad_gamma=[g.grad.numpy() for g in gs], grad_beta=[b.grad.numpy() for b in bs],
                running_mean=[m.numpy() for m in rm], running_var=[v.numpy() for v in rv])


def _tiles(R):
    return [(s, min(s + TILE, R)) for s in range(0, R, TILE)]


def tiled(x_rows, nsample, weights, gammas, betas, grad_out, running=None, momentum=0.1, eps=1e-5):
    """Same contract as `reference`, computed tile by tile with recomputation (see the module docstring).
    fp32 tile math, per-tile partial sums combined in float64 in tile order (what a deterministic kernel would do)."""
    f32 = np.float32
    R, _ = x_rows.shape
    G, L = R // nsample, len(weights)
    assert TILE % nsample == 0, "a tile holds whole centres"
    scale, shift, mean, var = [], [], [], []

    def forward_tile(rows, upto):
        """activations a_0..a_upto of a tile (a_0 = input rows) and the normalised pre-activations xhat_1..xhat_upto"""
        acts, xhats = [rows], []
        for l in range(upto):
            z = acts[-1] @ weights[l].T
            xh = (z - mean[l]) * (f32(1.0) / np.sqrt(var[l] + f32(eps)))
            xhats.append(xh.astype(f32))
            acts.append(np.maximum(z * scale[l] + shift[l], 0).astype(f32))
        return acts, xhats

    # ---- forward: one statistics pass per layer --------------------------------------------------------------------
    for l in range(L):
        s1 = np.zeros(weights[l].shape[0], np.float64)
        s2 = np.zeros(weights[l].shape[0], np.float64)
        for a0, a1 in _tiles(R):
            acts, _ = forward_tile(x_rows[a0:a1], l)
            z = (acts[-1] @ weights[l].T).astype(f32)
            s1 += z.sum(0, dtype=f32)                               # per-tile fp32 partials, fp64 combination
            s2 += (z * z).sum(0, dtype=f32)
        m = s1 / R
        v = np.maximum(s2 / R - m * m, 0.0)                         # biased variance, as BN normalises with
        mean.append(m.astype(f32)); var.append(v.astype(f32))
        sc = (gammas[l] / np.sqrt(v + eps)).astype(f32)
        scale.append(sc); shift.append((betas[l] - m * sc).astype(f32))
    # ---- forward: output pass (max + arg-max row per centre and channel) --------------------------------------------
    CL = weights[-1].shape[0]
    out = np.empty((G, CL), f32)
    arg = np.empty((G, CL), np.int64)
    for a0, a1 in _tiles(R):
        acts, _ = forward_tile(x_rows[a0:a1], L)
        y = acts[-1].reshape(-1, nsample, CL)
        g0 = a0 // nsample
        out[g0:g0 + y.shape[0]] = y.max(1)
        arg[g0:g0 + y.shape[0]] = y.argmax(1) + (np.arange(y.shape[0])[:, None] * nsample + a0)   # first maximum wins
    # ---- running statistics (torch: momentum, UNBIASED variance) -----------------------------------------------------
    rmean = [np.zeros_like(m) if running is None else running[0][l].copy() for l, m in enumerate(mean)]
    rvar = [np.ones_like(v) if running is None else running[1][l].copy() for l, v in enumerate(var)]
    for l in range(L):
        rmean[l] = ((1 - momentum) * rmean[l] + momentum * mean[l]).astype(f32)
        rvar[l] = ((1 - momentum) * rvar[l] + momentum * var[l] * (R / max(R - 1, 1))).astype(f32)

    # ---- backward ---------------------------------------------------------------------------------------------------
    grad_w = [np.zeros_like(w, dtype=np.float64) for w in weights]
    grad_gamma, grad_beta = [None] * L, [None] * L

    def dy_tile(a0, a1, l, chain):
        """gradient w.r.t. layer l's BN output (ReLU mask applied) for one tile; `chain` = finished (sum dy, sum dy xhat)
        of the layers above l, needed to push the gradient down through their BatchNorms"""
        acts, xhats = forward_tile(x_rows[a0:a1], L)
        g0 = a0 // nsample
        d = np.zeros((a1 - a0, CL), f32)                            # route grad_out to the arg-max rows
        for gi in range((a1 - a0) // nsample):
            rows = arg[g0 + gi] - a0
            d[rows, np.arange(CL)] = grad_out[g0 + gi]
        for k in range(L - 1, l - 1, -1):
            dy = d * (acts[k + 1] > 0)
            if k == l:
                return dy, xhats[k], acts[k]
            sdy, sdyx = chain[k]
            dz = scale[k] * (dy - (sdy / R).astype(f32) - xhats[k] * (sdyx / R).astype(f32))
            d = (dz @ weights[k]).astype(f32)
        raise AssertionError

    chain = {}
    grad_x = np.zeros_like(x_rows)
    for l in range(L - 1, -1, -1):
        sdy = np.zeros(weights[l].shape[0], np.float64)
        sdyx = np.zeros(weights[l].shape[0], np.float64)
        for a0, a1 in _tiles(R):                                    # pass 1: the sums BatchNorm's gradient needs
            dy, xh, _ = dy_tile(a0, a1, l, chain)
            sdy += dy.sum(0, dtype=f32)
            sdyx += (dy * xh).sum(0, dtype=f32)
        chain[l] = (sdy, sdyx)
        grad_beta[l] = sdy.astype(f32)
        grad_gamma[l] = sdyx.astype(f32)
        for a0, a1 in _tiles(R):                                    # pass 2: dz, weight gradient, gradient of the input
            dy, xh, a_in = dy_tile(a0, a1, l, chain)
            dz = scale[l] * (dy - (sdy / R).astype(f32) - xh * (sdyx / R).astype(f32))
            grad_w[l] += (dz.T @ a_in).astype(f32)
            if l == 0:
                grad_x[a0:a1] = dz @ weights[0]
    return dict(out=out, grad_x=grad_x, grad_w=[g.astype(f32) for g in grad_w], grad_gamma=grad_gamma,
                grad_beta=grad_beta, running_mean=rmean, running_var=rvar)
