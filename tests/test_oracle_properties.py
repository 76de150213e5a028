"""CPU: the C oracle against an independent, deliberately naive numpy restatement of the same reference semantics on
random small inputs (hypothesis), plus the algebra behind the factorised first layer of the fused SA kernel.

The numpy versions below are written from SURVEY.md appendix A, not from oracle/*.c: a second opinion on the checker."""
import numpy as np
from hypothesis import given, settings, strategies as st

import cases

F = np.float32


def _sqdist(a, b):
    """fma(dz,dz, fma(dx,dx, dy*dy)) in fp32, via float64 products rounded once per step (exact for fp32 inputs)."""
    d = (a.astype(F) - b.astype(F)).astype(F)
    t = F(np.float64(d[..., 1]) * np.float64(d[..., 1]))
    t = (np.float64(d[..., 0]) * np.float64(d[..., 0]) + np.float64(t)).astype(F)
    return (np.float64(d[..., 2]) * np.float64(d[..., 2]) + np.float64(t)).astype(F)


def _ball_query(new_xyz, xyz, radius, ns):
    r2 = F(radius) * F(radius)
    out = np.zeros((new_xyz.shape[0], ns), np.int32)
    for j, c in enumerate(new_xyz):
        hits = np.nonzero(_sqdist(c[None, :], xyz) < r2)[0][:ns]
        if len(hits):
            out[j, :] = hits[0]
            out[j, : len(hits)] = hits
    return out


def _three_nn(unknown, known):
    d = np.stack([_sqdist(u[None, :], known) for u in unknown])
    order = np.lexsort((np.broadcast_to(np.arange(known.shape[0]), d.shape), d), axis=1)[:, :3]  # (d asc, index asc)
    return np.take_along_axis(d, order, 1), order.astype(np.int32)


def _bitrev(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def _fps(xyz, m, bs):
    n = xyz.shape[0]
    bits = bs.bit_length() - 1
    temp = np.full(n, 1e10, F)
    mag = _sqdist(xyz, np.zeros_like(xyz))
    alive = ~(mag.astype(np.float64) <= 1e-3)
    idx = [0]
    for _ in range(1, m):
        d2 = np.minimum(_sqdist(xyz, xyz[idx[-1]][None, :]), temp)
        temp = np.where(alive, d2, temp).astype(F)          # skipped points keep their value but never compete
        cand = [(-float(temp[k]), _bitrev(k % bs, bits), k) for k in range(n) if alive[k]]
        idx.append(min(cand)[2] if cand else 0)
    return np.asarray(idx, np.int32)


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10_000), st.integers(5, 120), st.integers(1, 12), st.sampled_from([1, 3, 8, 16]),
       st.floats(0.2, 1.5))
def test_ball_query_matches_naive(orc, seed, n, m, ns, radius):
    xyz = cases.cloud(seed, 1, n, extent=(2.0, 2.0, 1.0), dup_frac=0.1)[0]
    centres = xyz[np.random.default_rng(seed).integers(0, n, m)] + F(0.05)
    got = orc.ball_query(centres[None], xyz[None], radius, ns)[0]
    assert np.array_equal(got, _ball_query(centres, xyz, radius, ns))


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 10_000), st.integers(1, 60), st.integers(3, 80))
def test_three_nn_matches_naive_including_ties(orc, seed, n, m):
    rng = np.random.default_rng(seed)
    known = np.round(rng.random((m, 3)) * 4).astype(F) / F(2)        # coarse lattice: many exact distance ties
    unknown = np.round(rng.random((n, 3)) * 4).astype(F) / F(2)
    d, i = orc.three_nn(unknown[None], known[None])
    rd, ri = _three_nn(unknown, known)
    assert np.array_equal(i[0], ri) and np.array_equal(d[0], rd)


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 10_000), st.sampled_from([7, 16, 33, 64, 130]), st.integers(2, 16))
def test_fps_matches_naive_with_ties_and_origin_points(orc, seed, n, m):
    rng = np.random.default_rng(seed)
    xyz = np.round(rng.random((n, 3)) * 3).astype(F)                 # lattice: exact ties everywhere, some points at 0
    m = min(m, n)
    bs = orc.opt_n_threads(n)
    assert np.array_equal(orc.furthest_point_sampling(xyz[None], m)[0], _fps(xyz, m, bs))


@settings(max_examples=20, deadline=None)
@given(st.integers(0, 10_000), st.integers(1, 5), st.integers(2, 40), st.integers(1, 30), st.integers(1, 6))
def test_scatter_gradients_are_adjoint_to_their_forward(orc, seed, C, N, M, ns):
    """<group(points), g> == <points, group_grad(g)> (and the same for gather / three_interpolate), in float64."""
    rng = np.random.default_rng(seed)
    pts = rng.standard_normal((1, C, N)).astype(F)
    idx = rng.integers(0, N, (1, M, ns)).astype(np.int32)
    g = rng.standard_normal((1, C, M, ns)).astype(F)
    lhs = np.sum(orc.group_points(pts, idx).astype(np.float64) * g)
    rhs = np.sum(pts.astype(np.float64) * orc.group_points_grad(g, idx, N))
    assert abs(lhs - rhs) <= 1e-4 * (1 + abs(lhs))
    gi = idx[:, :, 0].copy()
    g2 = rng.standard_normal((1, C, M)).astype(F)
    lhs = np.sum(orc.gather_points(pts, gi).astype(np.float64) * g2)
    rhs = np.sum(pts.astype(np.float64) * orc.gather_points_grad(g2, gi, N))
    assert abs(lhs - rhs) <= 1e-4 * (1 + abs(lhs))
    tidx = rng.integers(0, N, (1, M, 3)).astype(np.int32)
    w = rng.random((1, M, 3)).astype(F)
    lhs = np.sum(orc.three_interpolate(pts, tidx, w).astype(np.float64) * g2)
    rhs = np.sum(pts.astype(np.float64) * orc.three_interpolate_grad(g2, tidx, w, N))
    assert abs(lhs - rhs) <= 1e-4 * (1 + abs(lhs))


def test_first_layer_factorisation_algebra():
    """relu(s*(W1 [rel | f_j]) + t) == relu((s*(W1f f_j) + t) + (s*W1x) rel) -- the per-point term P and the
    per-row xyz term the fused kernel adds in its gather; equal to fp32 rounding (the kernel's 1e-5 bar)."""
    rng = np.random.default_rng(3)
    C, H, rows, pts = 128, 128, 500, 64
    layer = cases.mlp_params(5, [3 + C, H])[0]
    W = layer["weight"].astype(np.float64)
    s = (layer["gamma"] / np.sqrt(layer["var"] + 1e-5)).astype(np.float64)
    t = (layer["beta"] - layer["mean"] * s).astype(np.float64)
    f = rng.standard_normal((pts, C))
    src = rng.integers(0, pts, rows)
    rel = rng.standard_normal((rows, 3)) * 0.5
    direct = np.maximum(s * (np.concatenate([rel, f[src]], 1) @ W.T) + t, 0)
    P = (s * (f @ W[:, 3:].T) + t).astype(F)                          # pass 1, stored in fp32
    wx = (s[:, None] * W[:, :3]).astype(F)                            # [H][3]
    fused = np.maximum(P[src].astype(F) + (rel.astype(F) @ wx.T).astype(F), 0)
    assert np.abs(fused - direct).max() <= 1e-5 + 1e-5 * np.abs(direct).max()
    # three-neighbour blend (GridConv rows): blending P rows == P of the blended features when the weights sum to 1
    w3 = rng.random((rows, 3)); w3 /= w3.sum(1, keepdims=True)
    i3 = rng.integers(0, pts, (rows, 3))
    blend_f = (w3[:, :, None] * f[i3]).sum(1)
    direct = np.maximum(s * (np.concatenate([rel, blend_f], 1) @ W.T) + t, 0)
    blend_p = (w3[:, :, None].astype(F) * P[i3]).sum(1)
    fused = np.maximum(blend_p + (rel.astype(F) @ wx.T), 0)
    assert np.abs(fused - direct).max() <= 1e-5 + 1e-5 * np.abs(direct).max()


def _prefix_speculation_flag(xyz, m, bs, tile=None):
    """numpy restatement of the verify-then-skip test of 3dioumatch_b200/csrc/fps.cu (fps_prefix_values_kernel +
    fps_prefix_check_kernel): True = "the reference would not pick 0, 1, ..., m-1".  tile = None: the full rule in every
    column; tile = T: the shipped two-stage form (a tile of T columns is examined with the full rule only if some column
    had run >= V)."""
    n = xyz.shape[0]
    bits = bs.bit_length() - 1
    key = np.asarray([(_bitrev(k % bs, bits) << 22) | (k >> bits) for k in range(n)], np.int64)
    mag = _sqdist(xyz, np.zeros_like(xyz))
    run = np.where(mag.astype(np.float64) <= 1e-3, F(-1), F(1e10)).astype(F)     # temp[k]; skipped points pinned at -1
    R = np.empty((n, m), F)                                                       # R[k, j]: temp[k] as step j sees it
    for j in range(m):
        R[:, j] = run
        run = np.minimum(_sqdist(xyz, xyz[j][None, :]), run).astype(F)
    V = R[np.arange(m), np.arange(m)].copy()
    V[0] = np.inf                                                                 # pick 0 is never contested
    other = np.arange(n)[:, None] != np.arange(m)[None, :]
    beats = ((R > V[None, :]) | ((R == V[None, :]) & (key[:, None] < key[None, :m]))) & other
    beats[:, 0] = False
    if tile is None:
        return bool(beats.any())
    ge = (R >= V[None, :]) & other
    flag = False
    for j0 in range(0, m, tile):
        examined = ge[:, j0:j0 + tile].any(axis=1)                                # per point: replay this tile?
        flag = flag or bool(beats[examined, j0:j0 + tile].any())
    return flag


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 100_000), st.integers(4, 150), st.integers(2, 150), st.booleans(), st.sampled_from([0, 1, 2, 3]))
def test_fps_prefix_speculation_is_exact(orc, seed, n, m, lattice, damage):
    """The hierarchical FPS levels skip the serial kernel when three parallel kernels prove that the answer is
    0, 1, ..., m-1.  The proof rule, restated in numpy, must say "holds" exactly when the oracle
    (reference semantics: tie order by bit-reversed thread id, origin skip) returns 0..m-1 -- on clouds in furthest-point
    order, with exact ties, duplicated points, points inside the skip sphere and arbitrary permutations -- and the
    shipped two-stage form (cheap run >= V filter, full rule only for tiles that trip it) must decide identically."""
    rng = np.random.default_rng(seed)
    m = min(m, n)
    xyz = (np.round(rng.random((n, 3)) * 4) if lattice else rng.random((n, 3)) * 2 - 1).astype(F)
    order = orc.furthest_point_sampling(xyz[None], m)[0].astype(np.int64)
    rest = np.setdiff1d(np.arange(n), order)
    picks, first = np.unique(order, return_index=True)                          # a dead cloud repeats index 0
    pts = np.concatenate([xyz[order[np.sort(first)]], xyz[rest]])[:n]
    if pts.shape[0] < n:
        pts = np.concatenate([pts, xyz[: n - pts.shape[0]]])
    if damage == 1 and n > 3:                                                   # duplicate an early pick further back
        pts[rng.integers(n // 2, n)] = pts[rng.integers(0, max(1, min(m, n // 2)))]
    elif damage == 2:                                                           # a point inside the origin-skip sphere
        pts[rng.integers(0, n)] = F([0.01, -0.01, 0.005])
    elif damage == 3:                                                           # arbitrary order
        pts = pts[rng.permutation(n)]
    bs = orc.opt_n_threads(n)
    holds = bool(np.array_equal(orc.furthest_point_sampling(pts[None], m)[0], np.arange(m)))
    assert _prefix_speculation_flag(pts, m, bs) == (not holds)
    assert _prefix_speculation_flag(pts, m, bs, tile=16) == (not holds)
