"""CPU: the tiled, recompute-based training-mode SA algorithm (oracle/sa_train_ref.py: per-layer statistics passes, an
output pass with arg-max, two backward passes per layer) against the reference semantics run literally through torch
(conv1x1 -> BatchNorm2d(training) -> ReLU -> max_pool2d, autograd).  This is the oracle for SURVEY 8f row n4's fused
backward, pinned before the kernel exists."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _case(seed, G, ns, spec):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((G * ns, spec[0])).astype(np.float32)
    ws = [(rng.standard_normal((spec[i + 1], spec[i])) / np.sqrt(spec[i])).astype(np.float32) for i in range(len(spec) - 1)]
    gs = [(1 + 0.2 * rng.standard_normal(c)).astype(np.float32) for c in spec[1:]]
    bs = [(0.1 * rng.standard_normal(c)).astype(np.float32) for c in spec[1:]]
    go = rng.standard_normal((G, spec[-1])).astype(np.float32)
    run = ([rng.standard_normal(c).astype(np.float32) for c in spec[1:]], [(0.5 + rng.random(c)).astype(np.float32) for c in spec[1:]])
    return x, ws, gs, bs, go, run


@pytest.mark.parametrize("shape", [(40, 16, [7, 32, 32, 64]), (24, 32, [35, 64, 48]), (9, 8, [6, 16]), (130, 4, [10, 24, 24, 24])])
def test_tiled_training_algorithm_matches_autograd(shape):
    import sa_train_ref as T
    G, ns, spec = shape
    x, ws, gs, bs, go, run = _case(G + ns, G, ns, spec)
    ref = T.reference(x, ns, ws, gs, bs, go, running=run)
    got = T.tiled(x, ns, ws, gs, bs, go, running=run)

    def close(a, b, tol=2e-4):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert np.abs(a - b).max() <= tol * (1 + np.abs(b).max()), np.abs(a - b).max()

    close(got["out"], ref["out"], 1e-5)
    close(got["grad_x"], ref["grad_x"])
    for l in range(len(ws)):
        close(got["grad_w"][l], ref["grad_w"][l])
        close(got["grad_gamma"][l], ref["grad_gamma"][l])
        close(got["grad_beta"][l], ref["grad_beta"][l])
        close(got["running_mean"][l], ref["running_mean"][l], 1e-5)
        close(got["running_var"][l], ref["running_var"][l], 1e-5)


def test_tiled_algorithm_is_deterministic_and_tile_order_fixed():
    import sa_train_ref as T
    x, ws, gs, bs, go, run = _case(3, 64, 16, [12, 32, 32])
    a = T.tiled(x, 16, ws, gs, bs, go, running=run)
    b = T.tiled(x, 16, ws, gs, bs, go, running=run)
    for k in ("out", "grad_x"):
        assert np.array_equal(a[k], b[k])
    assert all(np.array_equal(p, q) for p, q in zip(a["grad_w"], b["grad_w"]))
