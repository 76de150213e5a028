"""GPU: end-to-end drop-in check.  The same VoteNet-with-IoU-branch module graph (3dioumatch_b200/harness.py, seeded
weights) is run once on this package's operator stack (fused sm_100a kernels) and once on the UNMODIFIED reference
operator stack of oracle/_ref (reference python modules + reference CUDA extensions, cuDNN convs with TF32 off).
Sampled indices must be identical; float outputs agree to the accumulated fp32 rounding of a 20-layer network."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(impl, out):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_run_harness.py"), "--impl", impl, "--out", out],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return np.load(out)


def test_votenet_path_matches_reference_stack(tmp_path):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "pointnet2", "_ext.so")):
        pytest.skip("oracle/_ref not present")
    ours = run("b200", str(tmp_path / "ours.npz"))
    ref = run("reference", str(tmp_path / "ref.npz"))
    assert np.array_equal(ours["seed_inds"], ref["seed_inds"])                      # FPS of SA1 (sa1_inds[:, :1024])
    assert np.array_equal(ours["aggregated_vote_inds"], ref["aggregated_vote_inds"])  # FPS on the seeds
    assert np.array_equal(ours["seed_xyz"], ref["seed_xyz"])
    for key, tol in (("vote_xyz", 2e-4), ("center", 5e-4), ("size", 5e-4), ("heading", 5e-4), ("iou_scores", 2e-3),
                     ("iou_labels", 2e-3)):
        err = np.abs(ours[key] - ref[key])
        # a neighbour flipping across a ball boundary after 1e-6 drift may change isolated rows; bound the bulk
        assert np.quantile(err, 0.99) <= tol, (key, float(np.quantile(err, 0.99)), float(err.max()))
    assert (ref["iou_labels"] > 0).any()
