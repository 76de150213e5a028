"""Helper process for test_gpu_harness_vs_reference.py: runs the VoteNet-path harness on one operator stack and saves
the outputs (the two stacks define modules with the same names, so they cannot share a process)."""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--impl", required=True)
ap.add_argument("--out", required=True)
ap.add_argument("--batch", type=int, default=2)
ap.add_argument("--points", type=int, default=20000)
a = ap.parse_args()

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
harness = importlib.import_module("3dioumatch_b200.harness")
ops = harness.stack_from_path(os.path.join(ROOT, "oracle", "_ref")) if a.impl == "reference" else harness.stack_b200()
net = harness.make_model(ops, seed=1)
pc, gt = harness.make_inputs(a.batch, a.points, 64, seed=0)
with torch.no_grad():
    out = net(torch.from_numpy(pc).cuda(), torch.from_numpy(gt).cuda())
torch.cuda.synchronize()
np.savez(a.out, **{k: v.detach().cpu().numpy() for k, v in out.items()})
