"""Kernel-time table of one harness step (torch.profiler / CUPTI), lanes=1, for development."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
harness = importlib.import_module("3dioumatch_b200.harness")
ops = harness.stack_b200()
net = harness.make_model(ops, seed=1)
pc, gt = harness.make_inputs(8, 40000, 64, seed=0)
pc, gt = torch.from_numpy(pc).cuda(), torch.from_numpy(gt).cuda()
with torch.no_grad():
    for _ in range(3):
        net(pc, gt)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            net(pc, gt)
        torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t:
        rows.append((t / 3.0, e.count / 3.0, e.key[:100]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print("total kernel time per step: %.1f us" % tot)
for t, c, k in rows[:40]:
    print("%9.1f us %5.1f  %s" % (t, c, k))
