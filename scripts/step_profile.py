"""Kernel-time table of one eager step of bench.py's workload (torch.profiler / CUPTI), one stream, for development.
  python scripts/step_profile.py [reference|fast]"""
import importlib
import os
import sys
from collections import defaultdict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ra = importlib.import_module("3dioumatch_b200.refapp")
callers = sys.argv[1] if len(sys.argv) > 1 else "reference"
ns = ra.load(ra.dropin_paths(fast_callers=(callers == "fast")), name="b200")
net, cfg = ra.build_votenet(ns, "scannet", 256, seed=1)
ns.pt.freeze_inference(net)
pc, labels = ra.make_inputs(8, 40000, seed=0, cfg=cfg)
pc = torch.from_numpy(pc).cuda()
lab = {k: torch.from_numpy(v).cuda() for k, v in labels.items()}
from torch.profiler import ProfilerActivity, profile
with torch.no_grad():
    for _ in range(3):
        ra.forward_with_iou_labels(ns, net, cfg, pc, lab)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ra.forward_with_iou_labels(ns, net, cfg, pc, lab)
        torch.cuda.synchronize()
agg = defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        t = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        agg[e.name[:110]][0] += t
        agg[e.name[:110]][1] += 1
tot = sum(v[0] for v in agg.values())
print("callers=%s  total kernel time %.3f ms, %d launches" % (callers, tot / 1e3, sum(v[1] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
    print("%9.1f us %5.1f%% x%-3d %s" % (v[0], 100 * v[0] / tot, v[1], k))
