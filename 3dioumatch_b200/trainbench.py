"""bench.py --config c4 | c5: the training steps of BASELINE.json configs[3] / configs[4] through the reference's own
training code, one process per GPU, scenes sharded, ONE gradient exchange per step on NCCL.

  c4  pretrain.py:310-345  net.train(); VoteNet.forward_with_pred_jitter (votenet_iou_branch.py:157-181) ->
      loss_helper_labeled.get_labeled_loss (:300-370) -> backward -> gradient all-reduce -> Adam.  SUN RGB-D-shaped
      synthetic scenes (N=20000, C=4, 10 classes / 12 heading bins / 10 size clusters, K=128), 4 scenes per GPU
      (16 on 4 GPUs, as BASELINE configs[3]).
  c5  train.py:305-371     both models .train(); teacher (EMA) forward under no_grad, student forward, labeled +
      unlabeled (pseudo-label, IoU filter, LHS) losses, backward, gradient all-reduce, Adam, EMA update.  ScanNet-shaped
      scenes (N=40000), 4 labeled + 8 unlabeled scenes per GPU (train.py:48), K=128.

Both arms run the SAME reference files (baseline/_ref) on their operator stack; BatchNorm runs on batch statistics, so
the SA layers take this package's training path (fused training kernels where available, the differentiable op-by-op
path on the same sm_100a kernels otherwise).  The collective is shard.GradientBuckets: flat buckets all-reduced
asynchronously from autograd hooks, overlapped with the tail of backward.
"""
import gc
import importlib
import json
import os
import sys
import time

import numpy as np


def _labels_to_device(torch, labels, dev):
    return {k: torch.from_numpy(v).to(dev) for k, v in labels.items()}


def main(a, ROOT, ClockSampler, load_stack):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and world > 1 and rank != 0:
        return 0
    distributed = world > 1 and a.impl != "reference"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if distributed:
        dist.init_process_group("nccl", device_id=dev)
    ns = load_stack(a.impl, a.callers, with_losses=True)
    if ns is None:
        print(json.dumps({"impl": a.impl, "unavailable": "baseline/_ref / oracle/_ref not built: python oracle/build_ref.py"}))
        return 0
    ra = importlib.import_module("3dioumatch_b200.refapp")
    shard = importlib.import_module("3dioumatch_b200.shard")
    cabi = importlib.import_module("3dioumatch_b200._cabi") if a.impl == "b200" else None
    c5 = a.config == "c5"
    dataset = "scannet" if c5 else "sunrgbd"
    N = a.points or (40000 if c5 else 20000)
    K = a.proposals or 128
    n_lab, n_unl = (4, 8) if c5 else (a.batch or 4, 0)
    if c5 and a.batch:
        n_lab, n_unl = max(a.batch // 3, 1), a.batch - max(a.batch // 3, 1)
    B = n_lab + n_unl
    room = (8.0, 8.0, 3.0) if c5 else (5.0, 5.0, 2.5)

    net, cfg = ra.build_votenet(ns, dataset, K, seed=1, device=dev, train=True)
    ema = None
    if c5:
        ema, _ = ra.build_votenet(ns, dataset, K, seed=1, device=dev, train=True)
        for p in ema.parameters():
            p.detach_()
    if distributed:
        shard.broadcast_parameters(net)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    buckets = shard.GradientBuckets(net, n_buckets=3)
    config_dict = {"dataset_config": cfg, "unlabeled_batch_size": n_unl, "dataset": dataset, "use_lhs": True,
                   "nms_iou": 0.25, "use_old_type_nms": False, "obj_threshold": 0.9, "cls_threshold": 0.9,
                   "iou_threshold": 0.25, "samecls_match": False, "view_stats": False}

    # ---- inputs: a few distinct batches per rank, device resident (labels are small) ----------------------------------
    n_rot = 4
    batches = []
    for i in range(n_rot):
        pc, labels = ra.make_inputs(B, N, seed=rank * 100 + i, room=room, cfg=cfg)
        d = _labels_to_device(torch, labels, dev)
        d["point_clouds"] = torch.from_numpy(pc).to(dev)
        d["supervised_mask"] = torch.cat([torch.ones(n_lab), torch.zeros(n_unl)]).long().to(dev)
        if c5:
            d["ema_point_clouds"] = d["point_clouds"].clone()
            d["flip_x_axis"] = torch.zeros(B, dtype=torch.long, device=dev)
            d["flip_y_axis"] = torch.zeros(B, dtype=torch.long, device=dev)
            d["rot_mat"] = torch.eye(3, device=dev).unsqueeze(0).repeat(B, 1, 1)
            d["scale"] = torch.ones(B, 1, 3, device=dev)
        batches.append(d)
    global_step = [0]
    exposed = []

    def step(i):
        d = {k: (v.clone() if k == "center_label" else v) for k, v in batches[i % n_rot].items()}
        opt.zero_grad(set_to_none=True)
        ema_end = None
        if c5:
            with torch.no_grad():
                ema_end = ema.forward_with_pred_jitter({"point_clouds": d["ema_point_clouds"]})
        end_points = net.forward_with_pred_jitter({"point_clouds": d["point_clouds"]})
        for k, v in d.items():
            end_points[k] = v
        loss, end_points = ns.loss_labeled.get_labeled_loss(end_points, cfg, config_dict)
        if c5:
            unl, end_points = ns.loss_unlabeled.get_unlabeled_loss(end_points, ema_end, cfg, config_dict)
            loss = loss + unl * 2.0      # --unlabeled_loss_weight default (train.py:56)
        buckets.begin()
        loss.backward()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        buckets.finish()                  # waits for the reductions still in flight after backward returned
        e1.record()
        exposed.append((e0, e1))
        opt.step()
        if c5:
            global_step[0] += 1
            alpha = min(1 - 1 / (global_step[0] + 1), 0.999)
            for ep, p in zip(ema.parameters(), net.parameters()):   # train.py:285-289
                ep.data.mul_(alpha).add_(p.data, alpha=1 - alpha)
        return loss

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    n_warm = max(a.warmup, 3)
    steps = a.steps if a.steps != 400 else (30 if c5 else 60)
    launches_per_step = 0
    for i in range(n_warm):
        c0 = cabi.launch_count() if cabi else 0
        loss = step(i)
        launches_per_step = (cabi.launch_count() - c0) if cabi else 0
    torch.cuda.synchronize()
    loss0 = float(loss.detach())
    gc.collect()
    gc.disable()
    sampler.arm()
    exposed.clear()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host = time.perf_counter()
    e0.record()
    for i in range(steps):
        loss = step(i)
    e1.record()
    torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t_host) * 1e3
    ms = e0.elapsed_time(e1)
    if distributed:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)
    gc.enable()
    device_time = None
    if not distributed and not a.no_extras:   # (one more step: every rank would have to take part in its collective)
        try:  # device-time view of one step: sum of kernel durations (CUPTI) and the heaviest kernels
            from collections import defaultdict
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                step(0)
                torch.cuda.synchronize()
            agg = defaultdict(lambda: [0.0, 0])
            for e in prof.events():
                if e.device_type == torch.autograd.DeviceType.CUDA:
                    t = e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
                    agg[e.name[:70]][0] += t
                    agg[e.name[:70]][1] += 1
            tot = sum(v[0] for v in agg.values())
            device_time = {"kernel_ms_sum_per_step": round(tot / 1e3, 3), "kernels_per_step": int(sum(v[1] for v in agg.values())),
                           "top": [[k, round(v[0] / 1e3, 3), v[1]] for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]]}
        except Exception as e:  # noqa: BLE001
            device_time = {"error": str(e)[:200]}
    n_gpus = world if distributed else 1
    scenes = n_gpus * B * steps
    exposed_ms = float(np.median([x.elapsed_time(y) for x, y in exposed])) if exposed else 0.0
    n_param = sum(p.numel() for p in net.parameters() if p.requires_grad)
    line = {
        "metric": "scenes/sec %s training step" % ("SSL teacher+student" if c5 else "pretrain"),
        "value": round(scenes / (ms / 1e3), 3), "unit": "scenes/s", "n_gpus": n_gpus, "steps": steps, "warmup": n_warm,
        "ms_per_step": round(ms / steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": a.impl,
        "config": {"workload": ("configs[4]: ScanNet-shaped SSL train.py step (teacher under no_grad + student "
                                "forward_with_pred_jitter, labeled + unlabeled losses, IoU filter + LHS, backward, Adam, EMA), "
                                "%d labeled + %d unlabeled scenes per GPU, N=%d, %d proposals" % (n_lab, n_unl, N, K)) if c5 else
                               ("configs[3]: SUN RGB-D-shaped pretrain.py step (forward_with_pred_jitter, get_labeled_loss, "
                                "backward, Adam), %d scenes per GPU, N=%d, %d proposals" % (B, N, K)),
                   "model": "models/votenet_iou_branch.py:VoteNet + models/loss_helper_%s.py (reference files; --callers fast swaps "
                            "in the module / loss-helper mirrors of SURVEY 8f n1-n3, see impl_options)" % (
                       "labeled/unlabeled" if c5 else "labeled"),
                   "scenes_per_gpu_per_step": B, "batchnorm": "training mode (batch statistics), per replica",
                   "parallelism": "scene-sharded data parallel, one flat-bucket gradient all-reduce per step"},
        "impl_options": {"callers": a.callers if a.impl == "b200" else "reference",
                         "collective": "NCCL all-reduce of %d fp32 gradient elements (%.2f MB) in %d flat buckets launched from "
                                       "autograd hooks (overlapped with the backward tail)" % (n_param, n_param * 4 / 1e6, len(buckets.buckets))
                         if distributed else "none (1 GPU)"},
        "collective": {"elements": n_param, "bytes": n_param * 4, "buckets": len(buckets.buckets),
                       "exposed_wait_ms_per_step": round(exposed_ms, 4),
                       "note": "time the step waits in GradientBuckets.finish() after backward returned (median, CUDA events): the "
                               "part of the exchange NOT hidden under backward"},
        "host_wall_ms_per_step": round(wall_ms / steps, 4),
        "gpu_launches": int(launches_per_step * steps), "gpu_launches_per_step": int(launches_per_step),
        "loss_first": loss0, "loss_last": float(loss.detach()),
        "e2e": {"value": round(scenes / (ms / 1e3), 3), "unit": "scenes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "note": "training batches are device resident (the reference's DataLoader is out of scope); the loss helpers "
                        "of the reference copy per-box values to the host inside the step (loss_helper_unlabeled.py:441-492)"},
        "clocks": sampler.summary(),
    }
    if device_time is not None:
        line["device_time"] = device_time
    if rank == 0:
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0
