#!/usr/bin/env python
"""bench.py -- scenes/sec of the VoteNet forward + IoU hot path (BASELINE.json `metric`, configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this package (hand-written sm_100a kernels)
  python bench.py --impl reference ...                            # the unmodified reference operator stack (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU, scenes sharded, no data-path collective

A step = one forward of the VoteNet-with-IoU-branch dataflow (3dioumatch_b200/harness.py) over one batch of
B=8 synthetic ScanNet-shaped scenes (N=40000 points, C=4, 256 proposals) + IoU labels against 64 padded GT boxes.
Prints ONE JSON line on rank 0 (keys: see the task contract; extra keys `breakdown_ms`, `reference_cuda`, `c3`,
`ssl_filter`).
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_SCENES, N_POINTS, N_PROPOSAL, N_GT = 8, 40000, 256, 64
N_ROTATE = 32  # distinct input batches cycled through the timed region: 32 x 5.1 MB = 164 MB > 126 MB of L2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref", action="store_true", help="skip the in-run timing of the reference CUDA ops")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--lanes", type=int, default=5,
                    help="consecutive steps alternate between this many CUDA streams (software pipelining across steps)")
    ap.add_argument("--no-prefetch", action="store_true", help="do not run the FPS index chain on a side stream")
    ap.add_argument("--graphs", type=int, default=-1,
                    help="replay each lane's step from a CUDA graph (default: on for --impl b200; the reference launches on "
                         "the legacy default stream and cannot be captured)")
    ap.add_argument("--room", default="8,8,3", help="synthetic room size in metres (SURVEY 8d C2: 8x8x3); a smaller room = denser cloud")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-cloud variant line (config.dense_variant)")
    ap.add_argument("--batch", type=int, default=B_SCENES)
    ap.add_argument("--points", type=int, default=N_POINTS)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region: NVML in-process every 20 ms (nvidia-smi every 200 ms
    as the fallback; B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.stop_flag, self.armed = gpu_index, False, False
        self.sm, self.max_sm, self.reasons, self.power = [], None, set(), []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES: map the torch device to its NVML handle through the UUID
            import torch
            uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
            h = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                hi = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(hi)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u:
                    h = hi
            self.handle = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1e3)
        except Exception:
            pass
        try:
            r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
        except Exception:
            r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
        for name, bit in (("hw_slowdown", 0x8), ("sw_power_cap", 0x4), ("sw_thermal_slowdown", 0x20),
                          ("hw_thermal_slowdown", 0x40)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout
        for line in out.strip().splitlines():
            r = [c.strip() for c in line.split(",")]
            if len(r) > 7:
                self.sm.append(float(r[1]))
                self.max_sm = float(r[2])
                for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6),
                                  ("sw_power_cap", 7)):
                    if r[col].lower().startswith("active"):
                        self.reasons.add(name)

    def run(self):
        # started before the warm-up (NVML's first queries take a driver lock for tens of milliseconds, which showed up as
        # a stalled launch thread at the start of the timed region); samples are kept from arm() on
        while not self.stop_flag:
            try:
                if self.nvml:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            if not self.armed:
                self.sm, self.power, self.reasons = [], [], set()
            time.sleep(0.05 if self.nvml else 0.2)

    def arm(self):
        self.armed = True

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.max_sm) if self.max_sm else None,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "power_w_max": round(max(self.power), 1) if self.power else None,
                "source": "nvml" if self.nvml else "nvidia-smi"}


def load_stack(impl):
    if impl == "reference":
        ref_root = os.path.join(ROOT, "oracle", "_ref")
        if not os.path.exists(os.path.join(ref_root, "pointnet2", "_ext.so")):
            return None
        harness = importlib.import_module("3dioumatch_b200.harness")
        return harness, harness.stack_from_path(ref_root)
    harness = importlib.import_module("3dioumatch_b200.harness")
    return harness, harness.stack_b200()


# ---- per-operator timing (serialised, CUDA events) for the breakdown and the roofline of the dominant kernel ----
def breakdown(net, ops, pcs, gt, torch, iters=3):
    import pointnet2._ext as ext
    records = {}

    def wrap(name, fn, bytes_fn):
        def inner(*a, **k):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            torch.cuda.synchronize()
            key = name + bytes_fn(*a, **k)[0]
            rec = records.setdefault(key, {"ms": 0.0, "calls": 0, "bytes": bytes_fn(*a, **k)[1], "flops": bytes_fn(*a, **k)[2]})
            rec["ms"] += e0.elapsed_time(e1)
            rec["calls"] += 1
            return out
        return inner

    def b_fps(points, m):
        B, N = points.shape[0], points.shape[1]
        return "[N=%d,m=%d]" % (N, m), B * (N * 12 + m * 4), 10.0 * B * N * m

    def b_bq(new_xyz, xyz, r, ns):
        B, M, N = new_xyz.shape[0], new_xyz.shape[1], xyz.shape[1]
        return "[N=%d,M=%d]" % (N, M), B * ((N + M) * 12 + M * ns * 4), 8.0 * B * N * M

    def b_sa(xyz, features, new_xyz, radius, nsample, layers, **kw):
        B, N, M = xyz.shape[0], xyz.shape[1], new_xyz.shape[1]
        C = features.shape[1] if features is not None else 0
        wsum = sum(int(w.shape[0]) * int(w.reshape(w.shape[0], -1).shape[1]) for w, _, _ in layers)
        cout = int(layers[-1][0].shape[0])
        by = B * (N * 12 + N * C * 4 + M * 12 + M * 4 + M * cout * 4) + wsum * 4
        return "[N=%d,M=%d,ns=%d]" % (N, M, nsample), by, 2.0 * B * M * nsample * wsum + 8.0 * B * N * M

    def b_nn(u, k):
        B, n, m = u.shape[0], u.shape[1], k.shape[1]
        return "[n=%d,m=%d]" % (n, m), B * ((n + m) * 12 + n * 24), 8.0 * B * n * m

    def b_ti(p, idx, w):
        B, C, n = p.shape[0], p.shape[1], idx.shape[1]
        return "[C=%d,n=%d]" % (C, n), B * n * (24 + 16 * C), 6.0 * B * C * n

    def b_gather(p, idx):
        B, C, m = p.shape[0], p.shape[1], idx.shape[1]
        return "[C=%d,m=%d]" % (C, m), B * m * (4 + 8 * C), 0.0

    saved = {}
    for name, bf in (("furthest_point_sampling", b_fps), ("ball_query", b_bq), ("sa_forward", b_sa),
                     ("three_nn", b_nn), ("three_interpolate", b_ti), ("gather_points", b_gather)):
        saved[name] = getattr(ext, name)
        setattr(ext, name, wrap(name, saved[name], bf))
    try:
        with torch.no_grad():
            for i in range(iters):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                net(pcs[i % len(pcs)], gt)
                e1.record()
                torch.cuda.synchronize()
                records.setdefault("_step_serialised", {"ms": 0.0, "calls": 0, "bytes": 0, "flops": 0})
                records["_step_serialised"]["ms"] += e0.elapsed_time(e1)
                records["_step_serialised"]["calls"] += 1
    finally:
        for name, fn in saved.items():
            setattr(ext, name, fn)
    out = {}
    for k, r in records.items():
        per = r["ms"] / max(r["calls"], 1)
        out[k] = {"ms": round(per, 4), "calls_per_step": r["calls"] // iters, "alg_bytes": int(r["bytes"]),
                  "alg_flops": float(r["flops"])}
    return out


# HBM traffic of the reference's UNFUSED pipeline for one scene at the ScanNet shape (grouped tensors and every
# conv / BN / ReLU round trip; BASELINE.md section 2, derivation in SURVEY.md 8d), keyed like the breakdown entries
UNFUSED_GB_PER_SCENE = {"sa_forward[N=40000,M=2048,ns=64]": 0.81, "sa_forward[N=2048,M=1024,ns=32]": 0.44,
                        "sa_forward[N=1024,M=512,ns=16]": 0.12, "sa_forward[N=512,M=256,ns=16]": 0.06,
                        "sa_forward[N=1024,M=256,ns=16]": 0.05}


def sa_hbm_view(sa, hbm_peak, scenes):
    """HBM view of the fused SA launches of one step (`sa`: breakdown entries named sa_forward[...]): algorithmic bytes
    over their serialised time, and -- what BASELINE.json's 60 %-of-roofline target can meaningfully be read against -- the
    bytes the unfused reference pipeline moves for the same layers over that same time."""
    ms = sum(v["ms"] * max(v["calls_per_step"], 1) for v in sa.values())
    alg = sum(v["alg_bytes"] * max(v["calls_per_step"], 1) for v in sa.values())
    out = {"kernel": "fused SA layers: %d launches per step, %.3f ms serialised (ball query + prepasses included)" % (
               sum(max(v["calls_per_step"], 1) for v in sa.values()), ms),
           "bound": "hbm", "achieved": round(alg / (ms / 1e3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
           "frac": round(alg / (ms / 1e3) / 1e9 / hbm_peak, 4),
           "note": "algorithmic (fused) bytes / time: small by construction -- the fused layers are compute-bound "
                   "(370-1860 FLOP/B), DESIGN.md section 3.3"}
    if all(k in UNFUSED_GB_PER_SCENE for k in sa):
        unf = sum(UNFUSED_GB_PER_SCENE[k] * max(v["calls_per_step"], 1) for k, v in sa.items()) * scenes  # GB per step
        out["unfused_equivalent"] = {"gb_per_step": round(unf, 2), "achieved": round(unf / (ms / 1e3), 1), "unit": "GB/s",
                                     "frac": round(unf / (ms / 1e3) / hbm_peak, 3),
                                     "note": "bytes the reference's unfused pipeline moves for these layers (BASELINE.md "
                                             "section 2) / the fused kernels' time: > 1 means faster than that pipeline "
                                             "could run at the HBM roofline"}
    return out


def cpu_baseline(torch, points):
    """The oracle port (oracle/*.c + fp32 torch-CPU MLPs) on the host cores: ONE scene of the same workload."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    import oracle as orc
    import torch_ref
    orc.build()
    pc = cases.scene_cloud(123, 1, points)
    xyz, feat = np.ascontiguousarray(pc[:, :, :3]), np.ascontiguousarray(pc[:, :, 3:].transpose(0, 2, 1))
    t0 = time.time()
    cfg = [(2048, 0.2, 64, [1 + 3, 64, 64, 128]), (1024, 0.4, 32, [131, 128, 128, 256]),
           (512, 0.8, 16, [259, 128, 128, 256]), (256, 1.2, 16, [259, 128, 128, 256])]
    levels = []
    for i, (m, r, ns, spec) in enumerate(cfg):
        inds = orc.furthest_point_sampling(xyz, m)
        new_xyz = np.take_along_axis(xyz, inds[:, :, None].astype(np.int64), 1)
        feat, _ = torch_ref.sa_forward(xyz, feat, new_xyz, r, ns, cases.mlp_params(i, spec), normalize_xyz=True)
        xyz = new_xyz
        levels.append((xyz, feat))
    f = torch_ref.fp_forward(levels[2][0], levels[3][0], levels[2][1], levels[3][1], cases.mlp_params(10, [512, 256, 256]))
    f = torch_ref.fp_forward(levels[1][0], levels[2][0], levels[1][1], f, cases.mlp_params(11, [512, 256, 256]))
    seed_xyz = levels[1][0]
    inds = orc.furthest_point_sampling(seed_xyz, N_PROPOSAL)
    agg_xyz = np.take_along_axis(seed_xyz, inds[:, :, None].astype(np.int64), 1)
    torch_ref.sa_forward(seed_xyz, f, agg_xyz, 0.3, 16, cases.mlp_params(12, [259, 128, 128, 128]), normalize_xyz=True)
    grid = (np.random.default_rng(0).random((1, N_PROPOSAL * 64, 3)) * [8, 8, 3] - [4, 4, 0]).astype(np.float32)
    d2, idx = orc.three_nn(grid, seed_xyz)
    w = np.full((1, N_PROPOSAL * 64, 3), 1 / 3, np.float32)
    interp = orc.three_interpolate(f, idx, w)
    x = torch.from_numpy(np.concatenate([np.zeros((1, 3, N_PROPOSAL * 64), np.float32), interp], 1)).view(1, 259, N_PROPOSAL, 64)
    torch_ref.shared_mlp(x, cases.mlp_params(13, [259, 128, 128, 128]))
    orc.boxes_iou3d(cases.boxes(0, N_PROPOSAL), cases.boxes(1, N_GT))
    dt = time.time() - t0
    return {"value": round(1.0 / dt, 4), "unit": "scenes/s", "cores": int(orc.num_threads()), "kind": "port",
            "sample": "1 scene (N=%d) through oracle/*.c index ops (OpenMP) + fp32 torch-CPU shared MLPs, %.1f s" % (points, dt)}


def main():
    a = parse()
    import numpy as np
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and world > 1 and rank != 0:
        return 0  # the reference arm runs on rank 0 only
    distributed = world > 1 and a.impl != "reference"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    loaded = load_stack(a.impl)
    if loaded is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref (reference CUDA extensions) not built"}))
        return 0
    harness, ops = loaded
    net = harness.make_model(ops, seed=1, num_proposal=N_PROPOSAL, device=dev)
    net.backbone.prefetch = not a.no_prefetch
    B, N = a.batch, a.points
    lanes = [torch.cuda.Stream() for _ in range(max(a.lanes, 1))]

    # ---- inputs: N_ROTATE distinct batches per rank, pinned on the host and resident on the device -------------
    room = tuple(float(x) for x in a.room.split(","))
    base_pc, base_gt = harness.make_inputs(B, N, N_GT, seed=rank, room=room)
    rng = np.random.default_rng(1000 + rank)
    host_pcs = []
    for i in range(N_ROTATE):
        pc = base_pc.copy()
        pc[:, :, :3] += rng.normal(0, 0.01, (B, 1, 3)).astype(np.float32)  # distinct data per batch, same geometry
        pc = pc[:, rng.permutation(N)] if i else pc
        host_pcs.append(torch.from_numpy(np.ascontiguousarray(pc)).pin_memory())
    host_gt = torch.from_numpy(base_gt).pin_memory()
    dev_pcs = [t.to(dev) for t in host_pcs]
    dev_gt = host_gt.to(dev)
    out_keys = ("iou_labels", "iou_scores", "center", "size", "heading", "objectness")
    cabi = importlib.import_module("3dioumatch_b200._cabi") if a.impl == "b200" else None
    if cabi and len(lanes) > 1 and not os.environ.get("B200_FPS_POLICY"):
        # several steps share the GPU: FPS takes the launch shape with the least SM-time instead of the shortest chain
        cabi.set_fps_policy("throughput")

    # ---- CUDA graphs: one captured step per lane (static input/output buffers), replayed with fresh inputs ----------
    use_graphs = (a.graphs == 1) or (a.graphs == -1 and a.impl == "b200")
    graphs = []
    if use_graphs:
        try:
            with torch.no_grad():
                for lane in lanes:
                    static_pc = torch.empty_like(dev_pcs[0])
                    static_pc.copy_(dev_pcs[0])
                    static_gt = dev_gt.clone()
                    with torch.cuda.stream(lane):
                        for _ in range(3):
                            net(static_pc, static_gt)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=lane):
                        static_out = net(static_pc, static_gt)
                    graphs.append((g, static_pc, static_out, static_gt))
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001 -- a failed capture must not hide the eager measurement
            print("bench.py: CUDA-graph capture failed (%s); running eagerly" % (str(e).splitlines()[0][:200]),
                  file=sys.stderr)
            graphs, use_graphs = [], False
            torch.cuda.synchronize()

    def step_resident(i):
        lane = i % len(lanes)
        with torch.cuda.stream(lanes[lane]):
            if use_graphs:
                g, static_pc, static_out, _ = graphs[lane]
                static_pc.copy_(dev_pcs[i % N_ROTATE], non_blocking=True)
                g.replay()
                return static_out
            return net(dev_pcs[i % N_ROTATE], dev_gt)

    host_out = [dict() for _ in lanes]

    def step_e2e(i):
        lane = i % len(lanes)
        with torch.cuda.stream(lanes[lane]):
            if use_graphs:
                g, static_pc, res, static_gt = graphs[lane]
                static_pc.copy_(host_pcs[i % N_ROTATE], non_blocking=True)   # H2D from pinned memory
                static_gt.copy_(host_gt, non_blocking=True)
                g.replay()
            else:
                pc = host_pcs[i % N_ROTATE].to(dev, non_blocking=True)
                gt = host_gt.to(dev, non_blocking=True)
                res = net(pc, gt)
            for k in out_keys:
                if k not in host_out[lane]:
                    host_out[lane][k] = torch.empty(res[k].shape, dtype=res[k].dtype, pin_memory=True)
                host_out[lane][k].copy_(res[k], non_blocking=True)
        return res

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        cur = torch.cuda.current_stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        for s_ in lanes:
            s_.wait_stream(cur)          # every lane starts after the start event
        t_host = time.perf_counter()
        for i in range(steps):
            fn(i)
        timed.enqueue_ms = (time.perf_counter() - t_host) * 1e3 / steps
        for s_ in lanes:
            cur.wait_stream(s_)          # the stop event waits for all lanes (and their side streams)
        e1.record(cur)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    launches_per_step = 0
    if cabi:
        with torch.no_grad():
            c0 = cabi.launch_count()
            net(dev_pcs[0], dev_gt)          # one eager step: how many libb200pc kernels a step launches
            torch.cuda.synchronize()
            launches_per_step = cabi.launch_count() - c0
    import gc
    with torch.no_grad():
        sampler = ClockSampler(local)
        sampler.start()
        n_warm = max(a.warmup, 3)
        if not use_graphs:
            # lane set-up, the eager counterpart of the three runs per lane that precede a graph capture: with W < lanes
            # the last lanes otherwise see their first step -- and the allocator its first cudaMallocs on that stream --
            # inside the timed region (measured on the reference arm: 68 ms/step instead of 25)
            for lane in lanes:
                with torch.cuda.stream(lane):
                    for _ in range(2):
                        net(dev_pcs[0], dev_gt)
            torch.cuda.synchronize()
        for i in range(n_warm):
            step_resident(i)
            step_e2e(i)
        torch.cuda.synchronize()
        gc.collect()
        gc.disable()          # no collector pause inside the timed regions
        sampler.arm()
        launches0 = cabi.launch_count() if cabi else 0
        ms_res = timed(step_resident, a.steps)
        enqueue_ms = timed.enqueue_ms
        launches = (cabi.launch_count() - launches0) if cabi else 0
        if use_graphs:
            launches = launches_per_step * a.steps   # replayed from the captured graphs (not re-counted by the library)
        ms_e2e = timed(step_e2e, a.steps)
        enqueue_e2e_ms = timed.enqueue_ms
        sampler.stop_flag = True
        sampler.join(timeout=2)
        gc.enable()

    scenes = world * B * a.steps if distributed else B * a.steps
    n_gpus = world if distributed else 1
    value = scenes / (ms_res / 1e3)
    e2e_value = scenes / (ms_e2e / 1e3)
    h2d = int(host_pcs[0].numel() * 4 + host_gt.numel() * 4)
    d2h = int(sum(v.numel() * v.element_size() for v in host_out[0].values()))
    line = {
        "metric": "scenes/sec VoteNet fwd+IoU (B=8, N=40000)", "value": round(value, 3), "unit": "scenes/s",
        "n_gpus": n_gpus, "steps": a.steps, "warmup": n_warm, "ms_per_step": round(ms_res / a.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": a.impl,
        "config": {"workload": "configs[1]: ScanNet-shaped synthetic (B=%d,N=%d,C=4) VoteNet-IoU-branch forward dataflow, "
                               "%d proposals, IoU labels vs %d GT slots; random-init weights, eval-mode BN" % (B, N, N_PROPOSAL, N_GT),
                   "scenes_per_gpu_per_step": B, "lanes": len(lanes), "cuda_graphs": bool(use_graphs),
                   "host_enqueue_ms_per_step": round(enqueue_ms, 3),
                   "fps_policy": ("throughput" if (a.impl == "b200" and len(lanes) > 1 and not os.environ.get("B200_FPS_POLICY")) else os.environ.get("B200_FPS_POLICY", "latency")), "prefetch_fps_chain": bool(net.backbone.prefetch), "parallelism": "scene-sharded x%d, no data-path collective" % n_gpus,
                   "l2": "%d rotating input batches (%.0f MB) > 126 MB L2" % (N_ROTATE, N_ROTATE * B * N * 16 / 1e6),
                   "tf32": "torch defaults (cudnn conv TF32 allowed) for the torch-side 1x1 convs; all libb200pc kernels fp32",
                   "room_m": a.room,
                   "neighbour_lists": "fused SA kernel feeds only the 16-slot units of each ball-query list that hold distinct "
                                      "neighbours through the MLP (the reference pads short lists with copies of the first hit; "
                                      "max over duplicated rows is unchanged) -- gain depends on point density, see dense_variant",
                   "first_layer": "SA layers with >= 32 feature channels run layer 1 factorised: W1f*f once per source point "
                                  "(tensor-core row GEMM), xyz FMAs + ReLU in the gather (B200_SA_TC_FACTOR=0 disables)"},
        "e2e": {"value": round(e2e_value, 3), "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / a.steps, 4), "host_enqueue_ms_per_step": round(enqueue_e2e_ms, 4)},
        "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        if a.impl == "b200" and not a.no_breakdown:
            with torch.no_grad():
                bd = breakdown(net, ops, dev_pcs, dev_gt, torch)
            line["breakdown_ms"] = {k: v for k, v in bd.items()}
            ours = {k: v for k, v in bd.items() if not k.startswith("_")}
            top = max(ours, key=lambda k: ours[k]["ms"] * max(ours[k]["calls_per_step"], 1))
            t = ours[top]
            ach = t["alg_bytes"] / (t["ms"] / 1e3) / 1e9
            traffic = None
            try:  # DRAM bytes per launch of the same kernel/shape from the committed `ncu --set full` capture
                tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
                traffic = tr.get(top, {}).get("dram_bytes_per_launch")
            except Exception:
                pass
            line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": round(ach, 3), "peak": hbm_peak,
                                "unit": "GB/s", "frac": round(ach / hbm_peak, 5), "traffic": traffic,
                                "peak_source": peak_src, "launch_ms": t["ms"],
                                "achieved_fp32_tflops": round(t["alg_flops"] / (t["ms"] / 1e3) / 1e12, 3),
                                "note": "algorithmic bytes / launch time; this kernel is latency/FP32-bound, not HBM-bound "
                                        "(DESIGN.md section 4)"}
            # second view: the tensor-core MLP kernel over all fused SA launches of a step (ball query and the
            # point-major transpose are inside these times, so the figure is conservative)
            sa = {k: v for k, v in ours.items() if k.startswith("sa_forward")}
            if sa:
                def mlp_flops(k, v):  # strip the brute-force query term b_sa adds: 8*B*N*M
                    n_, m_ = int(k.split("N=")[1].split(",")[0]), int(k.split("M=")[1].split(",")[0])
                    return v["alg_flops"] - 8.0 * B * n_ * m_
                fl = sum(mlp_flops(k, v) * max(v["calls_per_step"], 1) for k, v in sa.items())
                ms_sa = sum(v["ms"] * max(v["calls_per_step"], 1) for v in sa.values())
                bf16 = float(peaks.get("bf16_tflops", 2250.0))
                ach_tf = fl / (ms_sa / 1e3) / 1e12
                try:
                    line["roofline_sa_hbm"] = sa_hbm_view(sa, hbm_peak, B)
                except Exception as e:  # noqa: BLE001 -- a reporting extra must not cost the bench line
                    line["roofline_sa_hbm"] = {"error": str(e)[:200]}
                ncu_frac = None
                try:  # time-weighted sm__pipe_tensor_cycles_active of the sa_tcp_kernel launches in the committed capture
                    import csv
                    rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r1c_ncu_full_kernels.csv"))))
                    hdr = rows[0]
                    ti = [i for i, h in enumerate(hdr) if h.startswith("gpu__time_duration")][0]
                    pi = [i for i, h in enumerate(hdr) if h.startswith("sm__pipe_tensor_cycles_active")][0]
                    sel = [(float(r[ti]), float(r[pi])) for r in rows[1:] if "sa_tcp_kernel" in r[1]]
                    ncu_frac = round(sum(t * q for t, q in sel) / sum(t for t, _ in sel) / 100.0, 4)
                except Exception:
                    pass
                line["roofline_tensor"] = {
                    "kernel": "sa_tcp_kernel: %d fused SA launches per step (%.3f ms serialised, ball query + prepasses included)" % (
                        sum(max(v["calls_per_step"], 1) for v in sa.values()), ms_sa),
                    "bound": "tensor", "achieved": round(ach_tf, 2), "peak": round(bf16 / 2, 1), "unit": "TFLOP/s",
                    "frac": ncu_frac,
                    "note": "achieved = REFERENCE-EQUIVALENT fp32 FLOPs (every nsample row of every centre) / time: the kernel "
                            "skips duplicated neighbour rows, so this is delivered work, not tensor-pipe work; each executed "
                            "product costs 3 kind::tf32 MMAs (split precision for the 1e-5 bar).  frac = tensor-pipe active "
                            "fraction measured by ncu (profiles/r1c_ncu_full_kernels.csv, time-weighted over the sa_tcp_kernel "
                            "launches); peak = dense TF32 = half the measured bf16 cuBLAS figure in MEASURED_PEAKS.json"}
        # ---- configs[2]: 256 x 256 rotated 3D IoU + NMS (device-resident boxes, CUDA events) ------------------------
        try:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import cases
            ba = torch.from_numpy(cases.boxes(0, 256)).to(dev)
            bb = torch.from_numpy(cases.boxes(1, 256, jitter_of=cases.boxes(0, 256))).to(dev)
            sc = torch.rand(256, device=dev)

            def c3_time(fn, iters=20):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / iters * 1e3
            line["c3"] = {"workload": "configs[2]: boxes_iou3d_gpu 256x256 + nms_gpu(256 boxes, thresh 0.25)",
                          "iou3d_us": round(c3_time(lambda: ops.iou.boxes_iou3d_gpu(ba, bb)), 2)}
            try:
                line["c3"]["nms_us"] = round(c3_time(lambda: ops.iou.nms_gpu(ba, sc, 0.25)), 2)
            except Exception as e:  # the reference's nms_gpu passes a LongTensor to an int32 reader
                line["c3"]["nms_us"] = None
                line["c3"]["nms_error"] = str(e).splitlines()[0][:120]
        except Exception as e:  # noqa: BLE001
            line["c3"] = {"error": str(e)[:200]}
        # ---- pseudo-label filter (corners -> extents -> lower-half suppression), 8 scenes x 64 boxes ----
        if a.impl == "b200":
            try:
                nms = importlib.import_module("utils.nms")
                sys.path.insert(0, os.path.join(ROOT, "oracle"))
                import oracle as orc  # oracle/oracle.py (the module, not the directory as a namespace package)
                rb = np.stack([cases.aabb_boxes(i, 64, 18) for i in range(8)])
                cen = torch.from_numpy(((rb[:, :, 0:3] + rb[:, :, 3:6]) / 2).astype(np.float32)).to(dev)
                siz = torch.from_numpy(rb[:, :, 3:6] - rb[:, :, 0:3]).to(dev)
                hd = torch.zeros((8, 64), dtype=torch.float64, device=dev)
                tail = torch.from_numpy(rb[:, :, 6:8]).to(dev)

                def filt():
                    _, ext = nms.box_extents_batch(cen, siz, hd, return_corners=False)
                    return nms.suppress_batch(torch.cat([ext.double(), tail], -1), 0.25, use_cls=True, lhs=True)
                dev_us = c3_time(filt)
                t0 = time.perf_counter()
                for i in range(8):
                    _, e = orc.box_extents(cen[i].cpu().numpy(), siz[i].cpu().numpy(), np.zeros(64))
                    orc.aabb_suppress(np.concatenate([e.astype(np.float64), rb[i, :, 6:8]], 1), 0.25, True, True)
                line["ssl_filter"] = {"workload": "8 scenes x 64 boxes: box extents + lhs_3d_faster_samecls (thresh 0.25)",
                                      "device_us": round(dev_us, 2),
                                      "host_numpy_port_us": round((time.perf_counter() - t0) * 1e6, 1)}
            except Exception as e:  # noqa: BLE001
                line["ssl_filter"] = {"error": str(e)[:200]}
        if a.impl == "reference":
            line["cpu_baseline"] = {"value": line["value"], "unit": "scenes/s", "kind": "reference",
                                    "cores": os.cpu_count(),
                                    "sample": "unmodified reference CUDA ops (oracle/_ref, built from /root/reference) on "
                                              "cuda:0 -- the reference has no CPU implementation of this path; host cores "
                                              "only drive the launches"}
            line["e2e"]["h2d_bytes_per_step"] = h2d
        elif n_gpus == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(torch, N)
        if a.impl == "b200" and n_gpus == 1 and not a.no_dense and a.room == "8,8,3":
            try:  # same workload on a 6x denser cloud: SA1 balls hold ~55 of 64 distinct neighbours (little to compact)
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", "60", "--warmup", "5", "--room", "3.2,3.2,1.2",
                                    "--no-ref", "--no-cpu-baseline", "--no-breakdown", "--no-dense", "--lanes", str(a.lanes),
                                    "--batch", str(B), "--points", str(N)], capture_output=True, text=True, timeout=600,
                                   env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
                d = json.loads(r.stdout.strip().splitlines()[-1])
                line["config"]["dense_variant"] = {"room_m": "3.2,3.2,1.2", "value": d.get("value"), "e2e": d.get("e2e", {}).get("value"),
                                                   "unit": "scenes/s", "ms_per_step": d.get("ms_per_step")}
            except Exception as e:  # noqa: BLE001
                line["config"]["dense_variant"] = {"error": str(e)[:200]}
        if a.impl == "b200" and n_gpus == 1 and not a.no_ref:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6",
                                    "--warmup", "3", "--batch", str(B), "--points", str(N), "--lanes", str(a.lanes)] +
                                   (["--no-prefetch"] if a.no_prefetch else []),
                                   capture_output=True, text=True, timeout=600,
                                   env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
                ref = json.loads(r.stdout.strip().splitlines()[-1])
                line["reference_cuda"] = {"value": ref.get("value"), "e2e": ref.get("e2e", {}).get("value"),
                                          "c3": ref.get("c3"),
                                          "ms_per_step": ref.get("ms_per_step"), "unit": "scenes/s",
                                          "what": "unmodified reference pointnet2/_ext + iou3d_nms CUDA ops on the same GPU, same dataflow"}
                if ref.get("value"):
                    line["speedup_vs_reference_cuda"] = round(line["value"] / ref["value"], 3)
            except Exception as e:  # the reference arm is informational here
                line["reference_cuda"] = {"unavailable": str(e)[:200]}
        print(json.dumps(line))
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
