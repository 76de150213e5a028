#!/usr/bin/env python
"""bench.py -- scenes/sec of the VoteNet forward + IoU hot path (BASELINE.json `metric`).

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this package (hand-written sm_100a kernels)
  python bench.py --impl reference ...                            # the unmodified reference operator stack (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...               # one rank per GPU, scenes sharded
  python bench.py --config c4|c5 ...                              # training steps (configs[3], configs[4]): NCCL gradient all-reduce

configs[1] (default, `--config c2`): a step = ONE call of the reference's own model code --
models/votenet_iou_branch.py: VoteNet.forward (:139-151) + models/loss_helper_iou.py: compute_iou_labels (:52-112),
installed verbatim in baseline/_ref -- over one batch of B=8 synthetic ScanNet-shaped scenes (N=40000 points, C=4,
256 proposals, 64 padded GT slots).  Both arms run the SAME caller files; what differs is the operator stack under
them: 3dioumatch_b200/dropin + libb200pc.so (this package) or the reference's pointnet2/_ext + iou3d_nms_cuda
(oracle/_ref).  `--callers fast` additionally swaps in this package's mirrors of the voting / proposal / GridConv
modules, of compute_iou_labels and of the pseudo-label filter (same classes, same state-dict keys).
Prints ONE JSON line on rank 0.
"""
import argparse
import gc
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
REF_OPS = os.path.join(ROOT, "oracle", "_ref")
OUT_KEYS = ("iou_labels", "iou_scores", "center", "size", "heading", "objectness_scores")
CHECK_KEYS = ("sa1_inds", "aggregated_vote_inds", "seed_xyz", "fp2_features", "vote_xyz", "center", "iou_scores", "iou_labels")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c4", "c5"])
    ap.add_argument("--callers", default="reference", choices=["reference", "fast"],
                    help="b200 arm: run the reference's caller files unchanged (default) or with the drop-in mirrors of the "
                         "voting / proposal / GridConv modules and compute_iou_labels (SURVEY 8f n1, n2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref", action="store_true", help="skip the in-run timing of the reference arm")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-per-op", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="only the headline measurement (sub-runs use this)")
    ap.add_argument("--lanes", type=int, default=-1,
                    help="consecutive steps alternate between this many CUDA streams (default 5; the reference arm: 1, on the "
                         "legacy default stream its IoU kernel launches on)")
    ap.add_argument("--graphs", type=int, default=-1, help="replay each lane's step from a CUDA graph (default: b200 arm)")
    ap.add_argument("--room", default="8,8,3", help="synthetic room size in metres (SURVEY 8d C2: 8x8x3)")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--points", type=int, default=0)
    ap.add_argument("--proposals", type=int, default=0)
    ap.add_argument("--per-op", action="store_true", help="only print the per-operator table (both stacks, one process)")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled DURING the timed region: NVML in-process every 50 ms (nvidia-smi every 200 ms
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
            import torch
            uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)  # honour CUDA_VISIBLE_DEVICES
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
        # started before the warm-up (NVML's first queries take a driver lock for tens of milliseconds); samples kept from arm()
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


def refapp():
    return importlib.import_module("3dioumatch_b200.refapp")


def load_stack(impl, callers="reference", with_losses=False):
    ra = refapp()
    if not ra.available():
        return None
    if impl == "reference":
        if not os.path.exists(os.path.join(REF_OPS, "pointnet2", "_ext.so")):
            return None
        return ra.load([os.path.join(REF_OPS, "pointnet2"), REF_OPS], name="reference", with_losses=with_losses)
    return ra.load(ra.dropin_paths(fast_callers=(callers == "fast")), name="b200", with_losses=with_losses)


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        p = {}
    return (float(p.get("hbm_gbs", 6650.0)), float(p.get("bf16_tflops", 2250.0)),
            "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in p else "fallback (B200_PROFILING.md): 6.65 TB/s, 2250 TF/s bf16")


# ---- per-operator timing (serialised, CUDA events) for the breakdown and the roofline of the dominant kernel ----
def breakdown(step, ns, torch, iters=3):
    """Every call into the drop-in native module timed with CUDA events (device synchronised around each call), keyed by
    op + shape, plus the algorithmic bytes / FLOPs of SURVEY 8(d) for that shape."""
    ext = ns.ext
    records = {}

    def wrap(name, fn, bytes_fn):
        def inner(*a, **k):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            torch.cuda.synchronize()
            tag, by, fl = bytes_fn(*a, **k)
            rec = records.setdefault(name + tag, {"ms": 0.0, "calls": 0, "bytes": by, "flops": fl})
            rec["ms"] += e0.elapsed_time(e1)
            rec["calls"] += 1
            return out
        return inner

    def wsum_of(layers):
        return sum(int(w.shape[0]) * int(w.reshape(w.shape[0], -1).shape[1]) for w, _, _ in layers)

    def b_fps(points, m):
        Bq, Nq = points.shape[0], points.shape[1]
        return "[N=%d,m=%d]" % (Nq, m), Bq * (Nq * 12 + m * 4), 10.0 * Bq * Nq * m

    def b_bq(new_xyz, xyz, r, nsm):
        Bq, M, Nq = new_xyz.shape[0], new_xyz.shape[1], xyz.shape[1]
        return "[N=%d,M=%d]" % (Nq, M), Bq * ((Nq + M) * 12 + M * nsm * 4), 8.0 * Bq * Nq * M

    def b_sa(xyz, features, new_xyz, radius, nsample, layers, **kw):
        Bq, Nq, M = xyz.shape[0], xyz.shape[1], new_xyz.shape[1]
        C = features.shape[1] if features is not None else 0
        cout = int(layers[-1][0].shape[0])
        by = Bq * (Nq * 12 + Nq * C * 4 + M * 12 + M * 4 + M * cout * 4) + wsum_of(layers) * 4
        return "[N=%d,M=%d,ns=%d]" % (Nq, M, nsample), by, 2.0 * Bq * M * nsample * wsum_of(layers)

    def b_nn(u, k):
        Bq, n, m = u.shape[0], u.shape[1], k.shape[1]
        return "[n=%d,m=%d]" % (n, m), Bq * ((n + m) * 12 + n * 24), 8.0 * Bq * n * m

    def b_ti(p, idx, w):
        Bq, C, n = p.shape[0], p.shape[1], idx.shape[1]
        return "[C=%d,n=%d]" % (C, n), Bq * n * (24 + 16 * C), 6.0 * Bq * C * n

    def b_gather(p, idx):
        Bq, C, m = p.shape[0], p.shape[1], idx.shape[1]
        return "[C=%d,m=%d]" % (C, m), Bq * m * (4 + 8 * C), 0.0

    def b_fp(known_pm, skip_pm, idx3, w3, layers, **kw):
        Bq, m, C2 = known_pm.shape
        n = idx3.shape[1]
        C1 = skip_pm.shape[2] if skip_pm is not None else 0
        cout = int(layers[-1][0].shape[0])
        return "[n=%d,K=%d]" % (n, C1 + C2), Bq * (m * C2 * 4 + n * (C1 * 4 + 24 + cout * 4)) + wsum_of(layers) * 4, \
            2.0 * Bq * n * wsum_of(layers)

    def b_rows(x_pm, layers, **kw):
        S, R, ld = x_pm.shape
        cout = int(layers[-1][0].shape[0])
        return "[rows=%d,K=%d]" % (S * R, ld), S * R * (ld + cout) * 4 + wsum_of(layers) * 4, 2.0 * S * R * wsum_of(layers)

    def b_interp(known_feats, idx3, w3, rel, nsample, layers, **kw):
        Bq, C, m = known_feats.shape
        rows = idx3.shape[1]
        cout = int(layers[-1][0].shape[0])
        return "[rows=%d,K=%d]" % (rows, C + 3), Bq * (m * C * 4 + rows * 48 + rows // nsample * cout * 4) + wsum_of(layers) * 4, \
            2.0 * Bq * rows * wsum_of(layers)

    table = (("furthest_point_sampling", b_fps), ("ball_query", b_bq), ("sa_forward", b_sa), ("three_nn", b_nn),
             ("three_interpolate", b_ti), ("gather_points", b_gather), ("fp_rows_forward", b_fp),
             ("row_mlp_forward", b_rows), ("interp_mlp_forward", b_interp))
    saved = {}
    for name, bf in table:
        if hasattr(ext, name):
            saved[name] = getattr(ext, name)
            setattr(ext, name, wrap(name, saved[name], bf))
    cabi = importlib.import_module("3dioumatch_b200._cabi")
    try:
        with torch.no_grad():
            step()  # warm
            cabi.sa_tensor_work(reset=True)
            for i in range(iters):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                step()
                e1.record()
                torch.cuda.synchronize()
                rec = records.setdefault("_step_serialised", {"ms": 0.0, "calls": 0, "bytes": 0, "flops": 0})
                rec["ms"] += e0.elapsed_time(e1)
                rec["calls"] += 1
            executed_tf32_flops = cabi.sa_tensor_work(reset=True) / float(iters)
    finally:
        for name, fn in saved.items():
            setattr(ext, name, fn)
    out = {}
    for k, r in records.items():
        calls = max(r["calls"] // iters, 1)
        out[k] = {"ms": round(r["ms"] / max(r["calls"], 1), 4), "calls_per_step": r["calls"] // iters, "alg_bytes": int(r["bytes"]),
                  "alg_flops": float(r["flops"]), "ms_per_step": round(r["ms"] / iters, 4)}
        del calls
    return out, executed_tf32_flops


# HBM traffic of the reference's UNFUSED pipeline for one scene at the ScanNet shape (grouped tensors and every
# conv / BN / ReLU round trip; BASELINE.md section 2, derivation in SURVEY.md 8d), keyed like the breakdown entries
UNFUSED_GB_PER_SCENE = {"sa_forward[N=40000,M=2048,ns=64]": 0.81, "sa_forward[N=2048,M=1024,ns=32]": 0.44,
                        "sa_forward[N=1024,M=512,ns=16]": 0.12, "sa_forward[N=512,M=256,ns=16]": 0.06,
                        "sa_forward[N=1024,M=256,ns=16]": 0.05}

# Latency floor of ONE furthest-point-sampling iteration on this machine: the iteration is a
# dependent chain  distance update -> warp arg-max (2 redux) -> CTA arg-max (shared memory + barrier) -> [cluster
# exchange: DSMEM st.async + mbarrier wait -> arg-max of the records]  and none of its links can overlap the next
# iteration (the next distance update needs the winner's coordinates).  Measured link latencies on B200 at 1965 MHz
# (scripts/fps_profile.py): warp 45 + CTA 200 + [DSMEM push 125 + mbarrier 230 + cluster arg-max 200] cycles.
FPS_FLOOR_CYCLES_SINGLE_CTA = 245.0
FPS_FLOOR_CYCLES_CLUSTER = 800.0


def fps_latency_view(entry_key, entry, sm_mhz):
    n = int(entry_key.split("N=")[1].split(",")[0])
    m = int(entry_key.split("m=")[1].split("]")[0])
    us_iter = entry["ms"] * 1e3 / max(m - 1, 1)
    clustered = n > 2048
    floor_cycles = FPS_FLOOR_CYCLES_CLUSTER if clustered else FPS_FLOOR_CYCLES_SINGLE_CTA
    floor_us = floor_cycles / (sm_mhz or 1965.0)
    return {"bound": "latency", "iterations": m - 1, "us_per_iteration": round(us_iter, 4),
            "floor_us_per_iteration": round(floor_us, 4), "frac_of_floor": round(floor_us / us_iter, 4),
            "floor": "dependent chain per iteration without the distance update: %d cycles (%s), DESIGN.md 3.1" % (
                int(floor_cycles), "warp + CTA arg-max + DSMEM exchange + cluster arg-max" if clustered else "warp + CTA arg-max"),
            "points_per_iteration": n}


def cpu_baseline(torch, points):
    """The oracle port (oracle/*.c + fp32 torch-CPU MLPs) on the host cores: ONE scene of the same workload."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    synth = importlib.import_module("3dioumatch_b200.synth")
    import oracle as orc
    import torch_ref
    orc.build()
    pc = synth.scene_cloud(123, 1, points)
    xyz, feat = np.ascontiguousarray(pc[:, :, :3]), np.ascontiguousarray(pc[:, :, 3:].transpose(0, 2, 1))
    t0 = time.time()
    cfg = [(2048, 0.2, 64, [1 + 3, 64, 64, 128]), (1024, 0.4, 32, [131, 128, 128, 256]),
           (512, 0.8, 16, [259, 128, 128, 256]), (256, 1.2, 16, [259, 128, 128, 256])]
    levels = []
    for i, (m, r, nsm, spec) in enumerate(cfg):
        inds = orc.furthest_point_sampling(xyz, m)
        new_xyz = np.take_along_axis(xyz, inds[:, :, None].astype(np.int64), 1)
        feat, _ = torch_ref.sa_forward(xyz, feat, new_xyz, r, nsm, synth.mlp_params(i, spec), normalize_xyz=True, exact=False)
        xyz = new_xyz
        levels.append((xyz, feat))
    f = torch_ref.fp_forward(levels[2][0], levels[3][0], levels[2][1], levels[3][1], synth.mlp_params(10, [512, 256, 256]), exact=False)
    f = torch_ref.fp_forward(levels[1][0], levels[2][0], levels[1][1], f, synth.mlp_params(11, [512, 256, 256]), exact=False)
    seed_xyz = levels[1][0]
    inds = orc.furthest_point_sampling(seed_xyz, N_PROPOSAL)
    agg_xyz = np.take_along_axis(seed_xyz, inds[:, :, None].astype(np.int64), 1)
    torch_ref.sa_forward(seed_xyz, f, agg_xyz, 0.3, 16, synth.mlp_params(12, [259, 128, 128, 128]), normalize_xyz=True, exact=False)
    grid = (np.random.default_rng(0).random((1, N_PROPOSAL * 64, 3)) * [8, 8, 3] - [4, 4, 0]).astype(np.float32)
    d2, idx = orc.three_nn(grid, seed_xyz)
    w = np.full((1, N_PROPOSAL * 64, 3), 1 / 3, np.float32)
    interp = orc.three_interpolate(f, idx, w)
    x = torch.from_numpy(np.concatenate([np.zeros((1, 3, N_PROPOSAL * 64), np.float32), interp], 1)).view(1, 259, N_PROPOSAL, 64)
    torch_ref.shared_mlp(x, synth.mlp_params(13, [259, 128, 128, 128]), exact=False)
    orc.boxes_iou3d(synth.boxes(0, N_PROPOSAL), synth.boxes(1, N_GT))
    dt = time.time() - t0
    return {"value": round(1.0 / dt, 4), "unit": "scenes/s", "cores": int(orc.num_threads()), "kind": "port",
            "sample": "1 scene (N=%d) through oracle/*.c index ops (OpenMP) + fp32 torch-CPU shared MLPs, %.1f s" % (points, dt)}


# ---- per-operator table: this package vs the reference CUDA ops on the same inputs --------
def per_op_table():
    """Both operator stacks in ONE process (a comparison leg like cpu_baseline, run as a subprocess of the b200 arm so that
    the measured process never loads the reference extensions).  CUDA events, 3 warm-up + 10 timed calls, default stream."""
    import numpy as np
    import torch
    ours = load_stack("b200")
    ref = load_stack("reference")
    synth = importlib.import_module("3dioumatch_b200.synth")
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    pc = torch.from_numpy(synth.scene_cloud(0, B_SCENES, N_POINTS)).to(dev)
    xyz = pc[..., :3].contiguous()
    feats = pc[..., 3:].transpose(1, 2).contiguous()

    def t(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    rows = {}

    def add(name, f_ours, f_ref, same=None):
        a, b = t(f_ours), t(f_ref)
        rows[name] = {"b200_ms": round(a, 4), "reference_ms": round(b, 4), "speedup": round(b / a, 2)}
        if same is not None:
            rows[name]["identical"] = bool(same)

    with torch.no_grad():
        inds = ours.ext.furthest_point_sampling(xyz, 2048)
        add("furthest_point_sampling (8,40000)->2048", lambda: ours.ext.furthest_point_sampling(xyz, 2048),
            lambda: ref.ext.furthest_point_sampling(xyz, 2048), torch.equal(inds, ref.ext.furthest_point_sampling(xyz, 2048)))
        new_xyz = ours.utils.gather_operation(xyz.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
        idx = ours.ext.ball_query(new_xyz, xyz, 0.2, 64)
        add("ball_query (8,2048)x(8,40000) r=0.2 ns=64", lambda: ours.ext.ball_query(new_xyz, xyz, 0.2, 64),
            lambda: ref.ext.ball_query(new_xyz, xyz, 0.2, 64), torch.equal(idx, ref.ext.ball_query(new_xyz, xyz, 0.2, 64)))
        f128 = torch.randn(B_SCENES, 128, 2048, device=dev)
        idx2 = ours.ext.ball_query(new_xyz[:, :1024].contiguous(), new_xyz, 0.4, 32)
        add("group_points (8,128,2048) idx (8,1024,32)", lambda: ours.ext.group_points(f128, idx2),
            lambda: ref.ext.group_points(f128, idx2), torch.equal(ours.ext.group_points(f128, idx2), ref.ext.group_points(f128, idx2)))
        # SharedMLP + max = the whole SA2 layer: fused module vs the reference module (its ops + cuDNN fp32)
        def sa_pair(kw, x, f):
            torch.manual_seed(3)
            mo = ours.modules.PointnetSAModuleVotes(**{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()}).to(dev).eval()
            torch.manual_seed(3)
            mr = ref.modules.PointnetSAModuleVotes(**{k: (list(v) if isinstance(v, list) else v) for k, v in kw.items()}).to(dev).eval()
            ours.pt.freeze_inference(mo)
            return (lambda: mo(x, f)), (lambda: mr(x, f))
        fo, fr = sa_pair(dict(npoint=2048, radius=0.2, nsample=64, mlp=[1, 64, 64, 128], use_xyz=True, normalize_xyz=True), xyz, feats)
        add("SA1 module: FPS + ball_query + group + SharedMLP[4,64,64,128] + max", fo, fr)
        fo, fr = sa_pair(dict(npoint=1024, radius=0.4, nsample=32, mlp=[128, 128, 128, 256], use_xyz=True, normalize_xyz=True), new_xyz, f128)
        add("SA2 module: FPS + ball_query + group + SharedMLP[131,128,128,256] + max", fo, fr)
        grid = (torch.rand(B_SCENES, N_PROPOSAL * 64, 3, device=dev) * 6 - 3).contiguous()
        seeds = new_xyz[:, :1024].contiguous()
        d_o, i_o = ours.ext.three_nn(grid, seeds)
        d_r, i_r = ref.ext.three_nn(grid, seeds)
        add("three_nn (8,16384)x(8,1024)", lambda: ours.ext.three_nn(grid, seeds), lambda: ref.ext.three_nn(grid, seeds),
            torch.equal(i_o, i_r) and torch.equal(d_o, d_r))
        f256 = torch.randn(B_SCENES, 256, 1024, device=dev)
        w3 = torch.rand(B_SCENES, N_PROPOSAL * 64, 3, device=dev)
        add("three_interpolate (8,256,1024)->(8,256,16384)", lambda: ours.ext.three_interpolate(f256, i_o, w3),
            lambda: ref.ext.three_interpolate(f256, i_o, w3),
            torch.equal(ours.ext.three_interpolate(f256, i_o, w3), ref.ext.three_interpolate(f256, i_o, w3)))
        ba = torch.from_numpy(synth.boxes(0, 256)).to(dev)
        bb = torch.from_numpy(synth.boxes(1, 256, jitter_of=synth.boxes(0, 256))).to(dev)
        err = float((ours.iou.boxes_iou3d_gpu(ba, bb) - ref.iou.boxes_iou3d_gpu(ba, bb)).abs().max())
        add("boxes_iou3d_gpu 256x256", lambda: ours.iou.boxes_iou3d_gpu(ba, bb), lambda: ref.iou.boxes_iou3d_gpu(ba, bb), err <= 1e-5)
        big_a = torch.from_numpy(synth.boxes(2, 2048)).to(dev)
        big_b = torch.from_numpy(synth.boxes(3, 512)).to(dev)
        add("boxes_iou3d_gpu 2048x512 (B*K x B*64 of a training step)", lambda: ours.iou.boxes_iou3d_gpu(big_a, big_b),
            lambda: ref.iou.boxes_iou3d_gpu(big_a, big_b))
        # nms_gpu: the reference's python wrapper allocates a LongTensor keep and its C++ reads int32 (iou3d_nms_utils.py:97
        # vs iou3d_nms.cpp:98); the C++ entry is called directly with an int32 keep, as SURVEY 2a prescribes
        sc = torch.rand(256, device=dev)
        order = sc.sort(0, descending=True)[1]
        sorted_boxes = ba[order].contiguous()

        def ref_nms():
            keep = torch.zeros(256, dtype=torch.int32)
            n = ref.iou.iou3d_nms_cuda.nms_gpu(sorted_boxes, keep, 0.25)
            return order[keep[:n].long().to(dev)]

        def our_nms():
            return ours.iou.nms_gpu(ba, sc, 0.25)[0]
        same = torch.equal(our_nms(), ref_nms())
        add("nms_gpu 256 boxes thresh 0.25 (sort + mask + sweep + index)", our_nms, ref_nms, same)
    return rows


def main():
    a = parse()
    import numpy as np
    import torch

    if a.per_op:
        print(json.dumps({"per_op": per_op_table()}))
        return 0
    if a.config in ("c4", "c5"):
        train = importlib.import_module("3dioumatch_b200.trainbench")
        return train.main(a, ROOT, ClockSampler, load_stack)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference" and world > 1 and rank != 0:
        return 0  # contract: the reference arm runs on rank 0 only
    distributed = world > 1 and a.impl != "reference"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    ns = load_stack(a.impl, a.callers)
    if ns is None:
        print(json.dumps({"impl": a.impl, "unavailable": "baseline/_ref (reference callers) or oracle/_ref (reference CUDA "
                                                         "extensions) not built: python oracle/build_ref.py"}))
        return 0
    ra = refapp()
    B = a.batch or B_SCENES
    N = a.points or N_POINTS
    K = a.proposals or N_PROPOSAL
    net, cfg = ra.build_votenet(ns, "scannet", K, seed=1, device=dev)
    cabi = importlib.import_module("3dioumatch_b200._cabi") if a.impl == "b200" else None
    n_lanes = a.lanes if a.lanes > 0 else (5 if a.impl == "b200" else 1)
    use_graphs = (a.graphs == 1) or (a.graphs == -1 and a.impl == "b200")
    if a.impl == "reference":
        n_lanes, use_graphs = 1, False   # its IoU kernel launches on the legacy default stream (iou3d_nms_kernel.cu:396)

    # ---- inputs: N_ROTATE distinct batches per rank, pinned on the host and resident on the device -------------
    room = tuple(float(x) for x in a.room.split(","))
    base_pc, labels = ra.make_inputs(B, N, seed=rank, room=room, cfg=cfg, max_gt=N_GT)
    rng = np.random.default_rng(1000 + rank)
    host_pcs = []
    for i in range(N_ROTATE):
        pc = base_pc.copy()
        pc[:, :, :3] += rng.normal(0, 0.01, (B, 1, 3)).astype(np.float32)  # distinct data per batch, same geometry
        pc = pc[:, rng.permutation(N)] if i else pc
        host_pcs.append(torch.from_numpy(np.ascontiguousarray(pc)).pin_memory())
    label_keys = ra.LABEL_KEYS_IOU
    host_labels = [torch.from_numpy(labels[k]).pin_memory() for k in label_keys]
    dev_pcs = [t.to(dev) for t in host_pcs]
    dev_labels = [t.to(dev) for t in host_labels]

    def full_step(pc, *lab):
        return ra.forward_with_iou_labels(ns, net, cfg, pc, dict(zip(label_keys, lab)))

    def step_fn(pc, *lab):
        ep = full_step(pc, *lab)
        return {k: ep[k] for k in OUT_KEYS}

    # ---- step 0 with TF32 off: outputs saved by the reference arm, asserted by the b200 arm -----------------------
    check_path = os.path.join(ROOT, "gpurun_out", "bench_ref_step0_c2_B%d_N%d_K%d.npz" % (B, N, K))
    check = None
    tf32_default = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    launches_per_step = 0
    with torch.no_grad():
        c0 = cabi.launch_count() if cabi else 0
        ep0 = full_step(dev_pcs[0], *dev_labels)
        torch.cuda.synchronize()
        launches_per_step = (cabi.launch_count() - c0) if cabi else 0
        step0 = {k: ep0[k].detach().cpu().numpy() for k in CHECK_KEYS}
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_default
    if rank == 0:
        if a.impl == "reference":
            try:
                os.makedirs(os.path.dirname(check_path), exist_ok=True)
                np.savez(check_path, **step0)
            except Exception:
                pass
    del ep0

    if cabi:
        ns.pt.freeze_inference(net)      # final weights: BN affine folded and tensor-core weight images packed once per module
        if n_lanes > 1 and not os.environ.get("B200_FPS_POLICY"):
            cabi.set_fps_policy("throughput")  # several steps share the GPU: FPS takes the shape with the least SM-time
    runner_mod = importlib.import_module("3dioumatch_b200.runner")
    if a.impl == "reference":
        # stock code path: one (legacy default) stream, eager
        class _Eager:
            lanes, use_graphs, capture_error = [torch.cuda.default_stream()], False, None

            def run(self, i, inputs):
                return step_fn(*[t if t.is_cuda else t.to(dev, non_blocking=True) for t in inputs])

            def fork(self, cur):
                pass

            def join(self, cur):
                pass
        runner = _Eager()
        with torch.no_grad():
            for _ in range(2):
                step_fn(dev_pcs[0], *dev_labels)
        torch.cuda.synchronize()
    else:
        runner = runner_mod.LaneRunner(step_fn, [dev_pcs[0]] + dev_labels, lanes=n_lanes, graphs=use_graphs)
        if runner.capture_error:
            print("bench.py: CUDA-graph capture failed (%s); running eagerly" % runner.capture_error, file=sys.stderr)
    use_graphs = runner.use_graphs

    host_out = [dict() for _ in runner.lanes]

    def step_resident(i):
        return runner.run(i, [dev_pcs[i % N_ROTATE]] + dev_labels)

    def step_e2e(i):
        lane = i % len(runner.lanes)
        res = runner.run(i, [host_pcs[i % N_ROTATE]] + host_labels)   # H2D from pinned memory inside the step
        with torch.cuda.stream(runner.lanes[lane]):
            for k in OUT_KEYS:
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
        runner.fork(cur)             # every lane starts after the start event
        t_host = time.perf_counter()
        for i in range(steps):
            fn(i)
        timed.enqueue_ms = (time.perf_counter() - t_host) * 1e3 / steps
        runner.join(cur)             # the stop event waits for all lanes
        e1.record(cur)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if distributed:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    with torch.no_grad():
        sampler = ClockSampler(local)
        sampler.start()
        n_warm = max(a.warmup, 3)
        for i in range(n_warm):
            step_resident(i)
            step_e2e(i)
        torch.cuda.synchronize()
        gc.collect()
        gc.disable()          # no collector pause inside the timed regions
        sampler.arm()
        ms_res = timed(step_resident, a.steps)
        enqueue_ms = timed.enqueue_ms
        ms_e2e = timed(step_e2e, a.steps)
        enqueue_e2e_ms = timed.enqueue_ms
        sampler.stop_flag = True
        sampler.join(timeout=2)
        gc.enable()

    scenes = (world if distributed else 1) * B * a.steps
    n_gpus = world if distributed else 1
    value = scenes / (ms_res / 1e3)
    e2e_value = scenes / (ms_e2e / 1e3)
    h2d = int(host_pcs[0].numel() * 4 + sum(t.numel() * t.element_size() for t in host_labels))
    d2h = int(sum(v.numel() * v.element_size() for v in host_out[0].values()))
    clocks = sampler.summary()
    line = {
        "metric": "scenes/sec VoteNet fwd+IoU (B=8, N=40000)", "value": round(value, 3), "unit": "scenes/s",
        "n_gpus": n_gpus, "steps": a.steps, "warmup": n_warm, "ms_per_step": round(ms_res / a.steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "impl": a.impl,
        # identical in both arms (the driver compares them); everything arm-specific is under impl_options
        "config": {"workload": "configs[1]: ScanNet-shaped synthetic (B=%d,N=%d,C=4) through the reference's own "
                               "models/votenet_iou_branch.py VoteNet.forward + models/loss_helper_iou.py compute_iou_labels "
                               "(%d proposals, %d GT slots, all (B*K) x (B*%d) IoU pairs); random-init weights, eval-mode BN" % (
                                   B, N, K, N_GT, N_GT),
                   "model": "models/votenet_iou_branch.py:VoteNet (reference file, unmodified, baseline/_ref)",
                   "scenes_per_gpu_per_step": B, "room_m": a.room,
                   "l2": "%d rotating input batches (%.0f MB) > 126 MB L2" % (N_ROTATE, N_ROTATE * B * N * 16 / 1e6),
                   "tf32": "torch defaults in the timed region (cudnn conv TF32 allowed) for whatever torch convolutions a stack "
                           "runs; the step-0 output check runs with TF32 off",
                   "parallelism": "scene-sharded, no data-path collective"},
        "impl_options": {"callers": a.callers if a.impl == "b200" else "reference", "lanes": len(runner.lanes),
                         "cuda_graphs": bool(use_graphs), "frozen_plans": bool(cabi),
                         "fps_policy": ("throughput" if (cabi and n_lanes > 1 and not os.environ.get("B200_FPS_POLICY"))
                                        else os.environ.get("B200_FPS_POLICY", "latency")) if cabi else None,
                         "block_diagonal_iou": bool(a.impl == "b200" and a.callers == "fast"),
                         "fps_prefix_speculation": bool(cabi) and os.environ.get("B200_FPS_PREFIX", "1") != "0",
                         "ball_query_fused_into_gather": bool(cabi) and os.environ.get("B200_SA_TC_QUERY", "1") != "0",
                         "stream": "legacy default stream, eager (the reference's IoU kernel launches there)" if a.impl == "reference"
                                   else "%d non-blocking streams" % len(runner.lanes)},
        "host_enqueue_ms_per_step": round(enqueue_ms, 4),
        "e2e": {"value": round(e2e_value, 3), "unit": "scenes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": round(ms_e2e / a.steps, 4), "host_enqueue_ms_per_step": round(enqueue_e2e_ms, 4)},
        "gpu_launches": int(launches_per_step * a.steps),
        "gpu_launches_per_step": int(launches_per_step),
        "clocks": clocks,
    }

    if rank == 0 and not a.no_extras:
        hbm_peak, bf16_peak, peak_src = peaks()
        # ---- single-lane, single-batch latency (one step in flight, synchronised per step) ---------------------------
        try:
            with torch.no_grad():
                prev_policy = cabi.set_fps_policy("latency") if cabi else None
                # one lane; the FPS launch-shape policy is baked into a captured graph, so capture again under "latency"
                lat_runner = runner_mod.LaneRunner(step_fn, [dev_pcs[0]] + dev_labels, lanes=1, graphs=use_graphs) if cabi else runner
                lat = []
                for i in range(13):
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s = lat_runner.lanes[0]
                    e0.record(s)
                    lat_runner.run(0, [dev_pcs[i % N_ROTATE]] + dev_labels)
                    e1.record(s)
                    torch.cuda.synchronize()
                    lat.append(e0.elapsed_time(e1))
                if cabi:
                    cabi.set_fps_policy(prev_policy)
                lat = sorted(lat[3:])
                line["latency"] = {"ms_per_step_single_batch": round(lat[len(lat) // 2], 4),
                                   "scenes_per_s_single_batch": round(B / (lat[len(lat) // 2] / 1e3), 1),
                                   "how": "one step in flight, one lane, device-resident inputs, median of 10, CUDA events"
                                          + (", graph replay, latency FPS policy" if (cabi and use_graphs) else ", eager")}
        except Exception as e:  # noqa: BLE001
            line["latency"] = {"error": str(e)[:200]}
        # ---- device-time-only view: sum of kernel durations of one eager step (CUPTI via torch.profiler) ----------------
        try:
            from torch.profiler import ProfilerActivity, profile
            with torch.no_grad():
                step_fn(dev_pcs[1], *dev_labels)
                torch.cuda.synchronize()
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for i in range(3):
                        step_fn(dev_pcs[(2 + i) % N_ROTATE], *dev_labels)
                    torch.cuda.synchronize()
            evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
            kern_ms = sum(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total for e in evs) / 1e3 / 3
            line["device_time"] = {"kernel_ms_sum_per_step": round(kern_ms, 4), "kernels_per_step": len(evs) // 3,
                                   "how": "sum of GPU kernel + memcpy durations of one eager step on one stream (CUPTI), i.e. "
                                          "device work with launch gaps and overlap removed"}
        except Exception as e:  # noqa: BLE001
            line["device_time"] = {"error": str(e)[:200]}

        if a.impl == "b200" and not a.no_breakdown:
            with torch.no_grad():
                bd, tf32_flops = breakdown(lambda: full_step(dev_pcs[3], *dev_labels), ns, torch)
            line["breakdown_ms"] = bd
            ours_bd = {k: v for k, v in bd.items() if not k.startswith("_")}
            top = max(ours_bd, key=lambda k: ours_bd[k]["ms_per_step"])
            tt = ours_bd[top]
            ach = tt["alg_bytes"] / (tt["ms"] / 1e3) / 1e9
            traffic = None
            try:  # DRAM bytes per launch of the same kernel/shape from the committed `ncu --set full` capture
                tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
                traffic = tr.get(top, {}).get("dram_bytes_per_launch")
            except Exception:
                pass
            line["roofline"] = {"kernel": top, "bound": "hbm", "achieved": round(ach, 3), "peak": hbm_peak,
                                "unit": "GB/s", "frac": round(ach / hbm_peak, 5), "traffic": traffic,
                                "peak_source": peak_src, "launch_ms": tt["ms"],
                                "share_of_serialised_step": round(tt["ms_per_step"] / max(bd["_step_serialised"]["ms"], 1e-9), 3),
                                "note": "contract view (algorithmic bytes / launch time vs HBM peak).  This kernel is a serial "
                                        "dependency chain, not an HBM stream: its real bound is under `latency`"}
            if top.startswith("furthest_point_sampling"):
                line["roofline"]["latency"] = fps_latency_view(top, tt, clocks.get("sm_mhz"))
            sa = {k: v for k, v in ours_bd.items() if k.split("[")[0] in ("sa_forward", "interp_mlp_forward", "fp_rows_forward",
                                                                          "row_mlp_forward")}
            if sa:
                ms_sa = sum(v["ms_per_step"] for v in sa.values())
                fl = sum(v["alg_flops"] * max(v["calls_per_step"], 1) for v in sa.values())
                line["roofline_tensor"] = {
                    "kernel": "sa_tcp_kernel: %d fused-MLP launches per step, %.3f ms serialised (ball query, transposes and "
                              "prepasses inside these times)" % (sum(max(v["calls_per_step"], 1) for v in sa.values()), ms_sa),
                    "bound": "tensor", "unit": "TFLOP/s", "peak": round(bf16_peak / 2, 1),
                    "achieved": round(tf32_flops / (ms_sa / 1e3) / 1e12, 2),
                    "frac": round(tf32_flops / (ms_sa / 1e3) / 1e12 / (bf16_peak / 2), 4),
                    "executed_tf32_flop_per_step": tf32_flops,
                    "delivered_fp32_equivalent_tflops": round(fl / (ms_sa / 1e3) / 1e12, 2),
                    "note": "achieved = TF32 FLOPs the kernels actually issued in THIS run (every tcgen05.mma counted in-kernel: "
                            "2*128*N*8 each; b200pn2_sa_tensor_work) / CUDA-event time of the fused-MLP calls; each fp32 product "
                            "costs 3 TF32 MMAs (split precision for the 1e-5 bar), duplicated neighbour rows are skipped.  "
                            "peak = dense TF32 = half the measured bf16 cuBLAS figure.  delivered = reference-equivalent fp32 "
                            "FLOPs (every nsample row) / the same time"}
                sa_only = {k: v for k, v in sa.items() if k.startswith("sa_forward")}
                alg = sum(v["alg_bytes"] * max(v["calls_per_step"], 1) for v in sa_only.values())
                ms_only = sum(v["ms_per_step"] for v in sa_only.values())
                if sa_only:
                    view = {"bound": "hbm", "achieved": round(alg / (ms_only / 1e3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                            "frac": round(alg / (ms_only / 1e3) / 1e9 / hbm_peak, 4),
                            "note": "fused SA layers: algorithmic (fused) bytes / time -- small by construction, the fused layers "
                                    "are compute-bound (370-1860 FLOP/B), DESIGN.md 3.3"}
                    if all(k in UNFUSED_GB_PER_SCENE for k in sa_only):
                        unf = sum(UNFUSED_GB_PER_SCENE[k] * max(v["calls_per_step"], 1) for k, v in sa_only.items()) * B
                        view["unfused_equivalent"] = {"gb_per_step": round(unf, 2), "achieved": round(unf / (ms_only / 1e3), 1),
                                                      "unit": "GB/s", "frac": round(unf / (ms_only / 1e3) / hbm_peak, 3),
                                                      "note": "bytes the reference's unfused pipeline moves for these layers / the "
                                                              "fused kernels' time: > 1 = faster than that pipeline at the HBM roofline"}
                    line["roofline_sa_hbm"] = view
        # ---- configs[2]: 256 x 256 rotated 3D IoU + NMS (device-resident boxes, CUDA events) ------------------------
        synth = importlib.import_module("3dioumatch_b200.synth")

        def ev_time(fn, iters=20):
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
        try:
            ba = torch.from_numpy(synth.boxes(0, 256)).to(dev)
            bb = torch.from_numpy(synth.boxes(1, 256, jitter_of=synth.boxes(0, 256))).to(dev)
            sc = torch.rand(256, device=dev)
            iou_us = ev_time(lambda: ns.iou.boxes_iou3d_gpu(ba, bb))
            line["c3"] = {"workload": "configs[2]: boxes_iou3d_gpu 256x256 + nms_gpu(256 boxes, thresh 0.25)",
                          "iou3d_us": round(iou_us, 2),
                          "iou3d_roofline": {"bound": "hbm", "alg_bytes": (256 + 256) * 28 + 256 * 256 * 4,
                                             "achieved_gbs": round(((256 + 256) * 28 + 256 * 256 * 4) / (iou_us * 1e-6) / 1e9, 2),
                                             "peak": hbm_peak, "frac": round(((256 + 256) * 28 + 256 * 256 * 4) / (iou_us * 1e-6) / 1e9 / hbm_peak, 5),
                                             "note": "276 KB per call: launch-latency-bound (ALU/SFU work per pair, 600-1000 FLOP), "
                                                     "not an HBM stream"}}
            if a.impl == "reference":
                order = sc.sort(0, descending=True)[1]
                sb = ba[order].contiguous()

                def ref_nms():  # the C++ entry with the int32 keep it reads (iou3d_nms.cpp:98); the python wrapper passes int64
                    keep = torch.zeros(256, dtype=torch.int32)
                    n = ns.iou.iou3d_nms_cuda.nms_gpu(sb, keep, 0.25)
                    return order[keep[:n].long().to(dev)]
                line["c3"]["nms_us"] = round(ev_time(ref_nms), 2)
            else:
                line["c3"]["nms_us"] = round(ev_time(lambda: ns.iou.nms_gpu(ba, sc, 0.25)), 2)
        except Exception as e:  # noqa: BLE001
            line["c3"] = {"error": str(e)[:200]}
        # ---- pseudo-label filter (corners -> extents -> lower-half suppression), 8 scenes x 64 boxes ----
        if a.impl == "b200":
            try:
                sys.path.insert(0, os.path.join(ra.DROPIN_DIR))
                nms = importlib.import_module("utils.nms")
                rb = np.stack([synth.aabb_boxes(i, 64, 18) for i in range(8)])
                cen = torch.from_numpy(((rb[:, :, 0:3] + rb[:, :, 3:6]) / 2).astype(np.float32)).to(dev)
                siz = torch.from_numpy(rb[:, :, 3:6] - rb[:, :, 0:3]).to(dev)
                hd = torch.zeros((8, 64), dtype=torch.float64, device=dev)
                tail = torch.from_numpy(rb[:, :, 6:8]).to(dev)

                def filt():
                    _, ext_ = nms.box_extents_batch(cen, siz, hd, return_corners=False)
                    return nms.suppress_batch(torch.cat([ext_.double(), tail], -1), 0.25, use_cls=True, lhs=True)
                line["ssl_filter"] = {"workload": "8 scenes x 64 boxes: box extents + lhs_3d_faster_samecls (thresh 0.25)",
                                      "device_us": round(ev_time(filt), 2)}
            except Exception as e:  # noqa: BLE001
                line["ssl_filter"] = {"error": str(e)[:200]}
        if a.impl == "reference":
            line["cpu_baseline"] = {"value": line["value"], "unit": "scenes/s", "kind": "reference",
                                    "cores": os.cpu_count(),
                                    "sample": "unmodified reference CUDA ops (oracle/_ref, built from /root/reference) on "
                                              "cuda:0 -- the reference has no CPU implementation of this path; host cores "
                                              "only drive the launches"}
        elif n_gpus == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(torch, N)

        sub_env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT")}
        size_args = ["--batch", str(B), "--points", str(N), "--proposals", str(K)]
        if a.impl == "b200" and n_gpus == 1 and not a.no_ref:
            # ---- the reference arm in the same run (also writes the step-0 outputs the check below reads) ----------------
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "6", "--warmup", "3",
                                    "--no-extras"] + size_args, capture_output=True, text=True, timeout=900, env=sub_env)
                ref = json.loads(r.stdout.strip().splitlines()[-1])
                line["reference_cuda"] = {"value": ref.get("value"), "e2e": ref.get("e2e", {}).get("value"),
                                          "ms_per_step": ref.get("ms_per_step"), "unit": "scenes/s",
                                          "host_enqueue_ms_per_step": ref.get("host_enqueue_ms_per_step"),
                                          "what": "unmodified reference callers on the unmodified reference pointnet2/_ext + "
                                                  "iou3d_nms CUDA ops, same GPU, same inputs (stock single-stream eager path)"}
                if ref.get("value"):
                    line["speedup_vs_reference_cuda"] = round(line["value"] / ref["value"], 3)
            except Exception as e:  # the reference arm is informational here
                line["reference_cuda"] = {"unavailable": str(e)[:200]}
        if a.impl == "b200":
            # ---- in-run correctness: step-0 outputs (TF32 off) against the reference arm's saved step-0 outputs ----------
            try:
                refz = np.load(check_path)
                check = {"against": os.path.relpath(check_path, ROOT), "indices_exact": True, "max_abs_err": {}}
                for k in CHECK_KEYS:
                    if step0[k].dtype.kind in "iu":
                        same = bool(np.array_equal(step0[k], refz[k]))
                        check["indices_exact"] = check["indices_exact"] and same
                    elif k == "seed_xyz":
                        check["indices_exact"] = check["indices_exact"] and bool(np.array_equal(step0[k], refz[k]))
                    else:
                        err = np.abs(step0[k] - refz[k])
                        check["max_abs_err"][k] = float(err.max())
                        check.setdefault("frac_within_1e-3", {})[k] = float((err <= 1e-3 + 1e-3 * np.abs(refz[k])).mean())
                check["ok"] = bool(check["indices_exact"] and check["max_abs_err"].get("fp2_features", 1.0) <= 5e-4 and
                                   check["max_abs_err"].get("vote_xyz", 1.0) <= 5e-4 and
                                   min(check["frac_within_1e-3"].values()) >= 0.99)
                line["check"] = check
            except FileNotFoundError:
                line["check"] = {"skipped": "no reference-arm step-0 file (run `bench.py --impl reference` first)"}
            except Exception as e:  # noqa: BLE001
                line["check"] = {"error": str(e)[:200]}
        if a.impl == "b200" and n_gpus == 1 and not a.no_per_op:
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--per-op"], capture_output=True, text=True,
                                   timeout=900, env=sub_env)
                line["per_op"] = json.loads(r.stdout.strip().splitlines()[-1])["per_op"]
            except Exception as e:  # noqa: BLE001
                line["per_op"] = {"error": str(e)[:200]}
        if a.impl == "b200" and n_gpus == 1 and a.callers == "reference":
            # ---- the same model with this package's caller mirrors and on a 6x denser cloud ------------
            for key, extra in (("fast_callers", ["--callers", "fast"]), ("dense_variant", ["--room", "3.2,3.2,1.2"])):
                if (key == "dense_variant" and a.room != "8,8,3") or (key == "fast_callers" and not os.path.isdir(ra.FAST_CALLERS_DIR)):
                    continue
                try:
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--steps", "100", "--warmup", "5", "--no-extras"] +
                                       size_args + extra, capture_output=True, text=True, timeout=900, env=sub_env)
                    d = json.loads(r.stdout.strip().splitlines()[-1])
                    line[key] = {"value": d.get("value"), "e2e": d.get("e2e", {}).get("value"), "unit": "scenes/s",
                                 "ms_per_step": d.get("ms_per_step"), "gpu_launches_per_step": d.get("gpu_launches_per_step"),
                                 "impl_options": d.get("impl_options"), "room_m": d.get("config", {}).get("room_m")}
                except Exception as e:  # noqa: BLE001
                    line[key] = {"error": str(e)[:200]}
    if rank == 0:
        print(json.dumps(line))
        if a.impl == "b200" and isinstance(line.get("check"), dict) and line["check"].get("ok") is False:
            print("bench.py: step-0 outputs differ from the reference arm's: %s" % json.dumps(line["check"]), file=sys.stderr)
            if distributed:
                dist.destroy_process_group()
            return 3
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
