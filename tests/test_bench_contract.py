"""The committed bench lines (profiles/r2_bench_*.json, produced by bench.py on a B200) carry every key the measurement
contract names, and their derived numbers are consistent (no GPU needed: this checks the artefacts, not the device)."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip("%s not committed" % name)
    return json.loads(open(path).read().strip().splitlines()[-1])


def test_product_line_has_the_contract_keys():
    d = _line("r2_bench_1gpu.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "scenes/s" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["vs_baseline"] is None          # BASELINE.json publishes no number for this metric
    assert d["warmup"] >= 3 and d["n_gpus"] == 1 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and d["config"]["model"].startswith("models/votenet_iou_branch.py")
    assert "lanes" not in d["config"] and "host_enqueue_ms_per_step" not in d["config"]   # arm-specific -> impl_options
    assert d["impl_options"]["callers"] == "reference"
    assert d["check"]["ok"] is True and d["check"]["indices_exact"] is True
    t = d["roofline_tensor"]
    assert t["bound"] == "tensor" and abs(t["frac"] - t["achieved"] / t["peak"]) <= 1e-3 and t["executed_tf32_flop_per_step"] > 0
    assert d["roofline"]["latency"]["us_per_iteration"] > 0 and len(d["per_op"]) >= 8
    # value is whole-job throughput: scenes per step / time per step
    scenes = d["config"]["scenes_per_gpu_per_step"] * d["n_gpus"]
    assert abs(d["value"] - scenes / (d["ms_per_step"] / 1e3)) / d["value"] < 1e-3
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-4
    assert r["traffic"] is None or r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    clk = d["clocks"]
    assert not set(clk["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert clk["sm_mhz"] >= 0.9 * clk["sm_max_mhz"]


def test_reference_arm_line():
    d = _line("r2_bench_reference_arm.json")
    p = _line("r2_bench_1gpu.json")
    assert d["impl"] == "reference"
    for k in ("metric", "unit", "higher_is_better"):
        assert d[k] == p[k]
    assert d["config"] == p["config"]           # same workload, same model file: only impl_options differ
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["unit"] == d["unit"] and d["e2e"]["value"] > 0


def test_two_gpu_line_scales():
    d = _line("r2_bench_2gpu.json")
    p = _line("r2_bench_1gpu.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "weak"
    assert d["value"] > 1.8 * p["value"]     # scene-sharded, no data-path collective


def test_fps_latency_view_is_a_pure_function_of_the_measurement():
    """bench.fps_latency_view: microseconds per dependent iteration against the stated per-iteration latency floor."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    v = bench.fps_latency_view("furthest_point_sampling[N=40000,m=2048]", {"ms": 2.047}, 1965.0)
    assert v["bound"] == "latency" and v["iterations"] == 2047 and abs(v["us_per_iteration"] - 1.0) < 1e-6
    assert abs(v["floor_us_per_iteration"] - bench.FPS_FLOOR_CYCLES_CLUSTER / 1965.0) < 1e-4
    assert abs(v["frac_of_floor"] - v["floor_us_per_iteration"] / v["us_per_iteration"]) < 1e-3
    w = bench.fps_latency_view("furthest_point_sampling[N=1024,m=512]", {"ms": 0.1533}, 1965.0)
    assert abs(w["floor_us_per_iteration"] - bench.FPS_FLOOR_CYCLES_SINGLE_CTA / 1965.0) < 1e-4
