#!/bin/bash
# packed-fp32 three_nn + FPS update: parity, FPS timings, bench with FPS launch-shape variants
cd "$(dirname "$0")/.."
O=gpurun_out/c11; mkdir -p $O
echo "== pointops parity"; timeout 900 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_ref_cuda.py -m gpu -x -q --durations=6 2>&1 | tail -14 | tee $O/pytest_po.log
echo "== fps_one"; timeout 300 python scripts/op_sweep.py fps_one 2>&1 | tail -8 | tee $O/fps_one.txt
run() { name=$1; shift; echo "== $name"; env "$@" timeout 600 python bench.py $Q > $O/bench_$name.json 2> $O/bench_$name.err; tail -2 $O/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["host_enqueue_ms_per_step"], d["e2e"].get("host_enqueue_ms_per_step"))
    for k,v in d.get("breakdown_ms",{}).items():
        if "three_nn" in k or "furthest" in k: print("   ", k, v["ms"])
except Exception as e: print("$name FAILED", e)
PY
}
Q="--no-ref --no-cpu-baseline --no-dense"
run full X=1
Q="--no-ref --no-cpu-baseline --no-breakdown --no-dense"
run base2 X=1
run pack B200_FPS_PACK=1
run c4t512 B200_FPS_CLUSTER=4 B200_FPS_THREADS=512 B200_FPS_FORCE_MIN_N=8192
run pack_l7 B200_FPS_PACK=1 LANESX=1
echo done
