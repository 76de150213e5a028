#!/bin/bash
# direct first layer (SA1) + FPS policy: parity, op sweep, bench
cd "$(dirname "$0")/.."
O=gpurun_out/c12; mkdir -p $O
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_sa_fused.py tests/test_gpu_harness_vs_reference.py tests/test_gpu_pointops.py -m gpu -x -q -k "not every_cluster_size" 2>&1 | tail -8 | tee $O/pytest.log
echo "== op sweep"; timeout 300 python scripts/op_sweep.py sa 2>&1 | tail -6 | tee $O/sweep.txt
run() { name=$1; shift; echo "== $name"; env "$@" timeout 600 python bench.py $Q > $O/bench_$name.json 2> $O/bench_$name.err; tail -2 $O/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["host_enqueue_ms_per_step"], d["e2e"].get("host_enqueue_ms_per_step"), d["config"].get("fps_policy"))
except Exception as e: print("$name FAILED", e)
PY
}
Q="--no-ref --no-cpu-baseline --no-breakdown --no-dense"
run a X=1
run b X=1
run lat B200_FPS_POLICY=0
run l7 X=1 
echo done
