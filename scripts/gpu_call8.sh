#!/bin/bash
# merged pack launch + bench robustness: parity, op sweep, default bench twice
cd "$(dirname "$0")/.."
O=gpurun_out/c10; mkdir -p $O
echo "== sa parity"; timeout 900 python -m pytest tests/test_gpu_sa_fused.py tests/test_gpu_harness_vs_reference.py -m gpu -x -q 2>&1 | tail -8 | tee $O/pytest_sa.log
echo "== op sweep factor=1"; timeout 300 python scripts/op_sweep.py sa 2>&1 | tail -12 | tee $O/sweep_f1.txt
run() { name=$1; shift; echo "== $name"; env "$@" timeout 600 python bench.py $Q > $O/bench_$name.json 2> $O/bench_$name.err; tail -2 $O/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["e2e"]["value"], d["ms_per_step"], d["config"]["host_enqueue_ms_per_step"], d["e2e"].get("host_enqueue_ms_per_step"))
    for k,v in d.get("breakdown_ms",{}).items(): print("   ", k, v["ms"])
except Exception as e: print("$name FAILED", e)
PY
}
Q="--no-ref --no-cpu-baseline --no-dense"
run full X=1
Q="--no-ref --no-cpu-baseline --no-breakdown --no-dense"
run r2 X=1
run r3 X=1
run f0 B200_SA_TC_FACTOR=0
echo done
