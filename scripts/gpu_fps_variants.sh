#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/fps_variants; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 600 python bench.py $Q > $O/bench_$name.json 2> $O/bench_$name.err; tail -2 $O/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["e2e"]["value"], d["ms_per_step"])
except Exception as e: print("$name FAILED", e)
PY
}
Q="--no-ref --no-cpu-baseline --no-breakdown --no-dense"
run base X=1
run prune_all B200_FPS_PRUNE=1
run prune_big B200_FPS_PRUNE=1 B200_FPS_PRUNE_MIN_N=8192
run prune_big_c2 B200_FPS_PRUNE=1 B200_FPS_PRUNE_MIN_N=8192 B200_FPS_PRUNE_CLUSTER=2
run prune_big_l7 B200_FPS_PRUNE=1 B200_FPS_PRUNE_MIN_N=8192 LANES=7
echo "== fps_one prune"; B200_FPS_PRUNE=1 timeout 300 python scripts/op_sweep.py fps_one 2>&1 | tail -6
echo done
