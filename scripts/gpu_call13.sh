#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/c16; mkdir -p $O
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_sa_fused.py tests/test_gpu_harness_vs_reference.py -m gpu -x -q 2>&1 | tail -4 | tee $O/pytest.log
run() { name=$1; shift; env "$@" timeout 600 python bench.py $Q > $O/bench_$name.json 2> $O/bench_$name.err; tail -2 $O/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["e2e"]["value"], d["ms_per_step"])
except Exception as e: print("$name FAILED", e)
PY
}
Q="--no-ref --no-cpu-baseline --no-breakdown --no-dense"
H=$PWD/3dioumatch_b200/lib/libb200pc_head.so
run head1 B200_LIB_PATH=$H
run cur1 X=1
run head2 B200_LIB_PATH=$H
run cur2 X=1
echo "== op sweep head"; B200_LIB_PATH=$H timeout 300 python scripts/op_sweep.py sa 2>&1 | tail -5
echo "== op sweep cur"; timeout 300 python scripts/op_sweep.py sa 2>&1 | tail -5
echo done
