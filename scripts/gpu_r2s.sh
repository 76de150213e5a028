#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2s; mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_votenet_callers.py tests/test_gpu_sa_fused.py tests/test_gpu_sa_train.py -q -x 2>&1 | grep -v Warn | tail -4 | tee $O/t.log
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 300 --no-extras "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"; }
run ref X=1 --
run fast X=1 -- --callers fast
run ref2 X=1 --
run fast2 X=1 -- --callers fast
echo "== op sweep sa"; timeout 300 python scripts/op_sweep.py sa 2>&1 | grep -v Warn | tail -6
echo done
