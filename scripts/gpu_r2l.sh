#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2l; mkdir -p $O
echo "== sa tests (fused query v2)"; timeout 1200 python -m pytest tests/test_gpu_sa_fused.py -q 2>&1 | grep -v Warn | tail -6 | tee $O/t_sa.log
run() { # name, env..., -- args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 200 --no-extras "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'])"
}
run ref_q1 B200_SA_TC_QUERY=1 --
run ref_q0 B200_SA_TC_QUERY=0 --
run fast_q1 B200_SA_TC_QUERY=1 -- --callers fast
run fast_q0 B200_SA_TC_QUERY=0 -- --callers fast
for l in 1 2 3 8 12; do run ref_lanes$l X=1 -- --lanes $l; done
for l in 3 8 12; do run fast_lanes$l X=1 -- --callers fast --lanes $l; done
echo "== op sweep sa"; timeout 300 python scripts/op_sweep.py sa 2>&1 | grep -v Warn | tail -30 | tee $O/op_sweep_sa.txt
echo done
