#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2p; mkdir -p $O
echo "== fps tests"; timeout 600 python -m pytest tests/test_gpu_pointops.py -q -x 2>&1 | grep -v Warn | tail -8 | tee $O/t.log
echo "== sa + callers tests"; timeout 900 python -m pytest tests/test_gpu_sa_fused.py tests/test_gpu_votenet_callers.py -q -x 2>&1 | grep -v Warn | tail -4 | tee $O/t2.log
echo "== fps shapes"; timeout 600 python scripts/fps_shapes.py quick 2>&1 | grep -v Warn | tee $O/fps_shapes.txt
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 200 --no-extras "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"; }
run fast_default X=1 -- --callers fast
run fast_mintiles1 B200_SA_TC_MIN_TILES=1 -- --callers fast
run fast_mintiles2 B200_SA_TC_MIN_TILES=2 -- --callers fast
run fast_mintiles5 B200_SA_TC_MIN_TILES=5 -- --callers fast
run ref_default X=1 --
run ref_mintiles1 B200_SA_TC_MIN_TILES=1 --
echo done
