#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2m; mkdir -p $O
echo "== fps tests"; timeout 900 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_ref_cuda.py -q 2>&1 | grep -v Warn | tail -12 | tee $O/t_fps.log
echo "== votenet callers"; timeout 900 python -m pytest tests/test_gpu_votenet_callers.py -q 2>&1 | grep -v Warn | tail -4 | tee $O/t_votenet.log
echo "== fps one"; timeout 300 python scripts/op_sweep.py fps_one 2>&1 | grep -v Warn | tee $O/fps_one.txt
run() { name=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py --steps 200 --no-extras "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches_per_step'])"; }
run ref_pfx1 B200_FPS_PREFIX=1 --
run ref_pfx0 B200_FPS_PREFIX=0 --
run fast_pfx1 B200_FPS_PREFIX=1 -- --callers fast
run fast_pfx0 B200_FPS_PREFIX=0 -- --callers fast
echo "== bench full"; timeout 900 python bench.py --no-ref --no-cpu-baseline --no-per-op > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency'), d.get('fast_callers'))
for k,v in d.get('breakdown_ms',{}).items(): print(k, v['ms'])
PY
tail -3 $O/bench.err
for cfg in c4 c5; do
echo "== $cfg"; timeout 600 python bench.py --config $cfg --steps 30 --warmup 3 > $O/$cfg.json 2> $O/$cfg.err; python - $cfg <<'PY'
import json,sys
d=json.loads(open('gpurun_out/r2m/%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'], d['loss_first'], d['loss_last'], d['device_time']['kernel_ms_sum_per_step'])
PY
done
echo done
