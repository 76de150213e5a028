#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2k; mkdir -p $O
echo "== fps tests"; timeout 900 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_ref_cuda.py -q 2>&1 | grep -v Warn | tail -12 | tee $O/t_fps.log
echo "== fps shapes"; timeout 900 python scripts/fps_shapes.py quick 2>&1 | grep -v Warn | tee $O/fps_shapes.txt
echo "== sa tests (fused query)"; timeout 1200 python -m pytest tests/test_gpu_sa_fused.py tests/test_gpu_tc_gemm.py tests/test_gpu_sa_train.py -q 2>&1 | grep -v Warn | tail -12 | tee $O/t_sa.log
echo "== votenet callers"; timeout 900 python -m pytest tests/test_gpu_votenet_callers.py -q 2>&1 | grep -v Warn | tail -8 | tee $O/t_votenet.log
echo "== bench"; timeout 900 python bench.py --steps 200 --no-ref --no-cpu-baseline --no-per-op > $O/bench.json 2> $O/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2k/bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('latency'), d.get('fast_callers'))
for k,v in d.get('breakdown_ms',{}).items(): print(k, v['ms'])
print(d.get('roofline'))
PY
tail -3 $O/bench.err
echo done
