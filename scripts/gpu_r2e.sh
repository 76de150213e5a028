#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2e; mkdir -p $O
echo "== tc gemm"; timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q 2>&1 | tail -8 | tee $O/t_gemm.log
echo "== sa train"; timeout 900 python -m pytest tests/test_gpu_sa_train.py -q 2>&1 | tail -40 | tee $O/t_train.log
echo "== fp debug"; timeout 300 python scripts/fp_debug.py 2>&1 | tail -20 | tee $O/fp_debug.log
echo "== sa fused"; timeout 600 python -m pytest tests/test_gpu_sa_fused.py -q 2>&1 | tail -8 | tee $O/t_sa.log
echo "== votenet"; timeout 600 python -m pytest tests/test_gpu_votenet_callers.py -x -q 2>&1 | grep -v Warning | tail -12 | tee $O/t_votenet.log
echo "== c4"; timeout 600 python bench.py --config c4 --steps 30 --warmup 3 > $O/c4.json 2> $O/c4.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e/c4.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'], d['loss_first'], d['loss_last'])
print(d['device_time']['kernel_ms_sum_per_step'], d['device_time']['kernels_per_step'])
for r in d['device_time']['top']: print(r)
PY
tail -3 $O/c4.err
echo done
