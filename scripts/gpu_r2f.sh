#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2f; mkdir -p $O
echo "== tc gemm + sa train + sa fused"; timeout 900 python -m pytest tests/test_gpu_tc_gemm.py tests/test_gpu_sa_train.py tests/test_gpu_sa_fused.py -q 2>&1 | tail -8 | tee $O/t1.log
for cfg in c4 c5; do
echo "== $cfg"; timeout 600 python bench.py --config $cfg --steps 30 --warmup 3 > $O/$cfg.json 2> $O/$cfg.err; python - $cfg <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2f/%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'], d['loss_first'], d['loss_last'])
    print(d['device_time']['kernel_ms_sum_per_step'], d['device_time']['kernels_per_step'])
    for r in d['device_time']['top']: print(r)
except Exception as e: print('ERR', e)
PY
grep -c "timed out" $O/$cfg.err; tail -2 $O/$cfg.err
done
echo "== c4 unfused for comparison"; B200_SA_TRAIN_FUSED=0 timeout 600 python bench.py --config c4 --steps 30 --warmup 3 --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['loss_last'])"
echo done
