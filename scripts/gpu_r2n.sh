#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2n; mkdir -p $O
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_pseudo_labels.py -q 2>&1 | grep -v Warn | tail -12 | tee $O/t.log
echo "== fps one (speculation refuted: ordinary clouds)"; timeout 300 python scripts/op_sweep.py fps_one 2>&1 | grep -v Warn | tee $O/fps_one.txt
echo "== step profile fast"; timeout 300 python scripts/step_profile.py fast 2>&1 | grep -v Warn | tee $O/step_fast.txt | head -30
echo "== step profile reference"; timeout 300 python scripts/step_profile.py reference 2>&1 | grep -v Warn > $O/step_reference.txt
for cal in reference fast; do
echo "== c5 $cal"; timeout 600 python bench.py --config c5 --callers $cal --steps 30 --warmup 3 > $O/c5_$cal.json 2> $O/c5_$cal.err; python - $cal <<'PY'
import json,sys
try:
    d=json.loads(open('gpurun_out/r2n/c5_%s.json'%sys.argv[1]).read().strip().splitlines()[-1])
    print(d['value'], d['ms_per_step'], d['gpu_launches_per_step'], d['loss_first'], d['loss_last'], d['device_time']['kernel_ms_sum_per_step'])
    for r in d['device_time']['top'][:6]: print('   ', r)
except Exception as e: print('ERR', e)
PY
tail -2 $O/c5_$cal.err
done
echo "== c4 fast"; timeout 600 python bench.py --config c4 --callers fast --steps 30 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['loss_last'])"
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --lanes 1 --graphs 0 --no-extras > $O/b_ncu.log 2>&1; tail -1 $O/b_ncu.log | head -c 300; echo
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sa_tcp_kernel|fps_|bg_query|three_nn_kernel|pair_kernel" -o $O/prof_full python scripts/ncu_kernels.py > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
ls -la $O
echo done
