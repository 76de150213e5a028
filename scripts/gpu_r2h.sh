#!/bin/bash
# round 2: per-kernel step tables (both caller modes), ncu launch list of the eager step, ncu full capture of the hot kernels
cd "$(dirname "$0")/.."
O=gpurun_out/r2h; mkdir -p $O
echo "== step profile (reference callers)"; timeout 300 python scripts/step_profile.py reference 2>&1 | grep -v Warn | tee $O/step_reference.txt | head -50
echo "== step profile (fast callers)"; timeout 300 python scripts/step_profile.py fast 2>&1 | grep -v Warn | tee $O/step_fast.txt | head -40
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --lanes 1 --graphs 0 --no-extras > $O/b_ncu.log 2>&1; tail -1 $O/b_ncu.log | head -c 300; echo
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sa_tcp_kernel|fps_cluster|bg_query|three_nn_kernel|pair_kernel" --launch-skip 30 -c 22 -o $O/prof_full python scripts/ncu_kernels.py > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
ls -la $O
echo done
