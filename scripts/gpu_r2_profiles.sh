#!/bin/bash
# closing profiles on the final code: per-kernel step tables (both caller modes) and the ncu launch list of the eager step
cd "$(dirname "$0")/.."
O=gpurun_out/r2prof; mkdir -p $O
T0=$SECONDS
echo "== step profile (reference callers)"; timeout 120 python scripts/step_profile.py reference 2>&1 | grep -v Warn > $O/step_reference.txt; head -12 $O/step_reference.txt | cut -c1-140; echo "t=$((SECONDS-T0))"
echo "== step profile (fast callers)"; timeout 120 python scripts/step_profile.py fast 2>&1 | grep -v Warn > $O/step_fast.txt; head -8 $O/step_fast.txt | cut -c1-140; echo "t=$((SECONDS-T0))"
echo "== ncu launch list"; timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --lanes 1 --graphs 0 --no-extras > $O/b_ncu.log 2>&1; tail -1 $O/b_ncu.log | head -c 200; echo; wc -l $O/launches.csv; echo "t=$((SECONDS-T0))"
echo done
