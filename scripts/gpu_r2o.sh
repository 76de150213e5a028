#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2o; mkdir -p $O
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_pseudo_labels.py -q -x 2>&1 | grep -v Warn | tail -12 | tee $O/t.log
echo "== memcheck"; timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize_small.py > $O/memcheck.log 2>&1; tail -4 $O/memcheck.log
echo "== racecheck"; timeout 700 compute-sanitizer --tool racecheck --print-limit 20 python scripts/sanitize_small.py > $O/racecheck.log 2>&1; echo "rc=$?"; tail -6 $O/racecheck.log
echo done
