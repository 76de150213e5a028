#!/bin/bash
# round-2 closing evidence on one GPU, budgeted for <= 17 min: reference arm, default bench, full GPU suite, smoke
cd "$(dirname "$0")/.."
O=gpurun_out/r2final; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee $O/gpu.txt
T0=$SECONDS
echo "== bench reference arm"; timeout 240 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 300 $O/bench_ref.json; echo; echo "t=$((SECONDS-T0))"
echo "== bench (default flags)"; timeout 480 python bench.py > $O/bench.json 2> $O/bench.err; head -c 400 $O/bench.json; echo; tail -3 $O/bench.err; echo "t=$((SECONDS-T0))"
echo "== full gpu suite"; timeout 480 python -m pytest tests -m gpu -q -x --durations=12 2>&1 | grep -v Warn | tail -22 | tee $O/pytest_gpu.log; echo "t=$((SECONDS-T0))"
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log; echo "t=$((SECONDS-T0))"
ls -la $O
echo done
