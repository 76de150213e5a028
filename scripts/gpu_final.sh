#!/bin/bash
# round-end verification: full GPU suite, smoke, default bench, reference arm (no profiler)
cd "$(dirname "$0")/.."
O=gpurun_out/final; mkdir -p $O
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16 | tee $O/pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
echo "== bench (default flags)"; timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; head -c 300 $O/bench.json; echo; tail -3 $O/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 200 $O/bench_ref.json; echo
echo done
