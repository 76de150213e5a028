#!/bin/bash
# closing verification: GPU suite twice (fresh processes), smoke, default bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2verify; mkdir -p $O
T0=$SECONDS
for i in 1 2; do timeout 300 python -m pytest tests -m gpu -q 2>&1 | grep -v Warn > $O/pytest_$i.log; echo "== suite run $i: $(grep -E 'passed|failed' $O/pytest_$i.log | tail -1)"; grep -E "^FAILED|max err|fp diag" $O/pytest_$i.log | head -5; echo "t=$((SECONDS-T0))"; done
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log; echo "t=$((SECONDS-T0))"
echo "== bench reference arm"; timeout 240 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 200 $O/bench_ref.json; echo; echo "t=$((SECONDS-T0))"
echo "== bench (default flags)"; timeout 480 python bench.py > $O/bench.json 2> $O/bench.err; head -c 300 $O/bench.json; echo; tail -2 $O/bench.err; echo "t=$((SECONDS-T0))"
echo done
