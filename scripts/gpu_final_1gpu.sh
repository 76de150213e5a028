#!/bin/bash
# round-2 evidence, one GPU: full GPU suite, smoke, default bench, reference arm, training benches (both caller modes + reference arm)
cd "$(dirname "$0")/.."
O=gpurun_out/final1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee $O/gpu.txt
echo "== full gpu suite"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | grep -v Warn | tail -8 | tee $O/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; head -c 300 $O/bench_ref.json; echo
echo "== bench (default flags)"; timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; head -c 400 $O/bench.json; echo; tail -3 $O/bench.err
for cfg in c4 c5; do
  for cal in reference fast; do
    echo "== $cfg $cal"; timeout 600 python bench.py --config $cfg --callers $cal > $O/${cfg}_$cal.json 2> $O/${cfg}_$cal.err; head -c 200 $O/${cfg}_$cal.json; echo
  done
  echo "== $cfg reference arm"; timeout 900 python bench.py --config $cfg --impl reference --steps 8 --warmup 3 > $O/${cfg}_refarm.json 2> $O/${cfg}_refarm.err; head -c 200 $O/${cfg}_refarm.json; echo; tail -2 $O/${cfg}_refarm.err
done
ls -la $O
echo done
