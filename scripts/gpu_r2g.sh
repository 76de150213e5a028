#!/bin/bash
# round 2 re-entry: full GPU suite at HEAD, smoke, default bench (both arms), c4/c5 bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2g; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee $O/gpu.txt
echo "== gpu suite"; timeout 2400 python -m pytest tests -m gpu -q 2>&1 | grep -v Warning | tail -30 | tee $O/t_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/smoke.log
echo "== bench"; timeout 1500 python bench.py > $O/bench.json 2> $O/bench.err; head -c 3000 $O/bench.json; echo; tail -5 $O/bench.err
echo "== bench ref"; timeout 900 python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; head -c 1500 $O/bench_ref.json; echo; tail -3 $O/bench_ref.err
for cfg in c4 c5; do
echo "== $cfg"; timeout 600 python bench.py --config $cfg --steps 30 --warmup 3 > $O/$cfg.json 2> $O/$cfg.err; head -c 600 $O/$cfg.json; echo; tail -2 $O/$cfg.err
done
echo done
