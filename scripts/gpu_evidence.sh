#!/bin/bash
# round-1c evidence: full GPU suite, smoke, default bench, reference arm, ncu launch list, ncu full capture
cd "$(dirname "$0")/.."
O=gpurun_out/evidence; mkdir -p $O
echo "== full gpu suite"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/pytest_gpu.log
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/smoke.log
echo "== bench (default flags)"; timeout 900 python bench.py > $O/bench.json 2> $O/bench.err; head -c 400 $O/bench.json; echo; tail -3 $O/bench.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_ref.json 2> $O/bench_ref.err; head -c 300 $O/bench_ref.json; echo
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --lanes 1 --graphs 0 --no-ref --no-cpu-baseline --no-breakdown --no-dense > $O/b_ncu.log 2>&1; tail -1 $O/b_ncu.log | head -c 200; echo
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sa_tcp_kernel|fps_cluster|bg_query|three_nn_kernel|pair_kernel|sa_unit" --launch-skip 34 -c 18 -o $O/prof_full python scripts/ncu_kernels.py > $O/ncu_full.log 2>&1; tail -2 $O/ncu_full.log
ls -la $O
echo done
