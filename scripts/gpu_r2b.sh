#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2b; mkdir -p $O
echo "== votenet callers"; timeout 900 python -m pytest tests/test_gpu_votenet_callers.py -q -x 2>&1 | tail -30 | tee $O/t_votenet.log
echo "== profile strict"; timeout 300 python scripts/step_profile.py reference 2>&1 | tail -48 | tee $O/prof_strict.txt
echo "== profile fast"; timeout 300 python scripts/step_profile.py fast 2>&1 | tail -48 | tee $O/prof_fast.txt
echo "== bench fast only"; timeout 600 python bench.py --callers fast --steps 300 --no-extras > $O/bench_fast.json 2> $O/bench_fast.err; cat $O/bench_fast.json | head -c 1500; echo; tail -3 $O/bench_fast.err
echo "== c4 b200"; timeout 600 python bench.py --config c4 --steps 20 --warmup 3 > $O/c4.json 2> $O/c4.err; cat $O/c4.json | head -c 2500; echo; tail -5 $O/c4.err
echo "== c4 ref"; timeout 600 python bench.py --config c4 --impl reference --steps 6 --warmup 3 > $O/c4_ref.json 2> $O/c4_ref.err; cat $O/c4_ref.json | head -c 600; echo; tail -5 $O/c4_ref.err
echo "== c5 b200"; timeout 900 python bench.py --config c5 --steps 10 --warmup 3 > $O/c5.json 2> $O/c5.err; cat $O/c5.json | head -c 2500; echo; tail -5 $O/c5.err
echo "== c5 ref"; timeout 900 python bench.py --config c5 --impl reference --steps 4 --warmup 3 > $O/c5_ref.json 2> $O/c5_ref.err; cat $O/c5_ref.json | head -c 600; echo; tail -5 $O/c5_ref.err
echo done
