#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/c4; mkdir -p $O
echo "== parity (units on)"; timeout 600 python -m pytest tests/test_gpu_sa_fused.py tests/test_gpu_ref_cuda.py tests/test_gpu_harness_vs_reference.py -x -q 2>&1 | tail -15 | tee $O/pytest.log
echo "== parity (units off)"; B200_SA_TC_UNITS=0 timeout 600 python -m pytest tests/test_gpu_sa_fused.py -x -q 2>&1 | tail -5 | tee $O/pytest_u0.log
for cfg in "B200_SA_TC_UNITS=1" "B200_SA_TC_UNITS=0"; do
  echo "== op_sweep sa [$cfg]"; env $cfg timeout 200 python scripts/op_sweep.py sa 2>&1 | tail -8 | tee -a $O/sweep.log
done
echo "== bench default"; timeout 300 python bench.py --steps 100 --no-ref --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; head -c 400 $O/bench_default.json; echo
echo done
