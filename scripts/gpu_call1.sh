#!/bin/bash
# GPU session 1 (re-entry): persistent tensor-core SA kernel -- parity + A/B timing
cd "$(dirname "$0")/.."
O=gpurun_out/c1; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/gpu.txt 2>&1
echo "== parity (persist=1, final=1)"; timeout 600 python -m pytest tests/test_gpu_sa_fused.py tests/test_gpu_ref_cuda.py -x -q 2>&1 | tail -15 | tee $O/pytest_p1f1.log
echo "== parity (persist=1, final=0)"; B200_SA_TC_FINAL=0 timeout 400 python -m pytest tests/test_gpu_sa_fused.py -x -q 2>&1 | tail -8 | tee $O/pytest_p1f0.log
echo "== parity (persist=0)"; B200_SA_TC_PERSIST=0 timeout 400 python -m pytest tests/test_gpu_sa_fused.py -x -q 2>&1 | tail -5 | tee $O/pytest_p0.log
for cfg in "B200_SA_TC_PERSIST=0" "B200_SA_TC_PERSIST=1 B200_SA_TC_FINAL=0" "B200_SA_TC_PERSIST=1 B200_SA_TC_FINAL=1" "B200_SA_TC_PERSIST=1 B200_SA_TC_SLOTS=4"; do
  echo "== op_sweep sa [$cfg]"; env $cfg timeout 200 python scripts/op_sweep.py sa 2>&1 | tail -8 | tee -a $O/sweep.log
done
echo "== bench default"; timeout 300 python bench.py --steps 100 --no-ref --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
echo "== bench persist=0"; B200_SA_TC_PERSIST=0 timeout 300 python bench.py --steps 100 --no-ref --no-cpu-baseline --no-breakdown > $O/bench_p0.json 2> $O/bench_p0.err; head -c 300 $O/bench_p0.json
echo "== bench fps 4x512"; B200_FPS_FORCE_MIN_N=8192 B200_FPS_CLUSTER=4 B200_FPS_THREADS=512 timeout 300 python bench.py --steps 100 --no-ref --no-cpu-baseline --no-breakdown > $O/bench_fps4.json 2> $O/bench_fps4.err; head -c 300 $O/bench_fps4.json
echo "== bench lanes 4"; timeout 300 python bench.py --steps 100 --lanes 4 --no-ref --no-cpu-baseline --no-breakdown > $O/bench_l4.json 2> $O/bench_l4.err; head -c 300 $O/bench_l4.json
echo done
