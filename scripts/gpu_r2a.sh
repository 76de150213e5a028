#!/bin/bash
# round 2, first GPU pass: new kernels' parity, reference-caller parity at full size, smoke, first bench line
cd "$(dirname "$0")/.."
O=gpurun_out/r2a; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | tee $O/gpu.txt
echo "== tc gemm / row mlp"; timeout 300 python -m pytest tests/test_gpu_tc_gemm.py -x -q 2>&1 | tail -15 | tee $O/t_gemm.log
echo "== sa fused"; timeout 600 python -m pytest tests/test_gpu_sa_fused.py -x -q 2>&1 | tail -25 | tee $O/t_sa.log
echo "== votenet callers"; timeout 900 python -m pytest tests/test_gpu_votenet_callers.py -q 2>&1 | tail -40 | tee $O/t_votenet.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $O/smoke.log
echo "== bench"; timeout 1200 python bench.py --steps 200 > $O/bench.json 2> $O/bench.err; head -c 1500 $O/bench.json; echo; tail -5 $O/bench.err
echo "== rest of gpu suite"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_tc_gemm.py --deselect tests/test_gpu_sa_fused.py --deselect tests/test_gpu_votenet_callers.py 2>&1 | tail -15 | tee $O/t_rest.log
echo done
