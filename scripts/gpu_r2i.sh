#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2i; mkdir -p $O
echo "== fps tests"; timeout 900 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_ref_cuda.py -x -q 2>&1 | grep -v Warn | tail -8 | tee $O/t_fps.log
echo "== fps shapes"; timeout 900 python scripts/fps_shapes.py 2>&1 | grep -v Warn | tee $O/fps_shapes.txt
echo done
