#!/bin/bash
# FPS occupancy experiment: two CTAs per SM / 4x512 clusters vs the default 8x256, measured through bench.py
cd "$(dirname "$0")/.."
O=gpurun_out/c8; mkdir -p $O
Q="--no-ref --no-cpu-baseline --no-breakdown --no-dense"
echo "== fps parity with packing"; B200_FPS_PACK=1 timeout 600 python -m pytest tests/test_gpu_pointops.py tests/test_gpu_ref_cuda.py -m gpu -x -q -k "fps or furthest or FPS" 2>&1 | tail -3 | tee $O/pytest_fps_pack.log
run() { name=$1; shift; echo "== $name"; env "$@" timeout 300 python bench.py $Q $LANES > $O/bench_$name.json 2> $O/bench_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$name.json").read().strip().splitlines()[-1]); print("$name", d["value"], d["e2e"]["value"], d["ms_per_step"])
except Exception as e: print("$name FAILED", e)
PY
}
LANES="--lanes 5"
run base X=1
run pack B200_FPS_PACK=1
run c4t512 B200_FPS_CLUSTER=4 B200_FPS_THREADS=512 B200_FPS_FORCE_MIN_N=8192
LANES="--lanes 7"
run base_l7 X=1
run pack_l7 B200_FPS_PACK=1
run c4t512_l7 B200_FPS_CLUSTER=4 B200_FPS_THREADS=512 B200_FPS_FORCE_MIN_N=8192
LANES="--lanes 3"
run pack_l3 B200_FPS_PACK=1
echo done
