#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/c5; mkdir -p $O
for L in 2 4 5 6; do
  echo "== bench lanes $L"; timeout 300 python bench.py --steps 100 --lanes $L --no-ref --no-cpu-baseline --no-breakdown > $O/bench_l$L.json 2> $O/bench_l$L.err; head -c 180 $O/bench_l$L.json; echo
done
echo "== bench fps 4x512 lanes 3"; B200_FPS_FORCE_MIN_N=8192 B200_FPS_CLUSTER=4 B200_FPS_THREADS=512 timeout 300 python bench.py --steps 100 --no-ref --no-cpu-baseline --no-breakdown > $O/bench_fps4.json 2> $O/bench_fps4.err; head -c 180 $O/bench_fps4.json; echo
echo "== bench fps 4x512 lanes 5"; B200_FPS_FORCE_MIN_N=8192 B200_FPS_CLUSTER=4 B200_FPS_THREADS=512 timeout 300 python bench.py --steps 100 --lanes 5 --no-ref --no-cpu-baseline --no-breakdown > $O/bench_fps4l5.json 2> $O/bench_fps4l5.err; head -c 180 $O/bench_fps4l5.json; echo
echo "== bench fps 16x256 lanes 3"; B200_FPS_FORCE_MIN_N=8192 B200_FPS_CLUSTER=16 B200_FPS_THREADS=256 timeout 300 python bench.py --steps 100 --no-ref --no-cpu-baseline --no-breakdown > $O/bench_fps16.json 2> $O/bench_fps16.err; head -c 180 $O/bench_fps16.json; echo
B200_FPS_FORCE_MIN_N=8192 B200_FPS_CLUSTER=4 B200_FPS_THREADS=512 timeout 100 python scripts/op_sweep.py fps_one 2>&1 | head -3
B200_FPS_FORCE_MIN_N=8192 B200_FPS_CLUSTER=16 B200_FPS_THREADS=256 timeout 100 python scripts/op_sweep.py fps_one 2>&1 | head -3
echo done
