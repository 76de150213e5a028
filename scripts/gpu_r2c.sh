#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2c; mkdir -p $O
echo "== c4 2 GPUs (NCCL)"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config c4 --steps 30 --warmup 3 > $O/c4_2gpu.json 2> $O/c4_2gpu.err; cat $O/c4_2gpu.json | head -c 1800; echo; tail -3 $O/c4_2gpu.err
echo "== c5 2 GPUs (NCCL)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config c5 --steps 8 --warmup 3 > $O/c5_2gpu.json 2> $O/c5_2gpu.err; cat $O/c5_2gpu.json | head -c 1800; echo; tail -3 $O/c5_2gpu.err
echo "== test d"; timeout 300 python -m pytest tests/test_gpu_votenet_callers.py -q -x -k test_d 2>&1 | tail -5
echo done
