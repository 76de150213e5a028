#!/bin/bash
# what the driver's SCALE step runs at N=2 (default flags, rank-0 extras on) + the two-device two-thread test
cd "$(dirname "$0")/.."
O=gpurun_out/r2drv2; mkdir -p $O
T0=$SECONDS
echo "== c2, 2 GPUs, driver flags"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > $O/c2_2gpu.json 2> $O/c2_2gpu.err; echo "rc=$? t=$((SECONDS-T0))"; head -c 400 $O/c2_2gpu.json; echo; tail -4 $O/c2_2gpu.err
echo "== reference arm under torchrun (rank 0 only)"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > $O/ref_2gpu.json 2> $O/ref_2gpu.err; echo "rc=$? t=$((SECONDS-T0))"; head -c 300 $O/ref_2gpu.json; echo
echo "== two devices, two threads"; timeout 200 python -m pytest tests/test_gpu_multi_device.py -q 2>&1 | grep -v Warn | tail -4 | tee $O/t_multi_device.log; echo "t=$((SECONDS-T0))"
echo done
