#!/bin/bash
# round-2 evidence, N GPUs of one box (gpurun --gpus N): scene-sharded c2, NCCL gradient buckets for c4 / c5
N=${1:-2}
cd "$(dirname "$0")/.."
O=gpurun_out/final$N; mkdir -p $O
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@"; }
if [ "$N" = "2" ]; then
  echo "== two devices, two threads"; timeout 300 python -m pytest tests/test_gpu_multi_device.py -q 2>&1 | grep -v Warn | tail -4 | tee $O/t_multi_device.log
  echo "== c2"; timeout 600 bash -c "$(declare -f run); N=$N; run --no-extras" > $O/c2.json 2> $O/c2.err; head -c 300 $O/c2.json; echo; tail -2 $O/c2.err
fi
for cfg in c4 c5; do
  echo "== $cfg fast"; timeout 600 bash -c "$(declare -f run); N=$N; run --config $cfg --callers fast" > $O/${cfg}_fast.json 2> $O/${cfg}_fast.err; head -c 300 $O/${cfg}_fast.json; echo; tail -2 $O/${cfg}_fast.err
done
echo "== c5 reference callers"; timeout 600 bash -c "$(declare -f run); N=$N; run --config c5 --callers reference" > $O/c5_reference.json 2> $O/c5_reference.err; head -c 300 $O/c5_reference.json; echo
ls -la $O
echo done
