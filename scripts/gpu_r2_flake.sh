#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2flake; mkdir -p $O
export B200_TEST_DIAG=1
T0=$SECONDS
lscpu | grep -E "Model name|Flags" | cut -c1-2000 > $O/cpu.txt; grep -o "amx[a-z_0-9]*\|avx512_bf16\|avx512_fp16" $O/cpu.txt | sort -u | tr '\n' ' '; echo; head -1 $O/cpu.txt
run() { name=$1; shift; timeout 200 python -m pytest "$@" -q -m gpu 2>&1 | grep -v Warn > $O/$name.log; echo "== $name: $(grep -E 'passed|failed' $O/$name.log | tail -1)"; grep -A14 "fp diag" $O/$name.log | head -16; echo "t=$((SECONDS-T0))"; }
run file_only tests/test_gpu_sa_fused.py
run file_only2 tests/test_gpu_sa_fused.py
run pseudo_then tests/test_gpu_pseudo_labels.py tests/test_gpu_sa_fused.py
echo done
