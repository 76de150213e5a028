#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/r2r; mkdir -p $O
timeout 600 python scripts/tcp_profile.py 2>&1 | grep -v Warn | tee $O/tcp_profile.txt
echo done
