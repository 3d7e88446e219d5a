#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for r in 1 2 4 6 8 12 16 24; do
 echo "refill $r"; TB_REFILL=$r timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --classes 10:1024 2>&1 | grep "extended\": 0" | cut -c80-200
done
