#!/bin/bash
# Round 2, GPU session 20: the history-walk kernel (dynamic segments) on the config-2 shape, against the segment-walk kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 300 python tools/bench_configs.py --mib 1024 --mode 0 --v1-only --classes 10:1024 2>&1 | cut -c1-170
for plan in "10,11,16,32" "10,11,32,32" "10,11,16,64" "10,10,16,32" "9,11,16,32"; do
  echo "hwalk plan $plan"; TAMP_B200_HWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 1024 --mode 7 --v1-only --classes 10:1024 2>&1 | cut -c1-170
done 2>&1
