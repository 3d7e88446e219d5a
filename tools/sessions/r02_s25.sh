#!/bin/bash
# Round 2, GPU session 25: a few more cwalk shapes for windows 11..13
cd "$(dirname "$0")/../.."
for cls in 12:16384 13:32768 11:8192; do
for plan in "12,11,256" "13,11,256" "13,10,256" "12,10,256" "13,11,1024" "13,12,1024"; do
  echo -n "class $cls plan $plan: "; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes $cls 2>&1 | grep -o '"compress_GBps": [0-9.]*'
done; done
