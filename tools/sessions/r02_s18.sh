#!/bin/bash
# Round 2, GPU session 18: cwalk default plans after the sweep: parity, a few more shapes
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1500 python -m pytest tests -m gpu -x -q -k "history_walk or full_length or differential" 2>&1 | tail -8 ) > gpurun_out/s18_tests.log
tail -3 gpurun_out/s18_tests.log
timeout 600 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 11:8192,12:16384,13:32768,14:65536,15:65536 2>&1 | cut -c1-170 | tee gpurun_out/s18_cfg.log
for cls in 15:65536 14:65536; do
for plan in "13,12,1024" "13,11,1024" "13,14,1024"; do
  echo "class $cls plan $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes $cls 2>&1 | cut -c1-170
done; done 2>&1 | tee gpurun_out/s18_tune.log
for cls in 13:32768 12:16384; do
for plan in "12,10,256" "11,10,256" "13,11,512" "12,11,128"; do
  echo "class $cls plan $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes $cls 2>&1 | cut -c1-170
done; done 2>&1 | tee -a gpurun_out/s18_tune.log
