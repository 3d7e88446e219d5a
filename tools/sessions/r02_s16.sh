#!/bin/bash
# Round 2, GPU session 16: cwalk with third-byte sub-buckets: parity, plan sweep
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1500 python -m pytest tests -m gpu -x -q -k "history_walk or full_length or differential" 2>&1 | tail -8 ) > gpurun_out/s16_tests.log
tail -3 gpurun_out/s16_tests.log
for plan in "13,14,512,3" "13,13,512,0" "13,13,512,2" "13,13,512,3" "13,14,512,4" "13,14,512,2"; do
  echo "plan15 $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 15:65536 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s16_tune15.log
for plan in "12,13,256,2" "12,13,256,0" "12,13,256,3" "12,12,256,2" "13,13,512,2" "13,14,512,3"; do
  echo "plan14 $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 14:65536 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s16_tune14.log
for plan in "12,12,256,0" "12,12,256,2" "12,13,256,3" "12,11,256,0"; do
  echo "plan13 $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 13:32768 2>&1 | cut -c1-200
  echo "plan12 $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 12:16384 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s16_tune13.log
