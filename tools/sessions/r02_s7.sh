#!/bin/bash
# Round 2, GPU session 7: cooperative history walk (k_cwalk_compress, windows 11..15): parity, class sweep, plan tuning, ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1500 python -m pytest tests -m gpu -x -q -k "history_walk or full_length or differential" 2>&1 | tail -8 ) > gpurun_out/s7_tests.log
tail -3 gpurun_out/s7_tests.log
timeout 600 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 12:16384,15:65536,11:8192,13:32768,14:65536 2>&1 | cut -c1-300 | tee gpurun_out/s7_cfg.log
for plan in "13,13,512" "13,13,256" "13,12,512" "12,13,256" "12,13,512" "13,13,128"; do
  echo "plan15 $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 128 --mode 0 --v1-only --classes 15:65536 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s7_tune15.log
for plan in "12,12,256" "12,12,128" "12,13,256" "11,12,128" "13,12,512" "12,11,256" "11,12,256" "12,12,512"; do
  echo "plan12 $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 128 --mode 0 --v1-only --classes 12:16384 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s7_tune12.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_cwalk_compress' -c 1 -f \
   -o gpurun_out/s7_cwalk15 python tools/bench_configs.py --mib 64 --mode 0 --v1-only --classes 15:65536 > gpurun_out/s7_ncu15.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_cwalk_compress' -c 1 -f \
   -o gpurun_out/s7_cwalk12 python tools/bench_configs.py --mib 64 --mode 0 --v1-only --classes 12:16384 > gpurun_out/s7_ncu12.log 2>&1
