#!/bin/bash
# Round 2, GPU session 6: history-walk compressor (k_hwalk_compress): parity, class sweep, plan tuning, ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1500 python -m pytest tests -m gpu -x -q -k "history_walk or full_length or append_mode or differential or lap_variant" 2>&1 | tail -8 ) > gpurun_out/s6_tests.log
tail -3 gpurun_out/s6_tests.log
timeout 600 python tools/bench_configs.py --mib 256 --mode 0 --v1-only 2>&1 | cut -c1-300 | tee gpurun_out/s6_cfg.log
for plan in "13,13,16,512" "13,13,16,256" "13,13,32,256" "12,13,16,256" "12,13,32,128" "13,14,16,512"; do
  echo "plan15 $plan"; TAMP_B200_HWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 128 --mode 0 --v1-only --classes 15:65536 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s6_tune15.log
for plan in "12,12,16,256" "12,12,32,128" "12,13,16,256" "11,12,16,128" "13,12,16,512" "12,12,16,128"; do
  echo "plan12 $plan"; TAMP_B200_HWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 128 --mode 0 --v1-only --classes 12:16384 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s6_tune12.log
for plan in "11,11,32,64" "10,11,32,32" "12,11,32,128" "11,11,16,128" "12,12,16,256" "10,11,16,64"; do
  echo "plan10 $plan"; TAMP_B200_HWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 128 --mode 0 --v1-only --classes 10:4096,8:1024 2>&1 | cut -c1-200
done 2>&1 | tee gpurun_out/s6_tune10.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_hwalk_compress' -c 1 -f \
   -o gpurun_out/s6_hwalk15 python tools/bench_configs.py --mib 64 --mode 0 --v1-only --classes 15:65536 > gpurun_out/s6_ncu.log 2>&1
