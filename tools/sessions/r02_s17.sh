#!/bin/bash
# Round 2, GPU session 17: cwalk CTA shapes (1024 threads at 64 registers)
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
for cls in 15:65536 14:65536 13:32768 12:16384 11:8192; do
for plan in "13,13,1024" "13,13,512" "12,12,512" "12,11,512" "12,11,256" "12,12,1024" "11,11,512" "11,11,256"; do
  echo "class $cls plan $plan"; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes $cls 2>&1 | cut -c1-170
done; done 2>&1 | tee gpurun_out/s17_tune.log | grep -v "^$" | paste - - | awk '{print $2,$4, $0}' | sed 's/{"window.*"compress_ms"/ compress_ms/' | cut -c1-120
