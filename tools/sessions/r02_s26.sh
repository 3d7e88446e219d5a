#!/bin/bash
# Round 2, GPU session 26: half-warp walkers (gl = 16) for windows 11..13 against the 32-lane plan
cd "$(dirname "$0")/../.."
for cls in 11:8192 12:16384 13:32768; do
for plan in "12,11,256,32" "12,11,256,16" "12,11,128,16" "12,11,512,16" "11,11,256,16" "12,10,256,16"; do
  echo -n "class $cls plan $plan: "; TAMP_B200_CWALK_PLAN=$plan timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes $cls 2>&1 | grep -o '"compress_GBps": [0-9.]*\|"ratio": [0-9.]*\|"round_trip_ok": [a-z]*\|Error.*' | tr "\n" " "; echo
done; done
