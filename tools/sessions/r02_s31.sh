#!/bin/bash
# Round 2, GPU session 31: walk compressor, chain build fetching one block ahead (A/B on one box)
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
for rep in 1 2; do for ahead in 0 1; do
  echo -n "ahead=$ahead: "; TAMP_B200_WALK_AHEAD=$ahead timeout 600 python bench.py --no-extra-configs --no-e2e --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'ext compress_ms',round(l['other_format']['compress_ms'],3), l['cpu_baseline'])"
done; done 2>&1 | tee gpurun_out/s31_ab.log
