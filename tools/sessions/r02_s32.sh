#!/bin/bash
# Round 2, GPU session 32: walk compressor P2 with segments handed out to the lanes (TB_WALK_DYN=1, _build) against one
# 32-offset segment per lane (TB_WALK_DYN=0, _build_alt): A/B on one box, then parity of the new default
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
for rep in 1 2; do for dir in _build_alt _build; do
  echo -n "$dir: "; TAMP_B200_BUILD_DIR=$dir timeout 600 python bench.py --no-extra-configs --no-e2e 2>/dev/null | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'ext compress_ms',round(l['other_format']['compress_ms'],3), (l['cpu_baseline'] or {}).get('parity'))"
done; done 2>&1 | tee gpurun_out/s32_ab.log
( timeout 1500 python -m pytest tests -m gpu -q -x -k "fixtures or differential or no_longer or segmented or kats or traces" 2>&1 | tail -5 ) > gpurun_out/s32_tests.log
tail -3 gpurun_out/s32_tests.log
