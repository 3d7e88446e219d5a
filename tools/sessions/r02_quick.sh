#!/bin/bash
# quick GPU check of the compressor: parity subset, kernel-only bench line, optional ncu of the walk kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 python -m pytest tests -m gpu -x -q -k "no_longer or fixtures" 2>&1 | tail -3 ) > gpurun_out/q_tests.log
tail -1 gpurun_out/q_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-other-format > gpurun_out/q_bench.log 2>&1; tail -1 gpurun_out/q_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']), (l.get('cpu_baseline') or {}).get('parity'))"
if [ "$1" = "ncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/q_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format > gpurun_out/q_ncu.log 2>&1
fi
