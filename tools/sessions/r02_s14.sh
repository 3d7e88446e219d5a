#!/bin/bash
# Round 2, GPU session 14: packed host entry point, two-thread e2e over packed frames
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1200 python -m pytest tests -m gpu -x -q -k "packed or two_threads or host_pointer or compaction" 2>&1 | tail -12 ) > gpurun_out/s14_tests.log
tail -4 gpurun_out/s14_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-extra-configs > gpurun_out/s14_bench.log 2>&1; tail -1 gpurun_out/s14_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value'])); print(json.dumps(l['e2e'],indent=1))" || tail -20 gpurun_out/s14_bench.log
