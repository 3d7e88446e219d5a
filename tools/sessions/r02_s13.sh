#!/bin/bash
# Round 2, GPU session 13: concurrent host-pointer calls (two-thread e2e leg), binding fix
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1200 python -m pytest tests -m gpu -x -q -k "two_threads or host_pointer or binding or two_streams or window_bound" 2>&1 | tail -8 ) > gpurun_out/s13_tests.log
tail -3 gpurun_out/s13_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-extra-configs > gpurun_out/s13_bench.log 2>&1; tail -1 gpurun_out/s13_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value'])); print(json.dumps(l['e2e'],indent=1)); print(r.get('traffic'), r['decompress'].get('traffic'))" || tail -20 gpurun_out/s13_bench.log
