#!/bin/bash
# Round 2, GPU session 22: walk kernel with the 8-byte first-stage compare
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1500 python -m pytest tests -m gpu -x -q -k "no_longer or fixtures or differential or custom_dictionary or kats or traces" 2>&1 | tail -6 ) > gpurun_out/s22_tests.log
tail -3 gpurun_out/s22_tests.log
timeout 900 python bench.py --no-extra-configs --no-e2e > gpurun_out/s22_bench.log 2>&1; tail -1 gpurun_out/s22_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']), (l.get('cpu_baseline') or {}).get('parity'), l['other_format'])"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/s22_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s22_ncu1.log 2>&1
timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --classes 8:256,9:512,10:1024 2>&1 | cut -c1-200 | tee gpurun_out/s22_cfg.log
