#!/bin/bash
# Round 2, GPU session 8: extended format through the segment-walk compressor and the split decompressor; full suite, bench, ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 2000 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s8_tests.log
tail -3 gpurun_out/s8_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s8_bench.log 2>&1; tail -1 gpurun_out/s8_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']), 'e2e', l['e2e']['ms_per_step'], l['gpu_launches'], (l.get('cpu_baseline') or {}).get('parity'), l['other_format']); print(json.dumps(l.get('configs'))[:1500])"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/s8_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s8_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_split_decompress' -c 1 -f \
   -o gpurun_out/s8_split python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s8_ncu2.log 2>&1
timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --classes 8:256,9:512,10:1024,10:4096 2>&1 | cut -c1-300 | tee gpurun_out/s8_cfg.log
