#!/bin/bash
# Round 2, GPU session 9: long split decompressor (k_lsplit_decompress), lean split decompressor, stream-ordered scratch: full suite, bench, ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s9_tests.log
tail -3 gpurun_out/s9_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/s9_bench.log 2>&1; tail -1 gpurun_out/s9_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']), 'e2e', l['e2e']['ms_per_step'], l['gpu_launches'], (l.get('cpu_baseline') or {}).get('parity'), l['other_format']); print(json.dumps(l.get('configs'))[:1500])"
timeout 600 python tools/bench_configs.py --mib 256 --mode 0 2>&1 | cut -c1-300 | tee gpurun_out/s9_cfg.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/s9_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s9_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_split_decompress' -c 1 -f \
   -o gpurun_out/s9_split python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s9_ncu2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_lsplit_decompress' -c 1 -f \
   -o gpurun_out/s9_lsplit10 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 10:4096 > gpurun_out/s9_ncu3.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_lsplit_decompress' -c 1 -f \
   -o gpurun_out/s9_lsplit15 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 15:65536 > gpurun_out/s9_ncu4.log 2>&1
