#!/bin/bash
# Round 2, GPU session 21: split decompressor with the rare tokens out of line
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1500 python -m pytest tests -m gpu -x -q -k "no_longer or fixtures or differential or hostile or exact_capacity or custom_dictionary" 2>&1 | tail -6 ) > gpurun_out/s21_tests.log
tail -3 gpurun_out/s21_tests.log
timeout 900 python bench.py --no-extra-configs --no-e2e --no-cpu-baseline > gpurun_out/s21_bench.log 2>&1; tail -1 gpurun_out/s21_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']), l['other_format'])"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_split_decompress' -c 1 -f \
   -o gpurun_out/s21_split python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s21_ncu2.log 2>&1
