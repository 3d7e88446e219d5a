#!/bin/bash
# Round 2, GPU session 27: append mode in the batch kernels, the segmented single-stream calls; fresh ncu captures of the
# two bench kernels (their sources changed: header forms)
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 python -m pytest tests/test_gpu_segmented.py -x -q 2>&1 | tail -15 ) > gpurun_out/s27_seg_tests.log
tail -5 gpurun_out/s27_seg_tests.log
timeout 600 python tools/bench_segmented.py --mib 1024 2>&1 | tee gpurun_out/s27_segbench.log | cut -c1-420
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/s27_tests.log
tail -3 gpurun_out/s27_tests.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/s27_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s27_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_split_decompress' -c 1 -f \
   -o gpurun_out/s27_split python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s27_ncu2.log 2>&1
ls -la gpurun_out/s27_*.ncu-rep
timeout 900 python bench.py --no-extra-configs > gpurun_out/s27_bench.log 2>&1; tail -1 gpurun_out/s27_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']),'e2e',round(l['e2e']['ms_per_step'],2), l['other_format'])"
