#!/bin/bash
# Round 2, GPU session 28: the segmented single-stream calls (library rebuilt: session 27 ran a stale .so);
# fresh ncu captures of the two bench kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 python -m pytest tests/test_gpu_segmented.py -q 2>&1 | tail -15 ) > gpurun_out/s28_seg_tests.log
tail -5 gpurun_out/s28_seg_tests.log
timeout 600 python tools/bench_segmented.py --mib 1024 2>&1 | tee gpurun_out/s28_segbench.log | cut -c1-600
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/s28_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s28_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_split_decompress' -c 1 -f \
   -o gpurun_out/s28_split python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s28_ncu2.log 2>&1
ls -la gpurun_out/s28_*.ncu-rep
