#!/bin/bash
# Round 2, GPU session 2: first run of the segment-walk compressor (k_walk_compress): parity suite, bench, ncu.
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s2_tests.log
tail -4 gpurun_out/s2_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/s2_bench.log 2>&1; tail -1 gpurun_out/s2_bench.log | cut -c1-1500
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/s2_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format > gpurun_out/s2_ncu_walk.log 2>&1
timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --classes 8:256,9:512,10:1024,8:1024,10:4096 2>&1 | cut -c1-300 > gpurun_out/s2_cfg.log
for g in 0 1 2 3 4 5; do timeout 120 python tools/bench_configs.py --mib 64 --gen $g --classes 10:1024 2>&1 | head -1 | cut -c1-300 >> gpurun_out/s2_gens.log; done
cat gpurun_out/s2_cfg.log gpurun_out/s2_gens.log
