#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-format > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_ppar_compress|k_fast_decompress' -c 2 -f \
   -o gpurun_out/full_r01d python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-120
