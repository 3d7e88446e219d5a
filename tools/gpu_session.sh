#!/bin/bash
# Round profile captures: launch list of the bench command + ncu --set full of the two hot kernels at full size.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
tail -1 gpurun_out/launches_bench.log | cut -c1-200
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_ppar_compress|k_fast_decompress' -c 2 -f -o gpurun_out/full_r01c \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200
