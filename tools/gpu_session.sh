#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/t_all.log
tail -4 gpurun_out/t_all.log
timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --classes 8:1024,9:1024,10:1024,10:4096 2>&1 | grep "extended\": 0" | cut -c1-260
timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_m0.log 2>&1; tail -1 gpurun_out/bench_m0.log | cut -c1-1200
