#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -8 ) > gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log | cut -c1-300
timeout 120 python tools/bench_configs.py --mib 256 --mode 0 --classes 8:256,10:1024 2>&1 | cut -c1-250
