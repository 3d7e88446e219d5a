#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 2400 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -15 ) > gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --classes 8:256,9:512,10:1024 2>&1 | grep 'extended": 0' | cut -c1-200
timeout 300 python bench.py --steps 5 --no-cpu-baseline > gpurun_out/bench_m0.log 2>&1; tail -1 gpurun_out/bench_m0.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(l['value'], l['ms_per_step'], l['roofline']['kernel_ms'], l['roofline']['decompress']['kernel_ms'], l['e2e']['ms_per_step'], l['gpu_launches'])"
