#!/bin/bash
# Round 2, GPU session 3: tightened walk kernel; host copy ceiling; decompressor source-level capture.
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 python -m pytest tests -m gpu -x -q -k "no_longer or fixtures or batch" 2>&1 | tail -5 ) > gpurun_out/s3_tests.log
tail -2 gpurun_out/s3_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/s3_bench.log 2>&1; tail -1 gpurun_out/s3_bench.log | cut -c1-1800
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress|k_fast_decompress' -c 2 -f \
   -o gpurun_out/s3_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format > gpurun_out/s3_ncu.log 2>&1
( nvidia-smi topo -m; lscpu | head -25; numactl -H; free -g ) > gpurun_out/s3_topo.log 2>&1
timeout 300 python tools/e2e_probe.py --pin 0 2>&1 | tail -1 | tee gpurun_out/s3_probe.log
timeout 300 python tools/e2e_probe.py --chunk-mib 16 --pin 0 2>&1 | tail -1 | tee -a gpurun_out/s3_probe.log
timeout 300 python tools/e2e_probe.py --chunk-mib 256 --pin 0 2>&1 | tail -1 | tee -a gpurun_out/s3_probe.log
