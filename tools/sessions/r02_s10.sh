#!/bin/bash
# Round 2, GPU session 10: long split decompressor with the group mirror; mode 6 tests; config 4 / class sweep in modes 0 and 6
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s10_tests.log
tail -3 gpurun_out/s10_tests.log
for mode in 0 6; do
timeout 600 python tools/bench_configs.py --mib 1024 --mode $mode --v1-only --classes 8:1024,10:4096,12:16384,15:65536 2>&1 | cut -c1-300
done | tee gpurun_out/s10_cfg.log
timeout 600 python tools/bench_configs.py --mib 4096 --mode 6 --v1-only --classes 10:4096 2>&1 | cut -c1-300 | tee -a gpurun_out/s10_cfg.log
timeout 600 python tools/bench_configs.py --mib 4096 --mode 0 --v1-only --classes 10:4096 2>&1 | cut -c1-300 | tee -a gpurun_out/s10_cfg.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_lsplit_decompress' -c 1 -f \
   -o gpurun_out/s10_lsplit10 python tools/bench_configs.py --mib 1024 --mode 6 --v1-only --classes 10:4096 > gpurun_out/s10_ncu3.log 2>&1
