#!/bin/bash
# Round 2, GPU session 11: full suite on the final kernels, sanitizers (memcheck + racecheck), ncu captures for traffic.json
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s11_tests.log
tail -3 gpurun_out/s11_tests.log
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py 2>&1 | tail -30 ) > gpurun_out/s11_memcheck.log
tail -3 gpurun_out/s11_memcheck.log
( timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_check.py 2>&1 | tail -40 ) > gpurun_out/s11_racecheck.log
tail -4 gpurun_out/s11_racecheck.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_walk_compress' -c 1 -f \
   -o gpurun_out/s11_walk python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s11_ncu1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_split_decompress' -c 1 -f \
   -o gpurun_out/s11_split python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s11_ncu2.log 2>&1
timeout 300 python tools/bench_configs.py --mib 256 --mode 0 --v1-only --classes 10:4096,12:16384,15:65536 2>&1 | cut -c1-300 | tee gpurun_out/s11_cfg.log
