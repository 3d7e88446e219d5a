#!/bin/bash
# Round 2, GPU session 1: validate the kernel-mode-4 kernels, measure them, sanitizers, source-level ncu capture of ppar.
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( TAMP_B200_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/s1_tests.log
tail -5 gpurun_out/s1_tests.log
timeout 400 python tools/bench_configs.py --mib 256 --mode 0 2>&1 | cut -c1-300 > gpurun_out/s1_cfg_m0.log
timeout 400 python tools/bench_configs.py --mib 256 --mode 4 2>&1 | cut -c1-300 > gpurun_out/s1_cfg_m4.log
cat gpurun_out/s1_cfg_m4.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_ppar_compress' -c 1 -f \
   -o gpurun_out/s1_ppar python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/s1_ncu_ppar.log 2>&1
( TAMP_B200_EXPERIMENTAL=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py 2>&1 | tail -30 ) > gpurun_out/s1_memcheck.log
tail -3 gpurun_out/s1_memcheck.log
( TAMP_B200_EXPERIMENTAL=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_check.py 2>&1 | tail -30 ) > gpurun_out/s1_racecheck.log
tail -3 gpurun_out/s1_racecheck.log
ls -la gpurun_out | tail -8
