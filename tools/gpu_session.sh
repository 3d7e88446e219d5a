#!/bin/bash
# What a round's GPU check consists of, as one gpurun call:  gpurun --timeout 3000 -- 'bash tools/gpu_session.sh'
#   1. the GPU parity suite, 2. the secondary configs, 3. a bench line, 4. the launch list and the two
#   ncu --set full captures that profiles/ summarises (profiles/README.md says how they are read).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log
timeout 600 python tools/bench_configs.py --mib 256 --mode 0 2>&1 | cut -c1-250 | tee gpurun_out/cfg_all.log
# experimental: lap variant of the position-parallel compressor (kernel mode 4) — parity, then the same shapes
( TAMP_B200_EXPERIMENTAL=1 timeout 600 python -m pytest tests -m gpu -q -k 'lap_variant or no_longer or warp_per_stream or four_level' 2>&1 | tail -5 ) | tee gpurun_out/t_laps.log
timeout 600 python tools/bench_configs.py --mib 256 --mode 4 2>&1 | cut -c1-250 | tee gpurun_out/cfg_mode4.log
timeout 600 python bench.py > gpurun_out/bench_full.log 2>&1; tail -1 gpurun_out/bench_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'k_ppar_compress|k_fast_decompress' -c 2 -f \
   -o gpurun_out/full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out | tail -8
