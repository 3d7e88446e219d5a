#!/bin/bash
# One gpurun call: parity tests of the batch path, compressor variants side by side, a short bench line,
# and an ncu --set full capture of the dominant kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 1500 python -m pytest tests/test_gpu_batch.py -x -q 2>&1 | tail -25 ) > gpurun_out/t_batch.log
tail -8 gpurun_out/t_batch.log
for m in 0 2; do
  timeout 300 python tools/bench_configs.py --mib 256 --mode $m --classes 8:256,9:512,10:1024 > gpurun_out/cfg_m$m.log 2>&1
  cat gpurun_out/cfg_m$m.log
done
timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_m0.log 2>&1; tail -2 gpurun_out/bench_m0.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ppar_compress -c 1 -f -o gpurun_out/ppar \
   python bench.py --streams 131072 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_ppar.log 2>&1
tail -1 gpurun_out/ncu_ppar.log
ls -la gpurun_out
