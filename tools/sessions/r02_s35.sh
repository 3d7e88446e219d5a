#!/bin/bash
# Round 2, GPU session 35: long split decompressor with the history in shared memory (kernel mode 8) against modes 0 / 6
# (record only: kernel mode 8 and tools/bench_lsplit_hb.py were removed after this session — the variant brought nothing)
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 python -m pytest tests/test_gpu_batch.py -q -x -k "differential or hostile or exact_capacity or custom_dictionary or lap_variant or history_walk" 2>&1 | tail -5 ) > gpurun_out/s35_tests.log
tail -3 gpurun_out/s35_tests.log
timeout 600 python tools/bench_lsplit_hb.py 2>&1 | tee gpurun_out/s35_hb.log | cut -c1-500
