#!/bin/bash
# Round 2, GPU session 30: split decompressor with its reads ordered before its writes (racecheck warnings of session 29:
# over-read bytes of an independent token against the stores of its neighbours), closing-FLUSH frames
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 python -m pytest tests/test_gpu_segmented.py -q 2>&1 | tail -8 ) > gpurun_out/s30_seg_tests.log
tail -3 gpurun_out/s30_seg_tests.log
timeout 900 python bench.py --no-extra-configs --no-e2e --no-cpu-baseline > gpurun_out/s30_bench.log 2>&1; tail -1 gpurun_out/s30_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']), l['other_format'])"
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_check.py --segmented-only 2>&1 | tail -60 ) > gpurun_out/s30_racecheck.log
tail -4 gpurun_out/s30_racecheck.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:'k_split_decompress' -c 1 -f \
   -o gpurun_out/s30_split python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/s30_ncu2.log 2>&1
ls -la gpurun_out/s30_*.ncu-rep
