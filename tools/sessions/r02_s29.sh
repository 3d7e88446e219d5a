#!/bin/bash
# Round 2, GPU session 29: compute-sanitizer memcheck + racecheck on the kernels as shipped (half-warp walkers, append
# mode / segment headers, the segmented calls)
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py 2>&1 | tail -40 ) > gpurun_out/s29_memcheck.log
tail -3 gpurun_out/s29_memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_check.py 2>&1 | tail -40 ) > gpurun_out/s29_racecheck.log
tail -4 gpurun_out/s29_racecheck.log
