#!/bin/bash
# What a round's GPU check consists of, as one gpurun call:  tools/gpu.sh --timeout 3000 -- 'bash tools/gpu_session.sh'
#   1. smoke + the GPU parity suite, 2. the bench line (all configs) and the reference arm, 3. the launch list and the two
#   ncu --set full captures that profiles/ summarises (profiles/README.md says how they are read), 4. the sanitizers.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash tools/final_check.sh
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-configs > gpurun_out/launches_bench.log 2>&1
python profiles/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.csv
for k in k_walk_compress k_split_decompress; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -f -o gpurun_out/full_$k \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-other-format --no-extra-configs > gpurun_out/ncu_$k.log 2>&1
done
( timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize.py 2>&1 | tail -40 ) > gpurun_out/memcheck.log
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/race_check.py 2>&1 | tail -60 ) > gpurun_out/racecheck.log
tail -n 3 gpurun_out/memcheck.log gpurun_out/racecheck.log
