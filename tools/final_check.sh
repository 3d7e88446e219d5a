#!/bin/bash
# What the driver runs at round end, as one gpurun call: GPU parity suite, smoke(), one bench line.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/t_all.log
tail -3 gpurun_out/t_all.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_line.json | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(round(l['value']), round(l['ms_per_step'],2), round(l['e2e']['value']), l['gpu_launches'], l['cpu_baseline']['value'], l['other_format'])"
