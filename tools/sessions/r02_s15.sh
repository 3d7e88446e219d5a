#!/bin/bash
# Round 2, GPU session 15: final state: smoke, full GPU suite, default bench (all configs), reference arm
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/s15_smoke.log
( timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/s15_tests.log
tail -3 gpurun_out/s15_tests.log
timeout 900 python bench.py > gpurun_out/s15_bench.log 2>&1; tail -1 gpurun_out/s15_bench.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); r=l['roofline']; print('steps',l['steps'],'compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']), 'e2e', l['e2e']['ms_per_step'], l['e2e'].get('ms_per_step_two_host_threads'), l['gpu_launches'], (l.get('cpu_baseline') or {}).get('parity'), l['other_format']); print(json.dumps(l.get('configs'))[:1800]); print(l['clocks'], r.get('traffic'))" || tail -20 gpurun_out/s15_bench.log
timeout 900 python bench.py --impl reference > gpurun_out/s15_ref.log 2>&1; tail -1 gpurun_out/s15_ref.log | cut -c1-600
