#!/bin/bash
# Final state of the round on one B200: smoke, full GPU suite, default bench line (all configs), reference arm.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/final_smoke.log
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/final_tests.log
tail -3 gpurun_out/final_tests.log
timeout 900 python bench.py > gpurun_out/final_bench.log 2>&1; tail -1 gpurun_out/final_bench.log > gpurun_out/final_bench_line.json
python - <<'P'
import json
l=json.loads(open('gpurun_out/final_bench_line.json').read()); r=l['roofline']
print('steps',l['steps'],'compress_ms',round(r['kernel_ms'],3),'decompress_ms',round(r['decompress']['kernel_ms'],3),'value',round(l['value']),'frac',round(r['frac'],4),'e2e',round(l['e2e']['ms_per_step'],2),round(l['e2e']['value']),l['gpu_launches'],(l.get('cpu_baseline') or {}).get('parity'))
print('other',l['other_format']); print('clocks',l['clocks']); print('traffic',r.get('traffic'),r['decompress'].get('traffic'))
print({k:{kk:vv for kk,vv in v.items() if kk in('compress_MBps','decompress_MBps','MBps','parity')} for k,v in l['configs'].items()})
P
timeout 900 python bench.py --impl reference 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/final_ref.log
