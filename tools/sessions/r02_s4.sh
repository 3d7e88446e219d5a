#!/bin/bash
# Round 2, GPU session 4 (2 GPUs): bench.py with the extra configs at N=1 (small sizes first), then N=2 incl. config 5.
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 900 python bench.py --steps 3 --warmup 3 --c3-streams 4096 --c4-frames 262144 > gpurun_out/s4_bench1_small.log 2>&1; tail -1 gpurun_out/s4_bench1_small.log | cut -c1-300; tail -1 gpurun_out/s4_bench1_small.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(json.dumps(l['configs'])[:3000]); print(json.dumps(l['e2e']))"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --c4-frames 262144 --c5-mib 256 > gpurun_out/s4_bench2_small.log 2>&1; tail -1 gpurun_out/s4_bench2_small.log | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(round(l['value']), l['e2e']); print(json.dumps(l['configs'])[:4000])" || tail -30 gpurun_out/s4_bench2_small.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/e2e_probe.py 2>&1 | tail -1 | tee gpurun_out/s4_probe2.log
