#!/bin/bash
# Round 2, GPU session 23 (8 GPUs): config 5 alone with NCCL point-to-point channel settings
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
NG=${NG:-8}
run() { echo "== $1"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $2 tools/bench_sharded.py --c5-mib 2048 2>/dev/null | tail -1 | python -c "
import sys,json
l=json.loads(sys.stdin.read()); print(round(l['MBps_with_nccl']), round(l['MBps_kernels_only']), round(l['scatter_gather_efficiency'],3), [(x['window'],round(x['with_nccl_ms'],1),round(x['kernels_only_ms'],1)) for x in l['classes']])"; }
run "X=1" 29601
run "NCCL_MIN_P2P_NCHANNELS=8" 29602
run "NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32" 29603
run "NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=64 NCCL_MAX_NCHANNELS=64" 29604
