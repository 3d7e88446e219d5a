#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( timeout 2400 python -m pytest tests/test_gpu_batch.py tests/test_gpu_capi.py -x -q 2>&1 | tail -25 ) > gpurun_out/t_all.log
tail -5 gpurun_out/t_all.log
timeout 300 python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
from tamp_b200 import batch
x = batch.synth(0, 0, 1<<18, 1024)
for lazy in (False, True):
    for mode in (0, 1):
        batch.set_kernel_mode(mode)
        ts=[]
        for it in range(4):
            a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
            a.record(); r = batch.compress_batch(x, window=10, extended=False, lazy_matching=lazy); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print("lazy",lazy,"mode",mode,"ms",round(min(ts),2),"ratio",round(r.sizes.double().sum().item()/(x.numel()),4))
PY
