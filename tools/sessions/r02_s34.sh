#!/bin/bash
# Round 2, GPU session 34: the tree as committed — smoke() and the full GPU suite
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/s34_smoke.log
( timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/s34_tests.log
tail -3 gpurun_out/s34_tests.log
