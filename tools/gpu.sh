#!/bin/bash
# Build here (nvcc cross-compiles), then run the given command on the GPU box: the .so that travels is never stale.
#   tools/gpu.sh [--timeout S] [--gpus N] -- '<command>'
cd "$(dirname "$0")/.."
make -s -C tamp_b200/csrc -j8 2>&1 | grep -E "error|Error" && exit 1
exec /usr/local/graft/bin/gpurun "$@"
