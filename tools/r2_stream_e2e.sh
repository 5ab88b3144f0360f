#!/bin/bash
set -u
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/stream_e2e.py "$@" 2>&1 | grep -E "world|Error|error" ; }
{
run --call 16384 --threads 2
FDNN_SYNC=spin run --call 16384 --threads 2
run --call 16384 --threads 3
run --call 4096 --threads 3
run --call 2048 --threads 4
run --call 512 --threads 3 --frames 100000
} > gpurun_out/r2_stream_e2e_${N}gpu.log 2>&1
cat gpurun_out/r2_stream_e2e_${N}gpu.log
