#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/slab_worker.py 30000 8 graph 6 > gpurun_out/r02o_graph.log 2>&1
grep -v "^W1\|OMP_NUM\|^\*" gpurun_out/r02o_graph.log | grep -v "^$" | head -50
