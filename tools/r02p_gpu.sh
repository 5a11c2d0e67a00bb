#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SZ_DEBUG_POISON=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29613 tests/slab_worker.py 6000 2 voronoi 6 > gpurun_out/r02p.log 2>&1
grep -v "^W1\|OMP_NUM\|^\*" gpurun_out/r02p.log | grep -v "^$" | head -40
