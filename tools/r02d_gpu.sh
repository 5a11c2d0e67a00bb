#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 3 --master-addr 127.0.0.1 --master-port 29611 tests/slab_worker.py 9000 7 migrate 6 > gpurun_out/r02d_migrate.log 2>&1
grep -v "^W1\|OMP_NUM" gpurun_out/r02d_migrate.log | tail -60
