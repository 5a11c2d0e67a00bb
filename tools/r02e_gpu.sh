#!/bin/bash
# N GPUs: NCCL parity test + bench with --verify
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
{
echo "== nccl parity"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "nccl" 2>&1 | tail -8
echo "== bench $N gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02e_${N}gpu.json 2> gpurun_out/bench_r02e_${N}gpu.err
tail -c 3000 gpurun_out/bench_r02e_${N}gpu.json; grep -v "^W1\|OMP_NUM" gpurun_out/bench_r02e_${N}gpu.err | tail -15
} > gpurun_out/r02e_$N.log 2>&1
tail -c 6000 gpurun_out/r02e_$N.log
