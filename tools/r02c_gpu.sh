#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== slab tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25
echo "== bench 1 gpu"; timeout 900 python bench.py --steps 10 --no-cpu > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err; tail -c 1500 gpurun_out/bench_r02c.json; tail -5 gpurun_out/bench_r02c.err
} > gpurun_out/r02c.log 2>&1
tail -c 5000 gpurun_out/r02c.log
