#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "== bench 1 gpu"; timeout 900 python bench.py --steps 20 --no-cpu > gpurun_out/bench_r02f.json 2> gpurun_out/bench_r02f.err; tail -c 1800 gpurun_out/bench_r02f.json; tail -5 gpurun_out/bench_r02f.err
echo "== scale probe"; timeout 300 python tools/scale_probe.py 10000 125000 1000000
} > gpurun_out/r02f.log 2>&1
tail -c 7000 gpurun_out/r02f.log
