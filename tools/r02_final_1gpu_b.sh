#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02
{
echo "== gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | cut -c1-200
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_${T}_1gpu.json 2> gpurun_out/bench_${T}_1gpu.err; tail -c 600 gpurun_out/bench_${T}_1gpu.json; tail -3 gpurun_out/bench_${T}_1gpu.err
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1; tail -c 150 gpurun_out/${T}_ncu_launch.log
} > gpurun_out/${T}_final_1gpu_b.log 2>&1
tail -c 3500 gpurun_out/${T}_final_1gpu_b.log
