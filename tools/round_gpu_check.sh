#!/bin/bash
# Round check on the GPU box: full GPU suite, smoke, default bench, then the ncu evidence of the same bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r01c}
{
echo "== gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 2500 gpurun_out/bench_$TAG.json
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1; tail -c 300 gpurun_out/${TAG}_ncu_launch.log
echo "== ncu full (class C)"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:narrow_convex_kernel -s 4 -c 1 -f -o gpurun_out/${TAG}_narrow_C python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1; tail -c 300 gpurun_out/${TAG}_ncu_full.log
} > gpurun_out/round_check_$TAG.log 2>&1
tail -c 6000 gpurun_out/round_check_$TAG.log
