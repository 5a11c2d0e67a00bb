#!/bin/bash
# Round 2, last GPU call: the evidence at HEAD after the final narrow / broad phase changes -- ncu --set full of the five
# heaviest kernels (refreshes profiles/narrow_traffic.json's stamp), ncu launch list, the whole GPU suite, the default bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02h
{
date
echo "== ncu full (top kernels)"; timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:narrow_convex_kernel|broad_kernel|pair_classify_kernel|assemble_kernel' -s 15 -c 5 -f -o gpurun_out/${T}_top python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_full.log 2>&1; tail -c 200 gpurun_out/${T}_ncu_full.log
date
echo "== ncu launch list"; timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1; tail -c 200 gpurun_out/${T}_ncu_launch.log
date
echo "== gpu suite"; timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
date
echo "== bench"; timeout 400 python bench.py > gpurun_out/bench_${T}_1gpu.json 2> gpurun_out/bench_${T}_1gpu.err; tail -c 1500 gpurun_out/bench_${T}_1gpu.json; tail -3 gpurun_out/bench_${T}_1gpu.err
date
} > gpurun_out/${T}_head_1gpu.log 2>&1
tail -c 5000 gpurun_out/${T}_head_1gpu.log
