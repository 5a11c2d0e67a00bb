#!/bin/bash
# Round 2 final single-GPU evidence: bench line (CPU arm on the benchmark configuration), reference arm sanity, ncu launch list and
# --set full captures of the five heaviest kernels at HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=r02
{
echo "== bench"; timeout 1200 python bench.py > gpurun_out/bench_${T}_1gpu.json 2> gpurun_out/bench_${T}_1gpu.err; tail -c 1200 gpurun_out/bench_${T}_1gpu.json; tail -3 gpurun_out/bench_${T}_1gpu.err
echo "== reference arm (2 steps)"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${T}_reference.json 2> gpurun_out/bench_${T}_reference.err; tail -c 900 gpurun_out/bench_${T}_reference.json; tail -3 gpurun_out/bench_${T}_reference.err
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_launch.log 2>&1; tail -c 200 gpurun_out/${T}_ncu_launch.log
echo "== ncu full (top kernels)"; timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:narrow_convex_kernel|broad_kernel|pair_classify_kernel|assemble_kernel' -s 15 -c 5 -f -o gpurun_out/${T}_top python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${T}_ncu_full.log 2>&1; tail -c 200 gpurun_out/${T}_ncu_full.log
echo "== driver"; [ -x tools/sz_driver ] && timeout 600 tools/sz_driver 1000000 5 3 > gpurun_out/sz_driver_${T}.json 2> gpurun_out/sz_driver_${T}.err; tail -c 600 gpurun_out/sz_driver_${T}.json
} > gpurun_out/${T}_final_1gpu.log 2>&1
tail -c 5000 gpurun_out/${T}_final_1gpu.log
