#!/bin/bash
# Round 2: the second workload (the reference's concave outlines tiled) at HEAD, both forms, with their own cpu_baseline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
date
echo "== real_shapes"; timeout 200 python bench.py --workload real_shapes --steps 10 > gpurun_out/bench_r02h_real_shapes.json 2> gpurun_out/bench_r02h_real_shapes.err; tail -c 1200 gpurun_out/bench_r02h_real_shapes.json; tail -3 gpurun_out/bench_r02h_real_shapes.err
date
echo "== real_shapes_raw"; timeout 200 python bench.py --workload real_shapes_raw --steps 5 > gpurun_out/bench_r02h_real_shapes_raw.json 2> gpurun_out/bench_r02h_real_shapes_raw.err; tail -c 1200 gpurun_out/bench_r02h_real_shapes_raw.json; tail -3 gpurun_out/bench_r02h_real_shapes_raw.err
date
} > gpurun_out/r02_real_shapes_gpu.log 2>&1
tail -c 4000 gpurun_out/r02_real_shapes_gpu.log
