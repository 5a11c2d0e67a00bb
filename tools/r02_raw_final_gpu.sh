#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 80 python bench.py --workload real_shapes_raw --steps 4 > gpurun_out/bench_r02j_real_shapes_raw.json 2> gpurun_out/bench_r02j_real_shapes_raw.err; cut -c1-900 gpurun_out/bench_r02j_real_shapes_raw.json; tail -c 400 gpurun_out/bench_r02j_real_shapes_raw.json; tail -3 gpurun_out/bench_r02j_real_shapes_raw.err
