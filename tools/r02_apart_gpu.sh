#!/bin/bash
# Round 2: the edge-by-edge "apart" certificate of the classifier on the GPU -- whole GPU suite (incl. the on/off parity test),
# the headline step through the C++ driver (must not move), the two concave workloads with their cpu_baseline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
date
echo "== gpu suite"; timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
date
echo "== driver (1M convex floes)"; timeout 100 tools/sz_driver 1000000 6 3 2>&1 | tail -1 | cut -c1-420
echo "== real_shapes"; timeout 150 python bench.py --workload real_shapes --steps 10 > gpurun_out/bench_r02i_real_shapes.json 2> gpurun_out/bench_r02i_real_shapes.err; cut -c1-1100 gpurun_out/bench_r02i_real_shapes.json; tail -3 gpurun_out/bench_r02i_real_shapes.err
date
echo "== real_shapes_raw"; timeout 150 python bench.py --workload real_shapes_raw --steps 5 > gpurun_out/bench_r02i_real_shapes_raw.json 2> gpurun_out/bench_r02i_real_shapes_raw.err; cut -c1-1100 gpurun_out/bench_r02i_real_shapes_raw.json; tail -3 gpurun_out/bench_r02i_real_shapes_raw.err
date
} > gpurun_out/r02_apart_gpu.log 2>&1
tail -c 5000 gpurun_out/r02_apart_gpu.log
