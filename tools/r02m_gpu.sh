#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== full gpu suite (poisoned fresh buffers)"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== bench real_shapes"; timeout 900 python bench.py --workload real_shapes --steps 5 > gpurun_out/bench_r02m_real.json 2> gpurun_out/bench_r02m_real.err; tail -c 1800 gpurun_out/bench_r02m_real.json; tail -3 gpurun_out/bench_r02m_real.err
echo "== bench real_shapes_raw"; timeout 900 python bench.py --workload real_shapes_raw --steps 3 > gpurun_out/bench_r02m_raw.json 2> gpurun_out/bench_r02m_raw.err; tail -c 1800 gpurun_out/bench_r02m_raw.json; tail -3 gpurun_out/bench_r02m_raw.err
} > gpurun_out/r02m.log 2>&1
tail -c 6000 gpurun_out/r02m.log
