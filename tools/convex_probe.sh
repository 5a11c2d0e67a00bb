#!/bin/bash
# GPU probe of the class C fast path: parity subset, step time per build variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export SZ_CONVEX_FAST=1
{
echo "== parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "periodic_voronoi or shortcuts or uninflated or resident" 2>&1 | tail -5
echo "== default"; timeout 300 python tools/scale_probe.py 1000000 200000
for v in "$@"; do echo "== variant $v"; SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 300 python tools/scale_probe.py 1000000; done
} > gpurun_out/convex_probe.log 2>&1
grep -v " 0 pairs" gpurun_out/convex_probe.log | tail -40
