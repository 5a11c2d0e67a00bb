#!/bin/bash
# GPU probe of the class C fast path: parity subset, step time per build variant, then the full GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "periodic_voronoi or shortcuts or uninflated or resident" 2>&1 | tail -5
echo "== default (256x2)"; timeout 300 python tools/scale_probe.py 1000000 200000
echo "== no fast path"; SZ_NO_CONVEX_FAST=1 timeout 300 python tools/scale_probe.py 1000000
for v in m4 t128m6 t128m4; do echo "== variant $v"; SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 300 python tools/scale_probe.py 1000000; done
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
} > gpurun_out/convex_probe.log 2>&1
tail -40 gpurun_out/convex_probe.log
