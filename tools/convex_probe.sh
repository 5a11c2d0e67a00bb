#!/bin/bash
# GPU probe: parity subset, then step time at 1M floes per experiment switch / build variant
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity subset"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "periodic_voronoi or shortcuts or uninflated or resident or dead or small_periodic or boundary_floes" 2>&1 | tail -5
echo "== default"; timeout 300 python tools/scale_probe.py 1000000 200000
for v in "$@"; do
  echo "== variant $v"
  case "$v" in
    *=*) env $v timeout 300 python tools/scale_probe.py 1000000 ;;
    *) SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "class_c_equals or periodic_voronoi or shortcuts" 2>&1 | tail -2
       SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 300 python tools/scale_probe.py 1000000 ;;
  esac
done
} > gpurun_out/convex_probe.log 2>&1
grep -E "^==|1000000 2|200000 2|passed|failed|rror" gpurun_out/convex_probe.log | tail -40
