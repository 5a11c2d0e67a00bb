#!/bin/bash
# A/B of kernel changes: parity subset (incl. the 1M-floe test) and the 1M-floe step time, for the default library and for
# every build_exp/<variant> named on the command line (NAME=VALUE arguments are environment switches: timing only).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
K="periodic_voronoi or shortcuts or class_c_equals or many_candidate or small_periodic or boundary_floes or real_concave_shapes or simplified_concave or non_periodic or conservation or full_benchmark_size"
{
echo "== parity subset (default)"; timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | tail -4
echo "== default"; timeout 200 python tools/scale_probe.py 1000000 125000
for v in "$@"; do
  echo "== variant $v"
  case "$v" in
    *=*) env $v timeout 200 python tools/scale_probe.py 1000000 ;;
    *) SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -k "$K" 2>&1 | tail -4
       SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 200 python tools/scale_probe.py 1000000 125000 ;;
  esac
done
} > gpurun_out/r02r.log 2>&1
grep -E "^==|1000000 2|125000 2|passed|failed|rror" gpurun_out/r02r.log | cut -c1-900 | tail -40
