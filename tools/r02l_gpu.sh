#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== default"; timeout 300 python tools/scale_probe.py 1000000 | tail -1
for v in c512x3 c512x4 c640x2; do
  echo "== $v"; SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 300 python tools/scale_probe.py 1000000 | tail -1
done
} > gpurun_out/r02l.log 2>&1
cat gpurun_out/r02l.log
