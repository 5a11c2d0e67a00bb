#!/bin/bash
# timing-only experiments (some variants give wrong results on purpose): where class C's wall time goes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== new"; timeout 200 python tools/scale_probe.py 1000000
for v in "$@"; do echo "== variant $v"; SZ_LIB=$PWD/build_exp/$v/libsubzero_b200.so timeout 200 python tools/scale_probe.py 1000000; done
} > gpurun_out/r02s.log 2>&1
grep -E "^==|1000000 2|rror" gpurun_out/r02s.log | cut -c1-900 | tail -40
