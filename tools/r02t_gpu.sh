#!/bin/bash
# one ncu --set full capture (with source) of the class C launch at 1M floes, after a plain timing run of the same program
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== default"; timeout 200 python tools/scale_probe.py 1000000
echo "== ncu class C"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:narrow_convex_kernel -s 1 -c 1 -f -o gpurun_out/r02t_C python tools/scale_probe.py 1000000 2>&1 | tail -5
} > gpurun_out/r02t.log 2>&1
grep -E "^==|1000000 2|rror|==PROF==" gpurun_out/r02t.log | cut -c1-600 | tail -12
