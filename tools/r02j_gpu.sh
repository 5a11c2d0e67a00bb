#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
echo "== scale probe"; timeout 300 python tools/scale_probe.py 200 10000 125000 1000000
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 -k regex:"pair_classify|broad_kernel|assemble|convex" --csv --log-file gpurun_out/r02j_launches.csv python tools/scale_probe.py 1000000 > gpurun_out/r02j_ncu.log 2>&1; grep -E "pair_classify|broad_kernel|assemble|convex" gpurun_out/r02j_launches.csv | tail -6 | awk -F'","' '{print substr($5,1,40), $NF}'
} > gpurun_out/r02j.log 2>&1
tail -c 3500 gpurun_out/r02j.log
