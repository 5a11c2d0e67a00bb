#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "periodic_voronoi or shortcuts or class_c or uninflated or walls or real" 2>&1 | tail -4
echo "== default (CLS_G=1 + pruning)"; timeout 300 python tools/scale_probe.py 1000000
echo "== CLS_G=4"; SZ_LIB=$PWD/build_exp/cls4/libsubzero_b200.so timeout 300 python tools/scale_probe.py 1000000
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 -k regex:pair_classify --csv --log-file gpurun_out/r02h_launches.csv python tools/scale_probe.py 1000000 > gpurun_out/r02h_ncu.log 2>&1; grep pair_classify gpurun_out/r02h_launches.csv | tail -2 | cut -c1-60,400-
} > gpurun_out/r02h.log 2>&1
tail -c 3000 gpurun_out/r02h.log
