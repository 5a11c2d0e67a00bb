#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zy_known_answers_device.py -x -q -m gpu 2>&1 | tail -6
echo "== scale probe"; timeout 300 python tools/scale_probe.py 125000 1000000
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g_launches.csv python tools/scale_probe.py 1000000 > gpurun_out/r02g_ncu.log 2>&1; tail -2 gpurun_out/r02g_ncu.log
} > gpurun_out/r02g.log 2>&1
tail -c 3000 gpurun_out/r02g.log
