#!/bin/bash
# Round 2: the mex gateways executed on a GPU under the MATLAB stand-in (tests/test_zzzzz_mex_gateway.py), the stand-alone C++ driver at HEAD
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
date
echo "== mex gateway tests"; timeout 300 python -m pytest tests/test_zzzzz_mex_gateway.py -x -q 2>&1 | tail -25
date
echo "== driver"; timeout 120 tools/sz_driver 1000000 5 3 > gpurun_out/sz_driver_r02h.json 2> gpurun_out/sz_driver_r02h.err; tail -c 900 gpurun_out/sz_driver_r02h.json; tail -3 gpurun_out/sz_driver_r02h.err
date
} > gpurun_out/r02_mex_gpu.log 2>&1
tail -c 4000 gpurun_out/r02_mex_gpu.log
