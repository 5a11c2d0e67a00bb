#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
SZ_DEBUG_POISON=1 SZ_DEBUG_SYNC=1 timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k speculated 2>&1 | grep -E "\[sz\]|Error|passed|failed" | head -20 > gpurun_out/r02k.log
cat gpurun_out/r02k.log
