#!/bin/bash
# Round 2, second GPU call: the new device paths (slab list built on the device, pair searches, conservation replay, Nb in the integrator, drop-in mirror)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
echo "== slab world 1"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "world1" 2>&1 | tail -15
echo "== slab world 2/3"; timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -k "two_slabs_voronoi or three_slabs or walls or real_shapes" 2>&1 | tail -25
echo "== new tests"; timeout 900 python -m pytest tests/test_pair_searches.py tests/test_conservation.py tests/test_trajectory.py tests/test_gpu_parity.py -q -m gpu 2>&1 | tail -15
} > gpurun_out/r02b.log 2>&1
tail -c 6000 gpurun_out/r02b.log
