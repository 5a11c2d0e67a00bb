#!/bin/bash
# Builds the class C experiment variants next to the product library (build container, no GPU needed), for
#   gpurun --timeout 1500 -- 'bash tools/convex_probe.sh smem63 smem48 smem16'
# which runs the parity subset and then times a 1M-floe step per variant (tools/scale_probe.py, SZ_LIB).
#   smem63  all six 8-byte edge fields of the convex sweep in shared memory (96 KB per CTA)
#   smem48  curx + dx only (32 KB per CTA)        smem16  curx only (16 KB per CTA)
# See SZ_C_SMEM_EDGES in subzero_b200/csrc/sz_convex.cuh.  None of these has been measured yet.
set -e
cd "$(dirname "$0")/.."
for m in ${@:-63 48 16}; do
  d=build_exp/smem$m
  mkdir -p $d
  SZ_BUILD_DIR=$PWD/$d SZ_EXTRA_NVCC="-DSZ_C_SMEM_EDGES=$m" python -m subzero_b200.build > $d/build.log 2>&1 && echo "built $d/libsubzero_b200.so"
done
