#!/bin/bash
# Builds class C experiment variants next to the product library (build container, no GPU needed), for
#   gpurun --timeout 2100 -- 'bash tools/queued_gpu_check.sh r02a'      (or tools/convex_probe.sh <names> alone)
# which runs the class C parity tests and then times a 1M-floe step per variant (tools/scale_probe.py, SZ_LIB).
# Arguments: masks of SZ_C_SMEM_EDGES (a number N builds build_exp/smemN) or name:flags specs, e.g.
#   tools/class_c_variants.sh 63 48 16 "smem63_1024x1:-DSZ_C_SMEM_EDGES=63 -DSZ_C_TPB=1024 -DSZ_C_MINB=1"
#   smem63  all six 8-byte edge fields of the convex sweep in shared memory (96 KB per 512-thread CTA)
#   smem48  curx + dx only (32 KB per CTA)        smem16  curx only (16 KB per CTA)
# See SZ_C_SMEM_EDGES in subzero_b200/csrc/sz_convex.cuh.  None of these has been measured yet.
cd "$(dirname "$0")/.."
[ $# -eq 0 ] && set -- 63 48 16
for spec in "$@"; do
  case "$spec" in
    *:*) name=${spec%%:*}; flags=${spec#*:} ;;
    *)   name=smem$spec; flags="-DSZ_C_SMEM_EDGES=$spec" ;;
  esac
  d=build_exp/$name
  mkdir -p $d
  if SZ_BUILD_DIR=$PWD/$d SZ_EXTRA_NVCC="$flags" python -m subzero_b200.build > $d/build.log 2>&1; then echo "built $d/libsubzero_b200.so ($flags)"; else echo "FAILED $d (see $d/build.log)"; fi
done
