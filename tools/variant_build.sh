#!/bin/bash
# Experiment builds that differ from the product library in ONE translation unit: tools/variant_build.sh NAME TU.cu FLAGS...
# -> build_exp/NAME/libsubzero_b200.so (the other objects are taken from subzero_b200/_lib).  Timed on the GPU box with
#    LD_LIBRARY_PATH=build_exp/NAME tools/sz_driver 1000000 5 3      (the driver's RUNPATH yields to LD_LIBRARY_PATH)
cd "$(dirname "$0")/.."
name=$1; tu=$2; shift 2
d=build_exp/$name; mkdir -p $d
objs=""
for o in subzero_b200/_lib/*.o; do b=$(basename $o); [ "$b" = "$tu.o" ] && continue; objs="$objs $o"; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-O2 "$@" -c subzero_b200/csrc/$tu -o $d/$tu.o > $d/build.log 2>&1 \
 && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libsubzero_b200.so $objs $d/$tu.o -lpthread >> $d/build.log 2>&1 \
 && echo "built $d ($tu $*)" || echo "FAILED $d"
