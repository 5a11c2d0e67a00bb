#!/bin/bash
# Everything that was written without a GPU at the end of round 1, in ONE gpurun call (about 22 minutes of box time):
#   bash tools/class_c_variants.sh                      # here, in the build container (builds build_exp/smem*/)
#   gpurun --timeout 2100 -- 'bash tools/queued_gpu_check.sh r02a'
# 1. the two device paths that have never run: the corners mask and the coarse-grid averages (their tests sort last);
# 2. the round check (full GPU suite, smoke, default bench line, ncu launch list + full capture of class C) -- this also
#    times the SweepMem change of the class C sweep, which is in the default build but was never measured;
# 3. the same bench with the floes numbered along a Z-order curve (experiment: how much the gather-bound kernels gain);
# 4. the class C shared-memory variants against the default build (tools/convex_probe.sh).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02a}
{
echo "== new device paths and experiment switches"; timeout 900 python -m pytest tests/test_zz_fracture.py tests/test_zzz_eulerian_device.py tests/test_zzzz_experiments.py -q -m gpu 2>&1 | tail -15
} > gpurun_out/queued_$TAG.log 2>&1
bash tools/round_gpu_check.sh $TAG > /dev/null 2>&1
[ -x tools/sz_driver ] && timeout 600 tools/sz_driver 1000000 5 3 > gpurun_out/sz_driver_$TAG.json 2> gpurun_out/sz_driver_$TAG.err   # incl. corner mask / eulerian timings
timeout 600 python bench.py --opt convex_split=1 --no-cpu > gpurun_out/bench_${TAG}_split.json 2> gpurun_out/bench_${TAG}_split.err   # class C in two kernels, through the official bench
timeout 600 python bench.py --floe-order morton --no-cpu > gpurun_out/bench_${TAG}_morton.json 2> gpurun_out/bench_${TAG}_morton.err   # what a spatial numbering is worth
V=""; for d in build_exp/smem*/; do [ -f "$d/libsubzero_b200.so" ] && V="$V $(basename $d)"; done
bash tools/convex_probe.sh $V SZ_CONVEX_SPLIT=1 "SZ_LIB=$PWD/build_exp/smem63/libsubzero_b200.so SZ_CONVEX_SPLIT=1" > /dev/null 2>&1      # variants + class C split in two kernels (env form of the option)
cat gpurun_out/queued_$TAG.log; cat gpurun_out/sz_driver_$TAG.json 2>/dev/null; tail -c 700 gpurun_out/bench_${TAG}_morton.json 2>/dev/null; tail -c 700 gpurun_out/bench_${TAG}_split.json 2>/dev/null; tail -c 3000 gpurun_out/round_check_$TAG.log; grep -E "^==|1000000 2|passed|failed|rror" gpurun_out/convex_probe.log | tail -20
