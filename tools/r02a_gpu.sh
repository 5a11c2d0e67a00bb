#!/bin/bash
# Round 2, first GPU call: times the experiments round 1 left unmeasured (class C variants, split, Morton numbering)
# and takes ncu --set full captures of the four heaviest kernels at HEAD.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r02a}
V=""; for d in build_exp/smem*/; do [ -f "$d/libsubzero_b200.so" ] && V="$V $(basename $d)"; done
bash tools/convex_probe.sh $V SZ_CONVEX_SPLIT=1 > /dev/null 2>&1
cp gpurun_out/convex_probe.log gpurun_out/convex_probe_$TAG.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
timeout 600 python bench.py --floe-order morton --no-cpu > gpurun_out/bench_${TAG}_morton.json 2> gpurun_out/bench_${TAG}_morton.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:narrow_convex_kernel|broad_kernel|pair_classify_kernel|assemble_kernel' -s 18 -c 6 -f -o gpurun_out/${TAG}_top4 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1
grep -E "^==|1000000 2|passed|failed|rror" gpurun_out/convex_probe_$TAG.log | tail -30
tail -c 1500 gpurun_out/bench_${TAG}.json; tail -c 1500 gpurun_out/bench_${TAG}_morton.json; tail -c 300 gpurun_out/${TAG}_ncu_full.log
