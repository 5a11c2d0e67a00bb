#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
{
for w in real_shapes real_shapes_raw; do
echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 3 > gpurun_out/bench_r02n_$w.json 2> gpurun_out/bench_r02n_$w.err; python -c "
import json;d=json.loads(open('gpurun_out/bench_r02n_$w.json').read().strip().splitlines()[-1]);print(d['config']['floes'],d['config']['pairs_per_step'],d['ms_per_step'],d['value'],d['config']['class_ms'],d['config']['class_pairs'],d['cpu_baseline']['value'],d['e2e']['ms_per_step'])"; tail -2 gpurun_out/bench_r02n_$w.err
done
} > gpurun_out/r02n.log 2>&1
cat gpurun_out/r02n.log
