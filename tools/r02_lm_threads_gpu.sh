#!/bin/bash
# persistent threads per SM of the classes M / L on the raw concave workload (env switches SZ_M_TPSM / SZ_L_TPSM)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for cfg in "512 128" "512 160" "768 160" "768 192"; do
  set -- $cfg
  echo "== M_TPSM $1 L_TPSM $2"
  SZ_M_TPSM=$1 SZ_L_TPSM=$2 timeout 100 python bench.py --workload real_shapes_raw --steps 3 --no-cpu 2>&1 | python3 -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1]); c = d['config']
print('ms_per_step %.1f' % d['ms_per_step'], {k: round(v, 1) for k, v in c['class_ms'].items()}, c['class_pairs'])"
done 2>&1 | tee gpurun_out/r02_lm_threads.log
