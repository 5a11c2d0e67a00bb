#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
SZ_SLAB_TIMING=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29744 bench.py --gpus $N --steps 20 --warmup 3 --verify 0 > gpurun_out/bench_r02i_${N}gpu_timing.json 2> gpurun_out/bench_r02i_${N}gpu_timing.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29745 bench.py --gpus $N --steps 20 --warmup 3 --verify 0 > gpurun_out/bench_r02i_${N}gpu.json 2> gpurun_out/bench_r02i_${N}gpu.err
python - <<'PY'
import json,sys,glob
for f in sorted(glob.glob('gpurun_out/bench_r02i_*gpu*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['config'].get('wall_ms_per_step'), d['config'].get('phase_ms_rank0'), d['config'].get('slab_stage_ms_rank0'), d['e2e']['ms_per_step'], d['roofline']['kernel_ms'])
    except Exception as e: print(f, 'ERR', e)
PY
grep -v "^W1\|OMP_NUM\|^\*" gpurun_out/bench_r02i_${N}gpu.err | tail -5
