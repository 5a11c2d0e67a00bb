#!/bin/bash
# Round 2 multi-GPU evidence: NCCL parity tests and the bench line with --verify (20 coupled steps vs one GPU) at N GPUs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
{
echo "== nccl parity"; timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -k "nccl" 2>&1 | tail -4
echo "== bench $N gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['parity'], d['e2e']['ms_per_step'], d['config']['per_rank_ms_step_contact_narrow_pairs'], d['config']['phase_ms_rank0'])
PY
grep -v "^W1\|OMP_NUM\|^\*" gpurun_out/bench_r02_${N}gpu.err | tail -5
} > gpurun_out/r02_final_$N.log 2>&1
tail -c 4000 gpurun_out/r02_final_$N.log
