#!/bin/bash
# Round 2 multi-GPU evidence: the bench line with --verify (20 coupled steps vs one GPU) at N GPUs; optionally the NCCL parity tests first
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
{
if [ "$2" = "tests" ]; then echo "== nccl parity"; timeout 400 python -m pytest tests/test_gpu_multi.py -x -q -k "nccl" 2>&1 | tail -5; fi
echo "== bench $N gpus"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29733 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_r02_${N}gpu.json 2> gpurun_out/bench_r02_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r02_${N}gpu.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['parity']['ok'], d['e2e']['ms_per_step'], d['config']['per_rank_ms_step_contact_narrow_pairs'], d['config']['cuda_graph'], d['roofline']['kernel_ms'])
PY
grep -v "^W1\|OMP_NUM\|^\*" gpurun_out/bench_r02_${N}gpu.err | grep -i "error" | tail -5
} > gpurun_out/r02_final_$N.log 2>&1
tail -c 3000 gpurun_out/r02_final_$N.log
