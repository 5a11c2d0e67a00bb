"""multi-GPU probe: breakdown of the slab step (run under torchrun)"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, torch.distributed as dist
import subzero_b200 as sz
from subzero_b200 import slabs
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
job = slabs.SlabJob(n, 0, rank, world, local, dist)
step = job.step
def sync():
    torch.cuda.synchronize()
step.run()
for it in range(6):
    dist.barrier(); sync(); t0 = time.perf_counter()
    ok, dyn = slabs.refresh_local_state(step.st, step.prm.Lx, step.prm.Ly, True, step.local.plan, step.comm)
    sync(); t1 = time.perf_counter()
    s = step.run()
    sync(); t2 = time.perf_counter()
    ph = job.ctx.phase_ms()
    if it >= 3:
        print("rank %d it %d refresh %.3f ms | run %.3f ms | lib step %.3f (narrow %.3f broad %.3f asm %.3f ghosts %.3f) | local n %d owned pairs %d all pairs %d" % (
            rank, it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), s.ms_device, ph["narrow"], ph["broad"], ph["assembly"], ph["ghosts"], s.n, s.n_pairs_owned, s.n_pairs), flush=True)
dist.barrier(); dist.destroy_process_group()
