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
for it in range(5):
    dist.barrier(); sync(); t0 = time.perf_counter()
    L = slabs.build_local_list(step.st, step.prm.Lx, step.prm.Ly, True, step.reach, step.comm)
    sync(); t1 = time.perf_counter()
    s = step.run()
    sync(); t2 = time.perf_counter()
    ph = job.ctx.phase_ms()
    if it >= 2:
        print("rank %d it %d list %.2f ms | run(total incl. list) %.2f ms | lib step %.2f (narrow %.2f broad %.2f asm %.2f) | local n %d owned pairs %d all pairs %d halo sent %d (%.1f MB)" % (
            rank, it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), s.ms_device, ph["narrow"], ph["broad"], ph["assembly"], s.n, s.n_pairs_owned, s.n_pairs, L.halo_sent, L.halo_bytes / 1e6), flush=True)
dist.barrier(); dist.destroy_process_group()
