"""Worker of tests/test_gpu_multi.py: one rank of a world-size-W slab run.  Rank 0 gathers every rank's results,
runs the same field on one GPU and compares floe by floe, row by row."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import subzero_b200 as sz  # noqa: E402
from subzero_b200 import slabs  # noqa: E402


def main():
    n_floes, seed, real = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3] == "real"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    local = int(os.environ.get("LOCAL_RANK", "0")) % ngpu
    backend = "nccl" if ngpu >= world else "gloo"
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    dev = torch.device("cuda", local)
    if real:
        import scenarios
        prm, Floe = scenarios.real_shape_field(n_floes, seed=seed)
        field = sz.floes_to_soa(Floe)
    else:
        prm, field = sz.voronoi_field(n_floes, seed=seed)
    field, starts = slabs.sort_by_slab(field, prm.Lx, world)
    mine = slabs.take_range(field, int(starts[rank]), int(starts[rank + 1]))
    comm = slabs.Comm(dist, rank, world, dev)
    st = slabs.SlabState.from_soa(mine, int(starts[rank]), field.n, dev)
    ctx = sz.ContactContext(local)
    step = slabs.SlabStep(prm, st, comm, ctx)
    one = sz.ContactContext(local) if rank == 0 else None
    base_x, base_y = field.x.copy(), field.y.copy()
    gids = np.arange(field.n)
    ok = True
    # (label, displacement amplitude in units of the halo skin, expected path)
    scenarios = [("plan", 0.0, "plan"), ("same", 0.0, "fast"), ("moved", 0.2, "fast"), ("moved-more", 0.45, "fast"), ("stale", 3.0, "plan"), ("after", 3.0, "fast")]
    for label, amp, expect in scenarios:
        # floes away from the periodic boundary move (so the image set stays the planned one unless amp is large);
        # the slab interfaces are in the interior, so halo entries do move
        inner = (np.abs(base_x) < prm.Lx - 1.5 * step.reach) & (np.abs(base_y) < prm.Ly - 1.5 * step.reach) if amp < 1 else np.ones(field.n, bool)
        dx = amp * step.skin * np.sin(0.37 * gids + 1.0) * inner
        dy = amp * step.skin * np.cos(0.91 * gids + 2.0) * 0.999 * inner
        field.x[:], field.y[:] = base_x + dx, base_y + dy
        a, b = int(starts[rank]), int(starts[rank + 1])
        st.x, st.y = torch.as_tensor(field.x[a:b]).to(dev), torch.as_tensor(field.y[a:b]).to(dev)
        before = (step.plans, step.fast_steps)
        s = step.run()
        took = "plan" if step.plans > before[0] else "fast"
        out, row_off, rows = step.results()
        stats = np.array([s.n_pairs_owned, s.n_pairs_force, s.collision_count, s.n_pairs, step.local.halo_sent], dtype=np.float64)
        gathered = [None] * world
        dist.gather_object({"out": out, "row_off": row_off, "rows": rows, "stats": stats, "took": took}, gathered if rank == 0 else None, dst=0)
        if rank == 0:
            good = True
            s1 = one.step(prm, field, allow_pair_errors=True)
            o1 = one.floe_outputs()
            off1, rows1 = one.rows()
            for k in o1:
                got = np.concatenate([g["out"][k] for g in gathered])
                if not np.array_equal(got, o1[k], equal_nan=True):
                    good = False
                    print("MISMATCH per-floe", label, k, int((got != o1[k]).sum()))
            got_rows = np.concatenate([g["rows"] for g in gathered])
            cnt = np.concatenate([np.diff(g["row_off"]) for g in gathered])
            if not np.array_equal(cnt, np.diff(off1[:field.n + 1])):
                good = False
                print("MISMATCH row counts", label)
            elif not np.array_equal(got_rows, rows1[:off1[field.n]], equal_nan=True):
                good = False
                print("MISMATCH rows", label, int((got_rows != rows1[:off1[field.n]]).sum()))
            tot = np.sum([g["stats"] for g in gathered], 0)
            if (int(tot[0]), int(tot[1]), tot[2]) != (s1.n_pairs, s1.n_pairs_force, s1.collision_count):
                good = False
                print("MISMATCH totals", label, tot, s1.n_pairs, s1.n_pairs_force, s1.collision_count)
            paths = {g["took"] for g in gathered}
            if paths != {expect}:
                good = False
                print("UNEXPECTED path", label, paths, "expected", expect)
            ok = ok and good
            print("STEP %s %s path=%s pairs=%d rows=%d duplicated_pairs=%.3f halo_entries=%d" % (label, "OK" if good else "FAIL", took, s1.n_pairs, off1[field.n],
                                                                                            tot[3] / max(1, s1.n_pairs) - 1, int(tot[4])), flush=True)
    if rank == 0:
        print("RESULT %s world=%d backend=%s floes=%d kill_events=%d plans=%d fast_steps=%d" % (
            "OK" if ok else "FAIL", world, backend, field.n, int((o1["kill"] > 0).sum()), step.plans, step.fast_steps), flush=True)
        one.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
