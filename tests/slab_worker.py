"""Worker of tests/test_gpu_multi.py: one rank of a world-size-W run of the device-built slab step (subzero_b200.slabs.DeviceSlab)
over a COUPLED time loop -- contact step + integrator (nonzero ksi, thinning), floes crossing the periodic boundary and the
slab edges.  Rank 0 runs the same loop on one GPU and compares every step bit for bit: per-floe outputs, contact rows, and the
integrated state (positions, velocities, thickness, heading, rotated outlines)."""
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

RHO_ICE = 920.0


def main():
    n_floes, seed, kind, steps = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4])
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
    bnd = None
    if kind == "real":
        import scenarios
        prm, Floe = scenarios.real_shape_field(n_floes, seed=seed)
        field = sz.floes_to_soa(Floe)
    else:
        prm, field = sz.voronoi_field(n_floes, seed=seed)
        if kind == "walls":
            import scenarios
            prm.periodic = 0
            c2, fb = scenarios.domain(prm.Lx, prm.Ly)
            bnd = sz.Boundary(fb["c"][0], fb["c"][1], c2[0], c2[1], fb["area"], fb["h"])
    migrate = kind == "migrate"
    use_graph = kind == "graph"           # the whole step replayed from one CUDA graph (NCCL only)
    fast = kind != "real"
    if fast:
        field.u[:] *= 100.0      # tens of metres per step: floes cross the periodic boundary and the slab edges within the run
        field.v[:] *= 100.0
    prm.dt = 10.0
    field, starts = slabs.sort_by_slab(field, prm.Lx, world)
    a, b = int(starts[rank]), int(starts[rank + 1])
    mine = slabs.take_range(field, a, b)
    gid = np.arange(a + 1, b + 1, dtype=np.int32)
    comm = slabs.Comm(dist, rank, world, dev)
    ctx = sz.ContactContext(local)
    slab = slabs.DeviceSlab(prm, mine, gid, field.n, comm, ctx, bnd=bnd)
    if use_graph:
        slab.enable_graph(True)
    nz, hfo = 3, 2e-4
    L = prm.Lx
    bounds = (-1.5 * L, 1.5 * L, -1.5 * L, 1.5 * L)
    mass = field.area * field.h * RHO_ICE
    inertia = mass * field.rmax ** 2 / 4
    rng = np.random.default_rng(seed)
    torque_oa = rng.normal(0, 3e-3, field.n)           # keeps the floes spinning (ksi is clamped to 1e-5 rad/s, :214)
    slab.trajectory_init(mass[a:b], inertia[a:b], nz=nz, dXi_p=field.u[a:b], dYi_p=field.v[a:b], torqueOA=torque_oa[a:b])
    one = None
    if rank == 0:
        one = sz.ContactContext(local)
        one.upload(prm, field, bnd)
        one.trajectory_init(mass, inertia, nz=nz, dXi_p=field.u, dYi_p=field.v, torqueOA=torque_oa)
    ok = True
    moved = 0.0
    n_migrated = [0]
    for it in range(steps):
        s = slab.run(allow_pair_errors=True)
        if it == 0:
            # the list the kernels built against the torch reference implementation of the same logic (subzero_b200.slabs.build_local_list)
            st = slabs.SlabState.from_soa(mine, a, field.n, dev)
            reach = 2.0 * float(comm.all_gather(st.rmax.max().reshape(1)).max())
            Lt = slabs.build_local_list(st, prm.Lx, prm.Ly, bool(prm.periodic), reach, comm)
            Ld = slab.local_list()
            same = (np.array_equal(Ld["gid"], Lt.gid.cpu().numpy() + 1) and np.array_equal(Ld["floe_num"], Lt.floe_num.cpu().numpy()) and np.array_equal(Ld["owned"], Lt.owned.cpu().numpy())
                    and np.array_equal(Ld["x"], Lt.x.cpu().numpy(), equal_nan=True) and np.array_equal(Ld["y"], Lt.y.cpu().numpy(), equal_nan=True))
            flags = [None] * world
            dist.all_gather_object(flags, bool(same))
            if not all(flags):
                ok = False
                if rank == 0:
                    print("MISMATCH device-built list vs torch reference list", flags)
        out = slab.outputs()
        row_off, rows = slab.rows()
        n_sacked = slab.trajectory_step(prm.dt, hfo, *bounds)
        state = ctx.trajectory_state(nverts=slab.owned.vx.shape[0])
        stats = np.array([s.n_pairs_owned, s.n_pairs_force, s.collision_count, s.n_pairs, n_sacked, slab.plans], dtype=np.float64)
        gathered = [None] * world
        dist.gather_object({"out": out, "row_off": row_off, "rows": rows, "stats": stats, "state": state, "gid": slab.gid.copy(), "nv": np.diff(slab.owned.voff)},
                           gathered if rank == 0 else None, dst=0)
        if migrate and it in (1, 3):
            # floe migration: ownership follows the current centroids; the floes keep their numbers, so the comparison below
            # sorts every rank's results back into global order
            gave = slab.repartition(slabs.slab_edges(field.x, prm.Lx, world))
            tot_gave = [None] * world
            dist.all_gather_object(tot_gave, gave)
            if rank == 0:
                print("MIGRATED", it, tot_gave, flush=True)
                n_migrated[0] += sum(tot_gave)
        if rank == 0:
            good = True
            # global order of the gathered floes (ranks own arbitrary ascending subsets after a migration)
            all_gid = np.concatenate([g["gid"] for g in gathered])
            order = np.argsort(all_gid, kind="stable")
            assert np.array_equal(all_gid[order], np.arange(1, field.n + 1))

            def ragged(vals, counts):
                starts = np.concatenate([[0], np.cumsum(counts)[:-1]]).astype(np.int64)
                co = counts[order]
                idx = np.repeat(starts[order] - np.concatenate([[0], np.cumsum(co)[:-1]]), co) + np.arange(int(co.sum()))
                return vals[idx]
            for g in gathered:
                g["rowcnt"] = np.diff(g["row_off"])
            per = lambda key, sub: np.concatenate([g[sub][key] for g in gathered])[order]
            cat_rows = ragged(np.concatenate([g["rows"] for g in gathered]), np.concatenate([g["rowcnt"] for g in gathered]))
            cat_cnt = np.concatenate([g["rowcnt"] for g in gathered])[order]
            nvs = np.concatenate([g["nv"] for g in gathered])
            for g in gathered:
                g["out"] = dict(g["out"])
            gathered = [{"out": {k: per(k, "out") for k in gathered[0]["out"]}, "rows": cat_rows, "row_off": np.concatenate([[0], np.cumsum(cat_cnt)]),
                         "state": {k: (ragged(np.concatenate([g["state"][k] for g in gathered]), nvs) if k in ("cax", "cay") else per(k, "state")) for k in gathered[0]["state"]},
                         "stats": np.sum([g["stats"] for g in gathered], 0)}]
            s1 = one.step_resident(allow_pair_errors=True)
            o1 = one.floe_outputs()
            off1, rows1 = one.rows()
            ns1 = one.trajectory_step(prm.dt, hfo, *bounds)
            st1 = one.trajectory_state(nverts=field.vx.shape[0])
            for k in o1:
                got = np.concatenate([g["out"][k] for g in gathered])
                if not np.array_equal(got, o1[k], equal_nan=True):
                    good = False
                    print("MISMATCH per-floe", it, k, int((got != o1[k]).sum()))
            got_rows = np.concatenate([g["rows"] for g in gathered])
            cnt = np.concatenate([np.diff(g["row_off"]) for g in gathered])
            if not np.array_equal(cnt, np.diff(off1[:field.n + 1])):
                good = False
                print("MISMATCH row counts", it)
            elif not np.array_equal(got_rows, rows1[:off1[field.n]], equal_nan=True):
                good = False
                print("MISMATCH rows", it, int((got_rows != rows1[:off1[field.n]]).sum()))
            for k in ("x", "y", "u", "v", "ksi", "h", "alive", "mass", "alpha", "dUi_p", "dksi_p", "stress", "flags", "cax", "cay"):
                got = np.concatenate([g["state"][k] for g in gathered])
                if not np.array_equal(got, st1[k], equal_nan=True):
                    good = False
                    print("MISMATCH state", it, k, int((got != st1[k]).sum()))
            tot = np.sum([g["stats"] for g in gathered], 0)
            if (int(tot[0]), int(tot[1]), tot[2], int(tot[4])) != (s1.n_pairs, s1.n_pairs_force, s1.collision_count, ns1):
                good = False
                print("MISMATCH totals", it, tot, s1.n_pairs, s1.n_pairs_force, s1.collision_count, ns1)
            moved = max(moved, float(np.nanmax(np.abs(st1["alpha"]))))
            ok = ok and good
            print("STEP %d %s pairs=%d rows=%d duplicated_pairs=%.3f ghosts=%d kills=%d plans=%d" % (it, "OK" if good else "FAIL", s1.n_pairs, off1[field.n], tot[3] / max(1, s1.n_pairs) - 1,
                                                                                             s1.n - s1.n0, int((o1["kill"] > 0).sum()), int(tot[5])), flush=True)
    if rank == 0:
        if migrate and n_migrated[0] == 0:
            ok = False
            print("no floe migrated: the test did not exercise repartition()")
        if use_graph and slab.graph_replays == 0:
            ok = False
            print("the CUDA graph was never replayed")
        print("RESULT %s world=%d backend=%s floes=%d steps=%d max_alpha=%.3e kill_events=%d migrated=%d graph_replays=%d" % (
            "OK" if ok and moved > 0 else "FAIL", world, backend, field.n, steps, moved, int((o1["kill"] > 0).sum()), n_migrated[0], slab.graph_replays), flush=True)
        one.close()
    slab.graph = None                 # a captured graph holds NCCL kernels: let it go before the process group
    torch.cuda.synchronize()
    ctx.close()
    dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    if use_graph:
        os._exit(0 if ok else 1)      # tearing NCCL down after its collectives were captured into a CUDA graph has hung on this stack
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
