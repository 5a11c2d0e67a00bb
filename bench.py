"""bench.py -- floe-pair contacts resolved per second on the synthetic packed-Voronoi field
(BASELINE.json configs[4]), one process per GPU.

  python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus 1 --steps K ...  # the reference's CPU path (oracle port linked to the
                                                            # reference's own Clipper) on the host cores

A "step" is one pass of the contact loop (floe_interactions_all.m:9-285) over the whole field.
`value`  : candidate pairs taken through narrow phase + force law + reductions per second, inputs resident in HBM
           (sz_step_resident), timed with CUDA events on the library's stream, max over ranks.
`e2e`    : the same metric through the host-buffer entry point sz_contact_step + result read-back (per-floe
           outputs and all contact rows), host<->device copies inside the timed region (pinned host memory).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "floe-pair contacts resolved/s"
UNIT = "pairs/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst)"
        except Exception:
            pass
    return 6400.0, "fallback from B200_PROFILING.md (MEASURED_PEAKS.json absent)"


class ClockSampler(threading.Thread):
    """samples the SM clock and the throttle reasons during the timed region: NVML every 10 ms when pynvml is there
    (the timed region can be well under a second), else nvidia-smi every 200 ms"""

    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu, self.samples, self.reasons, self.max_mhz, self._stop_ev = gpu, [], set(), None, threading.Event()
        self.how = "nvidia-smi"
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[gpu]) if vis and vis.split(",")[0].isdigit() else gpu
            self._h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._nv = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self.how = "nvml"
        except Exception:
            self._nv = None

    def run(self):
        if self._nv is not None:
            nv = self._nv
            while not self._stop_ev.is_set():
                try:
                    self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    for nm, bit in self.REASONS.items():
                        if mask & bit:
                            self.reasons.add(nm)
                except Exception:
                    pass
                self._stop_ev.wait(0.01)
            return
        q = "clocks.sm,clocks.max.sm,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.hw_thermal_slowdown,clocks_throttle_reasons.sw_thermal_slowdown,clocks_throttle_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=6)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s), "how": self.how}


def algorithmic_bytes(floes, summary):
    """per candidate pair both outlines (16 B/vertex) + both state records (2 x 72 B) + pair id (8 B)
    + contact rows written (own + mirrored, 56 B each)."""
    nv_mean = floes.vx.shape[0] / max(1, floes.n)
    per_pair = 2 * nv_mean * 16 + 2 * 72 + 8
    return summary.n_pairs * per_pair + summary.n_rows * 56


_cpu_fields = {}


def cpu_sample(n_floes, seed, threads, order="site"):
    """the oracle (reference restatement + the reference's Clipper) on a bounded sample of the same workload;
    returns (pairs resolved, seconds of the contact step alone)"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import subzero_b200 as sz
    import oracle
    key = (n_floes, seed, order)
    if key not in _cpu_fields:
        _cpu_fields[key] = sz.voronoi_field(n_floes, seed=seed, order=order)
    prm, f = _cpu_fields[key]
    t = time.perf_counter()
    r = oracle.OracleStep(prm, f, nthreads=threads, broad_mode=1)
    dt = time.perf_counter() - t
    return r.summary.n_pairs, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    threads = max(1, oracle.lib().szo_hardware_threads())
    n_sample = args.cpu_floes
    for _ in range(args.warmup):
        cpu_sample(min(n_sample, 20000), args.seed, threads, args.floe_order)
    tot_pairs, tot_t = 0, 0.0
    for _ in range(args.steps):
        p, dt = cpu_sample(n_sample, args.seed, threads, args.floe_order)
        tot_pairs += p
        tot_t += dt
    v = tot_pairs / tot_t
    sample = "%d-floe periodic Voronoi field (same generator, density and physics as the %d-floe workload), whole contact step, cell-grid broad phase" % (n_sample, args.floes)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64+int64",
            "data": "synthetic", "config": {"workload": "configs[4] synthetic packed periodic Voronoi floe field, contact loop only", "floes": args.floes, "sample_floes": n_sample, "floe_order": args.floe_order,
                                            "same_config": n_sample == args.floes},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample + "; oracle = C++ restatement of the MATLAB path calling the reference's unmodified Clipper 6.4.2 (MATLAB itself is not installed)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)



def verify_against_single_gpu(job, steps, rank, world, local_rank, dist):
    """Parity of the multi-GPU job inside the bench run: `steps` coupled steps (contact step + integrator with nonzero ksi and
    thermodynamic thinning, so positions, outlines and thickness change every step) on the slab job, and the same loop on ONE
    GPU (rank 0, same field, same floe numbering).  Compared bit for bit, per rank's floes: the last step's per-floe outputs and
    contact rows, and the integrated state (x y u v ksi h alive alpha, rotated outlines).  Returns a dict for the JSON line."""
    import hashlib
    import numpy as np
    import subzero_b200 as sz
    prm, field, starts = job.prm, job.field, job.starts
    rho_ice, nz, hfo = 920.0, 1, 1e-4
    mass = field.area * field.h * rho_ice
    inertia = mass * field.rmax ** 2 / 4

    def digest(out, row_off, rows, st):
        h = hashlib.sha256()
        for k in ("fx", "fy", "torque", "overlap_area", "stress", "xi", "yi", "alive", "kill", "transfer"):
            h.update(np.ascontiguousarray(out[k]).tobytes())
        h.update(np.ascontiguousarray(np.diff(row_off)).astype(np.int64).tobytes())
        h.update(np.ascontiguousarray(rows).tobytes())
        for k in ("x", "y", "u", "v", "ksi", "h", "alive", "alpha", "cax", "cay"):
            h.update(np.ascontiguousarray(st[k]).tobytes())
        return h.hexdigest()

    a, b = int(starts[rank]), int(starts[rank + 1])
    slab = job.slab
    slab.trajectory_init(mass[a:b], inertia[a:b], nz=nz, dXi_p=field.u[a:b], dYi_p=field.v[a:b])
    pairs = 0
    for _ in range(steps):
        s = slab.run()
        pairs = int(s.n_pairs_owned)
        out, (row_off, rows) = slab.outputs(), slab.rows()
        slab.trajectory_step(prm.dt, hfo)
    st = job.ctx.trajectory_state(nverts=job.floes.vx.shape[0])
    mine = digest(out, row_off, rows, st)
    alpha_max = float(np.abs(st["alpha"]).max()) if st["alpha"].size else 0.0
    got = [None] * world
    dist.all_gather_object(got, (mine, pairs, alpha_max))
    res = None
    if rank == 0:
        with sz.ContactContext(local_rank) as one:
            one.upload(prm, field)
            one.trajectory_init(mass, inertia, nz=nz, dXi_p=field.u, dYi_p=field.v)
            for _ in range(steps):
                s1 = one.step_resident()
                o1, (off1, rows1) = one.floe_outputs(), one.rows()
                one.trajectory_step(prm.dt, hfo)
            st1 = one.trajectory_state(nverts=field.vx.shape[0])
        bad = []
        for r in range(world):
            ra, rb = int(starts[r]), int(starts[r + 1])
            va, vb = int(field.voff[ra]), int(field.voff[rb])
            sl = {k: (st1[k][va:vb] if k in ("cax", "cay") else st1[k][ra:rb]) for k in st1}
            ref = digest({k: v[ra:rb] for k, v in o1.items()}, off1[ra:rb + 1], rows1[off1[ra]:off1[rb]], sl)
            if ref != got[r][0]:
                bad.append(r)
        res = {"checked": True, "ok": not bad and sum(g[1] for g in got) == int(s1.n_pairs), "steps": steps, "ranks_differing": bad,
               "max_abs_heading_change_rad": max(g[2] for g in got),
               "what": "coupled contact + integrator steps (nonzero ksi, HFo %g) at N = %d vs N = 1 on the same field and numbering: sha256 of each rank's per-floe outputs, "
                       "contact rows (last step) and integrated state incl. rotated outlines" % (hfo, world)}
    return res



def run_real_shapes(args):
    """Second workload (--workload real_shapes / real_shapes_raw), one GPU: the reference's own floe outlines
    (test/test_conservation/FloeShapes.mat, 7..591 vertices, concave; decoded into tests/golden by tools/make_golden_floeshapes.py)
    tiled into a periodic field -- `real_shapes`: thinned to <= 30 vertices, what FloeSimplify leaves in a production run
    (Subzero.m:169-217; narrow-phase class T); `real_shapes_raw`: as they are (classes M and L).  Same JSON contract."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import scenarios
    import subzero_b200 as sz
    from subzero_b200 import abi
    raw = args.workload == "real_shapes_raw"
    n_side = args.tiles
    prm, Floe = scenarios.real_shape_field(n_side, seed=args.seed, max_vertices=None if raw else 30)
    soa = sz.floes_to_soa(Floe)
    ctx = sz.ContactContext(0)
    ctx.upload(prm, soa)
    for _ in range(max(3, args.warmup)):
        ctx.step_resident(allow_pair_errors=True)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = abi.lib().sz_launch_count()
    ms, cls_ms, cls_pairs = 0.0, {}, {}
    for _ in range(args.steps):
        s = ctx.step_resident(allow_pair_errors=True)
        ms += s.ms_device
        for k, (t, p) in ctx.narrow_class_ms().items():
            cls_ms[k] = cls_ms.get(k, 0.0) + t
            cls_pairs[k] = p
    torch.cuda.synchronize()
    launches = abi.lib().sz_launch_count() - l0
    clocks = sampler.stop()
    ms_per_step = ms / args.steps
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    pinned = abi.FloesSoA(*(pin(getattr(soa, k)) for k in abi.FloesSoA.FIELDS), pin(soa.alive), pin(soa.voff), pin(soa.vx), pin(soa.vy))
    h2d = sum(getattr(pinned, k).nbytes for k in abi.FloesSoA.FIELDS) + pinned.alive.nbytes + pinned.voff.nbytes + pinned.vx.nbytes + pinned.vy.nbytes
    e2e_t, d2h = 0.0, 0
    for it in range(4):
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        ctx.step(prm, pinned, allow_pair_errors=True)
        o = ctx.floe_outputs()
        off, rows = ctx.rows()
        dt = time.perf_counter() - w0
        if it:
            e2e_t += dt
        d2h = sum(v.nbytes for v in o.values()) + off.nbytes + rows.nbytes
    e2e_ms = 1e3 * e2e_t / 3
    top = max(cls_ms, key=lambda k: cls_ms[k])
    peak, peak_src = load_peaks()
    nv_mean = soa.vx.shape[0] / max(1, soa.n)
    alg = cls_pairs[top] * (2 * nv_mean * 16 + 2 * 72 + 8) + s.n_rows * 56 * (cls_pairs[top] / max(1, s.n_pairs))
    kms = cls_ms[top] / args.steps
    line = {"metric": METRIC, "value": s.n_pairs / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64+int64", "data": "synthetic",
            "config": {"workload": "%s: %d x %d tiles of the reference's FloeShapes.mat outlines%s, periodic, contact loop only" % (args.workload, n_side, n_side, "" if raw else " thinned to <= 30 vertices"),
                       "floes": soa.n, "floes_incl_ghosts": int(s.n), "pairs_per_step": int(s.n_pairs), "pairs_with_force": int(s.n_pairs_force), "rows_per_step": int(s.n_rows),
                       "vertices_per_outline_mean": nv_mean, "class_pairs": cls_pairs, "class_ms": {k: v / args.steps for k, v in cls_ms.items()},
                       "l2": "the same field every step; the per-pair arenas (class T: 35 KB of local memory per thread, M/L: HBM scratch) exceed L2", "seed": args.seed},
            "clocks": clocks, "e2e": {"value": s.n_pairs / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": 3},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "narrow phase class %s (general Clipper-exact sweep + force law, thread per pair)" % top, "achieved": alg / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (kms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "kernel_ms": kms, "kernel_pairs_per_launch": cls_pairs[top],
                         "note": "latency-bound sequential sweep per pair; no ncu capture for this workload"}}
    if not args.no_cpu:
        import oracle
        threads = max(1, oracle.lib().szo_hardware_threads())
        p, dt = 0, 0.0
        for _ in range(2):
            t0 = time.perf_counter()
            r = oracle.OracleStep(prm, soa, nthreads=threads, broad_mode=1)
            dt += time.perf_counter() - t0
            p += r.summary.n_pairs
        line["cpu_baseline"] = {"value": p / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "the same field, 2 whole contact steps (%.1f s on %d threads); oracle = C++ restatement of the MATLAB path calling the reference's unmodified Clipper 6.4.2" % (dt, threads)}
    print(json.dumps(line), flush=True)
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--floes", type=int, default=1000000)
    ap.add_argument("--cpu-floes", type=int, default=0, help="floes of the CPU arm's field; 0 (default) = --floes, i.e. the benchmark configuration itself")
    ap.add_argument("--verify", type=int, default=-1, metavar="STEPS",
                    help="after the timed region, run STEPS coupled steps (contact step + integrator, nonzero ksi, thinning) and compare every rank's per-floe outputs, contact rows and "
                         "integrated state with a single-GPU run of the same field bit for bit; prints parity in the line.  Default: 20 at N > 1, off at N = 1")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--workload", default="voronoi", choices=["voronoi", "real_shapes", "real_shapes_raw"],
                    help="voronoi (default): BASELINE.json configs[4], the headline; real_shapes / real_shapes_raw: the reference's own concave floe outlines tiled (one GPU)")
    ap.add_argument("--tiles", type=int, default=240, help="real_shapes: tiles per side (240 -> 57,600 floes)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-graph", action="store_true", help="N > 1: launch every kernel and collective of a step from the host instead of replaying the step's CUDA graph")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="run-time switch of the library (sz_set_option), e.g. convex_split=1; experiments only, recorded in config")
    ap.add_argument("--floe-order", default="site", choices=["site", "morton"],
                    help="numbering of the synthetic floes: the generator's site order (default, SURVEY.md 8d) or a Z-order curve (experiment; both arms)")
    args = ap.parse_args()
    if args.cpu_floes <= 0:
        args.cpu_floes = args.floes
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "voronoi":
        if rank == 0:
            run_real_shapes(args)
        return

    import numpy as np
    import torch
    import subzero_b200 as sz
    from subzero_b200 import abi
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from subzero_b200 import slabs
    job = slabs.SlabJob(args.floes, args.seed, rank, world, local_rank, dist, order=args.floe_order, graph=not args.no_graph)
    for o in args.opt:
        name, _, val = o.partition("=")
        job.ctx.set_option(name.strip(), int(val or 1))
    prm, floes = job.prm, job.floes

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput
    launches0 = abi.lib().sz_launch_count()
    warm_kern_ms, warm_narrow_ms, warm_pairs = 0.0, 0.0, 0
    for _ in range(max(3, args.warmup)):
        _, wph = job.step_resident()
        if wph["classes"]["C"][0] > 0:            # (steps replayed from a CUDA graph carry no per-kernel events: N > 1 reports the class C launch of a warm-up step)
            warm_kern_ms, warm_narrow_ms, warm_pairs = wph["classes"]["C"][0], wph["narrow"], wph["classes"]["C"][1]
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches1 = abi.lib().sz_launch_count()
    dev_ms, narrow_ms, kern_ms, kern_pairs, wall0 = 0.0, 0.0, 0.0, 0, time.perf_counter()
    for _ in range(args.steps):
        ms, ph = job.step_resident()
        dev_ms += ms
        narrow_ms += ph["narrow"]
        kern_ms += ph["classes"]["C"][0]          # the dominant kernel: class C of the narrow phase (CUDA events on the library's stream)
        kern_pairs = ph["classes"]["C"][1]
    phase_last = {k: v for k, v in ph.items() if k != "classes"}
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches = abi.lib().sz_launch_count() - launches1
    clocks = sampler.stop() if rank == 0 else None
    pairs_local = job.pairs_owned
    t = torch.tensor([dev_ms, wall_ms, narrow_ms], dtype=torch.float64, device="cuda")
    cnt = torch.tensor([pairs_local, job.rows_owned, launches, job.pairs_force_total, job.ext_entries_owned], dtype=torch.float64, device="cuda")
    per_rank = None
    if dist is not None:
        # every rank's own numbers (the step time is the slowest rank's): device ms per step, contact step alone, narrow phase, owned pairs
        mine = torch.tensor([dev_ms / args.steps, ph["total"], ph["narrow"], float(pairs_local)], dtype=torch.float64, device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(v), 4) for v in r.tolist()] for r in allr]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, narrow_ms = [float(v) for v in t.tolist()]
    total_pairs, total_rows, launches, total_force, total_ext = [int(v) for v in cnt.tolist()]
    ms_per_step = dev_ms / args.steps
    value = total_pairs / (ms_per_step * 1e-3)

    # ---- end to end: host buffers in (pinned), per-floe outputs + contact rows out, every step
    e2e_steps = max(1, min(args.steps, 3))
    job.e2e_step()   # warm
    barrier()
    w0 = time.perf_counter()
    h2d = d2h = 0
    for _ in range(e2e_steps):
        h2d, d2h = job.e2e_step()
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - w0) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
    tb = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tb, op=dist.ReduceOp.SUM)
    e2e_ms = float(te.item())
    e2e_value = total_pairs / (e2e_ms * 1e-3)

    # ---- parity of the multi-GPU path, inside this run
    parity = None
    n_verify = args.verify if args.verify >= 0 else (20 if world > 1 else 0)
    if world > 1 and n_verify > 0:
        parity = verify_against_single_gpu(job, n_verify, rank, world, local_rank, dist)
    # ---- timesteps/s with the integrator half of the step on the device (contact step + calc_trajectory, N = 1)
    ts_with_ab2 = None
    if world == 1:
        rho_ice = 920.0
        mass = floes.area * floes.h * rho_ice
        job.ctx.trajectory_init(mass, mass * floes.rmax ** 2 / 4, nz=1000)
        for _ in range(2):
            job.ctx.step_resident()
            job.ctx.trajectory_step(prm.dt)
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        nts = max(3, min(args.steps, 10))
        for _ in range(nts):
            job.ctx.step_resident()
            job.ctx.trajectory_step(prm.dt)
        torch.cuda.synchronize()
        ts_with_ab2 = nts / (time.perf_counter() - w0)
    if rank == 0:
        peak, peak_src = load_peaks()
        # rank 0's launch: the pairs class C received and the rows they produce (practically all of them on this field)
        alg_bytes = algorithmic_bytes(floes, job.summary) * (kern_pairs / max(1, job.summary.n_pairs))
        narrow_ms_per = narrow_ms / args.steps
        kern_ms_per = kern_ms / args.steps
        if kern_ms_per <= 0:
            kern_ms_per, narrow_ms_per, kern_pairs = warm_kern_ms, warm_narrow_ms, warm_pairs
        achieved = alg_bytes / (kern_ms_per * 1e-3) / 1e9
        traffic, traffic_note = None, "no ncu capture published (profiles/narrow_traffic.json)"
        tp = os.path.join(ROOT, "profiles", "narrow_traffic.json")
        if os.path.exists(tp):
            try:
                from subzero_b200.build import kernel_stamp
                tj = json.load(open(tp))
                if tj.get("kernel_stamp") == kernel_stamp():
                    traffic, traffic_note = tj.get("dram_bytes_per_launch"), "ncu --set full capture of this kernel's sources: " + str(tj.get("source"))
                else:
                    traffic_note = "stale: the published capture (%s) was taken from other sources of this kernel" % tj.get("source")
            except Exception as e:
                traffic, traffic_note = None, "unreadable: %r" % (e,)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64+int64",
                "data": "synthetic",
                "config": {"workload": "configs[4] synthetic packed periodic Voronoi floe field, contact loop only", "floes": args.floes,
                           "floes_incl_ghosts": total_ext, "pairs_per_step": total_pairs, "pairs_with_force": total_force,
                           "rows_per_step": total_rows, "timesteps_per_s": 1e3 / ms_per_step, "timesteps_per_s_with_trajectory_update": ts_with_ab2, "parallelism": job.describe(),
                           "l2": "inputs larger than L2 (state + vertex pool + pair buffers >> 126 MB at 1M floes); no flush", "seed": args.seed, "floe_order": args.floe_order, "options": args.opt,
                           "wall_ms_per_step": wall_ms / args.steps, "phase_ms_rank0": phase_last, "per_rank_ms_step_contact_narrow_pairs": per_rank,
                           "slab_stage_ms_rank0": (job.slab.stage_ms() if job.slab is not None else None),
                           "cuda_graph": (None if job.slab is None else {"enabled": job.slab.want_graph, "replays": job.slab.graph_replays, "launches_per_replay": job.slab.graph_launches})},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(tb[0].item()), "d2h_bytes_per_step": int(tb[1].item()), "ms_per_step": e2e_ms, "steps": e2e_steps},
                "gpu_launches": launches,
                "parity": parity if parity is not None else {"checked": False, "why": "N = 1 is the reference of the multi-GPU parity check; its own parity with the oracle is tests/ -m gpu and smoke()"},
                "roofline": {"bound": "hbm", "kernel": "narrow_convex_kernel<PairS> (class C of the narrow phase: Clipper-exact convex sweep + force law)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "kernel_ms": kern_ms_per, "kernel_pairs_per_launch": kern_pairs,
                             "narrow_phase_ms": narrow_ms_per, "algorithmic_bytes_per_launch": alg_bytes,
                             "note": "latency-bound sequential sweep per pair (thread per pair); the HBM roofline is reported as the contract asks, see DESIGN.md 4.2"}}
        if world == 1 and not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import oracle
            threads = max(1, oracle.lib().szo_hardware_threads())
            p, dt = 0, 0.0
            for _ in range(3):
                pp, dd = cpu_sample(args.cpu_floes, args.seed, threads, args.floe_order)
                p, dt = p + pp, dt + dd
            line["cpu_baseline"] = {"value": p / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d-floe field from the same generator, 3 whole contact steps (%.1f s of CPU work on %d threads); oracle = C++ restatement of the MATLAB path "
                                              "calling the reference's unmodified Clipper 6.4.2 (MATLAB is not installed)" % (args.cpu_floes, dt, threads)}
        print(json.dumps(line), flush=True)
    job.close()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)      # (no destroy_process_group: tearing NCCL down after CUDA-graph capture of its collectives has hung on this stack; the line is printed, every rank is past the barrier)


if __name__ == "__main__":
    main()
