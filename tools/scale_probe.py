"""GPU probe: step time by phase at several field sizes (not a bench; see bench.py)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import subzero_b200 as sz
sizes = [int(float(a)) for a in sys.argv[1:]] or [10000, 100000, 1000000]
for n in sizes:
    prm, f = sz.voronoi_field(n, seed=0)
    ctx = sz.ContactContext(0)
    ctx.upload(prm, f)
    for it in range(3):
        t = time.time(); s = ctx.step_resident(); w = time.time() - t
        print(n, it, 'pairs', s.n_pairs, 'force', s.n_pairs_force, 'rows', s.n_rows, 'ms', round(s.ms_device, 3), 'wall_ms', round(w * 1e3, 3),
              'pairs/s %.3e' % (s.n_pairs / s.ms_device * 1e3), {k: round(v, 3) for k, v in ctx.phase_ms().items()}, 'classC', round(ctx.narrow_class_ms()['C'][0], 3), flush=True)
    ctx.close()
