"""GPU probe: tiled real (concave, 7..591-vertex) floe shapes from the reference fixture -- classes M and L"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests")
import numpy as np
import subzero_b200 as sz, scenarios
side = int(sys.argv[1]) if len(sys.argv) > 1 else 60
t = time.time(); mv = int(os.environ.get('MAXV', '0')) or None
prm, Floe = scenarios.real_shape_field(side, seed=3, max_vertices=mv); soa = sz.floes_to_soa(Floe); print("field", soa.n, "floes", soa.vx.shape[0], "vertices", round(time.time() - t, 1), "s", flush=True)
ctx = sz.ContactContext(0); ctx.upload(prm, soa)
for it in range(3):
    s = ctx.step_resident(allow_pair_errors=True)
    print(it, "pairs", s.n_pairs, "force", s.n_pairs_force, "rows", s.n_rows, "ms", round(s.ms_device, 2), "pairs/s %.3e" % (s.n_pairs / s.ms_device * 1e3), {k: round(v, 2) for k, v in ctx.phase_ms().items()}, "fail", s.n_clipper_fail, flush=True)
if len(sys.argv) > 2:
    import oracle
    t = time.time(); r = oracle.OracleStep(prm, soa, broad_mode=1); dt = time.time() - t
    print("oracle %.2f s -> %.3e pairs/s on %d threads" % (dt, r.summary.n_pairs / dt, oracle.lib().szo_hardware_threads()))
