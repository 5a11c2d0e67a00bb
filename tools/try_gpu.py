import sys, time; import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np, subzero_b200 as sz, oracle
for n in (500, 5000):
    prm, f = sz.voronoi_field(n, seed=1)
    prm.want_clip_polys = 1
    t = time.time(); ref = oracle.OracleStep(prm, f, nthreads=8, broad_mode=0); t1 = time.time() - t
    print('oracle', n, ref.summary.as_dict(), 'sec', t1)
    ctx = sz.ContactContext(0)
    s = ctx.step(prm, f, allow_pair_errors=True)
    print('gpu', s.as_dict())
    print(oracle.compare_steps(ctx, ref))
