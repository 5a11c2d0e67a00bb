"""Synthetic input of BASELINE.json configs[4]: packed periodic Voronoi floe field (host generator in the
product library, subzero_b200/csrc/sz_field.cpp; recipe in SURVEY.md 8d)."""
import ctypes as C

import numpy as np

from . import abi


def voronoi_field(n_floes, seed=0, mean_area=4e6, inflate=0.02, dt=10.0):
    """Returns (params, FloesSoA): doubly periodic [-L,L]^2, L = 0.5*sqrt(N*mean_area)."""
    prm = abi.default_params()
    h = C.c_void_p()
    abi.check(abi.lib().sz_field_voronoi(C.byref(h), int(n_floes), int(seed), float(mean_area), float(inflate), C.byref(prm)))
    try:
        v = abi.SzFloesSoA()
        abi.check(abi.lib().sz_field_view(h, C.byref(v)))
        n, nv = v.n, v.nverts
        cp = lambda p, m, dt_: np.ctypeslib.as_array(p, shape=(m,)).astype(dt_, copy=True) if m else np.zeros(0, dt_)
        soa = abi.FloesSoA(*(cp(getattr(v, k), n, np.float64) for k in abi.FloesSoA.FIELDS),
                           cp(v.alive, n, np.uint8), cp(v.voff, n + 1, np.int32), cp(v.vx, nv, np.float64), cp(v.vy, nv, np.float64))
    finally:
        abi.lib().sz_field_free(h)
    prm.periodic, prm.collision, prm.dt, prm.Nb = 1, 1, float(dt), 0
    return prm, soa
