"""Synthetic input of BASELINE.json configs[4]: packed periodic Voronoi floe field (host generator in the
product library, subzero_b200/csrc/sz_field.cpp; recipe in SURVEY.md 8d)."""
import ctypes as C

import numpy as np

import os

from . import abi

_flib = None


def _field_lib():
    """the generator's own small host library (libsz_field.so, built next to the product library): the CPU reference arm uses
    the same synthetic field without mapping the CUDA library"""
    global _flib
    if _flib is None:
        path = os.path.join(os.path.dirname(abi.LIB_PATH), "libsz_field.so")
        if not os.path.exists(path):
            raise RuntimeError("subzero_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
        l = C.CDLL(path)
        for name in ("sz_field_voronoi", "sz_field_view", "sz_field_free"):
            fn = getattr(l, name)
            fn.restype, fn.argtypes = abi.PROTOTYPES[name]
        _flib = l
    return _flib


def morton_order(x, y, Lx, Ly):
    """indices that sort the centroids along a Z-order curve over [-Lx, Lx] x [-Ly, Ly] (16 bits per axis, stable)"""
    def spread(v):
        v = v.astype(np.uint64)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
        return v
    q = lambda c, L: np.clip(np.nan_to_num((np.asarray(c, np.float64) + L) / (2 * L), nan=0.0) * 65535.0, 0, 65535).astype(np.uint32)
    key = spread(q(x, Lx)) | (spread(q(y, Ly)) << np.uint64(1))
    return np.argsort(key, kind="stable")


def voronoi_field(n_floes, seed=0, mean_area=4e6, inflate=0.02, dt=10.0, order="site"):
    """Returns (params, FloesSoA): doubly periodic [-L,L]^2, L = 0.5*sqrt(N*mean_area).  order: "site" numbers the floes in the
    generator's (random) site order, the recipe of SURVEY.md 8d; "morton" renumbers the same floes along a Z-order curve (an
    experiment on how much a spatial numbering is worth to the gather-bound kernels; same floes, different input order)."""
    prm = abi.default_params()
    h = C.c_void_p()
    flib = _field_lib()
    abi.check(flib.sz_field_voronoi(C.byref(h), int(n_floes), int(seed), float(mean_area), float(inflate), C.byref(prm)))
    try:
        v = abi.SzFloesSoA()
        abi.check(flib.sz_field_view(h, C.byref(v)))
        n, nv = v.n, v.nverts
        cp = lambda p, m, dt_: np.ctypeslib.as_array(p, shape=(m,)).astype(dt_, copy=True) if m else np.zeros(0, dt_)
        soa = abi.FloesSoA(*(cp(getattr(v, k), n, np.float64) for k in abi.FloesSoA.FIELDS),
                           cp(v.alive, n, np.uint8), cp(v.voff, n + 1, np.int32), cp(v.vx, nv, np.float64), cp(v.vy, nv, np.float64))
    finally:
        flib.sz_field_free(h)
    prm.periodic, prm.collision, prm.dt, prm.Nb = 1, 1, float(dt), 0
    if order == "morton":
        soa = soa.take(morton_order(soa.x, soa.y, prm.Lx, prm.Ly))
    elif order != "site":
        raise ValueError("voronoi_field: order must be 'site' or 'morton'")
    return prm, soa
