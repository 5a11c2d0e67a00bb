"""ctypes binding of the C ABI declared in include/subzero_b200.h.

The library is the product: it is loaded from subzero_b200/_lib/libsubzero_b200.so (built in-tree by
subzero_b200/build.py) and there is no fallback of any kind -- a missing library or a missing sm_100a
device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SZ_LIB", os.path.join(_HERE, "_lib", "libsubzero_b200.so"))   # SZ_LIB: experiment builds only

SZ_OK, SZ_ERR_ARG, SZ_ERR_CUDA, SZ_ERR_CLIPPER, SZ_ERR_CAPACITY, SZ_ERR_STATE = 0, -1, -2, -3, -4, -5

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_bp = C.POINTER(C.c_uint8)


class SzParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "Lx", "Ly", "modulus", "dt", "nu", "mu", "merge_frac", "wall_frac", "amin_per_vertex", "vertex_match_tol",
        "on_edge_tol", "dl_min", "close_gap", "big_floe_r", "domain_area_frac")] + [
        (n, C.c_int32) for n in ("Nb", "periodic", "collision", "want_clip_polys", "pair_with_boundary_floes")]


class SzFloesSoA(C.Structure):
    _fields_ = [("n", C.c_int32), ("nverts", C.c_int64)] + [(n, c_dp) for n in ("x", "y", "rmax", "h", "area", "u", "v", "ksi")] + [
        ("alive", c_bp), ("voff", c_ip), ("vx", c_dp), ("vy", c_dp)]


class SzBoundary(C.Structure):
    _fields_ = [("x", c_dp), ("y", c_dp), ("n", C.c_int32), ("box_x", c_dp), ("box_y", c_dp), ("box_n", C.c_int32)] + [
        (n, C.c_double) for n in ("area", "h", "xi", "yi", "u", "v", "ksi")]


class SzExtendedList(C.Structure):
    _fields_ = [("gid", c_ip), ("floe_num", c_ip), ("root_x", c_dp), ("root_y", c_dp), ("owned", c_bp), ("parent", c_ip)]


class SzTrajectoryInit(C.Structure):
    NAMES = ("mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "FxOA", "FyOA", "torqueOA", "c0x", "c0y")
    _fields_ = [(n, c_dp) for n in NAMES] + [("nz", C.c_int32), ("stress_h", c_dp), ("stress_count", c_ip)]


class SzTrajectoryParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("dt", "HFo", "xo_min", "xo_max", "yo_min", "yo_max")]


class SzOcean(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("Xo", c_dp), ("Yo", c_dp), ("Uocn", c_dp), ("Vocn", c_dp), ("Uwinds", c_dp), ("Vwinds", c_dp),
                ("fCoriolis", C.c_double), ("turn_angle", C.c_double), ("rho0", C.c_double), ("Cd", C.c_double), ("rho_air", C.c_double), ("Cd_atm", C.c_double)]


class SzSummary(C.Structure):
    _fields_ = [("n0", C.c_int32), ("n", C.c_int32), ("n_pairs", C.c_int64), ("n_pairs_force", C.c_int64), ("n_rows", C.c_int64),
                ("n_clip_paths", C.c_int64), ("n_clip_verts", C.c_int64), ("collision_count", C.c_double),
                ("n_clipper_fail", C.c_int32), ("n_capacity_fail", C.c_int32), ("ms_device", C.c_float), ("n_pairs_owned", C.c_int64), ("n_kill_events", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/subzero_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "sz_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "sz_destroy": (None, [C.c_void_p]),
    "sz_last_error": (C.c_char_p, []),
    "sz_default_params": (None, [C.POINTER(SzParams)]),
    "sz_abi_version": (C.c_int, []),
    "sz_launch_count": (C.c_longlong, []),
    "sz_contact_step": (C.c_int, [C.c_void_p, C.POINTER(SzParams), C.POINTER(SzFloesSoA), C.POINTER(SzBoundary), C.POINTER(SzSummary)]),
    "sz_upload": (C.c_int, [C.c_void_p, C.POINTER(SzParams), C.POINTER(SzFloesSoA), C.POINTER(SzBoundary)]),
    "sz_upload_extended": (C.c_int, [C.c_void_p, C.POINTER(SzParams), C.POINTER(SzFloesSoA), C.POINTER(SzBoundary), C.POINTER(SzExtendedList)]),
    "sz_slab_meta_doubles": (C.c_int64, [C.c_int32]),
    "sz_slab_block_doubles": (C.c_int64, [C.c_int32, C.c_int32]),
    "sz_slab_upload": (C.c_int, [C.c_void_p, C.POINTER(SzParams), C.POINTER(SzFloesSoA), C.POINTER(SzBoundary), c_ip, C.c_int32, C.c_int32, C.c_int32]),
    "sz_slab_measure": (C.c_int, [C.c_void_p, c_dp]),
    "sz_slab_measure_halo": (C.c_int, [C.c_void_p, c_dp, c_lp, c_lp]),
    "sz_slab_configure": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]),
    "sz_slab_prepare": (C.c_int, [C.c_void_p, c_dp]),
    "sz_slab_pack": (C.c_int, [C.c_void_p, c_dp, c_dp]),
    "sz_slab_build": (C.c_int, [C.c_void_p, c_dp, c_ip]),
    "sz_slab_get_positions": (C.c_int, [C.c_void_p, c_ip, c_ip]),
    "sz_slab_get_rows": (C.c_int, [C.c_void_p, c_lp, c_dp, C.c_int64, c_lp]),
    "sz_slab_get_list": (C.c_int, [C.c_void_p, c_ip, c_ip, c_bp, c_dp, c_dp]),
    "sz_slab_get_outputs": (C.c_int, [C.c_void_p] + [c_dp] * 7 + [c_bp, c_ip, c_ip]),
    "sz_step_resident": (C.c_int, [C.c_void_p, C.POINTER(SzSummary)]),
    "sz_step_enqueue": (C.c_int, [C.c_void_p]),
    "sz_step_finish": (C.c_int, [C.c_void_p, C.POINTER(SzSummary)]),
    "sz_add_launches": (None, [C.c_longlong]),
    "sz_get_floe_outputs": (C.c_int, [C.c_void_p] + [c_dp] * 7 + [c_bp, c_ip, c_ip]),
    "sz_get_ghosts": (C.c_int, [C.c_void_p, c_ip, c_ip, c_dp, c_dp]),
    "sz_get_ghost_outputs": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp]),
    "sz_get_pairs": (C.c_int, [C.c_void_p, c_ip, c_ip, c_dp, c_ip, c_ip]),
    "sz_get_rows": (C.c_int, [C.c_void_p, c_lp, c_dp]),
    "sz_trajectory_init": (C.c_int, [C.c_void_p, C.POINTER(SzTrajectoryInit)]),
    "sz_trajectory_step": (C.c_int, [C.c_void_p, C.POINTER(SzTrajectoryParams), c_ip, c_ip]),
    "sz_get_stress_history": (C.c_int, [C.c_void_p, c_dp, c_ip]),
    "sz_slab_set_extent": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "sz_trajectory_set_ocean": (C.c_int, [C.c_void_p, C.POINTER(SzOcean)]),
    "sz_trajectory_set_points": (C.c_int, [C.c_void_p, C.c_int32, c_dp, c_dp, c_bp]),
    "sz_trajectory_ocean_forcing": (C.c_int, [C.c_void_p, C.POINTER(SzTrajectoryParams), C.c_int32, c_ip, c_ip]),
    "sz_get_trajectory_forcing": (C.c_int, [C.c_void_p, c_dp, c_dp, c_dp, c_dp]),
    "sz_fracture_deform": (C.c_int, [C.c_void_p, C.c_int32, c_ip, c_lp, c_lp]),
    "sz_get_fracture_deform": (C.c_int, [C.c_void_p, c_bp, c_dp, c_dp, c_dp, c_lp, c_dp, c_dp]),
    "sz_corner_mask": (C.c_int, [C.c_void_p, C.c_int32, c_ip, C.c_int32, c_lp]),
    "sz_get_corner_mask": (C.c_int, [C.c_void_p, c_lp, c_bp]),
    "sz_eulerian_data": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "sz_pair_search": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32, c_ip, c_lp]),
    "sz_get_pair_search": (C.c_int, [C.c_void_p, c_ip, c_lp, c_ip]),
    "sz_get_trajectory": (C.c_int, [C.c_void_p] + [c_dp] * 6 + [c_bp] + [c_dp] * 10 + [c_ip, c_dp, c_dp]),
    "sz_get_phase_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "sz_get_narrow_class_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
    "sz_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32]),
    "sz_get_stat": (C.c_int, [C.c_void_p, C.c_char_p, c_lp]),
    "sz_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sz_get_clip_polys": (C.c_int, [C.c_void_p, c_lp, c_lp, c_lp, c_lp]),
    "sz_clip_batch": (C.c_int, [C.c_void_p, C.c_int32, c_ip, c_lp, c_lp, c_lp, c_lp, c_lp, c_lp, c_lp, c_lp]),
    "sz_get_clip_batch": (C.c_int, [C.c_void_p, c_ip, c_lp, c_lp, c_lp, c_lp]),
    "sz_field_voronoi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int32, C.c_uint64, C.c_double, C.c_double, C.POINTER(SzParams)]),
    "sz_field_view": (C.c_int, [C.c_void_p, C.POINTER(SzFloesSoA)]),
    "sz_field_free": (None, [C.c_void_p]),
}

_lib = None


def lib():
    """The loaded product library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "subzero_b200: %s is missing -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)" % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


class SzError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (SzStatus %d)" % (msg, code))
        self.code = code


def check(code):
    if code != SZ_OK:
        raise SzError(code, lib().sz_last_error().decode("utf-8", "replace"))


def default_params(**kw):
    """sz_default_params of the C ABI restated on the host (the constants hard-coded in collisions/floe_interactions.m:20-21,37,
    55-58,79,99,127,141,15,54), so that building a parameter block does not need the CUDA library; tests/test_abi.py holds the
    two together"""
    p = SzParams()
    p.nu, p.mu, p.merge_frac, p.wall_frac, p.amin_per_vertex = 0.3, 0.2, 0.55, 0.75, 100.0 / 1.75
    p.vertex_match_tol, p.on_edge_tol, p.dl_min, p.close_gap, p.big_floe_r, p.domain_area_frac = 1.0, 1e-8, 0.1, 1.0, 1e5, 0.95
    p.dt, p.collision, p.periodic, p.Nb, p.want_clip_polys, p.pair_with_boundary_floes = 10.0, 1, 0, 0, 0, 0
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _ptr(a, typ):
    return a.ctypes.data_as(typ) if a is not None else None


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class FloesSoA:
    """Structure-of-arrays view of the hot-path fields of the reference's Floe struct array
    (Initialize_Model/initialize_floe_values.m:12-52).  Keeps the numpy arrays alive."""

    FIELDS = ("x", "y", "rmax", "h", "area", "u", "v", "ksi")

    def __init__(self, x, y, rmax, h, area, u, v, ksi, alive, voff, vx, vy):
        self.x, self.y, self.rmax, self.h, self.area = f64(x), f64(y), f64(rmax), f64(h), f64(area)
        self.u, self.v, self.ksi = f64(u), f64(v), f64(ksi)
        self.alive = np.ascontiguousarray(alive, dtype=np.uint8)
        self.voff = np.ascontiguousarray(voff, dtype=np.int32)
        self.vx, self.vy = f64(vx), f64(vy)
        n = self.x.shape[0]
        for nm in self.FIELDS + ("alive",):
            if getattr(self, nm).shape != (n,):
                raise ValueError("FloesSoA: field %s has the wrong length" % nm)
        if self.voff.shape != (n + 1,) or self.vx.shape != self.vy.shape or (n and int(self.voff[-1]) != self.vx.shape[0]):
            raise ValueError("FloesSoA: voff / vertex pool mismatch")

    @property
    def n(self):
        return self.x.shape[0]

    def take(self, order):
        """the floes `order` (indices), in that order, with their outlines"""
        order = np.asarray(order, np.int64)
        nv = (self.voff[1:] - self.voff[:-1]).astype(np.int64)
        new_nv = nv[order]
        new_off = np.zeros(order.shape[0] + 1, np.int64)
        np.cumsum(new_nv, out=new_off[1:])
        idx = np.repeat(self.voff[:-1].astype(np.int64)[order] - new_off[:-1], new_nv) + np.arange(int(new_off[-1]))
        return FloesSoA(*(getattr(self, k)[order] for k in self.FIELDS), self.alive[order], new_off.astype(np.int32), self.vx[idx], self.vy[idx])

    def struct(self):
        s = SzFloesSoA()
        s.n = self.n
        s.nverts = self.vx.shape[0]
        for nm in self.FIELDS + ("vx", "vy"):
            setattr(s, nm, _ptr(getattr(self, nm), c_dp))
        s.alive = _ptr(self.alive, c_bp)
        s.voff = _ptr(self.voff, c_ip)
        return s

    def outline(self, i):
        """closed c_alpha of floe i as (x, y)"""
        a, b = int(self.voff[i]), int(self.voff[i + 1])
        return self.vx[a:b], self.vy[a:b]


class Boundary:
    """The boundary 'floe' of the non-periodic wall call (floe_interactions_all.m:151)."""

    def __init__(self, x, y, box_x, box_y, area, h=0.0, xi=0.0, yi=0.0, u=0.0, v=0.0, ksi=0.0):
        self.x, self.y, self.box_x, self.box_y = f64(x), f64(y), f64(box_x), f64(box_y)
        self.area, self.h, self.xi, self.yi, self.u, self.v, self.ksi = area, h, xi, yi, u, v, ksi

    def struct(self):
        s = SzBoundary()
        s.x, s.y, s.n = _ptr(self.x, c_dp), _ptr(self.y, c_dp), self.x.shape[0]
        s.box_x, s.box_y, s.box_n = _ptr(self.box_x, c_dp), _ptr(self.box_y, c_dp), self.box_x.shape[0]
        s.area, s.h, s.xi, s.yi, s.u, s.v, s.ksi = self.area, self.h, self.xi, self.yi, self.u, self.v, self.ksi
        return s
