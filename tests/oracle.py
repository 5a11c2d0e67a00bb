"""ctypes binding of the CPU checker oracle/_ref/libsz_oracle.so (TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm)."""
import ctypes as C
import os
import subprocess

import numpy as np

from subzero_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "libsz_oracle.so")
CLIPPER_LIB = os.path.join(ORACLE_DIR, "_ref", "libclipper_ref.so")
REFERENCE = "/root/reference"

_lib = None
_clip = None


def build_if_possible():
    """(Re)build the checker when the reference sources are present (build container); on the GPU box the
    prebuilt files travel with the snapshot."""
    if os.path.isdir(os.path.join(REFERENCE, "private")):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)
    if not os.path.exists(LIB):
        raise RuntimeError("oracle library %s missing and /root/reference not available to build it" % LIB)


def lib():
    global _lib
    if _lib is None:
        build_if_possible()
        l = C.CDLL(LIB)
        l.szo_contact_step.restype = C.c_void_p
        l.szo_contact_step.argtypes = [C.POINTER(abi.SzParams), C.POINTER(abi.SzFloesSoA), C.POINTER(abi.SzBoundary), C.c_int, C.c_int]
        l.szo_free.argtypes = [C.c_void_p]
        l.szo_error.restype = C.c_char_p
        l.szo_error.argtypes = [C.c_void_p]
        l.szo_summary.argtypes = [C.c_void_p, C.POINTER(abi.SzSummary)]
        l.szo_get_floe_outputs.argtypes = [C.c_void_p] + [abi.c_dp] * 7 + [abi.c_bp, abi.c_ip, abi.c_ip]
        l.szo_get_ghosts.argtypes = [C.c_void_p, abi.c_ip, abi.c_ip, abi.c_dp, abi.c_dp]
        l.szo_get_pairs.argtypes = [C.c_void_p, abi.c_ip, abi.c_ip, abi.c_dp, abi.c_ip, abi.c_ip]
        l.szo_get_rows.argtypes = [C.c_void_p, abi.c_lp, abi.c_dp]
        l.szo_get_clip_polys.argtypes = [C.c_void_p, abi.c_lp, abi.c_lp, abi.c_lp, abi.c_lp]
        l.szo_polyclip.restype = C.c_int
        l.szo_polyclip.argtypes = [abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, abi.c_dp, C.c_int, C.c_int, abi.c_dp, abi.c_dp, C.c_int, abi.c_ip, C.c_int]
        l.szo_polyshape_area_centroid.argtypes = [abi.c_dp, abi.c_dp, C.c_int, abi.c_dp]
        l.szo_polyarea.restype = C.c_double
        l.szo_polyarea.argtypes = [abi.c_dp, abi.c_dp, C.c_int]
        l.szo_interx.restype = C.c_int
        l.szo_interx.argtypes = [abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, C.c_int]
        l.szo_inpolygon.argtypes = [abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, abi.c_dp, C.c_int, abi.c_bp]
        l.szo_p_poly_dist.restype = C.c_int
        l.szo_p_poly_dist.argtypes = [abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, abi.c_dp, C.c_int, abi.c_dp]
        l.szo_matlab_int64.restype = C.c_int64
        l.szo_matlab_int64.argtypes = [C.c_double]
        l.szo_floe_interactions.restype = C.c_int
        l.szo_floe_interactions.argtypes = [C.POINTER(abi.SzParams), abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, C.c_int,
                                            abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, C.c_int, abi.c_dp]
        l.szo_hardware_threads.restype = C.c_int
        l.szo_calc_trajectory.restype = None
        l.szo_calc_trajectory.argtypes = [C.c_int] + [C.c_double] * 6 + [C.c_int] + [abi.c_dp] * 4 + [abi.c_bp] + [abi.c_dp] * 7 + [abi.c_bp] + [abi.c_dp] * 12 + [
            abi.c_ip] + [abi.c_dp] * 4 + [abi.c_dp, abi.c_ip, abi.c_dp, abi.c_bp, abi.c_bp]
        l.szo_ocean_forcing.restype = None
        l.szo_ocean_forcing.argtypes = [C.c_int] + [C.c_double] * 6 + [C.c_int, abi.c_bp] + [abi.c_dp] * 9 + [abi.c_ip, abi.c_dp, abi.c_dp, C.c_int, abi.c_dp, abi.c_dp, abi.c_bp,
                                        C.c_int, C.c_int] + [abi.c_dp] * 6 + [C.c_double] * 6 + [abi.c_dp] * 3 + [abi.c_bp, abi.c_bp]
        l.szo_floe_strain.restype = None
        l.szo_floe_strain.argtypes = [C.c_int, abi.c_bp, abi.c_bp] + [abi.c_dp] * 4 + [abi.c_ip, abi.c_dp, abi.c_dp, abi.c_dp]
        l.szo_fracture_deform.restype = C.c_int
        l.szo_fracture_deform.argtypes = [C.POINTER(abi.SzFloesSoA), abi.c_lp, abi.c_dp, C.c_int, abi.c_ip, abi.c_bp, abi.c_dp, abi.c_dp, abi.c_dp, abi.c_lp, abi.c_dp, abi.c_dp, C.c_int64]
        l.szo_calc_eulerian_data.restype = C.c_int
        l.szo_calc_eulerian_data.argtypes = [C.POINTER(abi.SzFloesSoA)] + [abi.c_dp] * 6 + [C.c_int] * 3 + [C.c_double] * 4 + [C.c_int, abi.c_dp]
        l.szo_corner_eligibility.restype = C.c_int
        l.szo_corner_eligibility.argtypes = [C.POINTER(abi.SzFloesSoA), abi.c_lp, abi.c_dp, C.c_int, abi.c_ip, C.c_int, C.c_double, C.c_double, abi.c_dp, abi.c_dp, C.c_int,
                                             abi.c_lp, abi.c_bp, C.c_int64]
        l.szo_weld_search.restype = C.c_int64
        l.szo_weld_search.argtypes = [C.POINTER(abi.SzFloesSoA), C.c_int, C.c_int, C.c_int] + [C.c_double] * 4 + [abi.c_ip, abi.c_lp, abi.c_ip, C.c_int64]
        l.szo_simplify_search.restype = C.c_int64
        l.szo_simplify_search.argtypes = [C.POINTER(abi.SzFloesSoA), C.c_int, abi.c_ip, abi.c_lp, abi.c_ip, C.c_int64]
        _lib = l
    return _lib


def clipper():
    """the UNMODIFIED reference Clipper 6.4.2 behind oracle/clipper_ref_shim.cpp"""
    global _clip
    if _clip is None:
        lib()
        l = C.CDLL(CLIPPER_LIB)
        l.szref_clip.restype = C.c_int
        l.szref_clip.argtypes = [abi.c_lp, abi.c_lp, C.c_int, abi.c_lp, abi.c_lp, C.c_int, C.c_int, abi.c_lp, abi.c_lp, C.c_int, abi.c_ip, C.c_int]
        _clip = l
    return _clip


def ref_clip(subj, clip, method):
    """reference Clipper on (n,2) int64 arrays -> list of (m,2) int64 paths, or None on 'Clipper Error.'"""
    s = np.ascontiguousarray(subj, np.int64).reshape(-1, 2)
    c = np.ascontiguousarray(clip, np.int64).reshape(-1, 2)
    sx, sy, cx, cy = (np.ascontiguousarray(a) for a in (s[:, 0], s[:, 1], c[:, 0], c[:, 1]))
    cap = 4 * (len(s) + len(c)) + 64
    while True:
        ox, oy, off = np.empty(cap, np.int64), np.empty(cap, np.int64), np.empty(cap, np.int32)
        p = abi._ptr
        n = clipper().szref_clip(p(sx, abi.c_lp), p(sy, abi.c_lp), len(s), p(cx, abi.c_lp), p(cy, abi.c_lp), len(c), int(method),
                                 p(ox, abi.c_lp), p(oy, abi.c_lp), cap, p(off, abi.c_ip), cap)
        if n == -2:
            cap *= 4
            continue
        break
    if n < 0:
        return None
    return [np.stack([ox[off[k]:off[k + 1]], oy[off[k]:off[k + 1]]], 1) for k in range(n)]


class OracleStep:
    """One run of the oracle's floe_interactions_all restatement; same getters as ContactContext."""

    def __init__(self, prm, floes, boundary=None, nthreads=None, broad_mode=0):
        l = lib()
        fs = floes.struct()
        bs = boundary.struct() if boundary is not None else None
        if nthreads is None:
            nthreads = max(1, l.szo_hardware_threads())
        self._r = l.szo_contact_step(C.byref(prm), C.byref(fs), C.byref(bs) if bs is not None else None, int(nthreads), int(broad_mode))
        err = l.szo_error(self._r)
        if err:
            raise RuntimeError("oracle: " + err.decode())
        self.summary = abi.SzSummary()
        l.szo_summary(self._r, C.byref(self.summary))
        self._n0 = floes.n

    def __del__(self):
        if getattr(self, "_r", None):
            lib().szo_free(self._r)
            self._r = None

    def floe_outputs(self):
        n = self._n0
        o = {"fx": np.empty(n), "fy": np.empty(n), "torque": np.empty(n), "overlap_area": np.empty(n), "stress": np.empty((n, 2, 2)),
             "xi": np.empty(n), "yi": np.empty(n), "alive": np.empty(n, np.uint8), "kill": np.empty(n, np.int32), "transfer": np.empty(n, np.int32)}
        p = abi._ptr
        lib().szo_get_floe_outputs(self._r, p(o["fx"], abi.c_dp), p(o["fy"], abi.c_dp), p(o["torque"], abi.c_dp), p(o["overlap_area"], abi.c_dp),
                                   p(o["stress"], abi.c_dp), p(o["xi"], abi.c_dp), p(o["yi"], abi.c_dp), p(o["alive"], abi.c_bp), p(o["kill"], abi.c_ip), p(o["transfer"], abi.c_ip))
        return o

    def ghosts(self):
        g = self.summary.n - self.summary.n0
        o = {"parent": np.empty(g, np.int32), "floe_num": np.empty(g, np.int32), "x": np.empty(g), "y": np.empty(g)}
        p = abi._ptr
        lib().szo_get_ghosts(self._r, p(o["parent"], abi.c_ip), p(o["floe_num"], abi.c_ip), p(o["x"], abi.c_dp), p(o["y"], abi.c_dp))
        return o

    def pairs(self):
        n = self.summary.n_pairs
        o = {"i": np.empty(n, np.int32), "j": np.empty(n, np.int32), "overlap_state": np.empty(n), "n_regions": np.empty(n, np.int32), "status": np.empty(n, np.int32)}
        p = abi._ptr
        lib().szo_get_pairs(self._r, p(o["i"], abi.c_ip), p(o["j"], abi.c_ip), p(o["overlap_state"], abi.c_dp), p(o["n_regions"], abi.c_ip), p(o["status"], abi.c_ip))
        return o

    def rows(self):
        off = np.empty(self.summary.n + 1, np.int64)
        rows = np.empty((self.summary.n_rows, 7))
        lib().szo_get_rows(self._r, abi._ptr(off, abi.c_lp), abi._ptr(rows, abi.c_dp))
        return off, rows

    def clip_polys(self):
        s = self.summary
        ppo, pvo = np.empty(s.n_pairs + 1, np.int64), np.empty(s.n_clip_paths + 1, np.int64)
        x, y = np.empty(s.n_clip_verts, np.int64), np.empty(s.n_clip_verts, np.int64)
        p = abi._ptr
        lib().szo_get_clip_polys(self._r, p(ppo, abi.c_lp), p(pvo, abi.c_lp), p(x, abi.c_lp), p(y, abi.c_lp))
        return ppo, pvo, x, y


def calc_trajectory(step, floes, state, dt, HFo=0.0, bounds=(-np.inf, np.inf, -np.inf, np.inf), nz=1000, Nb=0):
    """the oracle's calc_trajectory restatement applied after an oracle contact step.  `state`: dict of per-floe arrays
    (mass inertia alpha dXi_p dYi_p dUi_p dVi_p dalpha_p dksi_p FxOA FyOA torqueOA, c0x c0y, stress_h [n,nz,4],
    stress_count, stress); floes: the FloesSoA the step ran on.  Everything is advanced in place; returns (sacked, unsupported).
    Nb: the timestepping loop is `parfor i=1+Nb:N0` (floe_interactions_all.m:249): the first Nb (topography) floes are left
    untouched -- no wrap, no thinning, no motion, no stress-history slot."""
    o = step.floe_outputs()
    off, _ = step.rows()
    n = floes.n
    has_rows = np.ascontiguousarray((off[1:n + 1] - off[:n]) > 0, np.uint8)
    floes.x[Nb:], floes.y[Nb:], floes.alive[Nb:] = o["xi"][Nb:], o["yi"][Nb:], o["alive"][Nb:]
    sacked, unsup = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    cfx, cfy, ctq, st = (np.ascontiguousarray(o[k]) for k in ("fx", "fy", "torque", "stress"))
    D, B, I = abi.c_dp, abi.c_bp, abi.c_ip
    p = lambda a, t: abi._ptr(a[Nb:], t)          # per-floe arrays from floe Nb on (contiguous views; voff keeps absolute vertex offsets)
    lib().szo_calc_trajectory(n - Nb, float(dt), float(HFo), *(float(b) for b in bounds), int(nz), p(cfx, D), p(cfy, D), p(ctq, D), p(st, D), p(has_rows, B),
                              p(floes.area, D), p(floes.x, D), p(floes.y, D), p(floes.u, D), p(floes.v, D), p(floes.ksi, D), p(floes.h, D), p(floes.alive, B),
                              *(p(state[k], D) for k in ("mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "FxOA", "FyOA", "torqueOA")),
                              p(floes.voff, I), abi._ptr(state["c0x"], D), abi._ptr(state["c0y"], D), abi._ptr(floes.vx, D), abi._ptr(floes.vy, D),
                              p(state["stress_h"], D), p(state["stress_count"], I), p(state["stress"], D), p(sacked, B), p(unsup, B))
    return sacked, unsup


def ocean_forcing(step, floes, state, ocean, points, dt, HFo=0.0, bounds=(-np.inf, np.inf, -np.inf, np.inf), do_int=True):
    """the oracle's restatement of calc_trajectory.m:94-166 on the state BEFORE calc_trajectory() advances it: call it
    right before calc_trajectory() of the same step.  ocean: dict Xo Yo Uocn Vocn Uwinds Vwinds [(ny, nx)] fCoriolis
    turn_angle (+ optional rho0 Cd rho_air Cd_atm); points: (X, Y, A) each (n, npts).  Updates state FxOA/FyOA/torqueOA in
    place; returns (evaluated, no_points)."""
    o = step.floe_outputs()
    n = floes.n
    x, y, alive = np.ascontiguousarray(o["xi"]), np.ascontiguousarray(o["yi"]), np.ascontiguousarray(o["alive"])
    X, Y, A = (np.ascontiguousarray(points[0], np.float64), np.ascontiguousarray(points[1], np.float64), np.ascontiguousarray(points[2], np.uint8))
    Xo, Yo = abi.f64(ocean["Xo"]), abi.f64(ocean["Yo"])
    cm = lambda a: np.ascontiguousarray(np.asarray(a, np.float64).T)
    U, V, Wu, Wv = (cm(ocean[k]) for k in ("Uocn", "Vocn", "Uwinds", "Vwinds"))
    ev, nop = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    p, D, B, I = abi._ptr, abi.c_dp, abi.c_bp, abi.c_ip
    lib().szo_ocean_forcing(n, float(dt), float(HFo), *(float(b) for b in bounds), int(bool(do_int)), p(alive, B), p(x, D), p(y, D), p(floes.u, D), p(floes.v, D), p(floes.ksi, D),
                            p(floes.h, D), p(state["mass"], D), p(floes.area, D), p(state["alpha"], D), p(floes.voff, I), p(floes.vx, D), p(floes.vy, D),
                            int(X.shape[1]), p(X, D), p(Y, D), p(A, B), int(Xo.shape[0]), int(Yo.shape[0]), p(Xo, D), p(Yo, D), p(U, D), p(V, D), p(Wu, D), p(Wv, D),
                            float(ocean["fCoriolis"]), float(ocean["turn_angle"]), float(ocean.get("rho0", 1027.0)), float(ocean.get("Cd", 3e-3)),
                            float(ocean.get("rho_air", 1.2)), float(ocean.get("Cd_atm", 1e-3)),
                            p(state["FxOA"], D), p(state["FyOA"], D), p(state["torqueOA"], D), p(ev, B), p(nop, B))
    return ev, nop


def floe_strain(floes, sacked, strain):
    """calc_trajectory.m:224-234 on the UPDATED state (after calc_trajectory()); strain (n, 2, 2) updated in place"""
    p, D, B, I = abi._ptr, abi.c_dp, abi.c_bp, abi.c_ip
    lib().szo_floe_strain(floes.n, p(floes.alive, B), p(np.ascontiguousarray(sacked, np.uint8), B), p(floes.area, D), p(floes.u, D), p(floes.v, D), p(floes.ksi, D),
                          p(floes.voff, I), p(floes.vx, D), p(floes.vy, D), p(strain, D))


def fracture_deform(step, floes, idx):
    """the oracle's restatement of fracture_floe.m:12-52 for the floes idx (1-based) of the list `floes`, whose contact
    rows come from `step`.  Returns dict changed xi yi area vert_off cx cy."""
    off, rows = step.rows()
    idx = np.ascontiguousarray(idx, np.int32)
    n = idx.shape[0]
    o = {"changed": np.zeros(n, np.uint8), "xi": np.zeros(n), "yi": np.zeros(n), "area": np.zeros(n), "vert_off": np.zeros(n + 1, np.int64)}
    cap = int(floes.vx.shape[0]) * 4 + 64
    cx, cy = np.zeros(cap), np.zeros(cap)
    p, D, B, I, Lp = abi._ptr, abi.c_dp, abi.c_bp, abi.c_ip, abi.c_lp
    off = np.ascontiguousarray(off, np.int64); rows = np.ascontiguousarray(rows)
    view = floes.struct()
    r = lib().szo_fracture_deform(C.byref(view), p(off, Lp), p(rows, D), n, p(idx, I), p(o["changed"], B), p(o["xi"], D), p(o["yi"], D), p(o["area"], D),
                                  p(o["vert_off"], Lp), p(cx, D), p(cy, D), cap)
    assert r >= 0 and r <= cap, r
    o["cx"], o["cy"] = cx[:r].copy(), cy[:r].copy()
    return o


def corner_eligibility(step, floes, idx, Lx, Ly, c2_boundary, Nb=0):
    """the oracle's restatement of the deterministic half of corners.m (the mask `da`); returns a list of uint8 arrays, one
    per selected floe (vertices of c_alpha without the closing duplicate)"""
    off, rows = step.rows()
    off = np.ascontiguousarray(off, np.int64); rows = np.ascontiguousarray(rows)
    idx = np.ascontiguousarray(idx, np.int32)
    bx, by = np.ascontiguousarray(c2_boundary[0], np.float64), np.ascontiguousarray(c2_boundary[1], np.float64)
    cap = int(floes.vx.shape[0]) + 8
    da_off, da = np.zeros(idx.shape[0] + 1, np.int64), np.zeros(cap, np.uint8)
    view = floes.struct()
    p = abi._ptr
    r = lib().szo_corner_eligibility(C.byref(view), p(off, abi.c_lp), p(rows, abi.c_dp), idx.shape[0], p(idx, abi.c_ip), int(Nb), float(Lx), float(Ly),
                                     p(bx, abi.c_dp), p(by, abi.c_dp), bx.shape[0], p(da_off, abi.c_lp), p(da, abi.c_bp), cap)
    assert r >= 0, r
    return [da[da_off[k]:da_off[k + 1]].copy() for k in range(idx.shape[0])]


EULERIAN_FIELDS = ("u", "v", "du", "dv", "stress", "stressxx", "stressyx", "stressxy", "stressyy", "strainux", "strainvx", "strainuy", "strainvy",
                   "c", "Over", "Mtot", "area", "h")


def calc_eulerian_data(floes, mass, Nx, Ny, box, periodic, overlap_area=None, dUi_p=None, dVi_p=None, stress=None, strain=None, Nb=0):
    """the oracle's restatement of calc_eulerian_data.m; box = (xmin, xmax, ymin, ymax) of c2_boundary.  Returns a dict of
    (Ny, Nx) arrays, row 0 = the top row (the reference flips y)."""
    n = floes.n
    z = lambda a, shape: np.zeros(shape) if a is None else np.ascontiguousarray(a, np.float64)
    ov, du, dv, st, en = z(overlap_area, n), z(dUi_p, n), z(dVi_p, n), z(stress, (n, 4)), z(strain, (n, 4))
    mass = np.ascontiguousarray(mass, np.float64)
    out = np.zeros((18, Ny, Nx))
    view = floes.struct()
    p, D = abi._ptr, abi.c_dp
    r = lib().szo_calc_eulerian_data(C.byref(view), p(mass, D), p(ov, D), p(du, D), p(dv, D), p(st, D), p(en, D), int(Nx), int(Ny), int(Nb),
                                     *(float(b) for b in box), int(bool(periodic)), p(out, D))
    assert r == 0, r
    return {k: out[i] for i, k in enumerate(EULERIAN_FIELDS)}


def weld_search(floes, Nb, Nx, Ny, xmin, xmax, ymin, ymax):
    """the oracle's literal restatement of weld.m:25-81 (bins + same-bin radius search); returns (bin, off, partner)"""
    nq = max(0, floes.n - Nb)
    b, off = np.zeros(nq, np.int32), np.zeros(nq + 1, np.int64)
    cap = 64 * max(1, nq) + 1024
    while True:
        pt = np.zeros(cap, np.int32)
        view = floes.struct()
        r = lib().szo_weld_search(C.byref(view), int(Nb), int(Nx), int(Ny), float(xmin), float(xmax), float(ymin), float(ymax),
                                  abi._ptr(b, abi.c_ip), abi._ptr(off, abi.c_lp), abi._ptr(pt, abi.c_ip), cap)
        if r >= 0:
            return b, off, pt[:r]
        cap *= 4


def simplify_search(floes, idx):
    """the oracle's restatement of FloeSimplify.m:13-31 for the floes idx (1-based); returns (off, partner)"""
    idx = np.ascontiguousarray(idx, np.int32)
    off = np.zeros(idx.shape[0] + 1, np.int64)
    cap = 64 * max(1, idx.shape[0]) + 1024
    while True:
        pt = np.zeros(cap, np.int32)
        view = floes.struct()
        r = lib().szo_simplify_search(C.byref(view), idx.shape[0], abi._ptr(idx, abi.c_ip), abi._ptr(off, abi.c_lp), abi._ptr(pt, abi.c_ip), cap)
        if r >= 0:
            return off, pt[:r]
        cap *= 4


def _rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    both_nan = np.isnan(a) & np.isnan(b)
    d = np.abs(np.where(both_inf | both_nan, 0.0, a - b))
    scale = np.maximum(np.abs(np.where(np.isfinite(b), b, 0.0)), 1e-300)
    return d / scale


def _elem_err(g, r, floor_frac):
    """per-element relative error |g - r| / max(|r|, floor) with floor = floor_frac * (largest finite magnitude of r): a small
    entry that is entirely wrong fails even next to a large one; only entries below the floor (cancellation residues) are
    judged on the absolute scale.  Non-finite entries must match exactly."""
    g, r = np.asarray(g, float), np.asarray(r, float)
    fin = np.isfinite(r)
    scale = max(np.abs(r[fin]).max(initial=0.0), 1e-300)
    same_nonfinite = (~fin) & ((g == r) | (np.isnan(g) & np.isnan(r)))
    den = np.maximum(np.abs(np.where(fin, r, 0.0)), floor_frac * scale)
    return np.nan_to_num(np.where(same_nonfinite, 0.0, np.abs(g - r)) / den, nan=np.inf)


def compare_steps(got, ref, rtol=1e-9, check_polys=True, bit_exact=False, floor_frac=1e-6):
    """Parity of a product step (ContactContext) with an oracle step: integer/index outputs bit-exact,
    FP64 outputs within rtol relative PER ELEMENT (BASELINE.json north_star: 1e-9; entries smaller than floor_frac of their
    column's largest magnitude are measured against that floor).  bit_exact: additionally demand that rows and per-floe
    outputs are bit-identical (the design's claim for Voronoi fields: same operation order, no FMA).  Returns a dict of
    measured maxima; raises AssertionError naming the first mismatch."""
    gs, rs = got.summary, ref.summary
    assert (gs.n0, gs.n) == (rs.n0, rs.n), "extended list size: got (%d,%d) ref (%d,%d)" % (gs.n0, gs.n, rs.n0, rs.n)
    gg, rg = got.ghosts(), ref.ghosts()
    for k in ("parent", "floe_num"):
        assert np.array_equal(gg[k], rg[k]), "ghost %s differs" % k
    for k in ("x", "y"):
        assert np.array_equal(gg[k], rg[k], equal_nan=True), "ghost centroid %s differs" % k
    assert gs.n_pairs == rs.n_pairs, "candidate pairs: got %d ref %d" % (gs.n_pairs, rs.n_pairs)
    gp, rp = got.pairs(), ref.pairs()
    for k in ("i", "j", "status", "n_regions"):
        if not np.array_equal(gp[k], rp[k]):
            bad = int(np.flatnonzero(gp[k] != rp[k])[0])
            raise AssertionError("pair list field %s differs first at pair %d (i=%d j=%d): got %d ref %d" % (k, bad, rp["i"][bad], rp["j"][bad], gp[k][bad], rp[k][bad]))
    assert np.array_equal(gp["overlap_state"], rp["overlap_state"]), "overlap_state differs"
    out = {"pairs": int(gs.n_pairs), "rows": int(gs.n_rows)}
    if check_polys:
        assert (gs.n_clip_paths, gs.n_clip_verts) == (rs.n_clip_paths, rs.n_clip_verts), "clip polygon counts differ: got (%d,%d) ref (%d,%d)" % (
            gs.n_clip_paths, gs.n_clip_verts, rs.n_clip_paths, rs.n_clip_verts)
        for a, b, nm in zip(got.clip_polys(), ref.clip_polys(), ("pair_path_off", "path_vert_off", "x", "y")):
            assert np.array_equal(a, b), "Clipper int64 polygons differ in %s" % nm
        out["clip_verts"] = int(gs.n_clip_verts)
    assert gs.n_rows == rs.n_rows, "row count: got %d ref %d" % (gs.n_rows, rs.n_rows)
    goff, grow = got.rows()
    roff, rrow = ref.rows()
    assert np.array_equal(goff, roff), "row offsets differ"
    assert np.array_equal(grow[:, 0], rrow[:, 0]), "row partner ids differ"
    mx = 0.0
    for col, nm in ((1, "Fx"), (2, "Fy"), (3, "Px"), (4, "Py"), (5, "torque"), (6, "overlap")):
        if len(rrow):
            e = _elem_err(grow[:, col], rrow[:, col], floor_frac)
            assert e.max(initial=0.0) <= rtol, "rows column %s: max per-element rel err %.3e (row %d)" % (nm, e.max(), int(e.argmax()))
            mx = max(mx, float(e.max(initial=0.0)))
    out["rows_max_rel"] = mx
    go, ro = got.floe_outputs(), ref.floe_outputs()
    for k in ("alive", "kill", "transfer"):
        assert np.array_equal(go[k], ro[k]), "per-floe %s differs" % k
    for k in ("xi", "yi"):
        assert np.array_equal(go[k], ro[k], equal_nan=True), "wrapped centroid %s differs" % k
    for k in ("fx", "fy", "torque", "overlap_area", "stress"):
        r = np.asarray(ro[k])
        g = np.asarray(go[k])
        e = _elem_err(g, r, floor_frac)        # sums of rows can cancel: the floor is what keeps a residue of 1e-20 N from being judged relatively
        assert e.max(initial=0.0) <= rtol, "per-floe %s: max per-element rel err %.3e (floe %d)" % (k, e.max(), int(e.argmax()) // max(1, r[0].size if r.ndim > 1 else 1))
        out[k + "_max_rel"] = float(e.max(initial=0.0))
        out[k + "_bit_exact"] = bool(np.array_equal(g, r, equal_nan=True))
        if bit_exact:
            assert out[k + "_bit_exact"], "per-floe %s is not bit-identical (max rel err %.3e)" % (k, e.max())
    assert gs.collision_count == rs.collision_count, "collision count: got %r ref %r" % (gs.collision_count, rs.collision_count)
    assert gs.n_pairs_force == rs.n_pairs_force and gs.n_clipper_fail == rs.n_clipper_fail
    out["rows_bit_exact"] = bool(np.array_equal(grow, rrow, equal_nan=True))
    if bit_exact:
        assert out["rows_bit_exact"], "contact rows are not bit-identical"
    return out
