"""Host-side mirror of the reference's contact step over the C ABI.

`ContactContext` is a thin object wrapper of the ABI (one context = one GPU = one host thread).
`floe_interactions_all(...)` keeps the call signature of the reference's floe_interactions_all.m:1 and
replaces its lines 9-285 (minus calc_trajectory) with one GPU step: it is what a Python host would call
where Subzero.m:133,301 call the MATLAB function.  The MATLAB-side drop-in (wrapper + mex shim) lives in
subzero_b200/matlab/.
"""
import ctypes as C

import numpy as np

from . import abi
from .abi import FloesSoA, Boundary, SzParams, SzSummary, default_params  # noqa: F401  (re-exported)


class ContactContext:
    def __init__(self, device=0):
        self._h = C.c_void_p()
        abi.check(abi.lib().sz_create(C.byref(self._h), int(device)))
        self.summary = None
        self._n0 = 0

    def close(self):
        if self._h:
            abi.lib().sz_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stepping
    def _finish(self, code, s, allow_pair_errors):
        self.summary = s
        if code in (abi.SZ_ERR_CLIPPER,) and allow_pair_errors:
            return s
        abi.check(code)
        return s

    def step(self, prm, floes, boundary=None, allow_pair_errors=False):
        """One contact step with host buffers in (sz_contact_step)."""
        s = SzSummary()
        fs = floes.struct()
        bs = boundary.struct() if boundary is not None else None
        self._n0 = floes.n
        code = abi.lib().sz_contact_step(self._h, C.byref(prm), C.byref(fs), C.byref(bs) if bs is not None else None, C.byref(s))
        return self._finish(code, s, allow_pair_errors)

    def upload(self, prm, floes, boundary=None):
        fs = floes.struct()
        bs = boundary.struct() if boundary is not None else None
        self._n0 = floes.n
        abi.check(abi.lib().sz_upload(self._h, C.byref(prm), C.byref(fs), C.byref(bs) if bs is not None else None))

    def step_resident(self, allow_pair_errors=False):
        s = SzSummary()
        code = abi.lib().sz_step_resident(self._h, C.byref(s))
        return self._finish(code, s, allow_pair_errors)

    def step_enqueue(self):
        """launch the step without waiting for it (sizes carried over from the step before); step_finish() completes it"""
        abi.check(abi.lib().sz_step_enqueue(self._h))

    def step_finish(self, allow_pair_errors=False):
        """returns the summary, or None when the enqueued step has to be repeated with step_resident()"""
        s = SzSummary()
        code = abi.lib().sz_step_finish(self._h, C.byref(s))
        if code == 1:
            return None
        return self._finish(code, s, allow_pair_errors)

    # ---- results
    def floe_outputs(self, into=None):
        n = self._n0
        o = into if into is not None else {
            "fx": np.empty(n), "fy": np.empty(n), "torque": np.empty(n), "overlap_area": np.empty(n), "stress": np.empty((n, 2, 2)),
            "xi": np.empty(n), "yi": np.empty(n), "alive": np.empty(n, np.uint8), "kill": np.empty(n, np.int32), "transfer": np.empty(n, np.int32)}
        p = abi._ptr
        abi.check(abi.lib().sz_get_floe_outputs(
            self._h, p(o["fx"], abi.c_dp), p(o["fy"], abi.c_dp), p(o["torque"], abi.c_dp), p(o["overlap_area"], abi.c_dp), p(o["stress"], abi.c_dp),
            p(o["xi"], abi.c_dp), p(o["yi"], abi.c_dp), p(o["alive"], abi.c_bp), p(o["kill"], abi.c_ip), p(o["transfer"], abi.c_ip)))
        return o

    def ghosts(self):
        g = self.summary.n - self.summary.n0
        o = {"parent": np.empty(g, np.int32), "floe_num": np.empty(g, np.int32), "x": np.empty(g), "y": np.empty(g)}
        p = abi._ptr
        abi.check(abi.lib().sz_get_ghosts(self._h, p(o["parent"], abi.c_ip), p(o["floe_num"], abi.c_ip), p(o["x"], abi.c_dp), p(o["y"], abi.c_dp)))
        return o

    def ghost_outputs(self):
        """collision_force, collision_torque and OverlapArea the reference leaves in the ghost structs Floe(N0+1:N)
        (floe_interactions_all.m:218-238,137,198)"""
        g = self.summary.n - self.summary.n0
        o = {k: np.empty(g) for k in ("fx", "fy", "torque", "overlap_area")}
        p = abi._ptr
        abi.check(abi.lib().sz_get_ghost_outputs(self._h, p(o["fx"], abi.c_dp), p(o["fy"], abi.c_dp), p(o["torque"], abi.c_dp), p(o["overlap_area"], abi.c_dp)))
        return o

    def pairs(self):
        n = self.summary.n_pairs
        o = {"i": np.empty(n, np.int32), "j": np.empty(n, np.int32), "overlap_state": np.empty(n), "n_regions": np.empty(n, np.int32), "status": np.empty(n, np.int32)}
        p = abi._ptr
        abi.check(abi.lib().sz_get_pairs(self._h, p(o["i"], abi.c_ip), p(o["j"], abi.c_ip), p(o["overlap_state"], abi.c_dp), p(o["n_regions"], abi.c_ip), p(o["status"], abi.c_ip)))
        return o

    def rows(self):
        """(row_off [n+1], rows [K,7]) -- rows[row_off[m]:row_off[m+1]] is Floe(m+1).interactions"""
        off = np.empty(self.summary.n + 1, np.int64)
        rows = np.empty((self.summary.n_rows, 7))
        abi.check(abi.lib().sz_get_rows(self._h, abi._ptr(off, abi.c_lp), abi._ptr(rows, abi.c_dp)))
        return off, rows

    def phase_ms(self):
        """device ms of the last step: ghosts, broad phase, narrow phase, assembly, total"""
        a = (C.c_float * 5)()
        abi.check(abi.lib().sz_get_phase_ms(self._h, a))
        return dict(zip(("ghosts", "broad", "narrow", "assembly", "total"), [float(v) for v in a]))

    def narrow_class_ms(self):
        """device ms of the narrow-phase kernel launch of every size class in the last step, and the pairs each received"""
        a, n = (C.c_float * 5)(), (C.c_int32 * 5)()
        abi.check(abi.lib().sz_get_narrow_class_ms(self._h, a, n))
        return {k: (float(a[i]), int(n[i])) for i, k in enumerate(("C", "S", "T", "M", "L"))}

    def set_stream(self, cuda_stream):
        """route the library's launches to the caller's CUDA stream: an integer handle such as torch's `cuda_stream`, where
        0 means the (legacy) default stream as it does in torch; None restores the context's own stream"""
        if cuda_stream is None:
            h = None
        else:
            h = int(cuda_stream) or 1                 # cudaStreamLegacy
        abi.check(abi.lib().sz_set_stream(self._h, C.c_void_p(h)))

    def stat(self, name):
        v = C.c_int64()
        abi.check(abi.lib().sz_get_stat(self._h, name.encode(), C.byref(v)))
        return v.value

    def set_option(self, name, value):
        abi.check(abi.lib().sz_set_option(self._h, name.encode(), int(value)))

    def clip_polys(self):
        s = self.summary
        ppo = np.empty(s.n_pairs + 1, np.int64)
        pvo = np.empty(s.n_clip_paths + 1, np.int64)
        x = np.empty(s.n_clip_verts, np.int64)
        y = np.empty(s.n_clip_verts, np.int64)
        p = abi._ptr
        abi.check(abi.lib().sz_get_clip_polys(self._h, p(ppo, abi.c_lp), p(pvo, abi.c_lp), p(x, abi.c_lp), p(y, abi.c_lp)))
        return ppo, pvo, x, y

    # ---- integrator half of the timestep (calc_trajectory.m, no-ocean branch)
    def trajectory_init(self, mass, inertia, nz=1000, **fields):
        """fields: alpha dXi_p dYi_p dUi_p dVi_p dalpha_p dksi_p FxOA FyOA torqueOA [n0], c0x c0y [nverts]; missing = zeros
        (c0 = the uploaded c_alpha)"""
        init = abi.SzTrajectoryInit()
        sh, sc = fields.pop("stress_h", None), fields.pop("stress_count", None)
        keep = {"mass": abi.f64(mass), "inertia": abi.f64(inertia)}
        keep.update({k: abi.f64(v) for k, v in fields.items()})
        for k in abi.SzTrajectoryInit.NAMES:
            setattr(init, k, abi._ptr(keep.get(k), abi.c_dp))
        init.nz = int(nz)
        if sh is not None:
            keep["stress_h"] = abi.f64(sh)
            keep["stress_count"] = np.ascontiguousarray(sc, np.int32)
            init.stress_h, init.stress_count = abi._ptr(keep["stress_h"], abi.c_dp), abi._ptr(keep["stress_count"], abi.c_ip)
        self._traj_nz = int(nz)
        abi.check(abi.lib().sz_trajectory_init(self._h, C.byref(init)))

    def stress_history(self):
        """(StressH [n0, nz, 4], StressCount [n0]) of the integrator (calc_trajectory.m:15-19)"""
        n = self._n0
        sh, sc = np.empty((n, self._traj_nz, 4)), np.empty(n, np.int32)
        abi.check(abi.lib().sz_get_stress_history(self._h, abi._ptr(sh, abi.c_dp), abi._ptr(sc, abi.c_ip)))
        return sh, sc

    def trajectory_step(self, dt, HFo=0.0, xo_min=-np.inf, xo_max=np.inf, yo_min=-np.inf, yo_max=np.inf):
        p = abi.SzTrajectoryParams(float(dt), float(HFo), float(xo_min), float(xo_max), float(yo_min), float(yo_max))
        ns, no = C.c_int32(), C.c_int32()
        abi.check(abi.lib().sz_trajectory_step(self._h, C.byref(p), C.byref(ns), C.byref(no)))
        return ns.value

    # ---- weld.m:29-81 / FloeSimplify.m:13-31: the bounding-radius pair searches on the resident floes
    def weld_search(self, Nb, Nx, Ny, xmin, xmax, ymin, ymax):
        """returns (bin [n0-Nb], off [n0-Nb+1], partner): bin number (Binx-1)*Ny+Biny of every floe of the cut list
        Floe(1+Nb:end) (0 = in no bin) and its same-bin partners as 1-based positions in the cut list, ascending"""
        npn = C.c_int64()
        abi.check(abi.lib().sz_pair_search(self._h, 0, int(Nb), int(Nx), int(Ny), float(xmin), float(xmax), float(ymin), float(ymax), 0, None, C.byref(npn)))
        nq = max(0, self._n0 - int(Nb))
        b, off, pt = np.empty(nq, np.int32), np.empty(nq + 1, np.int64), np.empty(npn.value, np.int32)
        abi.check(abi.lib().sz_get_pair_search(self._h, abi._ptr(b, abi.c_ip), abi._ptr(off, abi.c_lp), abi._ptr(pt, abi.c_ip)))
        return b, off, pt

    def simplify_search(self, idx):
        """idx: floe numbers (1-based) about to be simplified; returns (off [count+1], partner) with partners as 1-based
        positions in the resident list, ascending"""
        idx = np.ascontiguousarray(idx, np.int32)
        npn = C.c_int64()
        abi.check(abi.lib().sz_pair_search(self._h, 1, 0, 1, 1, 0.0, 1.0, 0.0, 1.0, idx.shape[0], abi._ptr(idx, abi.c_ip), C.byref(npn)))
        off, pt = np.empty(idx.shape[0] + 1, np.int64), np.empty(npn.value, np.int32)
        abi.check(abi.lib().sz_get_pair_search(self._h, None, abi._ptr(off, abi.c_lp), abi._ptr(pt, abi.c_ip)))
        return off, pt

    # ---- fracture_floe.m:12-52: deformation of the floes about to be fractured (consumer of the contact rows)
    def fracture_deform(self, idx):
        """idx: floe numbers (1-based) of the last contact step's list; returns dict changed xi yi area vert_off cx cy"""
        idx = np.ascontiguousarray(idx, np.int32)
        n = idx.shape[0]
        nc, nv = C.c_int64(), C.c_int64()
        abi.check(abi.lib().sz_fracture_deform(self._h, n, abi._ptr(idx, abi.c_ip), C.byref(nc), C.byref(nv)))
        o = {"changed": np.zeros(n, np.uint8), "xi": np.zeros(n), "yi": np.zeros(n), "area": np.zeros(n), "vert_off": np.zeros(n + 1, np.int64),
             "cx": np.zeros(nv.value), "cy": np.zeros(nv.value)}
        p = abi._ptr
        abi.check(abi.lib().sz_get_fracture_deform(self._h, p(o["changed"], abi.c_bp), p(o["xi"], abi.c_dp), p(o["yi"], abi.c_dp), p(o["area"], abi.c_dp),
                                                   p(o["vert_off"], abi.c_lp), p(o["cx"], abi.c_dp), p(o["cy"], abi.c_dp)))
        assert int(o["changed"].sum()) == nc.value
        return o

    # ---- corners.m:10-88: the contact mask `da` of the floes selected for corner grinding (consumer of the contact rows)
    def corner_mask(self, idx, Nb=0):
        """idx: the selection Floe(~keep) as floe numbers (1-based) of the last contact step's list; Nb: corners.m:54 skips
        the first Nb entries of the selection.  Returns a list of uint8 arrays, one per selected floe (one byte per vertex
        of c_alpha without its closing duplicate)"""
        idx = np.ascontiguousarray(idx, np.int32)
        n = idx.shape[0]
        nv = C.c_int64()
        abi.check(abi.lib().sz_corner_mask(self._h, n, abi._ptr(idx, abi.c_ip), int(Nb), C.byref(nv)))
        off, da = np.zeros(n + 1, np.int64), np.zeros(max(nv.value, 1), np.uint8)
        abi.check(abi.lib().sz_get_corner_mask(self._h, abi._ptr(off, abi.c_lp), abi._ptr(da, abi.c_bp)))
        assert off[n] == nv.value
        return [da[off[k]:off[k + 1]].copy() for k in range(n)]

    # ---- calc_eulerian_data.m: coarse-grid averages of the resident floe state
    EULERIAN_FIELDS = ("u", "v", "du", "dv", "stress", "stressxx", "stressyx", "stressxy", "stressyy", "strainux", "strainvx", "strainuy", "strainvy",
                       "c", "Over", "Mtot", "area", "h")

    def eulerian_data(self, mass, Nx, Ny, box, periodic, overlap_area=None, dUi_p=None, dVi_p=None, stress=None, strain=None):
        """box = (xmin, xmax, ymin, ymax) of c2_boundary; mass [n0]; stress / strain [n0][4]; None = zeros.  Returns a dict of
        (Ny, Nx) arrays, row 0 = the top row (the reference flips y)"""
        n = self._n0
        opt = lambda a, shape: None if a is None else np.ascontiguousarray(np.asarray(a, np.float64).reshape(shape))
        mass = np.ascontiguousarray(np.asarray(mass, np.float64).reshape(n))
        ov, du, dv, st, en = opt(overlap_area, n), opt(dUi_p, n), opt(dVi_p, n), opt(stress, (n, 4)), opt(strain, (n, 4))
        out = np.zeros((18, int(Ny), int(Nx)))
        p, D = abi._ptr, abi.c_dp
        abi.check(abi.lib().sz_eulerian_data(self._h, int(Nx), int(Ny), *(float(b) for b in box), int(bool(periodic)),
                                             p(mass, D), p(ov, D), p(du, D), p(dv, D), p(st, D), p(en, D), p(out, D)))
        return {k: out[i] for i, k in enumerate(self.EULERIAN_FIELDS)}

    # ---- ocean / atmosphere forcing (calc_trajectory.m:94-166) and strain (:224-234)
    def trajectory_set_ocean(self, Xo, Yo, Uocn, Vocn, Uwinds, Vwinds, fCoriolis, turn_angle, rho0=0.0, Cd=0.0, rho_air=0.0, Cd_atm=0.0):
        """Xo [nx], Yo [ny]; the four fields as (ny, nx) arrays like ocean.Uocn / winds.u in MATLAB"""
        Xo, Yo = abi.f64(Xo), abi.f64(Yo)
        cm = lambda a: np.ascontiguousarray(np.asarray(a, np.float64).T)          # column-major ny x nx
        f = [cm(a) for a in (Uocn, Vocn, Uwinds, Vwinds)]
        for a in f:
            assert a.shape == (Xo.shape[0], Yo.shape[0]), "ocean fields must be (ny, nx)"
        o = abi.SzOcean(int(Xo.shape[0]), int(Yo.shape[0]), abi._ptr(Xo, abi.c_dp), abi._ptr(Yo, abi.c_dp), *(abi._ptr(a, abi.c_dp) for a in f),
                        float(fCoriolis), float(turn_angle), float(rho0), float(Cd), float(rho_air), float(Cd_atm))
        abi.check(abi.lib().sz_trajectory_set_ocean(self._h, C.byref(o)))

    def trajectory_set_points(self, X, Y, A):
        """Floe.X, Floe.Y, Floe.A as (n0, npts) arrays"""
        X, Y = np.ascontiguousarray(X, np.float64), np.ascontiguousarray(Y, np.float64)
        A = np.ascontiguousarray(A, np.uint8)
        assert X.shape == Y.shape == A.shape and X.shape[0] == self._n0
        abi.check(abi.lib().sz_trajectory_set_points(self._h, int(X.shape[1]), abi._ptr(X, abi.c_dp), abi._ptr(Y, abi.c_dp), abi._ptr(A, abi.c_bp)))

    def trajectory_ocean_forcing(self, dt, HFo=0.0, xo_min=-np.inf, xo_max=np.inf, yo_min=-np.inf, yo_max=np.inf, do_int=True):
        """call between the contact step and trajectory_step (same parameters); returns (floes evaluated, floes without points)"""
        p = abi.SzTrajectoryParams(float(dt), float(HFo), float(xo_min), float(xo_max), float(yo_min), float(yo_max))
        ne, nn = C.c_int32(), C.c_int32()
        abi.check(abi.lib().sz_trajectory_ocean_forcing(self._h, C.byref(p), int(bool(do_int)), C.byref(ne), C.byref(nn)))
        return ne.value, nn.value

    def trajectory_forcing(self):
        n = self._n0
        o = {"FxOA": np.empty(n), "FyOA": np.empty(n), "torqueOA": np.empty(n), "strain": np.empty((n, 2, 2))}
        abi.check(abi.lib().sz_get_trajectory_forcing(self._h, *(abi._ptr(o[k], abi.c_dp) for k in ("FxOA", "FyOA", "torqueOA", "strain"))))
        return o

    def trajectory_state(self, nverts=0):
        n = self._n0
        o = {k: np.empty(n) for k in ("x", "y", "u", "v", "ksi", "h")}
        o["alive"] = np.empty(n, np.uint8)
        o.update({k: np.empty(n) for k in ("mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p")})
        o["stress"] = np.empty((n, 2, 2))
        o["flags"] = np.empty(n, np.int32)
        o["cax"], o["cay"] = np.empty(nverts), np.empty(nverts)
        p = abi._ptr
        order = ("x", "y", "u", "v", "ksi", "h")
        abi.check(abi.lib().sz_get_trajectory(self._h, *(p(o[k], abi.c_dp) for k in order), p(o["alive"], abi.c_bp),
                                              *(p(o[k], abi.c_dp) for k in ("mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p")),
                                              p(o["stress"], abi.c_dp), p(o["flags"], abi.c_ip), p(o["cax"], abi.c_dp) if nverts else None, p(o["cay"], abi.c_dp) if nverts else None))
        return o

    # ---- stand-alone clip (mex gateway semantics, private/mexclipper.cpp:204-305)
    def clip_batch(self, subjects, clips, methods):
        """subjects/clips: lists of (n,2) int64 arrays; methods: 0 dif / 1 int / 2 xor / 3 uni.
        Returns (status [count], list of lists of (m,2) int64 paths)."""
        cnt = len(subjects)
        soff = np.zeros(cnt + 1, np.int64)
        coff = np.zeros(cnt + 1, np.int64)
        for k in range(cnt):
            soff[k + 1] = soff[k] + len(subjects[k])
            coff[k + 1] = coff[k] + len(clips[k])
        sxy = np.concatenate([np.asarray(s, np.int64).reshape(-1, 2) for s in subjects]) if cnt else np.zeros((0, 2), np.int64)
        cxy = np.concatenate([np.asarray(s, np.int64).reshape(-1, 2) for s in clips]) if cnt else np.zeros((0, 2), np.int64)
        sx, sy = np.ascontiguousarray(sxy[:, 0]), np.ascontiguousarray(sxy[:, 1])
        cx, cy = np.ascontiguousarray(cxy[:, 0]), np.ascontiguousarray(cxy[:, 1])
        m = np.ascontiguousarray(methods, np.int32)
        npth, nvt = C.c_int64(), C.c_int64()
        p = abi._ptr
        abi.check(abi.lib().sz_clip_batch(self._h, cnt, p(m, abi.c_ip), p(soff, abi.c_lp), p(sx, abi.c_lp), p(sy, abi.c_lp),
                                          p(coff, abi.c_lp), p(cx, abi.c_lp), p(cy, abi.c_lp), C.byref(npth), C.byref(nvt)))
        status = np.empty(cnt, np.int32)
        ipo = np.empty(cnt + 1, np.int64)
        pvo = np.empty(npth.value + 1, np.int64)
        ox = np.empty(nvt.value, np.int64)
        oy = np.empty(nvt.value, np.int64)
        abi.check(abi.lib().sz_get_clip_batch(self._h, p(status, abi.c_ip), p(ipo, abi.c_lp), p(pvo, abi.c_lp), p(ox, abi.c_lp), p(oy, abi.c_lp)))
        out = []
        for k in range(cnt):
            out.append([np.stack([ox[pvo[q]:pvo[q + 1]], oy[pvo[q]:pvo[q + 1]]], 1) for q in range(ipo[k], ipo[k + 1])])
        return status, out


# ------------------------------------------------------------------------------------------------
def floes_to_soa(Floe):
    """Flatten a list of floe records (dicts with the reference's field names) to the SoA the ABI takes."""
    n = len(Floe)
    get = lambda k: np.array([float(f[k]) for f in Floe], np.float64)
    voff = np.zeros(n + 1, np.int32)
    xs, ys = [], []
    for i, f in enumerate(Floe):
        ca = np.asarray(f["c_alpha"], np.float64)          # 2 x (n+1), closed
        xs.append(ca[0])
        ys.append(ca[1])
        voff[i + 1] = voff[i] + ca.shape[1]
    vx = np.concatenate(xs) if n else np.zeros(0)
    vy = np.concatenate(ys) if n else np.zeros(0)
    return FloesSoA(get("Xi"), get("Yi"), get("rmax"), get("h"), get("area"), get("Ui"), get("Vi"), get("ksi_ice"),
                    np.array([int(f["alive"]) for f in Floe], np.uint8), voff, vx, vy)


_default_ctx = None


def floe_interactions_all(Floe, floebound, ocean, winds, c2_boundary, dt, HFo, min_floe_size, Nx, Ny, Nb, dissolvedNEW, doInt,
                          COLLISION, PERIODIC, RIDGING, RAFTING, Modulus=None, ctx=None, params=None):
    """The reference's call signature (floe_interactions_all.m:1; `Modulus` is the reference's global, :7).

    Floe: list of dicts with fields c_alpha (2 x n+1, closed), Xi, Yi, rmax, h, area, Ui, Vi, ksi_ice, alive.
    floebound: dict with c (2 x m hole vertices, floe_interactions.m:31), area, h, Xi, Yi, Ui, Vi, ksi_ice -- or None when PERIODIC.
    Writes, per floe, the fields the reference writes in :78-88,136-137,196-198,231,259-263,267-277:
      interactions (K x 7), OverlapArea, collision_force (1x2), collision_torque, Stress-sum input (`StressSum`, 2x2,
      calc_trajectory.m:12-13), alive, Xi, Yi, potentialInteractions (emptied, the reference rmfield-s it at :507);
    and returns (Floe, dissolvedNEW, kill, transfer).  calc_trajectory, ridging/rafting and the kill/fuse tail
    (:281,288-512) stay with the host model.  With PERIODIC, doInt.flag and RIDGING or RAFTING the returned list also holds
    the ghost floes Floe(N0+1:N) that tail indexes (N0 = len(kill)); the MATLAB twin is subzero_b200/matlab/floe_interactions_all.m.
    """
    global _default_ctx
    if Modulus is None:
        raise ValueError("Modulus (the reference's `global Modulus`) must be given")
    c2 = np.asarray(c2_boundary, np.float64)
    Floe = [f for f in Floe if int(f["alive"]) != 0]                        # :12-13
    prm = params if params is not None else default_params()
    prm.Lx, prm.Ly = float(c2[0].max()), float(c2[1].max())                    # :9-10
    prm.modulus, prm.dt, prm.Nb = float(Modulus), float(dt), int(Nb)
    prm.periodic, prm.collision = int(bool(PERIODIC)), int(bool(COLLISION))
    soa = floes_to_soa(Floe)
    bnd = None
    if not PERIODIC and floebound is not None:
        bc = np.asarray(floebound["c"], np.float64)
        bnd = Boundary(bc[0], bc[1], c2[0], c2[1], float(floebound["area"]), float(floebound.get("h", 0)), float(floebound.get("Xi", 0)),
                       float(floebound.get("Yi", 0)), float(floebound.get("Ui", 0)), float(floebound.get("Vi", 0)), float(floebound.get("ksi_ice", 0)))
    if ctx is None:
        if _default_ctx is None:
            _default_ctx = ContactContext(0)
        ctx = _default_ctx
    alive0 = [int(f["alive"]) for f in Floe]
    ctx.step(prm, soa, bnd)
    out = ctx.floe_outputs()
    off, rows = ctx.rows()
    N0 = len(Floe)
    for i, f in enumerate(Floe):
        if i < Nb:
            continue
        f["interactions"] = rows[off[i]:off[i + 1]].copy()
        f["OverlapArea"] = float(out["overlap_area"][i])
        f["collision_force"] = np.array([out["fx"][i], out["fy"][i]])
        f["collision_torque"] = float(out["torque"][i])
        f["StressSum"] = out["stress"][i].copy()
        f["alive"] = int(out["alive"][i])
        f["Xi"], f["Yi"] = float(out["xi"][i]), float(out["yi"][i])
        f["potentialInteractions"] = []
    # ghost structs Floe(N0+1:N) (floe_interactions_all.m:16-66 as data): only the reference's ridging / rafting tail reads
    # them -- Floe(partner) with partner > N0 (:312,327,401,416) -- so they are materialised only when that tail will run.
    # A ghost is its parent as it was when the step began, with the shifted centroid (:34,55), and carries what
    # :76-88,218-238 leave in it.  The caller's tail drops them again (:468, Floe = Floe(1:N0)).
    flag = doInt.get("flag", False) if isinstance(doInt, dict) else bool(getattr(doInt, "flag", False))
    if PERIODIC and flag and (RIDGING or RAFTING) and ctx.summary.n > N0:
        g, go = ctx.ghosts(), ctx.ghost_outputs()
        ghosts = []
        for k in range(ctx.summary.n - N0):
            p = int(g["parent"][k]) - 1
            if p < N0:
                d = dict(Floe[p])
                d["alive"] = alive0[p]
            else:
                d = dict(ghosts[p - N0])                 # a y-ghost of an x-ghost (:49-60 runs over the extended list)
            d["Xi"], d["Yi"] = float(g["x"][k]), float(g["y"][k])
            d["interactions"] = rows[off[N0 + k]:off[N0 + k + 1]].copy()
            d["OverlapArea"] = float(go["overlap_area"][k])
            d["collision_force"] = np.array([go["fx"][k], go["fy"][k]])
            d["collision_torque"] = float(go["torque"][k])
            d["StressSum"] = np.zeros((2, 2))
            d["potentialInteractions"] = []
            ghosts.append(d)
        Floe = Floe + ghosts
    return Floe, dissolvedNEW, out["kill"].copy(), out["transfer"].copy()
