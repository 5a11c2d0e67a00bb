"""The MATLAB boundary, executed: the unmodified mex gateways of subzero_b200/matlab/ (sz_contact_mex.cpp, the one call per
timestep behind floe_interactions_all.m, and sz_resident_mex.cpp, the device-resident timestep) are compiled against a stand-in
for MATLAB's mx/mex runtime (tests/host/mex_mock) and their mexFunction is called with the structs sz_contact_step.m /
sz_resident_timestep.m build.  MATLAB itself is not installed, so this is as close to `out = sz_contact_mex(prm, soa)` as the
image allows: argument checking and error raising (modelled on private/mexclipper.cpp:22-41,303-304), marshalling in and out
(column-major doubles, 0/1 doubles for logicals, 7 x K rows), the context that persists between calls and mexAtExit.

CPU tests: the stand-in itself, argument errors, and that the gateways fail loudly without a GPU (no CPU fallback).
GPU tests: every field a gateway returns equals, bit for bit, what the ctypes path (ContactContext, the one all the other parity
tests check against the oracle) returns for the same input.
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mexmock                      # noqa: E402
import scenarios                    # noqa: E402
import subzero_b200 as sz           # noqa: E402
from mexmock import MexError, col   # noqa: E402


@pytest.fixture(scope="module")
def ml():
    s = mexmock.Session()
    yield s
    s.clear()


def _has_gpu():
    import torch
    return torch.cuda.is_available()


# ------------------------------------------------------------------------------------------------------------ CPU
def test_stand_in_runtime_round_trips_matlab_values(ml):
    """doubles are column-major, structs keep their fields, a missing field reads as NULL, N x 1 columns come back as such"""
    a = np.arange(12.0).reshape(3, 4)
    p = ml.to_mx({"a": a, "v": np.array([1.0, 2.0, 3.0]), "s": 7.5, "inner": {"k": [4.0]}})
    back = ml.from_mx(p)
    assert np.array_equal(back["a"], a) and back["a"].flags.f_contiguous
    assert back["v"].shape == (3, 1) and np.array_equal(col(back["v"]), [1.0, 2.0, 3.0])
    assert back["s"].shape == (1, 1) and back["s"][0, 0] == 7.5 and back["inner"]["k"][0, 0] == 4.0
    assert not ml.mm.mxGetField(p, 0, b"absent") and not ml.mm.mxGetField(p, 1, b"a")
    pr = ml.mm.mxGetPr(ml.mm.mxGetField(p, 0, b"a"))
    assert [pr[k] for k in range(4)] == [0.0, 4.0, 8.0, 1.0]                   # column-major: a(1,1) a(2,1) a(3,1) a(1,2)


def test_contact_gateway_checks_its_arguments_before_touching_the_device(ml):
    """usage, output count, missing / non-double / short fields: raised as MATLAB errors with the gateway's identifiers, on any
    machine (the checks run before sz_create)"""
    prm, soa = sz.voronoi_field(50, seed=1)
    P, S = mexmock.prm_struct(prm), mexmock.soa_struct(soa)
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", P)
    assert e.value.identifier == "subzero_b200:arg" and "usage" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", P, S, nlhs=2)
    assert "one output" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", {k: v for k, v in P.items() if k != "modulus"}, S)
    assert "missing parameter 'modulus'" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", P, {k: v for k, v in S.items() if k != "rmax"})
    assert "missing field 'rmax'" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", P, dict(S, h=S["h"][:-1]))
    assert "field 'h' is too short" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", P, dict(S, vx=S["vx"][:10]))
    assert "field 'vx' is too short" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", P, dict(S, u="not numbers"))
    assert "must be a real double array" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_contact_mex", P, 3.0)
    assert "must be a struct" in e.value.message
    with pytest.raises(MexError) as e:                                        # a boundary struct without its hole vertices
        ml.call("sz_contact_mex", P, S, {"box_x": [1.0, 2.0, 3.0], "box_y": [1.0, 2.0, 3.0], "area": 1.0})
    assert "missing field 'x'" in e.value.message
    with pytest.raises(MexError) as e:
        ml.call("sz_resident_mex", 1.0)
    assert e.value.identifier == "subzero_b200:arg" and "usage" in e.value.message


def test_gateways_fail_loudly_without_a_gpu(ml):
    """no device, no answer: the error of sz_create reaches MATLAB as subzero_b200:error (there is no CPU fallback)"""
    if _has_gpu():
        pytest.skip("a GPU is present")
    prm, soa = sz.voronoi_field(50, seed=1)
    for call in (lambda: ml.call("sz_contact_mex", mexmock.prm_struct(prm), mexmock.soa_struct(soa)),
                 lambda: ml.call("sz_resident_mex", "upload", mexmock.prm_struct(prm), mexmock.soa_struct(soa)),
                 lambda: ml.call("sz_resident_mex", "step")):
        with pytest.raises(MexError) as e:
            call()
        assert e.value.identifier == "subzero_b200:error" and "no CUDA device" in e.value.message and "no CPU fallback" in e.value.message


# ------------------------------------------------------------------------------------------------------------ GPU
def _same(a, b, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a, b, equal_nan=True), what


def _check_contact_out(out, ctx, s, n):
    o = ctx.floe_outputs()
    for k in ("fx", "fy", "torque", "overlap_area", "xi", "yi", "alive", "kill", "transfer"):
        _same(col(out[k]), o[k], k)
    assert out["stress"].shape == (4, n)
    _same(out["stress"].T.reshape(n, 2, 2), o["stress"], "stress")
    off, rows = ctx.rows()
    _same(col(out["row_off"]), off, "row_off")
    assert out["rows"].shape == (7, rows.shape[0])
    _same(out["rows"].T, rows, "rows")
    assert out["n_ext"][0, 0] == s.n and out["n_pairs"][0, 0] == s.n_pairs and out["collision_count"][0, 0] == s.collision_count
    g, go = ctx.ghosts(), ctx.ghost_outputs()
    assert out["ghost_parent"].shape == (s.n - s.n0, 1)
    _same(col(out["ghost_parent"]), g["parent"], "ghost_parent")
    for k, src in (("ghost_x", g["x"]), ("ghost_y", g["y"]), ("ghost_fx", go["fx"]), ("ghost_fy", go["fy"]), ("ghost_torque", go["torque"]),
                   ("ghost_overlap_area", go["overlap_area"])):
        _same(col(out[k]), src, k)


@pytest.mark.gpu
def test_contact_gateway_returns_what_the_c_abi_returns(ml):
    """out = sz_contact_mex(prm, soa [, bnd]) on a periodic field with ghost partners, on the same field a step later (the
    context persists between calls), and on a walled field: every output field against the ctypes path"""
    prm, soa = sz.voronoi_field(3000, seed=3)
    with sz.ContactContext(0) as ctx:
        for rep in range(2):
            out = ml.call("sz_contact_mex", mexmock.prm_struct(prm), mexmock.soa_struct(soa))
            s = ctx.step(prm, soa)
            assert s.n > s.n0 and s.n_pairs > 1000 and s.n_rows > 1000
            _check_contact_out(out, ctx, s, soa.n)
            soa.x[:] = soa.x + 3.0 * np.cos(np.arange(soa.n))                 # new positions, same outlines: a second, different call
        prm2, soa2 = sz.voronoi_field(1500, seed=9)
        prm2.periodic = 0
        soa2.x[:20] += np.sign(soa2.x[:20]) * 1500.0
        bnd = scenarios.soa_and_boundary([], prm2, periodic=False)[1]
        out = ml.call("sz_contact_mex", mexmock.prm_struct(prm2), mexmock.soa_struct(soa2), mexmock.bnd_struct(bnd))
        s = ctx.step(prm2, soa2, bnd)
        assert np.isinf(out["rows"][0]).sum() > 10 and out["ghost_parent"].size == 0
        _check_contact_out(out, ctx, s, soa2.n)
        # parameters the struct may override reach the device (a stiffer friction coefficient changes the rows)
        out_mu = ml.call("sz_contact_mex", dict(mexmock.prm_struct(prm2), mu=0.35), mexmock.soa_struct(soa2), mexmock.bnd_struct(bnd))
        prm2.mu = 0.35
        s = ctx.step(prm2, soa2, bnd)
        _check_contact_out(out_mu, ctx, s, soa2.n)
        assert not np.array_equal(out_mu["rows"], out["rows"])


@pytest.mark.gpu
def test_resident_gateway_runs_the_timestep_like_the_c_abi(ml):
    """sz_resident_mex: upload, step, floe_outputs, rows, trajectory_init / trajectory_step / state over coupled steps, the weld /
    FloeSimplify searches, corner mask, fracture deformation and the coarse-grid averages, each against the ctypes path"""
    prm, soa = sz.voronoi_field(2500, seed=12, inflate=0.04)
    rng = np.random.default_rng(2)
    soa.ksi[:] = rng.normal(0, 2e-5, soa.n)
    n, nv = soa.n, soa.vx.shape[0]
    mass = soa.area * soa.h * 920.0
    inertia = mass * soa.rmax ** 2 / 4
    tp = {"dt": prm.dt, "HFo": 1e-4}
    with sz.ContactContext(0) as ctx:
        assert ml.call("sz_resident_mex", "upload", mexmock.prm_struct(prm), mexmock.soa_struct(soa)) is None
        ctx.upload(prm, soa)
        ml.call("sz_resident_mex", "trajectory_init", {"mass": mass, "inertia": inertia, "dXi_p": soa.u, "dYi_p": soa.v, "nz": 3.0})
        ctx.trajectory_init(mass, inertia, nz=3, dXi_p=soa.u, dYi_p=soa.v)
        for step in range(3):
            sm = ml.call("sz_resident_mex", "step")
            s = ctx.step_resident()
            assert (sm["n_ext"][0, 0], sm["n_pairs"][0, 0], sm["n_rows"][0, 0], sm["collision_count"][0, 0]) == (s.n, s.n_pairs, s.n_rows, s.collision_count)
            fo, o = ml.call("sz_resident_mex", "floe_outputs"), ctx.floe_outputs()
            for k in ("fx", "fy", "torque", "overlap_area", "xi", "yi", "alive", "kill", "transfer"):
                _same(col(fo[k]), o[k], k)
            _same(fo["stress"].T.reshape(n, 2, 2), o["stress"], "stress")
            r, (off, rows) = ml.call("sz_resident_mex", "rows"), ctx.rows()
            _same(col(r["row_off"]), off, "row_off")
            _same(r["rows"].T, rows, "rows")
            if step == 1:
                # consumers of the contact rows
                sel = np.flatnonzero(np.diff(off[:n + 1]) > 0)[:200] + 1
                cm, cm_ref = ml.call("sz_resident_mex", "corner_mask", sel.astype(np.float64), 0.0), ctx.corner_mask(sel)
                _same(col(cm["da"]), np.concatenate(cm_ref), "corner mask")
                _same(np.diff(col(cm["da_off"])), [len(m) for m in cm_ref], "corner mask offsets")
                fd, fd_ref = ml.call("sz_resident_mex", "fracture_deform", sel.astype(np.float64)), ctx.fracture_deform(sel)
                for k in ("changed", "xi", "yi", "area", "vert_off", "cx", "cy"):
                    _same(col(fd[k]), fd_ref[k], "fracture_deform " + k)
            ns = ml.call("sz_resident_mex", "trajectory_step", tp)
            assert ns[0, 0] == ctx.trajectory_step(prm.dt, 1e-4)
            st, st_ref = ml.call("sz_resident_mex", "state"), ctx.trajectory_state(nverts=nv)
            for k in ("x", "y", "u", "v", "ksi", "h", "alive", "mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "flags", "cax", "cay"):
                _same(col(st[k]), st_ref[k], "state " + k)
            _same(st["stress"].T.reshape(n, 2, 2), st_ref["stress"], "state stress")
            assert np.abs(st_ref["alpha"]).max() > 0                         # the outlines did rotate
        L = prm.Lx
        # weld.m:29-81 and FloeSimplify.m:13-31 searches (1-based partners)
        g = {"Nb": 0.0, "Nx": 3.0, "Ny": 4.0, "xmin": float(st_ref["x"].min()), "xmax": float(st_ref["x"].max()), "ymin": float(st_ref["y"].min()), "ymax": float(st_ref["y"].max())}
        w = ml.call("sz_resident_mex", "weld_search", g)
        b, off, pt = ctx.weld_search(0, 3, 4, g["xmin"], g["xmax"], g["ymin"], g["ymax"])
        assert pt.size > 0
        _same(col(w["bin"]), b, "weld bin"); _same(col(w["off"]), off, "weld off"); _same(col(w["partner"]), pt, "weld partner")
        idx = np.arange(1, n + 1, 7)
        sp = ml.call("sz_resident_mex", "simplify_search", idx.astype(np.float64))
        off, pt = ctx.simplify_search(idx)
        _same(col(sp["off"]), off, "simplify off"); _same(col(sp["partner"]), pt, "simplify partner")
        # calc_eulerian_data.m: Ny x Nx column-major planes
        kw = dict(overlap_area=rng.uniform(0, 1e5, n), dUi_p=rng.normal(0, 1e-3, n), dVi_p=rng.normal(0, 1e-3, n), stress=rng.normal(0, 1e3, (n, 4)), strain=rng.normal(0, 1e-6, (n, 4)))
        e = ml.call("sz_resident_mex", "eulerian_data", {"Nx": 7.0, "Ny": 5.0, "xmin": -L, "xmax": L, "ymin": -L, "ymax": L, "periodic": 1.0},
                    {"mass": mass, "overlap_area": kw["overlap_area"], "dUi_p": kw["dUi_p"], "dVi_p": kw["dVi_p"], "stress": kw["stress"].T, "strain": kw["strain"].T})
        e_ref = ctx.eulerian_data(mass, 7, 5, (-L, L, -L, L), True, **kw)
        for k in sz.ContactContext.EULERIAN_FIELDS:
            assert e[k].shape == (5, 7)
            _same(e[k], e_ref[k], "eulerian " + k)
        with pytest.raises(MexError) as err:
            ml.call("sz_resident_mex", "no_such_command")
        assert "unknown command 'no_such_command'" in err.value.message
    assert ml.clear() >= 1                                                    # `clear mex` releases the gateways' contexts
