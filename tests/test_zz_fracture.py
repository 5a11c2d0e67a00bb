"""Physical_Processes/fracture_floe.m:12-52 (first consumer of the contact rows): the deformation a
floe receives from its deepest contact before it is fractured.  The oracle's restatement against a hand-derived case
(CPU), and the device path (sz_fracture_deform) against the oracle on overlapping Voronoi fields (GPU)."""
import numpy as np
import pytest

import oracle
import scenarios
import subzero_b200 as sz


def two_squares(gap):
    """two 2 km squares, the second shifted by `gap` in x: an overlap strip of width 2000 - gap"""
    sq = np.array([[-1000.0, -1000.0], [-1000.0, 1000.0], [1000.0, 1000.0], [1000.0, -1000.0]])       # clockwise
    fl = [scenarios.floe_from_polygon(sq), scenarios.floe_from_polygon(sq + [gap, 0.0])]
    soa = sz.floes_to_soa(fl)
    prm = sz.default_params(Lx=1e5, Ly=1e5, modulus=1e7, dt=10.0, periodic=1, collision=1)
    return prm, soa


def test_oracle_fracture_deform_known_answer():
    """squares overlapping in a 100 m x 2000 m strip: region centroid (950, 0), 50 m from its outline; the partner is
    pushed 25 m along the contact force (-x on floe 1) and subtracted: floe 1 keeps [-1000, 875] x [-1000, 1000]
    = 93.75 % of its area, so it is replaced by that rectangle about its new centroid (-62.5, 0)"""
    prm, soa = two_squares(1900.0)
    step = oracle.OracleStep(prm, soa)
    off, rows = step.rows()
    assert off[1] - off[0] == 1 and rows[0, 0] == 2 and rows[0, 1] < 0 and rows[0, 2] == 0          # one contact row, force along -x
    assert rows[0, 6] == pytest.approx(100.0 * 2000.0, rel=1e-9)
    d = oracle.fracture_deform(step, soa, [1, 2])
    assert list(d["changed"]) == [1, 1]
    assert d["xi"][0] == pytest.approx(-62.5, abs=1e-6) and d["yi"][0] == pytest.approx(0.0, abs=1e-6) and d["area"][0] == pytest.approx(1875.0 * 2000.0, rel=1e-9)
    n0 = int(d["vert_off"][1])
    assert n0 == 4
    assert sorted(np.round(d["cx"][:n0] + d["xi"][0], 6)) == [-1000.0, -1000.0, 875.0, 875.0]
    assert sorted(np.round(d["cy"][:n0] + d["yi"][0], 6)) == [-1000.0, -1000.0, 1000.0, 1000.0]
    # the partner is deformed symmetrically: it keeps [1025, 2900] in world x
    n1 = int(d["vert_off"][2]) - n0
    assert n1 == 4 and sorted(np.round(d["cx"][n0:] + d["xi"][1], 6)) == [1025.0, 1025.0, 2900.0, 2900.0]
    # a deeper overlap removes more than 10 % of the floe: no deformation (:44)
    prm, soa = two_squares(1500.0)
    d = oracle.fracture_deform(oracle.OracleStep(prm, soa), soa, [1])
    assert list(d["changed"]) == [0] and d["area"][0] == soa.area[0] and d["vert_off"][1] == 0
    # no contact at all
    prm, soa = two_squares(2500.0)
    d = oracle.fracture_deform(oracle.OracleStep(prm, soa), soa, [1, 2])
    assert list(d["changed"]) == [0, 0]


def check_device_against_oracle(ctx, prm, soa, idx):
    ctx.step(prm, soa, allow_pair_errors=True)
    got = ctx.fracture_deform(idx)
    ref_step = oracle.OracleStep(prm, soa, broad_mode=1)
    want = oracle.fracture_deform(ref_step, soa, idx)
    assert np.array_equal(got["changed"], want["changed"])
    assert np.array_equal(got["vert_off"], want["vert_off"])
    for k in ("xi", "yi", "area", "cx", "cy"):
        scale = max(np.abs(want[k]).max(initial=0.0), 1e-300)
        assert np.abs(got[k] - want[k]).max(initial=0.0) <= 1e-9 * scale, k
    return want


@pytest.mark.gpu
def test_device_fracture_deform_matches_oracle():
    """the hand-derived square case, then every floe of an overlapping Voronoi field (shallow and deep overlaps, ghost
    partners across the periodic boundary, floes without contacts) and a field of real concave shapes"""
    with sz.ContactContext(0) as ctx:
        prm, soa = two_squares(1900.0)
        w = check_device_against_oracle(ctx, prm, soa, [1, 2])
        assert list(w["changed"]) == [1, 1]
        n_changed = 0
        for inflate, seed in ((0.02, 61), (0.1, 62)):
            prm, soa = sz.voronoi_field(3000, seed=seed, inflate=inflate)
            w = check_device_against_oracle(ctx, prm, soa, np.arange(1, soa.n + 1))
            n_changed += int(w["changed"].sum())
            assert 0 < w["changed"].sum() < soa.n
        assert n_changed > 1000
        # the reference's order: contact step -> calc_trajectory -> fracture: rows of the old positions, outlines of the new ones
        prm, soa = sz.voronoi_field(2000, seed=63, inflate=0.05)
        prm.dt = 10.0
        ref_soa = sz.FloesSoA(soa.x.copy(), soa.y.copy(), soa.rmax.copy(), soa.h.copy(), soa.area.copy(), soa.u.copy(), soa.v.copy(), soa.ksi.copy(), soa.alive.copy(),
                              soa.voff.copy(), soa.vx.copy(), soa.vy.copy())
        mass = soa.area * soa.h * 920.0
        st = {k: np.zeros(soa.n) for k in ("alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "FxOA", "FyOA", "torqueOA")}
        st.update(mass=mass.copy(), inertia=mass * soa.rmax ** 2 / 4, c0x=soa.vx.copy(), c0y=soa.vy.copy(), stress_h=np.zeros((soa.n, 2, 4)),
                  stress_count=np.ones(soa.n, np.int32), stress=np.zeros((soa.n, 2, 2)))
        ctx.upload(prm, soa)
        ctx.trajectory_init(st["mass"], st["inertia"], nz=2)
        ctx.step_resident()
        ctx.trajectory_step(prm.dt)
        idx = np.arange(1, soa.n + 1, 2)
        got = ctx.fracture_deform(idx)
        ref_step = oracle.OracleStep(prm, ref_soa, broad_mode=1)
        oracle.calc_trajectory(ref_step, ref_soa, st, prm.dt, nz=2)
        want = oracle.fracture_deform(ref_step, ref_soa, idx)
        assert np.array_equal(got["changed"], want["changed"]) and np.array_equal(got["vert_off"], want["vert_off"]) and want["changed"].sum() > 100
        for k in ("xi", "yi", "area", "cx", "cy"):
            assert np.abs(got[k] - want[k]).max() <= 1e-9 * np.abs(want[k]).max(), k
        # real floe shapes (7..591 vertices, concave), tiled so that neighbours overlap
        polys, _, modulus = scenarios.floe_shapes()
        rng = np.random.default_rng(3)
        fl, k = [], 0
        for gx in range(6):
            for gy in range(6):
                v = polys[(7 * k) % len(polys)]
                f = scenarios.floe_from_polygon(v, u=rng.uniform(-.1, .1), v=rng.uniform(-.1, .1))
                r = f["rmax"]
                f["Xi"], f["Yi"] = gx * 1.2e4 + rng.uniform(-1, 1) * 1e3, gy * 1.2e4 + rng.uniform(-1, 1) * 1e3
                fl.append(f); k += 1
        soa = sz.floes_to_soa(fl)
        prm = sz.default_params(Lx=2e5, Ly=2e5, modulus=modulus, dt=10.0, periodic=1, collision=1)
        w = check_device_against_oracle(ctx, prm, soa, np.arange(1, soa.n + 1))
        assert w["changed"].sum() > 0


def test_oracle_corner_eligibility_known_answers():
    """corners.m:54-88, the deterministic mask `da`: the vertex nearest to each contact point (first of equals) and the
    vertices lying in (or on) a partner's outline; against the wall, the vertices outside c2_boundary"""
    prm, soa = two_squares(1900.0)                       # overlap strip x in [900, 1000]; contact point (950, 0)
    c2, _ = scenarios.domain(1e5, 1e5)
    step = oracle.OracleStep(prm, soa)
    da = oracle.corner_eligibility(step, soa, [1, 2], prm.Lx, prm.Ly, c2)
    # floe 1: vertices (-1000,-1000) (-1000,1000) (1000,1000) (1000,-1000): the two on x = 1000 lie on the partner's outline
    assert list(da[0]) == [0, 0, 1, 1]
    # floe 2: vertices (900,-1000) (900,1000) (2900,1000) (2900,-1000): the two on x = 900 are inside floe 1
    assert list(da[1]) == [1, 1, 0, 0]
    # no contact: nothing is eligible; the first Nb floes of the selection are skipped
    prm, soa = two_squares(2500.0)
    assert all(d.sum() == 0 for d in oracle.corner_eligibility(oracle.OracleStep(prm, soa), soa, [1, 2], prm.Lx, prm.Ly, c2))
    prm, soa = two_squares(1900.0)
    da = oracle.corner_eligibility(oracle.OracleStep(prm, soa), soa, [1, 2], prm.Lx, prm.Ly, c2, Nb=1)
    assert da[0].sum() == 0 and list(da[1]) == [1, 1, 0, 0]
    # a floe across the east wall of a non-periodic domain: wall rows (partner Inf) flag the vertices outside c2_boundary
    L = 5000.0
    c2, fb = scenarios.domain(L, L)
    sq = np.array([[-1000.0, -1000.0], [-1000.0, 1000.0], [1000.0, 1000.0], [1000.0, -1000.0]]) + [L - 200.0, 0.0]
    soa = sz.floes_to_soa([scenarios.floe_from_polygon(sq)])
    prm = sz.default_params(Lx=L, Ly=L, modulus=1e7, dt=10.0, periodic=0, collision=1)
    bnd = sz.Boundary(fb["c"][0], fb["c"][1], c2[0], c2[1], fb["area"], fb["h"])
    step = oracle.OracleStep(prm, soa, bnd)
    off, rows = step.rows()
    assert off[1] > 0 and np.isinf(rows[0, 0])
    da = oracle.corner_eligibility(step, soa, [1], L, L, c2)
    assert list(da[0]) == [0, 0, 1, 1]                   # the two vertices at x = L + 800


# ---- corners.m:10-88, the contact mask: the product's per-floe core on the host (CPU) and the device path (GPU)
def _port_corner_mask(step, floes, idx, Lx, Ly, c2_boundary, Nb=0):
    """sz_corners.cuh compiled for the host (tests/host/pair_host.cpp), one lane per floe"""
    import ctypes as C
    import os
    from subzero_b200 import abi
    l = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "libpair_host.so"))
    l.szport_corner_mask.restype = C.c_int
    l.szport_corner_mask.argtypes = [C.POINTER(abi.SzFloesSoA), abi.c_lp, abi.c_dp, C.c_int, abi.c_ip, C.c_int, C.c_double, C.c_double, abi.c_dp, abi.c_dp, C.c_int,
                                     abi.c_lp, abi.c_bp, C.c_int64]
    off, rows = step.rows()
    off = np.ascontiguousarray(off, np.int64); rows = np.ascontiguousarray(rows)
    idx = np.ascontiguousarray(idx, np.int32)
    bx, by = np.ascontiguousarray(c2_boundary[0], np.float64), np.ascontiguousarray(c2_boundary[1], np.float64)
    cap = int(floes.vx.shape[0]) * 2 + 8
    da_off, da = np.zeros(idx.shape[0] + 1, np.int64), np.full(cap, 7, np.uint8)
    view = floes.struct()
    p = abi._ptr
    r = l.szport_corner_mask(C.byref(view), p(off, abi.c_lp), p(rows, abi.c_dp), idx.shape[0], p(idx, abi.c_ip), int(Nb), float(Lx), float(Ly),
                             p(bx, abi.c_dp), p(by, abi.c_dp), bx.shape[0], p(da_off, abi.c_lp), p(da, abi.c_bp), cap)
    assert r >= 0, r
    return [da[da_off[k]:da_off[k + 1]].copy() for k in range(idx.shape[0])]


def _same_masks(got, want):
    assert len(got) == len(want)
    for k, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape and np.array_equal(g, w), (k, g, w)
    return sum(int(w.sum()) for w in want)


def _wall_case():
    L = 5000.0
    c2, fb = scenarios.domain(L, L)
    sq = np.array([[-1000.0, -1000.0], [-1000.0, 1000.0], [1000.0, 1000.0], [1000.0, -1000.0]])
    soa = sz.floes_to_soa([scenarios.floe_from_polygon(sq + [L - 200.0, 0.0]), scenarios.floe_from_polygon(sq + [L - 2100.0, 100.0])])
    prm = sz.default_params(Lx=L, Ly=L, modulus=1e7, dt=10.0, periodic=0, collision=1)
    bnd = sz.Boundary(fb["c"][0], fb["c"][1], c2[0], c2[1], fb["area"], fb["h"])
    return prm, soa, bnd, c2


def test_corner_mask_core_on_host_matches_oracle():
    """the code the device runs per floe (sz_corners.cuh) against the oracle's restatement: hand-derived squares, the Nb
    rule, a wall contact, every floe of a periodic Voronoi field (partners that are periodic images), floes moved after the
    contact step (rows of the old positions, list of the new ones), real concave shapes, repeated selections"""
    c2, _ = scenarios.domain(1e5, 1e5)
    prm, soa = two_squares(1900.0)
    step = oracle.OracleStep(prm, soa)
    got = _port_corner_mask(step, soa, [1, 2], prm.Lx, prm.Ly, c2)
    assert list(got[0]) == [0, 0, 1, 1] and list(got[1]) == [1, 1, 0, 0]
    for Nb in (0, 1, 2):
        _same_masks(_port_corner_mask(step, soa, [2, 1, 2], prm.Lx, prm.Ly, c2, Nb), oracle.corner_eligibility(step, soa, [2, 1, 2], prm.Lx, prm.Ly, c2, Nb))
    prm, soa, bnd, c2w = _wall_case()
    step = oracle.OracleStep(prm, soa, bnd)
    want = oracle.corner_eligibility(step, soa, [1, 2], prm.Lx, prm.Ly, c2w)
    # floe 1 = [3800, 5800] x [-1000, 1000], floe 2 = [1900, 3900] x [-900, 1100]; contact point (3850, 50).  Floe 1: (3800, 1000) is
    # inside floe 2 and nearest to the contact point, the two vertices at x = 5800 are beyond the wall; floe 2: (3900, -900) only
    assert list(want[0]) == [0, 1, 1, 1] and list(want[1]) == [0, 0, 0, 1]
    _same_masks(_port_corner_mask(step, soa, [1, 2], prm.Lx, prm.Ly, c2w), want)
    flagged = 0
    for inflate, seed in ((0.02, 71), (0.1, 72)):
        prm, soa = sz.voronoi_field(1500, seed=seed, inflate=inflate)
        c2v, _ = scenarios.domain(prm.Lx, prm.Ly)
        step = oracle.OracleStep(prm, soa, broad_mode=1)
        idx = np.arange(1, soa.n + 1)
        flagged += _same_masks(_port_corner_mask(step, soa, idx, prm.Lx, prm.Ly, c2v, Nb=3), oracle.corner_eligibility(step, soa, idx, prm.Lx, prm.Ly, c2v, Nb=3))
        # the floes moved since the contact step (corners runs after calc_trajectory)
        rng = np.random.default_rng(seed)
        moved = sz.FloesSoA(soa.x + rng.uniform(-30, 30, soa.n), soa.y + rng.uniform(-30, 30, soa.n), soa.rmax, soa.h, soa.area, soa.u, soa.v, soa.ksi, soa.alive,
                            soa.voff, soa.vx, soa.vy)
        flagged += _same_masks(_port_corner_mask(step, moved, idx[::3], prm.Lx, prm.Ly, c2v), oracle.corner_eligibility(step, moved, idx[::3], prm.Lx, prm.Ly, c2v))
    assert flagged > 5000
    prm_r, Floe = scenarios.real_shape_field(5, seed=4)
    soa, _ = scenarios.soa_and_boundary(Floe, prm_r, periodic=True)
    c2r, _ = scenarios.domain(prm_r.Lx, prm_r.Ly)
    step = oracle.OracleStep(prm_r, soa, broad_mode=1)
    idx = np.arange(1, soa.n + 1)
    assert _same_masks(_port_corner_mask(step, soa, idx, prm_r.Lx, prm_r.Ly, c2r), oracle.corner_eligibility(step, soa, idx, prm_r.Lx, prm_r.Ly, c2r)) > 20


@pytest.mark.gpu
def test_device_corner_mask_matches_oracle():
    """sz_corner_mask against the oracle: the same cases as the host test, through the C ABI (8 lanes per floe on Voronoi
    fields, a warp per floe on real shapes)"""
    with sz.ContactContext(0) as ctx:
        c2, _ = scenarios.domain(1e5, 1e5)
        prm, soa = two_squares(1900.0)
        ctx.step(prm, soa)
        got = ctx.corner_mask([1, 2])
        assert list(got[0]) == [0, 0, 1, 1] and list(got[1]) == [1, 1, 0, 0]
        step = oracle.OracleStep(prm, soa)
        for Nb in (0, 1, 2):
            _same_masks(ctx.corner_mask([2, 1, 2], Nb), oracle.corner_eligibility(step, soa, [2, 1, 2], prm.Lx, prm.Ly, c2, Nb))
        assert ctx.corner_mask([]) == []
        prm, soa, bnd, c2w = _wall_case()
        ctx.step(prm, soa, bnd)
        _same_masks(ctx.corner_mask([1, 2]), oracle.corner_eligibility(oracle.OracleStep(prm, soa, bnd), soa, [1, 2], prm.Lx, prm.Ly, c2w))
        flagged = 0
        for inflate, seed in ((0.02, 71), (0.1, 72)):
            prm, soa = sz.voronoi_field(3000, seed=seed, inflate=inflate)
            c2v, _ = scenarios.domain(prm.Lx, prm.Ly)
            ctx.step(prm, soa, allow_pair_errors=True)
            step = oracle.OracleStep(prm, soa, broad_mode=1)
            idx = np.arange(1, soa.n + 1)
            flagged += _same_masks(ctx.corner_mask(idx, Nb=3), oracle.corner_eligibility(step, soa, idx, prm.Lx, prm.Ly, c2v, Nb=3))
            flagged += _same_masks(ctx.corner_mask(idx[::3]), oracle.corner_eligibility(step, soa, idx[::3], prm.Lx, prm.Ly, c2v))
        assert flagged > 10000
        prm_r, Floe = scenarios.real_shape_field(5, seed=4)
        soa, _ = scenarios.soa_and_boundary(Floe, prm_r, periodic=True)
        c2r, _ = scenarios.domain(prm_r.Lx, prm_r.Ly)
        ctx.step(prm_r, soa, allow_pair_errors=True)
        step = oracle.OracleStep(prm_r, soa, broad_mode=1)
        idx = np.arange(1, soa.n + 1)
        assert _same_masks(ctx.corner_mask(idx), oracle.corner_eligibility(step, soa, idx, prm_r.Lx, prm_r.Ly, c2r)) > 20
