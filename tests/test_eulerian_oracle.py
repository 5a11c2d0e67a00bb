"""calc_eulerian_data.m: the oracle's restatement against hand-derived answers, and the product's item /
reduction code (sz_euler.cuh) compiled for the host against the oracle.  CPU only; the device path (sz_eulerian_data) is
compared with the oracle in tests/test_zzz_eulerian_device.py (-m gpu)."""
import numpy as np
import pytest

import oracle
import scenarios
import subzero_b200 as sz


def square(cx, cy, half=1000.0, **kw):
    sq = np.array([[-half, -half], [-half, half], [half, half], [half, -half]]) + [cx, cy]
    return scenarios.floe_from_polygon(sq, **kw)


def test_one_floe_on_a_cell_corner():
    """a 2 km square centred on a grid corner of a 4 x 4 grid over [-4, 4] km: a quarter of it (1 km^2) in each of the four
    central cells, concentration 1/4 there, the cell averages equal to the floe's own values"""
    soa = sz.floes_to_soa([square(0.0, 0.0, u=0.3, v=-0.2)])
    mass = soa.area * soa.h * 920.0
    d = oracle.calc_eulerian_data(soa, mass, 4, 4, (-4000.0, 4000.0, -4000.0, 4000.0), periodic=False, overlap_area=[5.0], dUi_p=[1e-3], dVi_p=[2e-3],
                                  stress=[[3.0, 1.0, 1.0, -2.0]], strain=[[1e-6, 2e-6, 3e-6, 4e-6]])
    centre = np.zeros((4, 4), bool); centre[1:3, 1:3] = True
    assert np.allclose(d["c"][centre], 0.25, rtol=1e-12) and np.all(d["c"][~centre] == 0)
    assert np.allclose(d["area"][centre], 1e6, rtol=1e-12) and np.allclose(d["Mtot"][centre], mass[0] / 4, rtol=1e-12)
    for k, want in (("u", 0.3), ("v", -0.2), ("du", 1e-3), ("dv", 2e-3), ("h", soa.h[0]), ("Over", 5.0), ("stressxx", 3.0), ("stressyx", 1.0), ("stressxy", 1.0), ("stressyy", -2.0),
                    ("strainux", 1e-6), ("strainvx", 2e-6), ("strainuy", 3e-6), ("strainvy", 4e-6)):
        assert np.allclose(d[k][centre], want, rtol=1e-12), k
        assert np.all(d[k][~centre] == 0), k
    lam = 0.5 + np.sqrt(2.5 ** 2 + 1.0)                                # max eig of [3 1; 1 -2]
    assert np.allclose(d["stress"][centre], lam, rtol=1e-12)


def test_rows_run_from_the_top_and_mass_weighting():
    """two floes of different thickness sharing one cell: mass-weighted velocity; a floe in the upper half of the domain
    lands in the upper rows (the reference flips y); a dead floe does not count"""
    a, b = square(-2000.0, 2000.0, u=1.0, h=0.5), square(-2000.0, 2000.0, half=500.0, u=-1.0, h=2.0)
    dead = square(2000.0, -2000.0, u=9.0)
    dead["alive"] = 0
    soa = sz.floes_to_soa([a, b, dead])
    mass = soa.area * soa.h * 920.0
    d = oracle.calc_eulerian_data(soa, mass, 2, 2, (-4000.0, 4000.0, -4000.0, 4000.0), periodic=False)
    assert d["c"][0, 0] == pytest.approx((4e6 + 1e6) / 16e6, rel=1e-12) and np.count_nonzero(d["c"]) == 1        # top-left cell only
    assert d["u"][0, 0] == pytest.approx((mass[0] * 1.0 + mass[1] * -1.0) / (mass[0] + mass[1]), rel=1e-12)
    assert d["h"][0, 0] == pytest.approx((mass[0] * 0.5 + mass[1] * 2.0) / (mass[0] + mass[1]), rel=1e-12)


def test_periodic_ghosts_and_the_stale_polygon_quirk():
    """a floe poking through +Lx contributes its outside part to the opposite column through its x-ghost (:39-48); the y pass
    tests the polygon of the LAST floe for everybody (:56-65): y-ghosts appear for all floes or for none"""
    box = (-4000.0, 4000.0, -4000.0, 4000.0)
    east = square(3500.0, 1000.0)                                       # pokes 500 m through +Lx
    soa = sz.floes_to_soa([east])
    mass = soa.area * soa.h * 920.0
    d = oracle.calc_eulerian_data(soa, mass, 2, 2, box, periodic=True)
    assert d["area"][0, 1] == pytest.approx(1500.0 * 2000.0, rel=1e-12) and d["area"][0, 0] == pytest.approx(500.0 * 2000.0, rel=1e-12)
    assert d["area"][1].sum() == 0
    d0 = oracle.calc_eulerian_data(soa, mass, 2, 2, box, periodic=False)
    assert d0["area"][0, 0] == 0                                        # no ghost without PERIODIC
    # last floe inside in y: nobody gets a y-ghost, even the floe poking through +Ly
    north, inside = square(0.0, 3500.0), square(-2000.0, -2000.0)
    soa = sz.floes_to_soa([north, inside])
    mass = soa.area * soa.h * 920.0
    d = oracle.calc_eulerian_data(soa, mass, 2, 2, box, periodic=True)
    assert d["area"][1].sum() == pytest.approx(4e6, rel=1e-12)          # only `inside` in the bottom row
    # last floe pokes through +Ly: EVERY floe gets a y-ghost, also the one well inside (its ghost lies outside the grid)
    soa = sz.floes_to_soa([inside, north])
    mass = soa.area * soa.h * 920.0
    d = oracle.calc_eulerian_data(soa, mass, 2, 2, box, periodic=True)
    assert d["area"][1].sum() == pytest.approx(4e6 + 500.0 * 2000.0, rel=1e-12)      # `inside` + the part of `north` that re-enters at the bottom


# ---- the product's item / reduction code (subzero_b200/csrc/sz_euler.cuh) compiled for the host, against the oracle
def port_calc_eulerian_data(floes, mass, Nx, Ny, box, periodic, overlap_area=None, dUi_p=None, dVi_p=None, stress=None, strain=None, Nb=0):
    import ctypes as C
    import os
    from subzero_b200 import abi
    l = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "host", "libpair_host.so"))
    D = abi.c_dp
    l.szport_calc_eulerian_data.restype = C.c_int
    l.szport_calc_eulerian_data.argtypes = [C.POINTER(abi.SzFloesSoA), D, D, D, D, D, D, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, D]
    n = floes.n
    z = lambda a, shape: np.zeros(shape) if a is None else np.ascontiguousarray(a, np.float64)
    ov, du, dv, st, en = z(overlap_area, n), z(dUi_p, n), z(dVi_p, n), z(stress, (n, 4)), z(strain, (n, 4))
    mass = np.ascontiguousarray(mass, np.float64)
    out = np.full((18, Ny, Nx), 7.0)
    view = floes.struct()
    p = abi._ptr
    r = l.szport_calc_eulerian_data(C.byref(view), p(mass, D), p(ov, D), p(du, D), p(dv, D), p(st, D), p(en, D), int(Nx), int(Ny), int(Nb),
                                    *(float(b) for b in box), int(bool(periodic)), p(out, D))
    assert r == 0, r
    return {k: out[i] for i, k in enumerate(oracle.EULERIAN_FIELDS)}


def _same_fields(got, want):
    for k in oracle.EULERIAN_FIELDS:
        assert np.array_equal(got[k], want[k]), (k, np.abs(got[k] - want[k]).max())


def test_product_core_on_host_matches_oracle():
    """sz_euler.cuh (items floe by floe, one Clipper-exact clip per item, stable sort by cell, sequential reduction per cell)
    gives the oracle's 18 planes bit for bit: the hand-derived cases above, periodic and walled Voronoi fields on several
    grids with random per-floe state, dead floes and NaN velocities, and real concave shapes"""
    box = (-4000.0, 4000.0, -4000.0, 4000.0)
    soa = sz.floes_to_soa([square(0.0, 0.0, u=0.3, v=-0.2)])
    kw = dict(overlap_area=[5.0], dUi_p=[1e-3], dVi_p=[2e-3], stress=[[3.0, 1.0, 1.0, -2.0]], strain=[[1e-6, 2e-6, 3e-6, 4e-6]])
    mass = soa.area * soa.h * 920.0
    got = port_calc_eulerian_data(soa, mass, 4, 4, box, False, **kw)
    _same_fields(got, oracle.calc_eulerian_data(soa, mass, 4, 4, box, False, **kw))
    assert np.allclose(got["c"][1:3, 1:3], 0.25, rtol=1e-12) and got["u"][1, 1] == pytest.approx(0.3, rel=1e-12)
    for fl, per in (([square(3500.0, 1000.0)], True), ([square(0.0, 3500.0), square(-2000.0, -2000.0)], True), ([square(-2000.0, -2000.0), square(0.0, 3500.0)], True),
                    ([square(3500.0, 1000.0)], False)):
        soa = sz.floes_to_soa(fl)
        mass = soa.area * soa.h * 920.0
        _same_fields(port_calc_eulerian_data(soa, mass, 2, 2, box, per), oracle.calc_eulerian_data(soa, mass, 2, 2, box, per))
    # floes astronomically far from the grid, or with an infinite centre, are nobody's candidates
    soa = sz.floes_to_soa([square(0.0, 0.0), square(1e15, 0.0), square(-3e14, 2e15), square(500.0, 500.0)])
    soa.x[3] = np.inf
    mass = soa.area * soa.h * 920.0
    got = port_calc_eulerian_data(soa, mass, 4, 4, box, True)
    _same_fields(got, oracle.calc_eulerian_data(soa, mass, 4, 4, box, True))
    assert got["area"].sum() == pytest.approx(4e6, rel=1e-12)
    rng = np.random.default_rng(5)
    prm, soa = sz.voronoi_field(1500, seed=81, inflate=0.05)
    n = soa.n
    soa.alive[::17] = 0
    soa.u[5] = np.nan; soa.v[11] = np.nan
    mass = soa.area * soa.h * 920.0
    mass[3] = np.nan
    kw = dict(overlap_area=rng.uniform(0, 1e5, n), dUi_p=rng.normal(0, 1e-3, n), dVi_p=rng.normal(0, 1e-3, n), stress=rng.normal(0, 1e3, (n, 4)), strain=rng.normal(0, 1e-6, (n, 4)))
    L = prm.Lx
    for (Nx, Ny), per in (((7, 5), True), ((1, 1), True), ((20, 20), True), ((10, 10), False), ((3, 40), True)):
        got = port_calc_eulerian_data(soa, mass, Nx, Ny, (-L, L, -L, L), per, **kw)
        want = oracle.calc_eulerian_data(soa, mass, Nx, Ny, (-L, L, -L, L), per, **kw)
        _same_fields(got, want)
        assert np.count_nonzero(want["Mtot"]) == Nx * Ny and 0.9 < want["c"].mean() < 1.2
    # a grid that covers only part of the field, and one larger than it (empty cells stay zero)
    for b in ((-L / 3, L / 2, -L / 4, L / 5), (-2 * L, 2 * L, -3 * L, 3 * L)):
        _same_fields(port_calc_eulerian_data(soa, mass, 6, 9, b, False, **kw), oracle.calc_eulerian_data(soa, mass, 6, 9, b, False, **kw))
    prm_r, Floe = scenarios.real_shape_field(5, seed=4)
    soa, _ = scenarios.soa_and_boundary(Floe, prm_r, periodic=True)
    mass = soa.area * soa.h * 920.0
    L = prm_r.Lx
    got = port_calc_eulerian_data(soa, mass, 8, 8, (-L, L, -L, L), True)
    _same_fields(got, oracle.calc_eulerian_data(soa, mass, 8, 8, (-L, L, -L, L), True))
    assert np.count_nonzero(got["Mtot"]) > 40
