"""calc_eulerian_data.m: the oracle's restatement against hand-derived answers.  CPU only: the device
path of this row is not built yet; these known answers are what it will be checked against."""
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
