"""The oracle against every known answer the reference holds for this path, plus hand-derived
known answers for the MATLAB helpers.  CPU only."""
import ctypes as C

import numpy as np
import pytest

import oracle
import scenarios
from subzero_b200 import abi
import subzero_b200 as sz

S = 2.0 ** 32
P = abi._ptr


def polyclip(p1, p2, method):
    p1, p2 = np.asarray(p1, float), np.asarray(p2, float)
    x1, y1, x2, y2 = (np.ascontiguousarray(a) for a in (p1[:, 0], p1[:, 1], p2[:, 0], p2[:, 1]))
    ox, oy, off = np.empty(4096), np.empty(4096), np.empty(256, np.int32)
    n = oracle.lib().szo_polyclip(P(x1, abi.c_dp), P(y1, abi.c_dp), len(x1), P(x2, abi.c_dp), P(y2, abi.c_dp), len(x2), method,
                                  P(ox, abi.c_dp), P(oy, abi.c_dp), 4096, P(off, abi.c_ip), 256)
    assert n >= 0
    return [np.stack([ox[off[k]:off[k + 1]], oy[off[k]:off[k + 1]]], 1) for k in range(n)]


def test_clipper_version_is_the_reference():
    oracle.clipper().szref_version.restype = C.c_char_p
    assert oracle.clipper().szref_version() == b"6.4.2"          # private/clipper.hpp:37


def test_clipper_test_m_square_case():
    """private/clipper_test.m:2-11: unit square ∩ unit square shifted by 0.5 -> [0.5,1]^2"""
    p1 = [[0, 0], [1, 0], [1, 1], [0, 1]]
    p2 = [[0.5, 0.5], [1.5, 0.5], [1.5, 1.5], [0.5, 1.5]]
    r = polyclip(p1, p2, 1)
    assert len(r) == 1
    assert r[0].tolist() == [[1, 1], [0.5, 1], [0.5, 0.5], [1, 0.5]]   # SURVEY.md E.1 (the reference run here)
    x, y = np.ascontiguousarray(r[0][:, 0]), np.ascontiguousarray(r[0][:, 1])
    assert oracle.lib().szo_polyarea(P(x, abi.c_dp), P(y, abi.c_dp), 4) == 0.25


def test_clipper_kats():
    """behaviour known answers of the reference Clipper, scale 2^32, even-odd"""
    sq = lambda x0, y0, x1, y1: [[x0, y0], [x1, y0], [x1, y1], [x0, y1]]
    assert polyclip(sq(0, 0, 1, 1), sq(1, 0, 2, 1), 1) == []                       # shared edge
    assert polyclip(sq(0, 0, 1, 1), sq(1, 1, 2, 2), 1) == []                       # vertex touch
    r = polyclip(sq(0, 0, 5, 5), sq(2, 2, 3, 3), 1)
    assert [p.tolist() for p in r] == [[[3, 3], [2, 3], [2, 2], [3, 2]]]           # contained square
    r_cw = polyclip(sq(0, 0, 5, 5)[::-1], sq(2, 2, 3, 3)[::-1], 1)
    assert [p.tolist() for p in r_cw] == [p.tolist() for p in r]                   # orientation-independent
    u = [[0, 0], [3, 0], [3, 3], [2, 3], [2, 1], [1, 1], [1, 3], [0, 3]]
    r = polyclip(u, sq(-1, 2, 4, 2.5), 1)
    assert [p.tolist() for p in r] == [[[1, 2.5], [0, 2.5], [0, 2], [1, 2]], [[3, 2.5], [2, 2.5], [2, 2], [3, 2]]]
    r = polyclip(sq(0, 0, 3, 3), sq(1, 1, 2, 2), 0)                                # difference -> outer + hole
    assert len(r) == 2 and r[1].tolist() == [[1, 1], [1, 2], [2, 2], [2, 1]]
    r = polyclip(sq(90000, 0, 100500, 1000), sq(-1e5, -1e5, 1e5, 1e5), 0)          # floe poking through the wall
    assert [p.tolist() for p in r] == [[[100500, 1000], [100000, 1000], [100000, 0], [100500, 0]]]
    r = polyclip([[0, 0], [1, 0], [2, 0], [2, 2], [0, 2]], sq(-1, -1, 3, 3), 1)    # collinear vertex dropped
    assert len(r) == 1 and len(r[0]) == 4


def test_polyshape_area_centroid_pinned_by_floeshapes_mat():
    """462 known answers cached by MATLAB inside test/test_conservation/FloeShapes.mat (BoundaryInfo)"""
    polys, binfo, modulus = scenarios.floe_shapes()
    assert len(polys) == 462 and modulus == 90000000.0
    out = np.empty(3)
    worst_a = worst_c = 0.0
    for v, b in zip(polys, binfo):
        x, y = np.ascontiguousarray(v[:, 0]), np.ascontiguousarray(v[:, 1])
        oracle.lib().szo_polyshape_area_centroid(P(x, abi.c_dp), P(y, abi.c_dp), len(x), P(out, abi.c_dp))
        area, cx, cy = b[3], b[5], b[6]
        worst_a = max(worst_a, abs(out[0] - area) / area)
        worst_c = max(worst_c, np.hypot(out[1] - cx, out[2] - cy) / np.sqrt(area))
    assert worst_a < 5e-15 and worst_c < 1e-14, (worst_a, worst_c)


def test_matlab_int64_rounding():
    f = oracle.lib().szo_matlab_int64
    assert [f(v) for v in (0.5, 1.5, 2.5, -0.5, -1.5, 0.49999999999999994, -2.4, 2.6)] == [1, 2, 3, -1, -2, 0, -2, 3]
    assert f(float("nan")) == 0 and f(1e30) == 2 ** 63 - 1 and f(-1e30) == -2 ** 63


def interx(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    out = np.empty(512)
    x1, y1, x2, y2 = (np.ascontiguousarray(v) for v in (a[:, 0], a[:, 1], b[:, 0], b[:, 1]))
    n = oracle.lib().szo_interx(P(x1, abi.c_dp), P(y1, abi.c_dp), len(x1), P(x2, abi.c_dp), P(y2, abi.c_dp), len(x2), P(out, abi.c_dp), 256)
    return out[:2 * n].reshape(n, 2)


def test_interx_known_answers():
    sq = np.array([[0, 0], [2, 0], [2, 2], [0, 2], [0, 0]], float)
    p = interx(sq, sq + 1)                       # two crossings, sorted by x then y (unique(...,'rows'))
    assert p.tolist() == [[1, 2], [2, 1]]
    assert len(interx(sq, sq + 5)) == 0
    p = interx(sq, sq + [2, 0])                  # shared (parallel) edge: L == 0 pairs dropped; endpoint touches kept once each
    assert p.tolist() == [[2, 0], [2, 2]]
    diag = np.array([[-1, -1], [3, 3]], float)
    assert interx(sq, diag).tolist() == [[0, 0], [2, 2]]


def test_inpolygon_known_answers():
    xv, yv = np.array([0, 4, 4, 0.0]), np.array([0, 0, 4, 4.0])
    px, py = np.array([2, 0, 4, 5, 2, -1e-13, 4.0]), np.array([2, 2, 4, 2, 0, 2, 4.0000001])
    r = np.empty(len(px), np.uint8)
    oracle.lib().szo_inpolygon(P(px, abi.c_dp), P(py, abi.c_dp), len(px), P(xv, abi.c_dp), P(yv, abi.c_dp), 4, P(r, abi.c_bp))
    assert r.tolist() == [1, 1, 1, 0, 1, 0, 0]   # interior, edge, vertex, outside, edge, outside the bbox mask, outside


def test_p_poly_dist_known_answers():
    xv, yv = np.array([0, 4, 4, 0, 0.0]), np.array([0, 0, 4, 4, 0.0])
    px, py = np.array([2, 2, 6, 5, 2.0]), np.array([2, -1, 2, 5, 0.0])
    d = np.empty(5)
    assert oracle.lib().szo_p_poly_dist(P(px, abi.c_dp), P(py, abi.c_dp), 5, P(xv, abi.c_dp), P(yv, abi.c_dp), 5, P(d, abi.c_dp)) == 0
    np.testing.assert_allclose(d, [-2, 1, 2, np.sqrt(2), 0], atol=1e-15)          # negative inside; on the edge: -0 or 0
    bad = np.array([0, 0, 4, 4, 0.0])
    assert oracle.lib().szo_p_poly_dist(P(px, abi.c_dp), P(py, abi.c_dp), 5, P(bad, abi.c_dp), P(yv, abi.c_dp), 5, P(d, abi.c_dp)) == -1   # repeated vertex raises


def test_broad_phase_grid_equals_literal_double_loop():
    """the oracle's cell-grid broad phase (used beyond ~3e4 floes) is the literal O(N^2) loop, output for output"""
    import subzero_b200 as sz
    prm, f = sz.voronoi_field(3000, seed=3)
    prm.want_clip_polys = 1
    a = oracle.OracleStep(prm, f, broad_mode=0)
    b = oracle.OracleStep(prm, f, broad_mode=1)
    rep = oracle.compare_steps(b, a, rtol=0.0)
    assert rep["rows_bit_exact"] and a.summary.n > a.summary.n0 and a.summary.n_pairs > 12000


def test_two_block_head_on_force_is_equal_and_opposite():
    """conservation_test.m:22-26 set-up, advanced until the blocks overlap by 500 m: one rectangular overlap region,
    normal along x, rows mirrored (floe_interactions_all.m:196)"""
    import subzero_b200 as sz
    cases, modulus = scenarios.conservation_cases()
    Floe = scenarios.advance(cases["head_on"], 42000.0)          # closing speed 0.25 m/s -> gap 1e4 m closes by 10500 m
    prm = sz.default_params(Lx=1e5, Ly=1e5, modulus=modulus, dt=10.0, periodic=0, collision=1)
    soa, bnd = scenarios.soa_and_boundary(Floe, prm, periodic=False)
    r = oracle.OracleStep(prm, soa, bnd)
    off, rows = r.rows()
    assert r.summary.n_pairs == 1 and off.tolist() == [0, 1, 2]
    a, b = rows
    assert a[0] == 2 and b[0] == 1
    assert a[6] == b[6] == pytest.approx(500.0 * 3e4, rel=1e-12)  # overlap area
    np.testing.assert_array_equal(a[1:3], -b[1:3])
    assert a[1] < 0 and abs(a[2]) < abs(a[1])                    # floe 1 (left) is pushed to -x
    ff = modulus * 0.25 * 0.25 / (0.25 * 3e4 + 0.25 * 3e4)      # floe_interactions.m:12, sqrt(area) = 3e4
    assert abs(a[1]) <= ff * a[6] * (1 + 0.2) * (1 + 1e-12)     # normal + Coulomb-capped tangential
    assert r.summary.collision_count == 1.0


def periodic_image_case():
    """square A pokes 50 m through +Lx (Lx = 10 km); square B sits just inside -Lx (its edge ON the boundary: not beyond it)"""
    import subzero_b200 as sz
    L = 1e4
    sq = np.array([[-1000.0, -1000.0], [-1000.0, 1000.0], [1000.0, 1000.0], [1000.0, -1000.0]])
    soa = sz.floes_to_soa([scenarios.floe_from_polygon(sq + [L - 950.0, 0.0]), scenarios.floe_from_polygon(sq + [-L + 1000.0, 0.0])])
    return sz.default_params(Lx=L, Ly=L, modulus=1e7, dt=10.0, periodic=1, collision=1), soa


def check_periodic_image_answers(summary, ghosts, pairs, off, rows, out):
    """floe_interactions_all.m worked by hand for periodic_image_case():
      * :28-39   A has a vertex beyond +Lx (strict >): one image at Xi - 2 Lx = -10950, FloeNum -1, parent 1; B's edge lies ON -Lx: none
      * :76-120  the only candidate pair is (2, 3): |x_B - x_A'| = 1950 < rmax_B + rmax_A' = 2 sqrt(2) 1000
      * the image overlaps B in [-10000, -9950] x [-1000, 1000]: A = 1e5, Force_factor 625 -> 6.25e7 N, +x on B
      * :196     mirrored row on the image; :242-245 its force folded into floe 1 (-6.25e7) -- but not its OverlapArea
      * calc_trajectory.m:9-13 stress_xx of B = 2 (P_x - X_B) F_x / (2 area h) with P_x = -9975; calc_collisionNum.m: 1/2"""
    assert summary.n0 == 2 and summary.n == 3 and summary.n_pairs == 1 and summary.collision_count == 0.5
    assert ghosts["parent"].tolist() == [1] and ghosts["floe_num"].tolist() == [-1] and ghosts["x"].tolist() == [-10950.0] and ghosts["y"].tolist() == [0.0]
    assert pairs["i"].tolist() == [2] and pairs["j"].tolist() == [3]
    assert off.tolist() == [0, 0, 1, 2]
    F = 625.0 * 1e5
    assert rows[0][0] == 3 and rows[1][0] == 2
    assert rows[0][1] == pytest.approx(F, rel=1e-12) and rows[0][2] == 0 and rows[0][3] == pytest.approx(-9975.0, abs=1e-6) and rows[0][6] == pytest.approx(1e5, rel=1e-12)
    np.testing.assert_array_equal(rows[1][1:3], -rows[0][1:3])
    assert out["fx"].tolist() == pytest.approx([-F, F], rel=1e-12) and out["fy"].tolist() == [0.0, 0.0] and out["torque"].tolist() == [0.0, 0.0]
    assert out["overlap_area"].tolist() == pytest.approx([0.0, 1e5], rel=1e-12)
    assert out["stress"][1][0][0] == pytest.approx(2 * (-9975.0 + 9000.0) * F / (2 * 4e6 * 0.25), rel=1e-12) and np.all(out["stress"][0] == 0)


def test_periodic_image_contact_hand_derived():
    prm, soa = periodic_image_case()
    st = oracle.OracleStep(prm, soa)
    off, rows = st.rows()
    check_periodic_image_answers(st.summary, st.ghosts(), st.pairs(), off, rows, st.floe_outputs())


def run_threshold_checks(step):
    """thresholds of the contact loop worked by hand on pairs of squares (2 km, h = 0.25 m, Modulus 1e7, Force_factor 625);
    `step(prm, soa)` returns (row_off, rows, pairs, floe_outputs) of one contact step:
      * floe_interactions.m:79     regions smaller than Amin = min(N1, N2) * 100 / 1.75 = 228.57 m^2 are dropped: a 0.11 m strip
                                   (220 m^2) gives no row, a 0.12 m strip (240 m^2) gives 625 * 240 N
      * :54-60                     overlap / floe area > 0.55 (strict): exactly 55 % is an ordinary contact (625 * 2.2e6 N), 60 % is
                                   overlap = +Inf, no force; floe_interactions_all.m:138-141: kill(1) = 1, transfer(1) = 2
      * :57-58 and :142-143,175-179  a small floe 75 % inside a big one (big floe first): overlap = -Inf, kill(1) = 2, and the
                                   fix-up loop sets transfer(2) = 1"""
    import subzero_b200 as sz
    sq = np.array([[-1000.0, -1000.0], [-1000.0, 1000.0], [1000.0, 1000.0], [1000.0, -1000.0]])
    prm = sz.default_params(Lx=1e5, Ly=1e5, modulus=1e7, dt=10.0, periodic=1, collision=1)
    pair = lambda second: sz.floes_to_soa([scenarios.floe_from_polygon(sq), scenarios.floe_from_polygon(second)])
    off, rows, pairs, out = step(prm, pair(sq + [2000.0 - 0.11, 0.0]))
    assert off.tolist() == [0, 0, 0] and pairs["i"].tolist() == [1] and out["fx"].tolist() == [0.0, 0.0]
    off, rows, pairs, out = step(prm, pair(sq + [2000.0 - 0.12, 0.0]))
    assert off.tolist() == [0, 1, 2] and rows[0][6] == pytest.approx(240.0, rel=1e-8) and rows[0][1] == pytest.approx(-625.0 * 240.0, rel=1e-8)
    off, rows, pairs, out = step(prm, pair(sq + [900.0, 0.0]))
    assert pairs["overlap_state"].tolist() == [0.0] and out["fx"].tolist() == pytest.approx([-625.0 * 2.2e6, 625.0 * 2.2e6], rel=1e-12)
    assert out["kill"].tolist() == [0, 0] and out["transfer"].tolist() == [0, 0]
    off, rows, pairs, out = step(prm, pair(sq + [800.0, 0.0]))
    assert off.tolist() == [0, 0, 0] and pairs["overlap_state"].tolist() == [np.inf] and out["fx"].tolist() == [0.0, 0.0]
    assert out["kill"].tolist() == [1, 0] and out["transfer"].tolist() == [2, 0]
    soa = sz.floes_to_soa([scenarios.floe_from_polygon(sq * 3), scenarios.floe_from_polygon(sq + [2500.0, 0.0])])
    off, rows, pairs, out = step(prm, soa)
    assert pairs["overlap_state"].tolist() == [-np.inf] and out["kill"].tolist() == [2, 0] and out["transfer"].tolist() == [0, 1]


def test_contact_thresholds_hand_derived():
    def step(prm, soa):
        st = oracle.OracleStep(prm, soa)
        off, rows = st.rows()
        return off, rows, st.pairs(), st.floe_outputs()
    run_threshold_checks(step)


def test_topography_floes_pair_only_when_opted_in():
    """as written, floes 1..Nb can never be anybody's partner (floe_interactions_all.m:76,102-103: i >= Nb+1 and
    j > i) -- the reference behaviour and the default.  SzParams.pair_with_boundary_floes = 1 (opt-in, NOT the reference) lets
    a floe i > Nb record the topography floes below it; the force acts on i only.  Worked by hand: floe 1 (topography, Nb = 1)
    = [0,1000]^2, floe 2 = [900,1900] x [0,1000] overlapping it in a 100 m strip, floe 3 far away.  Default: no pairs at all
    between 1 and 2.  Opted in: pair (2,1); floe 2 gets one row [1 Fx 0 950 500 . 1e5] with Fx > 0 (pushed away from floe 1,
    the same magnitude floe 1 would get from the mirrored pair (1,2) if it were an ordinary floe); floe 1 gets nothing."""
    import scenarios
    sq = lambda x0: np.array([[x0, 0.0], [x0, 1000.0], [x0 + 1000.0, 1000.0], [x0 + 1000.0, 0.0]])
    Floe = [scenarios.floe_from_polygon(sq(0.0)), scenarios.floe_from_polygon(sq(900.0)), scenarios.floe_from_polygon(sq(5000.0))]
    soa = sz.floes_to_soa(Floe)
    prm = sz.default_params()
    prm.Lx = prm.Ly = 1e4
    prm.modulus, prm.dt, prm.periodic, prm.collision, prm.Nb = 9e7, 10.0, 1, 1, 1
    ref = oracle.OracleStep(prm, soa, broad_mode=0)
    assert ref.summary.n_pairs == 0 and ref.summary.n_rows == 0                       # the reference: topography never pairs
    prm.pair_with_boundary_floes = 1
    for mode in (0, 1):
        ref = oracle.OracleStep(prm, soa, broad_mode=mode)
        p = ref.pairs()
        assert p["i"].tolist() == [2] and p["j"].tolist() == [1]
        off, rows = ref.rows()
        assert (off[1] - off[0], off[2] - off[1], off[3] - off[2]) == (0, 1, 0)       # only floe 2 carries a row
        r = rows[0]
        assert r[0] == 1 and r[1] > 0 and r[2] == 0 and r[3] == pytest.approx(950.0) and r[4] == pytest.approx(500.0) and r[6] == pytest.approx(1e5)
        o = ref.floe_outputs()
        assert o["fx"][0] == 0 and o["fx"][1] == r[1] and o["overlap_area"][0] == 0 and o["overlap_area"][1] == pytest.approx(1e5)
    # the same contact between two ordinary floes: pair (1,2), floe 1 carries the row, floe 2 the mirrored one of equal magnitude
    prm.Nb, prm.pair_with_boundary_floes = 0, 0
    ref0 = oracle.OracleStep(prm, soa, broad_mode=0)
    off0, rows0 = ref0.rows()
    assert rows0[off0[1]][0] == 1 and rows0[off0[1]][1] == pytest.approx(r[1], rel=1e-12)
