"""Hand-derived known answers read straight off the device's outputs through the C ABI (the derivations are in
tests/test_host_port.py and tests/test_oracle_golden.py).  Sorted after the long-standing parity tests and before the device
paths that have not run on a GPU yet (pytest -x)."""
import numpy as np
import pytest

import scenarios
import subzero_b200 as sz

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    with sz.ContactContext(0) as c:
        yield c


def test_hand_derived_known_answers_on_the_device(ctx):
    """the closed-form answers of tests/test_host_port.py (test_force_law_hand_derived_known_answer,
    test_wall_force_hand_derived_known_answer) read straight off the device's outputs: two equal squares (general branch,
    dl = 2200/3), a taller partner turned by 30 degrees (two-point branch, through class C), a square through the east wall"""
    sq = np.array([[-1000.0, -1000.0], [-1000.0, 1000.0], [1000.0, 1000.0], [1000.0, -1000.0]])
    M, dt = 1e7, 10.0
    G = M / 2.6
    # equal squares, floe 2 sliding in +y at 0.01 m/s
    soa = sz.floes_to_soa([scenarios.floe_from_polygon(sq), scenarios.floe_from_polygon(sq + [1900.0, 0.0], v=0.01)])
    prm = sz.default_params(Lx=1e5, Ly=1e5, modulus=M, dt=dt, periodic=1, collision=1)
    ctx.step(prm, soa)
    off, rows = ctx.rows()
    Fn, Ft = 625.0 * 2e5, 1e-4 * (2200.0 / 3.0) * G * dt
    assert off.tolist()[:3] == [0, 1, 2] and rows[0][0] == 2 and rows[1][0] == 1
    assert rows[0][1] == pytest.approx(-Fn, rel=1e-12) and rows[0][2] == pytest.approx(Ft, rel=1e-12)
    assert rows[0][3] == pytest.approx(950.0, abs=1e-6) and rows[0][5] == pytest.approx(950.0 * Ft, rel=1e-12) and rows[0][6] == pytest.approx(2e5, rel=1e-12)
    assert np.array_equal(rows[0][1:3], -rows[1][1:3])
    # taller partner, everything turned by 30 degrees: strictly convex without horizontal edges -> class C
    th = np.pi / 6
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    tall = sq * [1.0, 1.2] + [1900.0, 0.0]
    soa = sz.floes_to_soa([scenarios.floe_from_polygon(sq @ R.T), scenarios.floe_from_polygon(tall @ R.T, u=-0.01 * np.sin(th), v=0.01 * np.cos(th))])
    s = ctx.step(prm, soa)
    off, rows = ctx.rows()
    Fn2 = M * 0.25 * 0.25 / (0.25 * np.sqrt(2000.0 * 2400.0) + 0.25 * 2000.0) * 2e5
    want = R @ np.array([-Fn2, 1e-4 * 2000.0 * G * dt])
    assert s.n_pairs == 1 and rows[0][1:3] == pytest.approx(want, rel=1e-9) and rows[0][6] == pytest.approx(2e5, rel=1e-9)
    # a square 800 m through the east wall, sliding in +y at 0.02 m/s
    L = 5000.0
    c2, fb = scenarios.domain(L, L)
    soa = sz.floes_to_soa([scenarios.floe_from_polygon(sq + [L - 200.0, 300.0], v=0.02)])
    prm = sz.default_params(Lx=L, Ly=L, modulus=M, dt=dt, periodic=0, collision=1)
    bnd = sz.Boundary(fb["c"][0], fb["c"][1], c2[0], c2[1], fb["area"], fb["h"])
    s = ctx.step(prm, soa, bnd)
    off, rows = ctx.rows()
    out = ctx.floe_outputs()
    Fw, Fy = 1250.0 * 1.6e6, -4e-4 * 2000.0 * G * dt
    assert np.isinf(rows[0][0]) and rows[0][1] == pytest.approx(-Fw, rel=1e-12) and rows[0][2] == pytest.approx(Fy, rel=1e-12)
    assert rows[0][3] == pytest.approx(5400.0, abs=1e-6) and rows[0][4] == pytest.approx(300.0, abs=1e-6) and rows[0][6] == pytest.approx(1.6e6, rel=1e-12)
    assert out["fx"][0] == pytest.approx(-Fw, rel=1e-12) and out["torque"][0] == pytest.approx(600.0 * Fy, rel=1e-12)
    assert out["stress"][0][0][0] == pytest.approx(2 * 600.0 * -Fw / (2 * 4e6 * 0.25), rel=1e-12) and s.collision_count == 1.0
    # a contact across the periodic boundary: image, pair list, mirrored row, fold into the parent (see the derivation there)
    from test_oracle_golden import periodic_image_case, check_periodic_image_answers
    prm, soa = periodic_image_case()
    s = ctx.step(prm, soa)
    off, rows = ctx.rows()
    check_periodic_image_answers(s, ctx.ghosts(), ctx.pairs(), off, rows, ctx.floe_outputs())
    # the thresholds of the loop: Amin, the strict 55 % merge rule, +Inf / -Inf with kill and transfer
    from test_oracle_golden import run_threshold_checks

    def step(prm, soa):
        ctx.step(prm, soa)
        off, rows = ctx.rows()
        return off, rows, ctx.pairs(), ctx.floe_outputs()
    run_threshold_checks(step)
