"""The reference's only assertion on the force law, replayed: test/conservation_test.m:22-54 sets up five collisions and demands
K(end)/K(1) < 1 for the kinetic energy history of each.  The function it calls (Subzero_conservation) is not in the reference
tree, so the loop is the reference's timestep restricted to what the test needs: floe_interactions_all (contact step) +
calc_trajectory per floe with no ocean / wind forcing and no thermodynamics, in the reference's default domain
(initialize_boundaries.m:4-6, +-1e5 m, walls), dt = 10 (Subzero.m:36), Modulus from the fixture's Modulus.mat, mass and
inertia_moment as initialize_floe_values.m:16,19 / PolygonMoments.m compute them.  The floes first fly freely (kinetic energy
unchanged) up to shortly before their first contact, then the loop runs until they have separated again.

CPU: the oracle.  GPU: the same loop with the state resident on the device (sz_step_resident + sz_trajectory_step), held to
the oracle's energy history."""
import numpy as np
import pytest

import oracle
import scenarios
import subzero_b200 as sz

RHO_ICE = 920.0
# free flight before the loop starts (s) and loop length (steps): contact begins 30-80 steps in and is over well before the end
PLAN = {"head_on": (39500.0, 500), "offset": (47000.0, 520), "triangle_between": (22000.0, 900), "complex_pair": (19500.0, 480), "complex_wall": (20800.0, 520)}


def polygon_moments(ca, h):
    """polygon_operations/PolygonMoments.m:19-32 on the closed outline c0: |Ixx + Iyy| h rho_ice"""
    x, y = ca[0], ca[1]
    w = x[:-1] * y[1:] - x[1:] * y[:-1]
    ixx = (w * ((y[:-1] + y[1:]) ** 2 - y[:-1] * y[1:])).sum() / 12
    iyy = (w * ((x[:-1] + x[1:]) ** 2 - x[:-1] * x[1:])).sum() / 12
    return abs(ixx + iyy) * h * RHO_ICE


def setup(name):
    cases, modulus = scenarios.conservation_cases()
    Floe = scenarios.advance(cases[name], PLAN[name][0])
    prm = sz.default_params()
    prm.Lx = prm.Ly = 1e5
    prm.modulus, prm.dt, prm.periodic, prm.collision = modulus, 10.0, 0, 1
    soa, bnd = scenarios.soa_and_boundary(Floe, prm, False)
    n = soa.n
    st = {k: np.zeros(n) for k in ("alpha", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "FxOA", "FyOA", "torqueOA")}
    st["dXi_p"], st["dYi_p"] = soa.u.copy(), soa.v.copy()                    # steady free flight
    st["mass"] = soa.area * soa.h * RHO_ICE
    st["inertia"] = np.array([polygon_moments(f["c_alpha"], f["h"]) for f in Floe])
    return prm, soa, bnd, st


def kinetic(mass, inertia, u, v, ksi):
    return float((0.5 * mass * (u ** 2 + v ** 2) + 0.5 * inertia * ksi ** 2).sum())


def run_oracle(name, nz=4):
    prm, soa, bnd, st = setup(name)
    n = soa.n
    st.update(c0x=soa.vx.copy(), c0y=soa.vy.copy(), stress_h=np.zeros((n, nz, 4)), stress_count=np.ones(n, np.int32), stress=np.zeros((n, 2, 2)))
    K = [kinetic(st["mass"], st["inertia"], soa.u, soa.v, soa.ksi)]
    rows = []
    for _ in range(PLAN[name][1]):
        step = oracle.OracleStep(prm, soa, bnd, broad_mode=0)
        rows.append(int(step.summary.n_rows))
        sacked, unsup = oracle.calc_trajectory(step, soa, st, prm.dt, 0.0, nz=nz)
        assert not sacked.any() and not unsup.any()
        K.append(kinetic(st["mass"], st["inertia"], soa.u, soa.v, soa.ksi))
    return np.array(K), np.array(rows), soa


@pytest.mark.parametrize("name", list(PLAN))
def test_oracle_collisions_do_not_create_energy(name):
    K, rows, soa = run_oracle(name)
    assert rows[:10].sum() == 0 and rows.max() > 0 and rows[-100:].sum() == 0        # free flight, a collision, separated again
    assert soa.alive.all()
    assert K[-1] / K[0] < 1                                                          # conservation_test.m:26,33,41,48,54
    assert K[-1] / K[0] > 0.2 and K.min() / K[0] < 0.5                               # a real rebound: energy went into the contact and came back out


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(PLAN))
def test_device_collisions_do_not_create_energy(name):
    nz = 4
    Kref, rows_ref, soa_ref = run_oracle(name, nz)
    prm, soa, bnd, st = setup(name)
    K = [kinetic(st["mass"], st["inertia"], soa.u, soa.v, soa.ksi)]
    with sz.ContactContext(0) as ctx:
        ctx.upload(prm, soa, bnd)
        ctx.trajectory_init(st["mass"], st["inertia"], nz=nz, dXi_p=st["dXi_p"], dYi_p=st["dYi_p"])
        rows = []
        for it in range(PLAN[name][1]):
            s = ctx.step_resident()
            rows.append(int(s.n_rows))
            assert ctx.trajectory_step(prm.dt) == 0
            if it % 20 == 19 or it == PLAN[name][1] - 1:
                g = ctx.trajectory_state()
                K.append(kinetic(g["mass"], g["inertia"], g["u"], g["v"], g["ksi"]))
                assert K[-1] == pytest.approx(Kref[it + 1], rel=1e-6), it
        g = ctx.trajectory_state()
    assert np.array_equal(np.array(rows) > 0, rows_ref > 0)                          # contact begins and ends on the same steps
    assert K[-1] / K[0] < 1                                                          # the reference's assertion, on the device
    np.testing.assert_allclose(g["x"], soa_ref.x, rtol=1e-9)
    np.testing.assert_allclose(g["y"], soa_ref.y, rtol=1e-9)
