"""calc_trajectory.m (no-ocean branch): the oracle's restatement against hand-derived answers (CPU),
and the device integrator against the oracle over several coupled contact + trajectory steps (GPU)."""
import numpy as np
import pytest

import oracle
import subzero_b200 as sz

RHO_ICE = 920.0


def make_state(soa, nz, rng=None):
    n, nv = soa.n, soa.vx.shape[0]
    mass = soa.area * soa.h * RHO_ICE                                    # initialize_floe_values.m:16
    inertia = mass * soa.rmax ** 2 / 4                                   # a plausible PolygonMoments value (input data, not part of the path)
    st = {k: np.zeros(n) for k in ("alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p", "FxOA", "FyOA", "torqueOA")}
    if rng is not None:
        st["dXi_p"], st["dYi_p"] = soa.u + rng.normal(0, 1e-3, n), soa.v + rng.normal(0, 1e-3, n)
        st["FxOA"], st["FyOA"], st["torqueOA"] = rng.normal(0, 1e-3, n), rng.normal(0, 1e-3, n), rng.normal(0, 1e-2, n)
    st.update(mass=mass.copy(), inertia=inertia.copy(), c0x=soa.vx.copy(), c0y=soa.vy.copy(), stress_h=np.zeros((n, nz, 4)), stress_count=np.ones(n, np.int32), stress=np.zeros((n, 2, 2)))
    return st


def copy_soa(s):
    return sz.FloesSoA(s.x.copy(), s.y.copy(), s.rmax.copy(), s.h.copy(), s.area.copy(), s.u.copy(), s.v.copy(), s.ksi.copy(), s.alive.copy(), s.voff.copy(), s.vx.copy(), s.vy.copy())


def test_oracle_free_floe_known_answers():
    """one floe, no contacts, no forcing: AB2 drift x += 1.5 dt U - 0.5 dt dXi_p, heading from ksi, outline rotated by alpha,
    velocities unchanged; then a thin floe is reported, a floe outside the ocean grid is sacked and keeps its state"""
    prm, soa = sz.voronoi_field(16, seed=0)
    one = sz.FloesSoA(soa.x[:1], soa.y[:1], soa.rmax[:1], soa.h[:1], soa.area[:1], np.array([0.2]), np.array([-0.1]), np.array([4e-6]), soa.alive[:1],
                      soa.voff[:2], soa.vx[:soa.voff[1]], soa.vy[:soa.voff[1]])
    prm.collision = 0
    st = make_state(one, 5)
    st["dXi_p"][:] = 0.1
    x0, y0, c0 = one.x.copy(), one.y.copy(), (one.vx.copy(), one.vy.copy())
    step = oracle.OracleStep(prm, one)
    sacked, unsup = oracle.calc_trajectory(step, one, st, dt=10.0, nz=5)
    assert not sacked[0] and not unsup[0]
    assert one.x[0] == x0[0] + (1.5 * 10 * 0.2 - 0.5 * 10 * 0.1) and one.y[0] == y0[0] + 1.5 * 10 * -0.1
    assert st["alpha"][0] == 1.5 * 10 * 4e-6 and st["dalpha_p"][0] == 4e-6 and st["dXi_p"][0] == 0.2
    assert one.u[0] == 0.2 and one.v[0] == -0.1 and one.ksi[0] == 4e-6
    a = st["alpha"][0]
    np.testing.assert_allclose(one.vx, np.cos(a) * c0[0] - np.sin(a) * c0[1], rtol=0, atol=1e-9)
    assert st["stress_count"][0] == 2 and np.all(st["stress"] == 0)
    # out of the ocean grid: sacked, state untouched
    before = (one.x.copy(), st["mass"].copy(), st["stress_count"].copy())
    sacked, _ = oracle.calc_trajectory(oracle.OracleStep(prm, one), one, st, dt=10.0, bounds=(-1.0, 1.0, -1.0, 1.0), nz=5)
    assert sacked[0] and one.x[0] == before[0][0] and st["mass"][0] == before[1][0] and st["stress_count"][0] == before[2][0]
    # thermodynamic thinning below 0.1 m needs the ocean: reported
    _, unsup = oracle.calc_trajectory(oracle.OracleStep(prm, one), one, st, dt=10.0, HFo=0.02, nz=5)
    assert unsup[0]


def test_oracle_force_clamp_and_velocity_limiter():
    """a huge collision force is divided by 10 until max|F| <= mass/(5 dt) (:42-46) and the velocity change is limited to
    0.5 h / dt with the same fraction applied to the spin (:184-213)"""
    prm, soa = sz.voronoi_field(16, seed=1)
    one = sz.FloesSoA(soa.x[:1], soa.y[:1], soa.rmax[:1], soa.h[:1], soa.area[:1], np.zeros(1), np.zeros(1), np.zeros(1), soa.alive[:1],
                      soa.voff[:2], soa.vx[:soa.voff[1]], soa.vy[:soa.voff[1]])
    st = make_state(one, 3)
    m, h, dt = st["mass"][0], one.h[0], 10.0

    def run(fx, fy, tq):
        class FakeStep:                   # collision_force / torque as the contact step would hand them over
            def floe_outputs(self):
                return {"xi": one.x.copy(), "yi": one.y.copy(), "alive": one.alive.copy(), "fx": np.array([fx]), "fy": np.array([fy]), "torque": np.array([tq]),
                        "stress": np.ones((1, 2, 2))}

            def rows(self):
                return np.array([0, 1]), np.zeros((1, 7))
        s2 = make_state(one, 3)
        o2 = copy_soa(one)
        oracle.calc_trajectory(FakeStep(), o2, s2, dt=dt, nz=3)
        return o2, s2
    # clamp: 7 m -> 0.7 m -> 0.07 m -> 0.007 m <= m / (5 dt) = 0.02 m; the torque follows; then no limiter (dt dU = 0.07 < 0.5 h)
    o2, s2 = run(m * 7.0, 0.0, 5e8)
    assert s2["dUi_p"][0] == pytest.approx(0.007, rel=1e-12) and s2["dksi_p"][0] == pytest.approx(5e8 / 1000 / s2["inertia"][0], rel=1e-12)
    assert o2.u[0] == pytest.approx(1.5 * dt * 0.007, rel=1e-12)
    # limiter: dU = 0.018 (dt dU = 0.18 > 0.5 h = 0.125), dV small -> dU cut to 0.5 h / dt, dV and the spin by the same fraction
    o2, s2 = run(m * 0.018, m * 1e-4, 3e6)
    frac = (0.5 * h / dt) / 0.018
    assert s2["dUi_p"][0] == pytest.approx(0.5 * h / dt, rel=1e-12)
    assert s2["dVi_p"][0] == pytest.approx(frac * 1e-4, rel=1e-12)
    assert s2["dksi_p"][0] == pytest.approx(frac * 3e6 / s2["inertia"][0], rel=1e-12)
    assert np.allclose(s2["stress"][0], 1.0 / 3)                               # mean over a 3-deep history holding one sample


@pytest.mark.gpu
@pytest.mark.parametrize("n,nz,hfo", [(3000, 4, 0.0), (800, 1000, 1e-4)])
def test_device_trajectory_matches_oracle_over_coupled_steps(n, nz, hfo):
    """contact step -> trajectory step, three times, state resident on the device, against the same sequence on the oracle.
    The first step is bit-identical; later steps agree to 1e-9 (the outline rotation uses cos/sin, whose last bit differs
    between CUDA and the host libm and is amplified by nothing but the contact law itself)."""
    rng = np.random.default_rng(n)
    prm, soa = sz.voronoi_field(n, seed=5)
    prm.dt = 10.0
    ref_soa = copy_soa(soa)
    st = make_state(soa, nz, rng)
    L = prm.Lx
    bounds = (-1.2 * L, 1.2 * L, -1.2 * L, 1.2 * L)
    with sz.ContactContext(0) as ctx:
        ctx.upload(prm, soa)
        ctx.trajectory_init(st["mass"], st["inertia"], nz=nz, **{k: st[k] for k in ("dXi_p", "dYi_p", "FxOA", "FyOA", "torqueOA")})
        for it in range(3):
            ctx.step_resident()
            ns = ctx.trajectory_step(prm.dt, hfo, *bounds)
            ref = oracle.OracleStep(prm, ref_soa, broad_mode=1)
            sacked, unsup = oracle.calc_trajectory(ref, ref_soa, st, prm.dt, hfo, bounds, nz)
            assert ns == int(sacked.sum()) and not unsup.any()
            got = ctx.trajectory_state(nverts=soa.vx.shape[0])
            tol = 0.0 if it == 0 else 1e-9
            for k, want in (("x", ref_soa.x), ("y", ref_soa.y), ("u", ref_soa.u), ("v", ref_soa.v), ("ksi", ref_soa.ksi), ("h", ref_soa.h), ("mass", st["mass"]),
                            ("inertia", st["inertia"]), ("alpha", st["alpha"]), ("dXi_p", st["dXi_p"]), ("dUi_p", st["dUi_p"]), ("dVi_p", st["dVi_p"]), ("dksi_p", st["dksi_p"]),
                            ("stress", st["stress"])):
                scale = max(np.abs(want).max(), 1e-300)
                err = np.abs(got[k] - want).max() / scale
                assert err <= tol, (it, k, err)
            assert np.array_equal(got["alive"], ref_soa.alive)
            np.testing.assert_allclose(got["cax"], ref_soa.vx, rtol=0, atol=1e-9 * np.abs(ref_soa.vx).max())
            assert (got["flags"] & 1).sum() == sacked.sum()
    assert np.abs(st["alpha"]).max() > 0 and np.abs(st["dUi_p"]).max() > 0
