"""calc_eulerian_data.m on the device: sz_eulerian_data against the oracle.  The item / reduction code it
runs is the source tests/test_eulerian_oracle.py checks on the host; this adds the list, item and sort kernels and the C ABI."""
import numpy as np
import pytest

import oracle
import scenarios
import subzero_b200 as sz
from test_eulerian_oracle import square


def _check(ctx, soa, mass, Nx, Ny, box, periodic, **kw):
    got = ctx.eulerian_data(mass, Nx, Ny, box, periodic, **kw)
    want = oracle.calc_eulerian_data(soa, mass, Nx, Ny, box, periodic, **kw)
    for k in oracle.EULERIAN_FIELDS:
        assert got[k].shape == want[k].shape == (Ny, Nx)
        assert np.array_equal(got[k] == 0, want[k] == 0), k                      # the same cells were processed
        scale = max(np.abs(want[k]).max(), 1e-300)
        assert np.abs(got[k] - want[k]).max() <= 1e-12 * scale, (k, np.abs(got[k] - want[k]).max(), scale)
    return want


@pytest.mark.gpu
def test_device_eulerian_data_matches_oracle():
    box = (-4000.0, 4000.0, -4000.0, 4000.0)
    with sz.ContactContext(0) as ctx:
        # hand-derived: a 2 km square on a grid corner, a quarter of it in each central cell
        soa = sz.floes_to_soa([square(0.0, 0.0, u=0.3, v=-0.2)])
        prm = sz.default_params(Lx=4000.0, Ly=4000.0, modulus=1e7, dt=10.0, periodic=0, collision=1)
        mass = soa.area * soa.h * 920.0
        ctx.upload(prm, soa)
        w = _check(ctx, soa, mass, 4, 4, box, False, overlap_area=[5.0], dUi_p=[1e-3], dVi_p=[2e-3], stress=[[3.0, 1.0, 1.0, -2.0]], strain=[[1e-6, 2e-6, 3e-6, 4e-6]])
        assert np.allclose(w["c"][1:3, 1:3], 0.25, rtol=1e-12)
        # periodic images and the stale-polygon quirk of the y pass
        for fl, per in (([square(3500.0, 1000.0)], True), ([square(0.0, 3500.0), square(-2000.0, -2000.0)], True), ([square(-2000.0, -2000.0), square(0.0, 3500.0)], True),
                        ([square(3500.0, 1000.0)], False)):
            soa = sz.floes_to_soa(fl)
            ctx.upload(prm, soa)
            _check(ctx, soa, soa.area * soa.h * 920.0, 2, 2, box, per)
        # a Voronoi field with random per-floe state, dead floes and NaN entries, on several grids
        rng = np.random.default_rng(5)
        prm, soa = sz.voronoi_field(3000, seed=81, inflate=0.05)
        n = soa.n
        soa.alive[::17] = 0
        soa.u[5] = np.nan; soa.v[11] = np.nan
        mass = soa.area * soa.h * 920.0
        mass[3] = np.nan
        kw = dict(overlap_area=rng.uniform(0, 1e5, n), dUi_p=rng.normal(0, 1e-3, n), dVi_p=rng.normal(0, 1e-3, n), stress=rng.normal(0, 1e3, (n, 4)), strain=rng.normal(0, 1e-6, (n, 4)))
        ctx.upload(prm, soa)
        L = prm.Lx
        for (Nx, Ny), per in (((7, 5), True), ((1, 1), True), ((20, 20), True), ((10, 10), False), ((3, 40), True)):
            w = _check(ctx, soa, mass, Nx, Ny, (-L, L, -L, L), per, **kw)
            assert np.count_nonzero(w["Mtot"]) == Nx * Ny
        for b in ((-L / 3, L / 2, -L / 4, L / 5), (-2 * L, 2 * L, -3 * L, 3 * L)):
            _check(ctx, soa, mass, 6, 9, b, False, **kw)
        # the two forms of the per-cell reduction (a warp per cell = default, a thread per cell) give the same bits
        warp = ctx.eulerian_data(mass, 2, 3, (-L, L, -L, L), True, **kw)          # ~500 items per cell: many 32-item rounds per warp
        ctx.set_option("euler_cell_warp", 0)
        one = ctx.eulerian_data(mass, 2, 3, (-L, L, -L, L), True, **kw)
        _check(ctx, soa, mass, 7, 5, (-L, L, -L, L), True, **kw)
        ctx.set_option("euler_cell_warp", 1)
        for k in oracle.EULERIAN_FIELDS:
            assert np.array_equal(warp[k], one[k]), k
        # real concave shapes: items that overflow the class S arena run in class L
        prm_r, Floe = scenarios.real_shape_field(5, seed=4)
        soa, _ = scenarios.soa_and_boundary(Floe, prm_r, periodic=True)
        ctx.upload(prm_r, soa)
        L = prm_r.Lx
        w = _check(ctx, soa, soa.area * soa.h * 920.0, 8, 8, (-L, L, -L, L), True)
        assert np.count_nonzero(w["Mtot"]) > 40
