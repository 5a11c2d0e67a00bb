"""Parity of the CUDA path (called through the C ABI) with the oracle on identical inputs.
Bars (BASELINE.json north_star): ghosts, candidate-pair lists and Clipper int64 polygons bit-exact;
forces, torques, overlap areas, stress within 1e-9 relative (FP64).  Needs a B200."""
import numpy as np
import pytest

import oracle
import scenarios
import subzero_b200 as sz

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    with sz.ContactContext(0) as c:
        yield c


def run_both(ctx, prm, soa, bnd=None, broad_mode=0, bit_exact=False):
    prm.want_clip_polys = 1
    before = sz.abi.lib().sz_launch_count()
    ctx.step(prm, soa, bnd, allow_pair_errors=True)
    assert sz.abi.lib().sz_launch_count() > before            # our kernels ran
    ref = oracle.OracleStep(prm, soa, bnd, broad_mode=broad_mode)
    return oracle.compare_steps(ctx, ref, rtol=RTOL, bit_exact=bit_exact), ref


@pytest.mark.parametrize("n,seed", [(64, 0), (500, 1), (5000, 2), (20000, 3)])
def test_periodic_voronoi_field(ctx, n, seed):
    prm, soa = sz.voronoi_field(n, seed=seed)
    # same operation order, no FMA: rows and per-floe outputs are demanded bit for bit, not just within 1e-9
    rep, ref = run_both(ctx, prm, soa, broad_mode=0 if n <= 5000 else 1, bit_exact=True)
    assert ref.summary.n > ref.summary.n0                       # ghosts exist
    assert rep["pairs"] > 4 * n and rep["rows"] > 0


@pytest.mark.parametrize("inflate,seed", [(0.0003, 41), (0.002, 42), (0.1, 43), (0.35, 44)])
def test_shortcuts_thin_and_deep_overlaps(ctx, inflate, seed):
    """margin-certified shortcuts vs the oracle's full evaluation: 0.5 m strips (the 1 m nudge can empty the re-clip)
    up to deep overlaps with merges (kill/transfer)"""
    prm, soa = sz.voronoi_field(6000, seed=seed, inflate=inflate)
    rep, ref = run_both(ctx, prm, soa, broad_mode=1)
    assert rep["pairs"] > 20000
    if inflate >= 0.35:
        assert (ref.floe_outputs()["kill"] > 0).sum() > 0


@pytest.mark.parametrize("inflate,seed", [(0.02, 51), (0.0, 52), (0.0003, 53), (0.35, 54)])
def test_class_c_equals_class_s(ctx, inflate, seed):
    """the convex fast path (class C: four-edge sweep + certified sign test) against the general sweep (class S) on the
    same field: Clipper polygons, pair states and contact rows must be identical bit for bit, and both equal the oracle.
    inflate 0.02 is the benchmark geometry (class C keeps practically every pair); 0.0 (exactly shared edges) and 0.0003
    (strips thinner than the 1 m nudge) make class C decline pairs, which class S then re-runs."""
    prm, soa = sz.voronoi_field(8000, seed=seed, inflate=inflate)
    prm.want_clip_polys = 1
    res = {}
    for fast in (1, 0):
        ctx.set_option("convex_fast", fast)
        try:
            ctx.step(prm, soa, allow_pair_errors=True)
        finally:
            ctx.set_option("convex_fast", 1)
        cls = ctx.narrow_class_ms()
        res[fast] = (ctx.pairs(), ctx.rows(), ctx.clip_polys(), cls)
        if fast:
            ref = oracle.OracleStep(prm, soa, broad_mode=1)
            oracle.compare_steps(ctx, ref, rtol=RTOL)
    (p1, r1, c1, k1), (p0, r0, c0, k0) = res[1], res[0]
    for key in ("i", "j", "status", "n_regions", "overlap_state"):
        assert np.array_equal(p1[key], p0[key]), key
    assert np.array_equal(r1[0], r0[0]) and np.array_equal(r1[1], r0[1], equal_nan=True)
    for a, b in zip(c1, c0):
        assert np.array_equal(a, b)
    n_c, n_s = k1["C"][1], k1["S"][1]
    assert n_c > 10000 and k0["C"][1] == n_c                       # with the switch off the same list runs in class S
    if inflate == 0.02:
        assert n_s < 0.01 * n_c, (n_c, n_s)                        # the benchmark geometry stays on the fast path
    if inflate == 0.0:
        assert n_s > 0                                             # shared edges: some pairs are declined and re-run


def test_uninflated_voronoi_shared_edges_give_no_polygons(ctx):
    """exactly shared edges: every Clipper intersection must come back empty, as in the reference"""
    prm, soa = sz.voronoi_field(3000, seed=4, inflate=0.0)
    rep, ref = run_both(ctx, prm, soa)
    assert rep["pairs"] > 10000
    assert ref.summary.n_clip_paths <= 0.05 * rep["pairs"]        # only rounding slivers survive


def test_dead_and_nan_floes_and_collision_off(ctx):
    prm, soa = sz.voronoi_field(2000, seed=5)
    soa.alive[::7] = 0
    soa.x[5] = np.nan
    soa.y[11] = np.nan
    run_both(ctx, prm, soa)
    prm.collision = 0
    rep, ref = run_both(ctx, prm, soa)
    assert rep["pairs"] == 0 and ref.summary.n > ref.summary.n0


def test_boundary_floes_nb(ctx):
    """floes 1..Nb (topography) never start a pair (floe_interactions_all.m:76)"""
    prm, soa = sz.voronoi_field(1500, seed=6)
    prm.Nb = 40
    rep, ref = run_both(ctx, prm, soa)
    assert ref.pairs()["i"].min() > 40


def test_boundary_floes_opt_in_pairing(ctx):
    """SzParams.pair_with_boundary_floes (opt-in, not the reference): floes i > Nb also record the topography
    floes below them, ahead of their other partners; no mirrored rows, no kill / transfer from such pairs"""
    prm, soa = sz.voronoi_field(1500, seed=6)
    prm.Nb = 40
    prm.pair_with_boundary_floes = 1
    rep, ref = run_both(ctx, prm, soa, broad_mode=1)
    p = ref.pairs()
    assert (p["j"] <= 40).sum() > 50 and p["i"].min() > 40
    off, rows = ref.rows()
    assert off[40] == 0                                         # the topography floes carry no rows


def test_small_periodic_domain_big_floes(ctx):
    """2(rmax_i+rmax_j) > min(2Lx,2Ly): the ghost de-dup exemption of floe_interactions_all.m:103"""
    prm, soa = sz.voronoi_field(12, seed=8)
    rep, ref = run_both(ctx, prm, soa)
    assert ref.summary.n > ref.summary.n0 and rep["pairs"] > 12


def test_non_periodic_with_walls(ctx):
    """same field, PERIODIC = 0: floes poking through the domain wall get wall rows (partner Inf), floes whose
    centroid left the domain die (floe_interactions_all.m:150-172)"""
    prm, soa = sz.voronoi_field(1500, seed=9)
    prm.periodic = 0
    soa.x[:20] += np.sign(soa.x[:20]) * 1500.0                   # push a few across / onto the wall
    soa, bnd = soa, scenarios.soa_and_boundary([], prm, periodic=False)[1]
    rep, ref = run_both(ctx, prm, soa, bnd)
    off, rows = ref.rows()
    assert np.isinf(rows[:, 0]).sum() > 10
    assert ref.floe_outputs()["alive"].sum() < soa.n


@pytest.mark.parametrize("max_vertices", [None, 30])
def test_apart_certificate_saves_sweeps_and_changes_no_output(ctx, max_vertices):
    """the classifier's edge-by-edge certificate for concave outlines (sz_apart.cuh, option "apart"): with it the step equals the
    oracle like any other, and equals the step without it bit for bit -- pair states, Clipper polygons, rows, per-floe outputs --
    while a good part of the candidate pairs is answered without a sweep"""
    prm, Floe = scenarios.real_shape_field(10, seed=6, max_vertices=max_vertices)
    soa = sz.floes_to_soa(Floe)
    prm.want_clip_polys = 1
    res = {}
    for on in (0, 1):
        ctx.set_option("apart", on)
        s = ctx.step(prm, soa, allow_pair_errors=True)
        off, rows = ctx.rows()
        res[on] = (ctx.floe_outputs(), off.copy(), rows.copy(), ctx.pairs(), ctx.clip_polys(), ctx.stat("classifier_answered"), s.n_pairs)
    o0, off0, rows0, p0, c0, a0, np0 = res[0]
    o1, off1, rows1, p1, c1, a1, np1 = res[1]
    assert np0 == np1 and a1 >= a0 + 0.2 * np1, (a0, a1, np1)           # the certificate answers a fifth of all pairs at least
    assert np.array_equal(off0, off1) and np.array_equal(rows0, rows1)
    for k in o0:
        assert np.array_equal(o0[k], o1[k], equal_nan=True), k
    for k in p0:
        assert np.array_equal(p0[k], p1[k], equal_nan=True), k
    for a, b in zip(c0, c1):
        assert np.array_equal(a, b)
    rep, ref = run_both(ctx, prm, soa)                                   # option on (the default): against the oracle
    assert rep["pairs"] == np1


def test_real_concave_shapes_periodic(ctx):
    """tiled FloeShapes.mat polygons (7..591 vertices): size classes M and L, multi-region contacts, merges"""
    prm, Floe = scenarios.real_shape_field(12, seed=1)
    soa = sz.floes_to_soa(Floe)
    rep, ref = run_both(ctx, prm, soa)
    pr = ref.pairs()
    assert rep["pairs"] > 200 and (pr["n_regions"] > 1).sum() > 5
    assert np.isinf(pr["overlap_state"]).sum() > 0               # kill / transfer path exercised
    o = ref.floe_outputs()
    assert (o["kill"] > 0).sum() > 0


def test_simplified_concave_shapes_class_t(ctx):
    """real shapes thinned to <= 30 vertices (what FloeSimplify leaves, Subzero.m:169-217): the local-memory class T"""
    prm, Floe = scenarios.real_shape_field(20, seed=5, max_vertices=30)
    soa = sz.floes_to_soa(Floe)
    assert (soa.voff[1:] - soa.voff[:-1]).max() <= 31
    rep, ref = run_both(ctx, prm, soa)
    assert rep["pairs"] > 1000 and (ref.pairs()["n_regions"] > 1).sum() > 10


def test_real_concave_shapes_with_walls(ctx):
    prm, Floe = scenarios.real_shape_field(8, seed=2, periodic=False)
    soa, bnd = scenarios.soa_and_boundary(Floe, prm, periodic=False)
    rep, ref = run_both(ctx, prm, soa, bnd)
    off, rows = ref.rows()
    assert np.isinf(rows[:, 0]).sum() > 3


def test_conservation_test_scenarios(ctx):
    """the five set-ups of the reference's test/conservation_test.m, advanced to contact"""
    cases, modulus = scenarios.conservation_cases()
    for name, t in (("head_on", 42000.0), ("offset", 50000.0), ("triangle_between", 30000.0), ("complex_pair", 30000.0), ("complex_wall", 60000.0)):
        Floe = scenarios.advance(cases[name], t)
        prm = sz.default_params(Lx=1e5, Ly=1e5, modulus=modulus, dt=10.0, periodic=0, collision=1)
        soa, bnd = scenarios.soa_and_boundary(Floe, prm, periodic=False)
        rep, ref = run_both(ctx, prm, soa, bnd)
        assert rep["rows"] > 0, name


def test_empty_and_single_floe(ctx):
    prm, soa = sz.voronoi_field(8, seed=0)
    one = sz.FloesSoA(soa.x[:1], soa.y[:1], soa.rmax[:1], soa.h[:1], soa.area[:1], soa.u[:1], soa.v[:1], soa.ksi[:1], soa.alive[:1],
                      soa.voff[:2], soa.vx[:soa.voff[1]], soa.vy[:soa.voff[1]])
    run_both(ctx, prm, one)
    z = np.zeros(0)
    none = sz.FloesSoA(z, z, z, z, z, z, z, z, np.zeros(0, np.uint8), np.zeros(1, np.int32), z, z)
    s = ctx.step(prm, none)
    assert (s.n0, s.n, s.n_pairs, s.n_rows) == (0, 0, 0, 0)


def test_resident_step_equals_host_step(ctx):
    prm, soa = sz.voronoi_field(4000, seed=12)
    ctx.step(prm, soa)
    a = ctx.floe_outputs()
    ra = ctx.rows()
    ctx.upload(prm, soa)
    for _ in range(2):
        ctx.step_resident()
    b = ctx.floe_outputs()
    rb = ctx.rows()
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    np.testing.assert_array_equal(ra[1], rb[1])                   # idempotent and deterministic, row for row


def test_clip_batch_fuzz_vs_reference_clipper(ctx):
    """the stand-alone clip entry point (mex gateway semantics) against the unmodified reference Clipper:
    random stars, shared edges, grid-degenerate and rectilinear polygons, 3..400 vertices, all four clip types"""
    rng = np.random.default_rng(3)
    S = 2.0 ** 32

    def star(cx, cy, r, n, jit):
        a = rng.uniform(0, 2 * np.pi) - 2 * np.pi * np.arange(n) / n
        rr = r * (1 - jit + jit * rng.uniform(size=n))
        return np.stack([np.rint((cx + rr * np.cos(a)) * S), np.rint((cy + rr * np.sin(a)) * S)], 1).astype(np.int64)

    subj, clip, meth = [], [], []
    for t in range(3000):
        kind = t % 6
        if kind == 0:
            d, a = 1000 + 1500 * rng.uniform(), rng.uniform(0, 6.28)
            ox, oy = rng.uniform(-1e6, 1e6, 2)
            s, c = star(ox, oy, 1000, rng.integers(3, 10), 0.3), star(ox + d * np.cos(a), oy + d * np.sin(a), 1000, rng.integers(3, 10), 0.3)
        elif kind == 1:
            d = 3000 * rng.uniform()
            s, c = star(0, 0, 2000, rng.integers(8, 60), 0.7), star(d, 0.2 * d, 2000, rng.integers(8, 60), 0.7)
        elif kind == 2:
            s, c = rng.integers(0, 7, (rng.integers(3, 9), 2)), rng.integers(0, 7, (rng.integers(3, 9), 2))
        elif kind == 3:
            s = star(0, 0, 1500, rng.integers(4, 12), 0.4)
            c = s[::-1].copy() + rng.integers(-3, 4, 2) * (1 << 31)
        elif kind == 4:
            s, c = star(0, 0, 3000, rng.integers(100, 400), 0.8), star(2000 * rng.uniform(), 500, 3000, rng.integers(100, 400), 0.8)
        else:
            s, c = rng.integers(0, 9, (rng.integers(3, 12), 2)) << 32, rng.integers(0, 9, (rng.integers(3, 12), 2)) << 32
        subj.append(np.asarray(s, np.int64)); clip.append(np.asarray(c, np.int64)); meth.append(1 if t % 5 else (t // 5) % 4)
    status, out = ctx.clip_batch(subj, clip, meth)
    nonempty = 0
    for k in range(len(subj)):
        ref = oracle.ref_clip(subj[k], clip[k], meth[k])
        if ref is None:
            assert status[k] == sz.abi.SZ_ERR_CLIPPER, k
            continue
        assert status[k] == 0, (k, status[k])
        assert len(ref) == len(out[k]), k
        for a, b in zip(ref, out[k]):
            assert np.array_equal(a, b), k
        nonempty += len(ref) > 0
    assert nonempty > 1500


def test_full_size_properties_100k(ctx):
    """at a size the oracle's literal loop cannot do: Newton's third law over the periodic field (every mirrored
    row cancels its source exactly, so the column sums over the extended list vanish to rounding), row bookkeeping
    invariants, and agreement with the oracle's grid mode on the same input"""
    prm, soa = sz.voronoi_field(100000, seed=0)
    rep, ref = run_both(ctx, prm, soa, broad_mode=1)
    off, rows = ctx.rows()
    s = ctx.summary
    assert rows.shape[0] == s.n_rows == 2 * (ref.pairs()["n_regions"].sum())
    scale = np.abs(rows[:, 1:3]).sum()
    assert abs(rows[:, 1].sum()) < 1e-9 * scale and abs(rows[:, 2].sum()) < 1e-9 * scale
    assert np.all(np.diff(off) >= 0) and off[-1] == s.n_rows
    assert s.collision_count <= s.n_rows / 2


def test_full_benchmark_size_1m_floes(ctx):
    """BASELINE.json's headline size (bench.py's field: 1M floes, seed 0): size-independent properties -- Newton's third
    law over the periodic field, row bookkeeping, run-to-run determinism -- and, since the oracle finishes it in seconds on
    the host cores, the complete comparison as well: 4.5M candidate pairs, 6.1M contact rows, every per-floe output"""
    prm, soa = sz.voronoi_field(1000000, seed=0)
    prm.want_clip_polys = 0
    s = ctx.step(prm, soa)
    assert s.n_pairs > 4.4e6 and s.n_pairs_force > 2.9e6 and s.n_capacity_fail == 0 and s.n_clipper_fail == 0
    off, rows = ctx.rows()
    out = ctx.floe_outputs()
    scale = np.abs(rows[:, 1:3]).sum()
    assert abs(rows[:, 1].sum()) < 1e-9 * scale and abs(rows[:, 2].sum()) < 1e-9 * scale         # every mirrored row cancels its source
    assert rows.shape[0] == s.n_rows and off[-1] == s.n_rows and np.all(np.diff(off) >= 0)
    assert s.collision_count * 2 == np.isfinite(rows[:off[s.n0], 0]).sum()
    ctx.upload(prm, soa)
    ctx.step_resident()
    off2, rows2 = ctx.rows()
    assert np.array_equal(off, off2) and np.array_equal(rows, rows2)                                 # deterministic, row for row
    del rows2, off2
    ref = oracle.OracleStep(prm, soa, broad_mode=1)
    rep = oracle.compare_steps(ctx, ref, rtol=RTOL, check_polys=False)
    assert rep["rows_bit_exact"] and rep["fx_bit_exact"] and rep["torque_bit_exact"] and rep["stress_bit_exact"]



def soa_to_floe_dicts(soa):
    """the reference's struct array as a list of dicts (the fields the contact loop reads, initialize_floe_values.m:12-52)"""
    out = []
    for i in range(soa.n):
        a, b = int(soa.voff[i]), int(soa.voff[i + 1])
        out.append({"c_alpha": np.stack([soa.vx[a:b], soa.vy[a:b]]), "Xi": soa.x[i], "Yi": soa.y[i], "rmax": soa.rmax[i], "h": soa.h[i], "area": soa.area[i],
                    "Ui": soa.u[i], "Vi": soa.v[i], "ksi_ice": soa.ksi[i], "alive": int(soa.alive[i])})
    return out


def test_drop_in_signature_with_ghost_structs(ctx):
    """floe_interactions_all with the reference's argument list (floe_interactions_all.m:1; the Python twin of
    subzero_b200/matlab/floe_interactions_all.m) on a periodic field: per-floe fields against the oracle, and -- with RIDGING
    on -- the ghost structs Floe(N0+1:N) the reference's tail indexes (:312,327,401,416): parents, shifted centroids, their
    own rows and column sums, and partner numbers > N0 in the originals' interactions resolving to them."""
    from subzero_b200.contact import floe_interactions_all
    prm, soa = sz.voronoi_field(900, seed=12)
    c2 = np.array([[-prm.Lx, -prm.Lx, prm.Lx, prm.Lx, -prm.Lx], [-prm.Ly, prm.Ly, prm.Ly, -prm.Ly, -prm.Ly]])
    ref = oracle.OracleStep(prm, soa, broad_mode=0)
    roff, rrow = ref.rows()
    ro, rg = ref.floe_outputs(), ref.ghosts()
    N0, N = ref.summary.n0, ref.summary.n
    assert N > N0
    for ridging in (False, True):
        Floe, dis, kill, transfer = floe_interactions_all(soa_to_floe_dicts(soa), None, None, None, c2, prm.dt, 0.0, 0.0, 1, 1, 0, 0.0, {"flag": True},
                                                          True, True, ridging, False, Modulus=prm.modulus, ctx=ctx)
        assert len(Floe) == (N if ridging else N0) and len(kill) == N0
        for i in range(N0):
            f = Floe[i]
            assert np.array_equal(f["interactions"], rrow[roff[i]:roff[i + 1]], equal_nan=True)
            assert f["OverlapArea"] == ro["overlap_area"][i] and f["collision_torque"] == ro["torque"][i]
            assert f["collision_force"][0] == ro["fx"][i] and f["collision_force"][1] == ro["fy"][i]
            assert f["Xi"] == ro["xi"][i] and f["Yi"] == ro["yi"][i] and f["potentialInteractions"] == []
        assert np.array_equal(kill, ro["kill"]) and np.array_equal(transfer, ro["transfer"])
    seen_ghost_partner = False
    for k in range(N - N0):
        g = Floe[N0 + k]
        p = int(rg["parent"][k]) - 1
        assert g["Xi"] == rg["x"][k] and g["Yi"] == rg["y"][k]
        assert g["area"] == Floe[p]["area"] and g["h"] == Floe[p]["h"] and np.array_equal(g["c_alpha"], Floe[p]["c_alpha"])
        rows = rrow[roff[N0 + k]:roff[N0 + k + 1]]
        assert np.array_equal(g["interactions"], rows, equal_nan=True)
        sx = sy = st = 0.0
        for r in rows:                                  # MATLAB's column sum, top to bottom (:234-235)
            sx, sy, st = sx + r[1], sy + r[2], st + r[5]
        assert g["collision_force"][0] == sx and g["collision_force"][1] == sy and g["collision_torque"] == st
        assert g["OverlapArea"] == pytest.approx(rows[:, 6].sum(), rel=1e-12, abs=0)
    for i in range(N0):
        partners = Floe[i]["interactions"][:, 0]
        for q in partners[np.isfinite(partners)]:
            if q > N0:
                seen_ghost_partner = True
                assert abs(int(rg["floe_num"][int(q) - 1 - N0])) >= 1 and Floe[int(q) - 1]["area"] > 0     # Floe(partner) exists for the tail
    assert seen_ghost_partner


def test_speculated_sizes_give_the_same_step_and_recover_from_overflow():
    """option "speculate": steps that take list length, grid, pair and row capacities from the step before must reproduce the
    fully measured step bit for bit; a step whose field outgrew the carried-over capacities (same floe count, far more pairs
    and rows) is flagged on the device and repeated; floes that drift (integrator) keep the speculation valid."""
    def run(spec, fields, traj):
        outs = []
        c = sz.ContactContext(0)
        try:
            c.set_option("speculate", spec)
            for prm, soa in fields:
                c.upload(prm, soa)
                if traj:
                    mass = soa.area * soa.h * 920.0
                    c.trajectory_init(mass, mass * soa.rmax ** 2 / 4, nz=2, dXi_p=soa.u, dYi_p=soa.v)
                for it in range(4):
                    s = c.step_resident(allow_pair_errors=True)
                    off, rows = c.rows()
                    outs.append((s.n, s.n_pairs, s.n_rows, s.collision_count, c.floe_outputs(), off, rows, c.pairs(), c.ghosts()))
                    if traj:
                        c.trajectory_step(prm.dt, 1e-4)
            return outs, c.stat("speculated_steps"), c.stat("repeated_steps")
        finally:
            c.close()
    prm_a, a = sz.voronoi_field(9000, seed=31, inflate=0.01)
    prm_b, b = sz.voronoi_field(9000, seed=32, inflate=0.2)          # same floe count: the plan survives the upload ...
    b.rmax[:] *= 2.0                                                 # ... but the candidate pairs quadruple (and rows grow with the deeper overlaps)
    for traj in (False, True):
        if traj:
            a.u[:] *= 80; a.v[:] *= 80; b.u[:] *= 80; b.v[:] *= 80
        want, f0, r0 = run(0, [(prm_a, a), (prm_b, b)], traj)
        got, f1, r1 = run(1, [(prm_a, a), (prm_b, b)], traj)
        assert f0 == 0 and r0 == 0 and f1 >= 5 and r1 >= 1, (f0, r0, f1, r1)
        for k, (w, g) in enumerate(zip(want, got)):
            assert w[:4] == g[:4], (traj, k, w[:4], g[:4])
            for key in w[4]:
                assert np.array_equal(w[4][key], g[4][key], equal_nan=True), (traj, k, key)
            assert np.array_equal(w[5], g[5]) and np.array_equal(w[6], g[6], equal_nan=True), (traj, k)
            for key in w[7]:
                assert np.array_equal(w[7][key], g[7][key], equal_nan=True), (traj, k, key)
            for key in w[8]:
                assert np.array_equal(w[8][key], g[8][key], equal_nan=True), (traj, k, key)


def test_floes_with_many_candidate_partners(ctx):
    """a few floes whose bounding radius is many times their neighbours' (rmax is an input: initialize_floe_values.m:21) collect
    40..300 candidate partners each: the in-warp ranked sort of 33..256 partners and the serial one beyond, against the oracle's
    ascending lists (floe_interactions_all.m:100-116)"""
    prm, soa = sz.voronoi_field(6000, seed=17)
    soa.rmax[100] *= 4.0
    soa.rmax[2500] *= 7.0
    soa.rmax[10] *= 20.0
    rep, ref = run_both(ctx, prm, soa, broad_mode=1)
    p = ref.pairs()
    per_i = np.bincount(p["i"], minlength=6001)
    assert 32 < per_i[101] <= 256 or 32 < per_i[2501] <= 256          # the ranked path ...
    assert per_i.max() > 256                                          # ... and the serial one both ran
