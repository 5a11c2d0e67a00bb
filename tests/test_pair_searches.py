"""SURVEY.md 8f row f4, second half: the bounding-radius pair searches of Physical_Processes/weld.m:29-81 and
polygon_operations/FloeSimplify.m:13-31.  The oracle's literal restatement against answers worked by hand from the MATLAB
source (CPU), and the device searches (cell grid, warp per floe) against the oracle through the C ABI (GPU)."""
import numpy as np
import pytest

import oracle
import subzero_b200 as sz


def discs(xy, r, alive=None):
    """floes as squares of half-diagonal r (only Xi, Yi, rmax and alive matter to the searches)"""
    n = len(xy)
    xy = np.asarray(xy, float)
    r = np.broadcast_to(np.asarray(r, float), (n,)).copy()
    vx = np.concatenate([[ri, 0, -ri, 0, ri] for ri in r])
    vy = np.concatenate([[0, ri, 0, -ri, 0] for ri in r])
    al = np.ones(n, np.uint8) if alive is None else np.asarray(alive, np.uint8)
    return sz.FloesSoA(xy[:, 0].copy(), xy[:, 1].copy(), r, np.full(n, 0.25), 2 * r * r, np.zeros(n), np.zeros(n), np.zeros(n), al,
                       (5 * np.arange(n + 1)).astype(np.int32), vx, vy)


def test_weld_search_hand_derived():
    """weld.m:31-81 by hand.  Domain [-100, 100]^2, Nx = Ny = 2, Nb = 1 (the first floe is a boundary floe and is cut, :25).
    Cut list (1-based): 1 (-50,-50) r 30 | 2 (-20,-50) r 10 | 3 (-50,-50.5) r 5 | 4 (10,-50) r 40 | 5 (-60,-60) r 30 dead | 6 (150, 0) r 500 | 7 (-50, 50) r 80
    Binx = fix((X+100)/200*2+1): floes 1 2 3 5 -> 1; 4 -> 2; 6 -> 3 (outside); 7 -> 1.  Biny: 1 2 3 4 5 -> 1; 7 -> 2; 6 -> 2.
    bin numbers (Binx-1)*Ny+Biny: [1 1 1 3 1 0 2].
    In bin 1 = {1, 2, 3, 5}: d(1,2) = 30 < 30+10 -> partners; d(1,3) = 0.5, not > 1 -> no (the "same floe" guard of :67);
    d(1,5) = 14.1 < 60 but floe 5 is dead: 5 is nobody's partner, yet 5 itself still records its live neighbours (alive(j) only);
    d(2,3) = sqrt(900.25) = 30.004 > 15 -> no;  d(5,2) = 41.2 > 40 -> no;  d(5,3) = 13.8 < 35 -> 5 records 3.
    Floe 4 is alone in bin 3 although floe 2 is 30 m away (d < 40+10): other bins are never searched (:56-81).  Floe 7 alone in bin 2."""
    f = discs([(0, 0), (-50, -50), (-20, -50), (-50, -50.5), (10, -50), (-60, -60), (150, 0), (-50, 50)], [5, 30, 10, 5, 40, 30, 500, 80],
              alive=[1, 1, 1, 1, 1, 0, 1, 1])
    b, off, pt = oracle.weld_search(f, 1, 2, 2, -100.0, 100.0, -100.0, 100.0)
    assert b.tolist() == [1, 1, 1, 3, 1, 0, 2]
    got = [pt[off[q]:off[q + 1]].tolist() for q in range(7)]
    assert got == [[2], [1], [], [], [1, 3], [], []]


def test_simplify_search_hand_derived():
    """FloeSimplify.m:13-31: the query floe against the WHOLE list -- no bins, no Nb cut, itself excluded by d > 1, dead floes
    excluded as partners: with the floes of the weld case, floe 3 (-20,-50) r 10 sees 2 (d 30 < 40) and 5 (d 30 < 50), not 4
    (d 30.004 > 15), not 6 (dead, d 41.2 > 40 anyway), and 7 (150,0) r 500 (d 177.2 < 510)."""
    f = discs([(0, 0), (-50, -50), (-20, -50), (-50, -50.5), (10, -50), (-60, -60), (150, 0), (-50, 50)], [5, 30, 10, 5, 40, 30, 500, 80],
              alive=[1, 1, 1, 1, 1, 0, 1, 1])
    off, pt = oracle.simplify_search(f, [3, 6])
    assert pt[off[0]:off[1]].tolist() == [2, 5, 7]
    # the dead floe 6 as a query still finds its live neighbours: 1 (d 84.9 > 35 no), 2 (14.1 yes), 3 (41.2 > 40 no), 4 (13.8 yes), 5 (70.7 > 70 no), 7 (218 < 530), 8 (110.5 > 110 no)
    assert pt[off[1]:off[2]].tolist() == [2, 4, 7]


@pytest.mark.gpu
@pytest.mark.parametrize("n,Nb,Nx,Ny", [(3000, 0, 4, 3), (20000, 25, 10, 10), (500, 3, 1, 1)])
def test_device_weld_search_matches_oracle(n, Nb, Nx, Ny):
    prm, soa = sz.voronoi_field(n, seed=n)
    rng = np.random.default_rng(n)
    soa.alive[rng.integers(0, n, n // 50)] = 0
    soa.x[7] = np.nan
    soa.x[11] = prm.Lx * 1.5                                   # outside the grid: in no bin
    soa.rmax[13] *= 12                                         # one big floe (wider search than the cell size)
    L = prm.Lx
    with sz.ContactContext(0) as ctx:
        ctx.upload(prm, soa)
        b, off, pt = ctx.weld_search(Nb, Nx, Ny, -L, L, -L, L)
    rb, roff, rpt = oracle.weld_search(soa, Nb, Nx, Ny, -L, L, -L, L)
    assert np.array_equal(b, rb) and np.array_equal(off, roff) and np.array_equal(pt, rpt)
    assert len(rpt) > 4 * (n - Nb) and (rb == 0).sum() >= 2


@pytest.mark.gpu
def test_device_simplify_search_matches_oracle():
    n = 6000
    prm, soa = sz.voronoi_field(n, seed=3)
    soa.alive[::9] = 0
    soa.rmax[100] *= 20
    idx = np.concatenate([np.arange(1, 400, 7), [101, n, 1]]).astype(np.int32)
    with sz.ContactContext(0) as ctx:
        ctx.upload(prm, soa)
        off, pt = ctx.simplify_search(idx)
    roff, rpt = oracle.simplify_search(soa, idx)
    assert np.array_equal(off, roff) and np.array_equal(pt, rpt)
    k = list(idx).index(101)
    assert roff[k + 1] - roff[k] > 40                          # the big floe reaches far beyond its own cell
