"""The classifier's "these outlines cannot intersect" certificate for outlines of any shape (subzero_b200/csrc/sz_apart.cuh),
compiled for the host (tests/host/libpair_host.so) and held against the reference's own Clipper and against the oracle: whenever
the certificate says yes, the reference's clip #1 (floe_interactions.m:29, polyclip 'int' on the 2^32-scaled int64 outlines)
must return nothing, and the oracle's pair must carry no rows and overlap 0 -- the zero-force branch (:43-44,71-74) that the
classifier then answers without a sweep.  CPU only; the device uses the same header (pair_classify_kernel)."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle                   # noqa: E402
import scenarios                # noqa: E402
import subzero_b200 as sz       # noqa: E402

HOST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host")
_dp = C.POINTER(C.c_double)
_lib = None


def apart(a, A, b, B):
    """a, b: (n, 2) outlines about their centroids; A, B: the centroids"""
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(HOST, "libpair_host.so"))
        _lib.szport_rings_apart.restype = C.c_int
        _lib.szport_rings_apart.argtypes = [_dp, _dp, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_int, C.c_double, C.c_double]
    a, b = np.ascontiguousarray(a, np.float64), np.ascontiguousarray(b, np.float64)
    ax, ay, bx, by = (np.ascontiguousarray(v) for v in (a[:, 0], a[:, 1], b[:, 0], b[:, 1]))
    p = lambda v: v.ctypes.data_as(_dp)
    return bool(_lib.szport_rings_apart(p(ax), p(ay), len(a), float(A[0]), float(A[1]), p(bx), p(by), len(b), float(B[0]), float(B[1])))


def to_clipper(a, A):
    """polyclip.m:66: int64(x * 2^32), MATLAB rounding (half away from zero)"""
    w = (np.asarray(a, np.float64) + np.asarray(A, np.float64)) * 4294967296.0
    return (np.sign(w) * np.floor(np.abs(w) + 0.5)).astype(np.int64)


def ref_intersection_is_empty(a, A, b, B):
    r = oracle.ref_clip(to_clipper(a, A), to_clipper(b, B), 1)
    assert r is not None
    return len(r) == 0


SQ = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]]) * 50.0            # a 100 m square
CSHAPE = np.array([[0, 0], [300, 0], [300, 100], [100, 100], [100, 200], [300, 200], [300, 300], [0, 300.0]]) - [150.0, 150.0]   # a "C", mouth to the right


def test_hand_cases():
    z = (0.0, 0.0)
    assert apart(SQ, z, SQ, (100.002, 0.0))                        # 2 mm between the facing edges
    assert not apart(SQ, z, SQ, (100.0005, 0.0))                   # 0.5 mm: inside the margin, left to the sweep
    assert not apart(SQ, z, SQ, (100.0, 0.0))                      # touching
    assert not apart(SQ, z, SQ, (60.0, 30.0))                      # overlapping
    assert not apart(SQ, z, SQ * 0.2, (3.0, -4.0))                 # B inside A, boundaries far apart: the intersection is B
    assert not apart(SQ * 0.2, (3.0, -4.0), SQ, z)                 # A inside B
    assert apart(SQ, z, SQ, (100.002, 100.002))                    # corner to corner
    # a square in the mouth of the C: bounding boxes nested, outlines disjoint
    small = SQ * 0.6                                               # 60 m square; the mouth is 100 m high and 200 m deep
    assert apart(CSHAPE, z, small, (60.0, 0.0)) and ref_intersection_is_empty(CSHAPE, z, small, (60.0, 0.0))
    assert not apart(CSHAPE, z, small, (60.0, 25.0))               # pokes into the upper jaw
    assert not apart(CSHAPE, z, small, (-100.0, 0.0))              # inside the C's spine: contained
    # closed rings (first point repeated) and either orientation give the same answers
    closed = np.vstack([SQ, SQ[:1]])
    assert apart(closed, z, closed[::-1], (100.002, 0.0)) and not apart(closed, z, closed[::-1], (100.0005, 0.0))
    # degenerate input is never certified
    assert not apart(SQ[:2], z, SQ, (500.0, 0.0))
    nan = SQ.copy(); nan[2, 0] = np.nan
    assert not apart(nan, z, SQ, (100.002, 0.0)) and not apart(SQ, z, nan, (100.002, 0.0))


def _star(rng, n, r):
    """a random simple concave outline: radii jittered around a circle"""
    th = np.sort(rng.uniform(0, 2 * np.pi, n))
    rr = r * rng.uniform(0.35, 1.0, n)
    return np.stack([rr * np.cos(th), rr * np.sin(th)], 1)


def test_certified_pairs_have_an_empty_reference_intersection():
    """random concave pairs at random offsets, from interlocked to far apart: a certified pair never intersects in the reference
    Clipper, and most of the bounding-box-overlapping pairs whose intersection IS empty are certified (the rule earns its keep)"""
    rng = np.random.default_rng(11)
    certified = empty_not_certified = nonempty = 0
    for case in range(1500):
        a, b = _star(rng, int(rng.integers(5, 60)), 1000.0), _star(rng, int(rng.integers(5, 60)), 1000.0)
        A = rng.uniform(-2e5, 2e5, 2)
        d = rng.uniform(600.0, 2100.0) * np.array([np.cos(t := rng.uniform(0, 2 * np.pi)), np.sin(t)])
        if case % 7 == 0:
            b, d = b * 0.15, d * 0.2                      # a small outline near or inside the big one
        if case % 2:
            b = b[::-1]
        B = A + d
        ap = apart(a, A, b, B)
        emp = ref_intersection_is_empty(a, A, b, B)
        assert not ap or emp, case
        assert ap == apart(b, B, a, A), case              # symmetric
        certified += ap; empty_not_certified += (emp and not ap); nonempty += (not emp)
    assert certified > 300 and nonempty > 300
    assert empty_not_certified < 0.05 * certified, (certified, empty_not_certified)


def test_near_misses_around_the_margin():
    """the same pair slid towards contact: certified while the gap is above the margin, never once the outlines touch"""
    rng = np.random.default_rng(3)
    for case in range(60):
        a, b = _star(rng, 24, 500.0), _star(rng, 31, 500.0)
        A = rng.uniform(-1e5, 1e5, 2)
        u = np.array([np.cos(t := rng.uniform(0, 2 * np.pi)), np.sin(t)])
        lo, hi = 0.0, 1200.0                               # bisect the offset along u at which the reference intersection appears
        for _ in range(60):
            mid = 0.5 * (lo + hi)
            if ref_intersection_is_empty(a, A, b, A + mid * u):
                hi = mid
            else:
                lo = mid
        assert not apart(a, A, b, A + lo * u)
        assert not apart(a, A, b, A + (hi + 2e-4) * u)      # 0.2 mm beyond first contact: inside the margin
        assert apart(a, A, b, A + (hi + 0.5) * u)           # half a metre clear (these stars are not re-entrant along u at this scale)


def test_certified_pairs_of_a_real_shape_field_are_zero_force_pairs_of_the_oracle():
    """the reference's own floe outlines (7..591 vertices, concave) tiled with overlaps: over the oracle's candidate pairs, a
    certified pair has no rows and overlap 0, and the rule answers most of the force-free pairs the bounding boxes cannot"""
    prm, Floe = scenarios.real_shape_field(7, seed=2)
    soa = sz.floes_to_soa(Floe)
    ref = oracle.OracleStep(prm, soa)
    pr = ref.pairs()
    ex = oracle_extended_centroids(ref, soa)
    off, rows = ref.rows()
    n_cert = n_free = n_box = 0
    for i1, j1, st, ov in zip(pr["i"], pr["j"], pr["status"], pr["overlap_state"]):
        i, j = int(i1) - 1, int(j1) - 1                          # the pair list holds 1-based positions in the extended list
        si, sj = ex["src"][i], ex["src"][j]
        a = np.stack(soa.outline(si), 1); b = np.stack(soa.outline(sj), 1)
        A, B = (ex["x"][i], ex["y"][i]), (ex["x"][j], ex["y"][j])
        wa, wb = a + A, b + B
        boxes_apart = wa[:, 0].max() < wb[:, 0].min() or wb[:, 0].max() < wa[:, 0].min() or wa[:, 1].max() < wb[:, 1].min() or wb[:, 1].max() < wa[:, 1].min()
        if boxes_apart:
            n_box += 1
            continue
        emp = ref_intersection_is_empty(a, A, b, B)
        n_free += emp
        if apart(a, A, b, B):
            n_cert += 1
            assert emp and st == 0 and ov == 0, (i, j)
            assert pair_row_count(ref, pr, i, j, off, rows) == 0, (i, j)
    assert n_free > 20 and n_cert >= 0.8 * n_free, (n_cert, n_free, n_box)


def oracle_extended_centroids(ref, soa):
    """centroid and source floe of every entry of the oracle's extended list (originals, then the periodic images)"""
    g = ref.ghosts()
    n0 = soa.n
    x, y = np.concatenate([soa.x, g["x"]]), np.concatenate([soa.y, g["y"]])
    src = np.arange(n0 + len(g["x"]))
    for k, p in enumerate(g["parent"]):                     # parent: 1-based position in the extended list
        src[n0 + k] = src[int(p) - 1]
    return {"x": x, "y": y, "src": src}


def pair_row_count(ref, pr, i, j, off, rows):
    """rows floe i (0-based position) holds about partner j: column 0 = the partner's 1-based position in the extended list"""
    r = rows[off[i]:off[i + 1]]
    return int((r[:, 0] == j + 1).sum()) if len(r) else 0
