"""Offline fuzz of the classifier's edge-by-edge certificate (sz_apart.cuh, host build) against the reference's own Clipper:
random concave stars, the reference's FloeShapes.mat outlines at random poses, thin gaps, near-touching vertices, nested
outlines.  certified => the reference intersection is empty.  usage: python tools/apart_fuzz.py [cases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_apart as T
import scenarios

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
polys, _, _ = scenarios.floe_shapes()
stats = {}
bad = 0


def shape(kind):
    if kind == "star":
        return T._star(rng, int(rng.integers(3, 80)), 10 ** rng.uniform(0.5, 4.0))
    v = np.asarray(polys[int(rng.integers(0, len(polys)))], np.float64)
    v = v - v.mean(0)
    th = rng.uniform(0, 2 * np.pi)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    return v @ R.T * 10 ** rng.uniform(-1.0, 0.5)


for case in range(cases):
    fam = ("star", "real", "mixed", "nested", "slide")[case % 5]
    a = shape("real" if fam in ("real", "mixed") else "star")
    b = shape("real" if fam == "real" else "star")
    ra, rb = np.abs(a).max(), np.abs(b).max()
    A = rng.uniform(-5e5, 5e5, 2)
    t = rng.uniform(0, 2 * np.pi); u = np.array([np.cos(t), np.sin(t)])
    if fam == "nested":
        b = b * (0.05 * ra / max(rb, 1e-9)); d = rng.uniform(0, 0.9) * ra * u
    elif fam == "slide":
        # slide b along u until first contact (bisection on the reference), then stand off by a gap around the margin
        lo, hi = 0.0, 2.5 * (ra + rb)
        for _ in range(50):
            mid = 0.5 * (lo + hi)
            if T.ref_intersection_is_empty(a, A, b, A + mid * u): hi = mid
            else: lo = mid
        d = (hi + 10 ** rng.uniform(-6, 0)) * u
    else:
        d = rng.uniform(0.2, 2.2) * (ra + rb) * 0.7 * u
    if case % 2: b = b[::-1]
    if case % 3 == 0: a = np.vstack([a, a[:1]])          # closed ring
    B = A + d
    ap = T.apart(a, A, b, B)
    emp = T.ref_intersection_is_empty(a, A, b, B)
    s = stats.setdefault(fam, [0, 0, 0, 0])
    s[0] += 1; s[1] += ap; s[2] += emp; s[3] += (ap and not emp)
    if ap and not emp:
        bad += 1
        print("MISMATCH case", case, fam)
print("cases=%d mismatches=%d (certified although the reference intersection is not empty)" % (cases, bad))
for k, s in stats.items():
    print("  %-7s cases %6d  certified %6d  reference-empty %6d  certified-but-not-empty %d" % (k, *s))
