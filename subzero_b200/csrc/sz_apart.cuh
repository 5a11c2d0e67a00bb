// sz_apart.cuh -- "these two outlines cannot intersect": a certificate for ANY pair of outlines (concave, any length), used by
// the pair classifier (sz_contact.cu) next to the bounding-box rule and the convex separating-axis rule.
//
// floe_interactions.m:25-30 clips the two world outlines; when the clip returns nothing, Ar = 0 (:43-44), no merge (:54-60) and
// `isempty(Xi)` sends the pair to the zero-force branch (:71-74): no rows, overlap 0.  For the packed convex field the classifier
// already answers such pairs (1.48M of 4.5M); for concave outlines it only had the bounding boxes, and a 400 x 400-vertex pair
// that merely shares a bounding-box corner paid a whole sweep.
//
// rings_apart(A, B) returns true only when
//   (1) every edge of A is farther than SZ_APART_M (1 mm) from every edge of B -- for each edge pair one of: their boxes are more
//       than the margin apart; both ends of one lie on the same side of the other's LINE, farther than the margin from it (then
//       every point of that edge is, the distance to a line being linear along a segment); and
//   (2) A's first vertex is not inside B and B's first vertex is not inside A (crossing number = the even-odd rule the clip
//       uses; safe because by (1) neither vertex is within the margin of the other boundary).
// (1) says the boundaries do not meet, so membership in B's region is constant along A's boundary and vice versa; with (2) no
// component of one region can lie in the other: the even-odd regions are disjoint -- also for a self-intersecting outline.
// Clipper works on coordinates rounded to 2^-32 m (polyclip.m:66) and rounds intersection points to the same grid
// (clipper.cpp:466-531): six orders of magnitude below the margin, as are the FP64 roundings of the products below (coordinates
// up to 1e6 m, edges up to 1e4 m: absolute error of a cross product < 1e-5 m^2, i.e. < 1e-5 m of distance for edges >= 1 m; edges
// shorter than 1 m are only ever answered by the box rule).  Anything not certified is swept as before: the rule can only turn a
// sweep that returns nothing into no sweep.  tests/test_apart.py checks the certificate against the reference's own Clipper.
#pragma once
#include "sz_clip.cuh"

namespace szapart {

#ifndef SZ_APART_M
#define SZ_APART_M 1e-3
#endif

// ring: n points (px[k] + X, py[k] + Y), closed implicitly (a repeated first point at the end gives a zero-length edge, skipped).
// (axmin .. aymax), (bxmin .. bymax): the rings' world bounding boxes; an edge farther than the margin from the other ring's box is apart from all of it.
SZ_HD bool rings_apart(const double* ax, const double* ay, int na, double AX, double AY, double axmin, double axmax, double aymin, double aymax,
                       const double* bx, const double* by, int nb, double BX, double BY, double bxmin, double bxmax, double bymin, double bymax)
{
    const double m = SZ_APART_M;
    if (na < 3 || nb < 3) return false;
    // only finite, moderate coordinates are certified (int64(NaN * 2^32) is 0 in MATLAB: such a vertex is somewhere else entirely)
    for (int i = 0; i < na; ++i) if (!(fabs(ax[i] + AX) < 1e12 && fabs(ay[i] + AY) < 1e12)) return false;
    for (int j = 0; j < nb; ++j) if (!(fabs(bx[j] + BX) < 1e12 && fabs(by[j] + BY) < 1e12)) return false;
    // ---- (1) edges
    for (int i = 0; i < na; ++i) {
        const int i1 = (i + 1 == na) ? 0 : i + 1;
        const double p0x = ax[i] + AX, p0y = ay[i] + AY, p1x = ax[i1] + AX, p1y = ay[i1] + AY;
        const double pxmin = p0x < p1x ? p0x : p1x, pxmax = p0x < p1x ? p1x : p0x, pymin = p0y < p1y ? p0y : p1y, pymax = p0y < p1y ? p1y : p0y;
        if (pxmin > bxmax + m || pxmax < bxmin - m || pymin > bymax + m || pymax < bymin - m) continue;      // apart from all of B
        const double dpx = p1x - p0x, dpy = p1y - p0y;
        const double lp = sqrt(dpx * dpx + dpy * dpy);
        if (lp == 0) continue;                                   // the closing duplicate: its point belongs to the neighbouring edges
        for (int j = 0; j < nb; ++j) {
            const int j1 = (j + 1 == nb) ? 0 : j + 1;
            const double q0x = bx[j] + BX, q0y = by[j] + BY, q1x = bx[j1] + BX, q1y = by[j1] + BY;
            const double qxmin = q0x < q1x ? q0x : q1x, qxmax = q0x < q1x ? q1x : q0x, qymin = q0y < q1y ? q0y : q1y, qymax = q0y < q1y ? q1y : q0y;
            if (qxmin > pxmax + m || qxmax < pxmin - m || qymin > pymax + m || qymax < pymin - m) continue;
            const double dqx = q1x - q0x, dqy = q1y - q0y;
            const double lq = sqrt(dqx * dqx + dqy * dqy);
            if (lq == 0) continue;
            bool apart = false;
            if (lp >= 1.0) {                                     // both ends of Q beyond P's line, same side
                const double c0 = dpx * (q0y - p0y) - dpy * (q0x - p0x), c1 = dpx * (q1y - p0y) - dpy * (q1x - p0x);
                const double lim = m * lp;
                apart = (c0 > lim && c1 > lim) || (c0 < -lim && c1 < -lim);
            }
            if (!apart && lq >= 1.0) {                           // both ends of P beyond Q's line, same side
                const double c0 = dqx * (p0y - q0y) - dqy * (p0x - q0x), c1 = dqx * (p1y - q0y) - dqy * (p1x - q0x);
                const double lim = m * lq;
                apart = (c0 > lim && c1 > lim) || (c0 < -lim && c1 < -lim);
            }
            if (!apart) return false;
        }
    }
    // ---- (2) containment, by the crossing number of a horizontal ray (half-open rule: a vertex on the ray counts for one edge only)
    {
        const double x = ax[0] + AX, y = ay[0] + AY;
        if (x >= bxmin && x <= bxmax && y >= bymin && y <= bymax) {
            bool in = false;
            for (int j = 0; j < nb; ++j) {
                const int j1 = (j + 1 == nb) ? 0 : j + 1;
                const double q0x = bx[j] + BX, q0y = by[j] + BY, q1x = bx[j1] + BX, q1y = by[j1] + BY;
                if ((q0y > y) != (q1y > y)) { if (x < q0x + (y - q0y) / (q1y - q0y) * (q1x - q0x)) in = !in; }
            }
            if (in) return false;
        }
    }
    {
        const double x = bx[0] + BX, y = by[0] + BY;
        if (x >= axmin && x <= axmax && y >= aymin && y <= aymax) {
            bool in = false;
            for (int i = 0; i < na; ++i) {
                const int i1 = (i + 1 == na) ? 0 : i + 1;
                const double p0x = ax[i] + AX, p0y = ay[i] + AY, p1x = ax[i1] + AX, p1y = ay[i1] + AY;
                if ((p0y > y) != (p1y > y)) { if (x < p0x + (y - p0y) / (p1y - p0y) * (p1x - p0x)) in = !in; }
            }
            if (in) return false;
        }
    }
    return true;
}

}  // namespace szapart
