// sz_corners.cuh -- the deterministic half of Physical_Processes/corners.m: the contact mask `da`
// that says which vertices of a floe may be ground off.  The random half (break1 = rand > angle/Anorm, :72, and
// frac_corner) stays with the host.
//
// corners.m:54-88 for one selected floe:
//   for every contact row [partner Fx Fy Px Py tau A] of the floe
//     partner == Inf            -> the floe touches the wall (:83)
//     partner <= N0             -> da(nearest vertex to (Px, Py)) = 1 (dsearchn, :74: the first of equals), and every vertex
//                                  in or on the partner's outline is flagged (inpolygon, :78-82); the partner is looked up in
//                                  the periodic list corners.m rebuilds from the CURRENT positions (:13-51)
//   wall contact                -> every vertex outside c2_boundary is flagged (:83-86)
// Vertices are those of polyshape(c_alpha'): c_alpha without its closing duplicate, in the stored order.
//
// Mapping: a group of G lanes (8 for Voronoi-sized outlines, 32 for real shapes) owns one selected floe; lane l owns the
// vertices l, l + G, ... (coalesced reads of the floe's outline, no write races on da), the partner's outline is read
// with group-uniform addresses (one broadcast load per edge), nearest-vertex and bounding-box reductions are shuffles
// inside the group.  The same code runs on the host with G = 1 for the CPU-side parity test (tests/host/pair_host.cpp).
#pragma once
#include "sz_pairforce.cuh"

namespace szcorn {

struct CornerArgs {
    int count; const int* idx;           // selected floes: 1-based positions in the floe list of the last contact step
    int nb_skip;                         // corners.m:54 skips the first Nb entries of the SELECTION
    const double* x; const double* y; const int* voff; const double* vx; const double* vy;
    const int* row_off; const double* rows;                                   // Floe(i).interactions: [K][7]
    const double* cex; const double* cey; const int* cesrc; const int* n_ext; // corners.m's own periodic list (:13-51)
    const double* boxx; const double* boxy; int nbox;                         // c2_boundary
    const int* da_off; unsigned char* da;                                     // da of selected floe q at [da_off[q], da_off[q+1])
};

// number of polyshape vertices of floe i (the closing duplicate of c_alpha dropped)
SZ_HD int open_count(const CornerArgs& a, int i)
{
    const int o = a.voff[i]; int nv = a.voff[i + 1] - o;
    if (nv > 1 && a.vx[o] == a.vx[o + nv - 1] && a.vy[o] == a.vy[o + nv - 1]) --nv;
    return nv;
}

// ---- reductions inside a lane group (identity on the host, where G = 1)
template <int G> SZ_HD double grp_min(double v, unsigned mask)
{
#if defined(__CUDA_ARCH__)
    for (int d = G / 2; d > 0; d >>= 1) { const double o = __shfl_xor_sync(mask, v, d, G); if (o < v) v = o; }
#endif
    return v;
}
template <int G> SZ_HD double grp_max(double v, unsigned mask)
{
#if defined(__CUDA_ARCH__)
    for (int d = G / 2; d > 0; d >>= 1) { const double o = __shfl_xor_sync(mask, v, d, G); if (o > v) v = o; }
#endif
    return v;
}
// smallest distance, the lowest index among equals (a sequential `if (d < best)` scan finds the same vertex)
template <int G> SZ_HD void grp_argmin(double& d, int& idx, unsigned mask)
{
#if defined(__CUDA_ARCH__)
    for (int s = G / 2; s > 0; s >>= 1) {
        const double od = __shfl_xor_sync(mask, d, s, G); const int oi = __shfl_xor_sync(mask, idx, s, G);
        if (od < d || (od == d && oi < idx)) { d = od; idx = oi; }
    }
#endif
}

// A polygon given as `m` stored vertices (px[s] + sx, py[s] + sy); polygon_operations/inpolygon.m closes it when its
// ends differ (close_loops, :226-236).
struct Ring {
    const double* px; const double* py; double sx, sy; int m;
    bool closed_by_us; int Nv;                     // Nv = vertices of the closed loop
    double xmin, xmax, ymin, ymax;
    SZ_HD double X(int s) const { return px[(s >= m) ? 0 : s] + sx; }
    SZ_HD double Y(int s) const { return py[(s >= m) ? 0 : s] + sy; }
};
// bounding box (:66-71) and loop closing, the group's lanes striding over the ring's vertices
template <int G> SZ_HD void ring_prepare(Ring& r, int lane, unsigned mask)
{
    double xmn = SZ_INF, xmx = -SZ_INF, ymn = SZ_INF, ymx = -SZ_INF;
    for (int s = lane; s < r.m; s += G) {
        const double x = r.px[s] + r.sx, y = r.py[s] + r.sy;
        if (x < xmn) xmn = x;
        if (x > xmx) xmx = x;
        if (y < ymn) ymn = y;
        if (y > ymx) ymx = y;
    }
    r.xmin = grp_min<G>(xmn, mask); r.xmax = grp_max<G>(xmx, mask); r.ymin = grp_min<G>(ymn, mask); r.ymax = grp_max<G>(ymx, mask);
    r.closed_by_us = r.m >= 3 && (r.X(0) != r.X(r.m - 1) || r.Y(0) != r.Y(r.m - 1));
    r.Nv = r.m + (r.closed_by_us ? 1 : 0);
}
// inpolygon.m:150-224 for one query point: quadrant-change winding number, on-edge tolerance 3*eps*max(|xm|,|ym|,|xm*ym|)
// of the edge's midpoint; boundary points count as inside.  The caller handles the empty ring / Nv < 2 cases.
SZ_HD bool in_ring(const Ring& r, double x, double y)
{
    if (!(x >= r.xmin && x <= r.xmax && y >= r.ymin && y <= r.ymax)) return false;      // mask (:71)
    double sumdq = 0; bool on = false;
    double ax = r.X(0), ay = r.Y(0);
    double vx0 = ax - x, vy0 = ay - y;
    bool px0 = vx0 > 0, py0 = vy0 > 0;
    double q0 = (double)((!px0 && py0) + 2 * (!px0 && !py0) + 3 * (px0 && !py0));
    for (int e = 0; e + 1 < r.Nv; ++e) {
        const double bx = r.X(e + 1), by = r.Y(e + 1);
        const double avx = fabs(0.5 * (ax + bx)), avy = fabs(0.5 * (ay + by));
        double sf = avx > avy ? avx : avy; const double pr = avx * avy; if (pr > sf) sf = pr;
        const double seps = sf * SZ_EPS * 3;
        const double vx1 = bx - x, vy1 = by - y;
        const bool px1 = vx1 > 0, py1 = vy1 > 0;
        const double q1 = (double)((!px1 && py1) + 2 * (!px1 && !py1) + 3 * (px1 && !py1));
        const double cross = vx0 * vy1 - vx1 * vy0;
        double sgn = (double)((cross > 0) - (cross < 0));
        if (fabs(cross) < seps) sgn = 0;
        const double dot = vx0 * vx1 + vy0 * vy1;
        double dq = q1 - q0;
        if (fabs(dq) == 3) dq = -dq / 3; else if (fabs(dq) == 2) dq = 2 * sgn;
        sumdq += dq;
        if (sgn == 0 && dot <= 0) on = true;
        ax = bx; ay = by; vx0 = vx1; vy0 = vy1; q0 = q1;
    }
    return (sumdq != 0) || on;
}

// The mask of selected floe q.  All G lanes of the group call it together (`mask` names them); lane = 0..G-1.
template <int G>
SZ_HD void corner_mask_floe(const CornerArgs& a, int q, int lane, unsigned mask)
{
    const int i = a.idx[q] - 1;
    const int o = a.voff[i], nv = open_count(a, i);
    unsigned char* da = a.da + a.da_off[q];
    for (int t = lane; t < nv; t += G) da[t] = 0;
    const int r0 = a.row_off[i], r1 = a.row_off[i + 1];
    if (q + 1 < 1 + a.nb_skip || r1 == r0 || nv <= 0) return;                            // :54-55
    const double Xi = a.x[i], Yi = a.y[i];
    const double N0 = (double)(*a.n_ext);                                                // :51
    bool bnd = false;
    for (int r = r0; r < r1; ++r) {
        const double* row = a.rows + (size_t)r * 7;
        const double partner = row[0];
        if (partner == SZ_INF || partner == -SZ_INF) { bnd = true; continue; }
        if (!(partner <= N0) || !(partner >= 1)) continue;
        // break2 = dsearchn(polytrue.Vertices, [Xi Yi]) (:74)
        const double cpx = row[3], cpy = row[4];
        double bd = SZ_INF; int best = 0x7fffffff;
        for (int t = lane; t < nv; t += G) {
            const double vxw = a.vx[o + t] + Xi, vyw = a.vy[o + t] + Yi;
            const double d = (vxw - cpx) * (vxw - cpx) + (vyw - cpy) * (vyw - cpy);
            if (d < bd) { bd = d; best = t; }
        }
        grp_argmin<G>(bd, best, mask);
        if (best == 0x7fffffff) best = 0;
        if (best % G == lane) da[best] = 1;
        // the floe's vertices in or on the partner's outline (:78-82)
        const int g = (int)partner - 1;
        const int src = a.cesrc[g];
        Ring ring; ring.px = a.vx + a.voff[src]; ring.py = a.vy + a.voff[src]; ring.sx = a.cex[g]; ring.sy = a.cey[g]; ring.m = a.voff[src + 1] - a.voff[src];
        if (ring.m <= 0) continue;
        ring_prepare<G>(ring, lane, mask);
        if (ring.Nv < 2) continue;
        for (int t = lane; t < nv; t += G)
            if (in_ring(ring, a.vx[o + t] + Xi, a.vy[o + t] + Yi)) da[t] = 1;
    }
    if (bnd) {                                                                           // :83-86
        Ring ring; ring.px = a.boxx; ring.py = a.boxy; ring.sx = 0; ring.sy = 0; ring.m = a.nbox;
        bool usable = ring.m > 0;
        if (usable) { ring_prepare<G>(ring, lane, mask); usable = ring.Nv >= 2; }
        for (int t = lane; t < nv; t += G)
            if (!usable || !in_ring(ring, a.vx[o + t] + Xi, a.vy[o + t] + Yi)) da[t] = 1;
    }
}

}  // namespace szcorn
