// sz_pairforce.cuh -- the pair force law of collisions/floe_interactions.m for ONE (floe, partner)
// or (floe, wall) pair, written for one CUDA thread per pair on top of the Clipper-exact sweep in
// sz_clip.cuh.  Everything the reference does through MATLAB temporaries (cell arrays of regions,
// dense InterX matrices, per-edge vectors of the general branch) is streamed through a
// fixed-capacity workspace instead; floating-point operations are kept in the reference's order and
// individually rounded (this translation unit is compiled with -fmad=false / -ffp-contract=off).
//
// Reference map (file:line of /root/reference):
//   spring constants                floe_interactions.m:10-21      -> pair_force() head
//   clip #1 (+ wall 0.75 test)      :25-41, polyclip.m:63-73       -> run_clip(), Workspace::ra*
//   region areas, merge test        :43-60                          -> ring_area_centroid()
//   InterX                          InterX.m:54-77                  -> interx()
//   per-region contact + direction  :92-150                         -> region loop
//   sign test (clips #2, #3..)      :151-165
//   normal + tangential force       :167-187
//   p_poly_dist / inpolygon         polygon_operations/p_poly_dist.m:111-288, inpolygon.m:150-224
#pragma once
#include "sz_clip.cuh"
#include "sz_convex.cuh"

namespace szpf {

using szclip::i64;
using szclip::P64;

#define SZ_INF (__builtin_huge_val())
#define SZ_EPS 2.220446049250313e-16
#define SZ_SCALE 4294967296.0

struct Params {            // device copy of SzParams (include/subzero_b200.h)
    double Lx, Ly, modulus, dt, nu, mu, merge_frac, wall_frac, amin_per_vertex, vertex_match_tol,
           on_edge_tol, dl_min, close_gap, big_floe_r, domain_area_frac;
    int Nb, periodic, collision;
    // c2_boundary extent + area for the merge-test guard (floe_interactions.m:54); has_box = 0 when absent
    double bxmin, bxmax, bymin, bymax, barea; int has_box;
};

struct Body { double h, area, Xi, Yi, Ui, Vi, ksi; };   // the fields of floe1 / floe2 the law reads

enum PairStatus { PS_OK = 0, PS_CLIPPER_FAIL = -3 /* SZ_ERR_CLIPPER */, PS_CAPACITY = -4 /* SZ_ERR_CAPACITY */, PS_BAD_POLY = -1 /* SZ_ERR_ARG */,
                  PS_BAIL = -7 /* internal: the convex fast path declined the pair, re-run it with the general sweep */ };

// Per-pair hints computed once per floe (ext_prep_kernel / the host test shim).  convex: both outlines are strictly
// convex in Clipper's coordinates.  For convex outlines the sweep is fed the OPEN ring (no1/no2 vertices, closing
// duplicate dropped -- AddPath strips it anyway, clipper.cpp:1056) rotated to start at its bottom vertex (rot1/rot2):
// Clipper's result does not depend on where a closed path starts (one local minimum per convex path; checked on 110k
// convex pairs incl. shared edges and rectangles against the reference itself), and lanes that sweep pairs with the
// same vertex counts then touch the same arena words at the same time.
struct PairHints { bool convex; int rot1, no1, rot2, no2; };

struct PairResult {
    int status;            // PairStatus
    int n_rows;            // contact rows produced (0 when the pair exerts no force)
    double overlap_state;  // 0, +Inf, -Inf  (floe_interactions.m:38,56,58)
};

// polyclip.m:66  int64(x*scale): round half away from zero, saturating, NaN -> 0
SZ_HD i64 matlab_int64_general(double v)
{
    if (v != v) return 0;
    if (v >= 9223372036854775807.0) return 0x7FFFFFFFFFFFFFFFLL;
    if (v <= -9223372036854775808.0) return (i64)0x8000000000000000ULL;
    double t = trunc(v), f = v - t;
    i64 r = (i64)t;
    if (f >= 0.5) r += 1; else if (f <= -0.5) r -= 1;
    return r;
}
// The same value for |v| < 2^51 (every coordinate of a real field: 2^51 units are 5e5 km) from one round-to-nearest-even
// conversion: only an exact tie can differ from "half away from zero", and there v - r is exactly +-0.5.
SZ_HD i64 matlab_int64(double v)
{
    if (!(fabs(v) < 2251799813685248.0)) return matlab_int64_general(v);
#if defined(__CUDA_ARCH__)
    i64 r = __double2ll_rn(v);
#else
    i64 r = (i64)nearbyint(v);          // default rounding mode: to nearest even
#endif
    const double d = v - (double)r;
    if (d == 0.5 && v > 0) r += 1; else if (d == -0.5 && v < 0) r -= 1;
    return r;
}

template <class ClipC, int NV_, int RV_, int RP_, int NP_, int ROWS_>
struct PairCaps { typedef ClipC Clip; enum { NV = NV_, RV = RV_, RP = RP_, NP = NP_, ROWS = ROWS_ }; };

template <class C>
struct Workspace {
    szclip::ClipEngine<typename C::Clip> eng;
    double c1x[C::NV], c1y[C::NV], c2x[C::NV], c2y[C::NV];   // world outlines (closed); caller fills, n1/n2 points
    int n1, n2;
    i64 rax[C::RV], ray[C::RV]; int ra_off[C::RP + 1]; int ra_n;   // clip #1 regions, Clipper coordinates
    double ar[C::RP];
    i64 rbx[C::RV], rby[C::RV]; int rb_off[C::RP + 1]; int rb_n;   // clip #2 regions
    double px[C::NP], py[C::NP]; int np;                          // InterX points
};

// The same buffers without the general sweep's arena: workspace of the convex fast path (sz_convex.cuh), whose
// sweep state is a few hundred bytes of its own.
template <class C>
struct WorkspaceLite {
    double c1x[C::NV], c1y[C::NV], c2x[C::NV], c2y[C::NV];
    int n1, n2;
    i64 rax[C::RV], ray[C::RV]; int ra_off[C::RP + 1]; int ra_n;
    double ar[C::RP];
    i64 rbx[C::RV], rby[C::RV]; int rb_off[C::RP + 1]; int rb_n;
    // InterX points; the same bytes hold the int64 outlines of the convex sweep, which is over before InterX runs
    union { double px[C::NP]; i64 svx[C::NP]; };
    union { double py[C::NP]; i64 svy[C::NP]; };
    int np;
};

// ------------------------------------------------------------------------------------------------
template <class C>
struct RegionSink {       // collects emitted paths into (x,y,off) pools; flags overflow
    i64* x; i64* y; int* off; int n_paths, n_pts; bool overflow;
    SZ_HD RegionSink(i64* x_, i64* y_, int* off_) : x(x_), y(y_), off(off_), n_paths(0), n_pts(0), overflow(false) { off[0] = 0; }
    SZ_HD void begin_path(int cnt) { if (n_paths >= C::RP || n_pts + cnt > C::RV) overflow = true; else { ++n_paths; off[n_paths] = off[n_paths - 1]; } }
    SZ_HD void point(P64 p) { if (overflow) return; x[n_pts] = p.x; y[n_pts] = p.y; ++n_pts; off[n_paths] = n_pts; }
};
struct CountSink { int n; SZ_HD void begin_path(int) { ++n; } SZ_HD void point(P64) {} };
// ------------------------------------------------------------------------------------------------
// Warp-synchronous execution.  On the GPU the 32 lanes of a warp resolve 32 independent pairs.  Left to
// itself the hardware lets the lanes drift apart in the branchy sweep (the first profile showed 4 of 32
// lanes active on average), so the force law below is written as a per-lane phase machine around ONE
// clip site, and the lanes re-converge explicitly before every clip and between the phases of every
// scanbeam.  All 32 lanes of a warp must call pair_force() together (lanes without a pair pass
// valid = false).  On the host the macros collapse and the same code handles one pair.
#if !defined(SZ_WARP_SYNC_ONLY)
#define SZ_BLOCK_SYNC 1
#endif
#if defined(__CUDA_ARCH__) && defined(SZ_BLOCK_SYNC)
// block-synchronous variant: every warp of the CTA walks the phases together, so the instruction lines of a
// phase are fetched once per CTA instead of once per warp (the sweep's SASS is far larger than the I-cache)
#define SZ_WARP_ANY(p) __syncthreads_or((p))
#define SZ_WARP_SYNC() __syncthreads()
#define SZ_LANE_SYNC() __syncwarp()
#elif defined(__CUDA_ARCH__)
#define SZ_WARP_ANY(p) __any_sync(0xffffffffu, (p))
#define SZ_WARP_SYNC() __syncwarp()
#define SZ_LANE_SYNC() __syncwarp()
#else
#define SZ_WARP_ANY(p) (p)
#define SZ_WARP_SYNC() ((void)0)
#define SZ_LANE_SYNC() ((void)0)
#endif

// the subject / clip path of the next clip: either a world outline shifted by (dx,dy) and packed like
// polyclip.m:66, or a Clipper result ring fed back in (int64(double(X)/2^32*2^32))
struct ClipInput {
    const double* x; const double* y; double dx, dy;
    const i64* ix; const i64* iy;
    int n; int ring;
    int rot;     // convex outlines are fed to the sweep starting at their bottom vertex (see PairHints); 0 = as stored
    SZ_HD P64 operator()(int i) const
    {
        P64 p;
        if (rot) { i += rot; if (i >= n) i -= n; }
        if (ring) { p.x = matlab_int64(((double)ix[i] / SZ_SCALE) * SZ_SCALE); p.y = matlab_int64(((double)iy[i] / SZ_SCALE) * SZ_SCALE); }
        else { p.x = matlab_int64((x[i] + dx) * SZ_SCALE); p.y = matlab_int64((y[i] + dy) * SZ_SCALE); }
        return p;
    }
};

// One polyclip() call per participating lane (want = false: the lane only keeps the warp company).
// Returns PS_OK / PS_CAPACITY / PS_CLIPPER_FAIL; the solution stays in eng (read it with emit()).
template <class E>
SZ_HDN int run_sweep(E& eng, bool want, int method, const ClipInput& subj, const ClipInput& clip)
{
    bool run = false;
    if (want) {
        eng.begin(method);
        eng.add_path(subj, subj.n, 0);
        eng.add_path(clip, clip.n, 1);
        run = eng.sweep_begin();
    }
    // SZ_SWEEP_SYNCS: CTA barriers per scanbeam besides the loop vote (0..3)
#ifndef SZ_SWEEP_SYNCS
#define SZ_SWEEP_SYNCS 0
#endif
    for (;;) {
        if (SZ_SWEEP_SYNCS >= 3) SZ_WARP_SYNC(); else SZ_LANE_SYNC();
        if (run) run = eng.sweep_next();
        if (!SZ_WARP_ANY(run)) break;
        if (run) run = eng.sweep_intersections();
        if (SZ_SWEEP_SYNCS >= 1) SZ_WARP_SYNC(); else SZ_LANE_SYNC();
        if (run) run = eng.sweep_top();
        if (SZ_SWEEP_SYNCS >= 2) SZ_WARP_SYNC(); else SZ_LANE_SYNC();
        if (run) run = eng.sweep_minima();
    }
    if (!want) return PS_OK;
    eng.sweep_finish();
    if (eng.status == szclip::ST_OVERFLOW) return PS_CAPACITY;
    if (eng.status != szclip::ST_OK) return PS_CLIPPER_FAIL;
    return PS_OK;
}

// area(polyshape) / centroid(polyshape): shoelace relative to vertex 0 (pinned by FloeShapes.mat)
SZ_HD void ring_area_centroid(const i64* X, const i64* Y, int n, double& area, double& cx, double& cy)
{
    if (n < 3) { area = 0; cx = cy = SZ_INF - SZ_INF; return; }
    const double x0 = (double)X[0] / SZ_SCALE, y0 = (double)Y[0] / SZ_SCALE;
    double a2 = 0, sx = 0, sy = 0;
    double xi = 0, yi = 0;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1 == n) ? 0 : i + 1;
        double xj = (double)X[j] / SZ_SCALE - x0, yj = (double)Y[j] / SZ_SCALE - y0;
        double c = xi * yj - xj * yi;
        a2 += c; sx += (xi + xj) * c; sy += (yi + yj) * c;
        xi = xj; yi = yj;
    }
    area = fabs(a2) / 2;
    cx = x0 + sx / (3 * a2);
    cy = y0 + sy / (3 * a2);
}
// polyarea(): abs(sum((x([2:end 1])-x).*(y([2:end 1])+y))/2), untranslated
SZ_HD double ring_polyarea(const i64* X, const i64* Y, int n)
{
    double s = 0;
    for (int i = 0; i < n; ++i) {
        int j = (i + 1 == n) ? 0 : i + 1;
        double xi = (double)X[i] / SZ_SCALE, yi = (double)Y[i] / SZ_SCALE, xj = (double)X[j] / SZ_SCALE, yj = (double)Y[j] / SZ_SCALE;
        s += (xj - xi) * (yj + yi);
    }
    return fabs(s / 2);
}

// Certificate used to answer clip #3 ("does the new region meet region k?", floe_interactions.m:158-159) without a
// sweep: the point (cX, cY) (Clipper units) lies on the inner side of EVERY edge line of the ring by at least
// SZ_CERT_MARGIN units.  A point in the kernel of a simple polygon with that margin is the centre of a disc of that
// radius inside the polygon; if one point certifies for both rings their intersection contains a disc of radius 2^20
// units (0.24 mm), which Clipper -- whose vertices are rounded by at most a few units -- cannot return as empty.
// FP64 suffices: |cross| = len * distance, its rounding error is below len * 2^-7 units for rings smaller than 2^43.
#define SZ_CERT_MARGIN 1048576.0
SZ_HD bool certify_inside(const i64* X, const i64* Y, int n, double cX, double cY)
{
    if (n < 3) return false;
    double sign = 0;
    double ax = (double)X[n - 1], ay = (double)Y[n - 1];
    for (int i = 0; i < n; ++i) {
        const double bx = (double)X[i], by = (double)Y[i];
        const double ex = bx - ax, ey = by - ay;
        const double cr = ex * (cY - ay) - ey * (cX - ax);
        const double len2 = ex * ex + ey * ey;
        if (!(cr * cr >= (SZ_CERT_MARGIN * SZ_CERT_MARGIN) * len2) || len2 == 0) return false;
        const double sg = cr > 0 ? 1.0 : -1.0;
        if (sign == 0) sign = sg; else if (sg != sign) return false;
        ax = bx; ay = by;
    }
    return true;
}

// Exact convexity of an outline in Clipper's coordinates (polyclip.m:66): every turn has the same strict sign.
// `get(i)` returns vertex i of a closed ring given with n distinct-position slots (first vertex NOT repeated).
template <class Getter>
SZ_HD bool ring_is_strictly_convex(const Getter& get, int n)
{
    if (n < 3) return false;
    int sign = 0;
    P64 a = get(n - 2), b = get(n - 1);
    for (int i = 0; i < n; ++i) {
        const P64 c = get(i);
        const i64 ux = b.x - a.x, uy = b.y - a.y, vx = c.x - b.x, vy = c.y - b.y;
        // sign of ux*vy - uy*vx, exactly
#if defined(__CUDA_ARCH__)
        const i64 h1 = __mul64hi(ux, vy), h2 = __mul64hi(uy, vx);
        const unsigned long long l1 = (unsigned long long)ux * (unsigned long long)vy, l2 = (unsigned long long)uy * (unsigned long long)vx;
        const int sg = (h1 != h2) ? (h1 > h2 ? 1 : -1) : (l1 != l2 ? (l1 > l2 ? 1 : -1) : 0);
#else
        const __int128 cr = (__int128)ux * vy - (__int128)uy * vx;
        const int sg = cr > 0 ? 1 : (cr < 0 ? -1 : 0);
#endif
        if (sg == 0) return false;
        if (sign == 0) sign = sg; else if (sg != sign) return false;
        a = b; b = c;
    }
    return true;
}

// index of the bottom vertex of an open ring in Clipper's sense (largest Y, then smallest X)
template <class Getter>
SZ_HD int ring_bottom_vertex(const Getter& get, int n)
{
    int best = 0; P64 b = get(0);
    for (int i = 1; i < n; ++i) { const P64 p = get(i); if (p.y > b.y || (p.y == b.y && p.x < b.x)) { b = p; best = i; } }
    return best;
}

// FP64 Sutherland-Hodgman clip of the convex polygon S by the convex polygon K (open rings).  Used ONLY to decide
// the sign test of floe_interactions.m:151-165 when the decision has a wide margin (see convex_sign_test); never to
// produce a polygon that is output.  Returns the vertex count with (*rx, *ry) pointing at the result (one of the two
// buffers the caller passed), or -1 when a buffer would overflow.
//
// One pass per edge line of K.  A line cuts a convex ring in at most two places, so for rings of up to 32 vertices the
// pass first takes the side of every vertex into a bit mask (a uniform loop: the lanes of a warp clip different pairs),
// and then: all inside -> the ring is left where it is (most edges of K do not reach the overlap strip at all);
// all outside -> empty; one inside run -> [entry point, the run, exit point] goes to the other buffer, the two
// intersection points computed by all lanes at once instead of wherever each lane's loop happens to meet them (the
// one-lane-at-a-time divisions were 4 % of class C's instructions).  Anything else (a ring that rounding left
// non-convex, more than 32 vertices) takes the vertex-by-vertex loop.  The ring may come out rotated against that loop's;
// its area is compared with a margin six orders of magnitude above the rounding either way.
SZ_HD int sz_popc32(unsigned v)
{
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
SZ_HD int sz_ctz32(unsigned v)     // v != 0
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
SZ_HD int sh_clip_convex(const double* sx, const double* sy, int ns, double shx, double shy, const double* kx, const double* ky, int nk,
                         double* ax, double* ay, double* bx, double* by, int cap, const double** rx, const double** ry)
{
    if (ns > cap) return -1;
    for (int i = 0; i < ns; ++i) { ax[i] = sx[i] + shx; ay[i] = sy[i] + shy; }
    int n = ns;
    double area2 = 0;
    for (int i = 0; i < nk; ++i) { const int j = (i + 1 == nk) ? 0 : i + 1; area2 += (kx[i] - kx[0]) * (ky[j] - ky[0]) - (kx[j] - kx[0]) * (ky[i] - ky[0]); }
    const double sK = area2 >= 0 ? 1.0 : -1.0;
    for (int e = 0; e < nk && n > 0; ++e) {
        const int e1 = (e + 1 == nk) ? 0 : e + 1;
        const double px = kx[e], py = ky[e], dx = kx[e1] - px, dy = ky[e1] - py;
#if !defined(SZ_SH_PLAIN)
        if (n <= 32) {
            unsigned in = 0;
            for (int i = 0; i < n; ++i) { if (sK * (dx * (ay[i] - py) - dy * (ax[i] - px)) >= 0) in |= 1u << i; }
            const unsigned full = (n == 32) ? 0xffffffffu : ((1u << n) - 1u);
            if (in == full) continue;
            if (in == 0) { n = 0; break; }
            const unsigned prev = ((in << 1) | (in >> (n - 1))) & full;      // bit i: the side of vertex i - 1
            const unsigned tr = in ^ prev;
            if (sz_popc32(tr) == 2) {
                const int s = sz_ctz32(tr & in), t = sz_ctz32(tr & ~in);     // first vertex of the inside run, first vertex behind it
                int run = t - s; if (run < 0) run += n;
                if (run + 2 > cap) return -1;
                {
                    const int q = (s == 0) ? n - 1 : s - 1;
                    const double qx = ax[q], qy = ay[q], rx_ = ax[s], ry_ = ay[s];
                    const double dq = sK * (dx * (qy - py) - dy * (qx - px)), dr = sK * (dx * (ry_ - py) - dy * (rx_ - px));
                    const double u = dq / (dq - dr);
                    bx[0] = qx + u * (rx_ - qx); by[0] = qy + u * (ry_ - qy);
                }
                for (int k = 0, i = s; k < run; ++k) { bx[1 + k] = ax[i]; by[1 + k] = ay[i]; if (++i == n) i = 0; }
                {
                    const int q = (t == 0) ? n - 1 : t - 1;
                    const double qx = ax[q], qy = ay[q], rx_ = ax[t], ry_ = ay[t];
                    const double dq = sK * (dx * (qy - py) - dy * (qx - px)), dr = sK * (dx * (ry_ - py) - dy * (rx_ - px));
                    const double u = dq / (dq - dr);
                    bx[run + 1] = qx + u * (rx_ - qx); by[run + 1] = qy + u * (ry_ - qy);
                }
                double* t1 = ax; ax = bx; bx = t1; t1 = ay; ay = by; by = t1;
                n = run + 2;
                continue;
            }
        }
#endif
        int m = 0;
        double qx = ax[n - 1], qy = ay[n - 1];
        double dq = sK * (dx * (qy - py) - dy * (qx - px));
        for (int i = 0; i < n; ++i) {
            const double rx_ = ax[i], ry_ = ay[i];
            const double dr = sK * (dx * (ry_ - py) - dy * (rx_ - px));
            if ((dq >= 0) != (dr >= 0)) {
                if (m >= cap) return -1;
                const double u = dq / (dq - dr);
                bx[m] = qx + u * (rx_ - qx); by[m] = qy + u * (ry_ - qy); ++m;
            }
            if (dr >= 0) { if (m >= cap) return -1; bx[m] = rx_; by[m] = ry_; ++m; }
            qx = rx_; qy = ry_; dq = dr;
        }
        double* t1 = ax; ax = bx; bx = t1; t1 = ay; ay = by; by = t1;
        n = m;
    }
    *rx = ax; *ry = ay;
    return n;
}

// Decides the sign test of floe_interactions.m:151-165 for a convex pair without clips #2/#3 when it can be decided
// with margin.  The reference re-clips floe 1 nudged by force_dir (1 m) and flips force_dir once for every new
// region that meets region k and is larger than it (:158-163).  For convex outlines the new intersection is one convex
// region Q; this computes its area in FP64 (error ~1e-8 m^2) and answers
//   -1  no flip:  area(Q) < Ak - tol  (no region of the new clip can exceed Ak, whatever its exact shape)
//   +1  flip:     area(Q) > Ak + tol, and one disc of radius >= 1 mm lies inside both Q and region k
//    0  undecided -> the caller runs the reference's clips.
// tol = 1e-6 Ak + 1 m^2 is six orders of magnitude above either computation's rounding.
template <class C, class W>
SZ_HD int convex_sign_test(W& w, double fdx, double fdy, const i64* RX, const i64* RY, int nr, double Ak, double pcx, double pcy)
{
    // open rings: drop the closing vertex (and a second closing vertex added by :62-67)
    int n1 = w.n1, n2 = w.n2;
    while (n1 > 1 && w.c1x[n1 - 1] == w.c1x[0] && w.c1y[n1 - 1] == w.c1y[0]) --n1;
    while (n2 > 1 && w.c2x[n2 - 1] == w.c2x[0] && w.c2y[n2 - 1] == w.c2y[0]) --n2;
    const double tol = 1e-6 * Ak + 1.0;
    const int cap = C::RV / 2;
    double* ax = reinterpret_cast<double*>(w.rbx); double* ay = ax + cap;
    double* bx = reinterpret_cast<double*>(w.rby); double* by = bx + cap;
    const double* qx = ax; const double* qy = ay;
    const int n = sh_clip_convex(w.c1x, w.c1y, n1, fdx, fdy, w.c2x, w.c2y, n2, ax, ay, bx, by, cap, &qx, &qy);
    if (n < 0) return 0;
    if (n < 3) return (0.0 < Ak - tol) ? -1 : 0;
    double a2 = 0;
    for (int i = 0; i < n; ++i) { const int j = (i + 1 == n) ? 0 : i + 1; a2 += (qx[i] - qx[0]) * (qy[j] - qy[0]) - (qx[j] - qx[0]) * (qy[i] - qy[0]); }
    const double aQ = fabs(a2) / 2;
    if (aQ < Ak - tol) return -1;
    if (!(aQ > Ak + tol)) return 0;
    // flip needs "the new region meets region k": a disc around region k's centroid inside both
    if (!certify_inside(RX, RY, nr, pcx * SZ_SCALE, pcy * SZ_SCALE)) return 0;
    const double sg = a2 > 0 ? 1.0 : -1.0;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1 == n) ? 0 : i + 1;
        const double ex = qx[j] - qx[i], ey = qy[j] - qy[i];
        const double cr = sg * (ex * (pcy - qy[i]) - ey * (pcx - qx[i]));
        if (!(cr >= 1e-3 * sqrt(ex * ex + ey * ey))) return 0;
    }
    return 1;
}

// unique(...,'rows') of InterX.m:77 on the collected points: insertion sort by (x, then y), duplicates dropped
template <class W>
SZ_HD void interx_sort_unique(W& w, int np)
{
    for (int i = 1; i < np; ++i) {
        double vx = w.px[i], vy = w.py[i]; int k = i - 1;
        while (k >= 0 && (w.px[k] > vx || (w.px[k] == vx && w.py[k] > vy))) { w.px[k + 1] = w.px[k]; w.py[k + 1] = w.py[k]; --k; }
        w.px[k + 1] = vx; w.py[k + 1] = vy;
    }
    int m = 0;
    for (int i = 0; i < np; ++i) if (m == 0 || !(w.px[i] == w.px[m - 1] && w.py[i] == w.py[m - 1])) { w.px[m] = w.px[i]; w.py[m] = w.py[i]; ++m; }
    w.np = m;
}

// InterX.m:54-77 for outlines of at most 32 segments each (classes C and S), arranged so that the lanes of a warp -- which
// resolve different pairs -- stay together.  The reference evaluates, for every (segment i of curve 1, segment j of curve 2),
//   test a  (a0 - S1)(a1 - S1) <= 0   the ends of segment j lie on different sides of the line of segment i,
//   test b  (b0 - S2)(b1 - S2) <= 0   the same the other way round,
// and computes a point where both hold and the segments are not parallel.  Written as one doubly nested loop with early
// `continue`s (interx_general below) the few lanes that pass test a run test b, and the one or two that pass both run the two
// divisions, while the rest of the warp waits: 2.1 of 32 lanes in the profile of class C.  Here the two tests are two uniform
// passes -- test a into one bit mask per segment of curve 1, then test b, recording the (j, i) that pass both in the
// reference's order (j outer, i inner) -- and the points are computed in a third loop in which almost every lane has the
// same two hits.  Every expression is the reference's own (a1 of segment j is a0 of segment j + 1: the same operands, the
// same rounding), so the points are bit-identical.
#ifndef SZ_INTERX_HITS
#define SZ_INTERX_HITS 12
#endif
template <class C, class W>
SZ_HD bool interx_general(W& w);
template <class C, class W>
SZ_HD bool interx_masked(W& w)
{
    const int n1 = w.n1 - 1, n2 = w.n2 - 1;
    if (n1 > 32 || n2 > 32 || n1 > C::NV || n2 > C::NV) return interx_general<C>(w);
    // scratch in the clip #2 region buffers, idle between clip #1 and the sign test -- and in their MIDDLE, where the convex
    // sweep's output deque has just been (class C's time follows the distinct local-memory lines a thread touches)
    static_assert(C::RV >= 40 && (C::RV / 2 - 4) * sizeof(i64) + 32 * sizeof(int) <= C::RV * sizeof(i64), "rbx holds the masks");
    static_assert((C::RV / 2 - 4) * sizeof(i64) + SZ_INTERX_HITS * sizeof(int) <= C::RV * sizeof(i64), "rby holds the hit list");
    unsigned* passA = reinterpret_cast<unsigned*>(w.rbx + (C::RV / 2 - 4));
    for (int i = 0; i < n1; ++i) {
        const double x1a = w.c1x[i], y1a = w.c1y[i], x1b = w.c1x[i + 1], y1b = w.c1y[i + 1];
        const double dx1 = x1b - x1a, dy1 = y1b - y1a;
        const double S1 = dx1 * y1a - dy1 * x1a;
        double prev = (dx1 * w.c2y[0] - dy1 * w.c2x[0]) - S1;
        unsigned m = 0;
        for (int j = 0; j < n2; ++j) {
            const double nxt = (dx1 * w.c2y[j + 1] - dy1 * w.c2x[j + 1]) - S1;
            if ((prev * nxt) <= 0) m |= 1u << j;
            prev = nxt;
        }
        passA[i] = m;
    }
    int* hits = reinterpret_cast<int*>(w.rby + (C::RV / 2 - 4)); int nh = 0;
    for (int j = 0; j < n2; ++j) {
        const double x2a = w.c2x[j], y2a = w.c2y[j], x2b = w.c2x[j + 1], y2b = w.c2y[j + 1];
        const double dx2 = x2b - x2a, dy2 = y2b - y2a;
        const double S2 = dx2 * y2a - dy2 * x2a;
        double prev = (w.c1y[0] * dx2 - w.c1x[0] * dy2) - S2;
        for (int i = 0; i < n1; ++i) {
            const double nxt = (w.c1y[i + 1] * dx2 - w.c1x[i + 1] * dy2) - S2;
            if (((prev * nxt) <= 0) && ((passA[i] >> j) & 1u)) { if (nh < SZ_INTERX_HITS) hits[nh] = (j << 8) | i; ++nh; }
            prev = nxt;
        }
    }
    if (nh > SZ_INTERX_HITS) return interx_general<C>(w);       // many touching segments: the plain loop handles any number
    int np = 0;
    for (int t = 0; t < nh; ++t) {
        const int j = hits[t] >> 8, i = hits[t] & 255;
        const double x1a = w.c1x[i], y1a = w.c1y[i], x2a = w.c2x[j], y2a = w.c2y[j];
        const double dx1 = w.c1x[i + 1] - x1a, dy1 = w.c1y[i + 1] - y1a, dx2 = w.c2x[j + 1] - x2a, dy2 = w.c2y[j + 1] - y2a;
        const double S1 = dx1 * y1a - dy1 * x1a, S2 = dx2 * y2a - dy2 * x2a;
        const double L = dy2 * dx1 - dy1 * dx2;
        if (L == 0) continue;
        if (np >= C::NP) return false;
        w.px[np] = (dx2 * S1 - dx1 * S2) / L;
        w.py[np] = (dy2 * S1 - dy1 * S2) / L;
        ++np;
    }
    interx_sort_unique(w, np);
    return true;
}

// InterX.m:54-77 (two-curve form).  Points are collected, sorted (x, then y) and de-duplicated.
template <class C, class W>
SZ_HD bool interx(W& w)
{
#if !defined(SZ_INTERX_PLAIN)
    if constexpr (C::NV <= 33) return interx_masked<C>(w);
#endif
    return interx_general<C>(w);
}
template <class C, class W>
SZ_HD bool interx_general(W& w)
{
    const int n1 = w.n1 - 1, n2 = w.n2 - 1;
    int np = 0;
    for (int j = 0; j < n2; ++j) {
        const double x2a = w.c2x[j], y2a = w.c2y[j], x2b = w.c2x[j + 1], y2b = w.c2y[j + 1];
        const double dx2 = x2b - x2a, dy2 = y2b - y2a;
        const double S2 = dx2 * y2a - dy2 * x2a;
        for (int i = 0; i < n1; ++i) {
            const double x1a = w.c1x[i], y1a = w.c1y[i], x1b = w.c1x[i + 1], y1b = w.c1y[i + 1];
            const double dx1 = x1b - x1a, dy1 = y1b - y1a;
            const double S1 = dx1 * y1a - dy1 * x1a;
            const double a0 = dx1 * y2a - dy1 * x2a, a1 = dx1 * y2b - dy1 * x2b;
            if (!(((a0 - S1) * (a1 - S1)) <= 0)) continue;
            const double b0 = y1a * dx2 - x1a * dy2, b1 = y1b * dx2 - x1b * dy2;
            if (!(((b0 - S2) * (b1 - S2)) <= 0)) continue;
            const double L = dy2 * dx1 - dy1 * dx2;
            if (L == 0) continue;
            if (np >= C::NP) return false;
            w.px[np] = (dx2 * S1 - dx1 * S2) / L;
            w.py[np] = (dy2 * S1 - dy1 * S2) / L;
            ++np;
        }
    }
    interx_sort_unique(w, np);
    return true;
}

// inpolygon.m:150-224 for a single query point against the closed ring [X;X(1)] of a region
SZ_HD bool in_region(double x, double y, const i64* X, const i64* Y, int n, double xmin, double xmax, double ymin, double ymax)
{
    if (!(x >= xmin && x <= xmax && y >= ymin && y <= ymax)) return false;
    double sumdq = 0; bool on = false;
    double ax = (double)X[0] / SZ_SCALE, ay = (double)Y[0] / SZ_SCALE;
    double vx0 = ax - x, vy0 = ay - y;
    bool px0 = vx0 > 0, py0 = vy0 > 0;
    double q0 = (double)((!px0 && py0) + 2 * (!px0 && !py0) + 3 * (px0 && !py0));
    for (int m = 0; m < n; ++m) {
        int j = (m + 1 == n) ? 0 : m + 1;
        double bx = (double)X[j] / SZ_SCALE, by = (double)Y[j] / SZ_SCALE;
        double avx = fabs(0.5 * (ax + bx)), avy = fabs(0.5 * (ay + by));
        double sf = avx > avy ? avx : avy; double pr = avx * avy; if (pr > sf) sf = pr;
        double seps = sf * SZ_EPS * 3;
        double vx1 = bx - x, vy1 = by - y;
        bool px1 = vx1 > 0, py1 = vy1 > 0;
        double q1 = (double)((!px1 && py1) + 2 * (!px1 && !py1) + 3 * (px1 && !py1));
        double cross = vx0 * vy1 - vx1 * vy0;
        double sgn = (double)((cross > 0) - (cross < 0));
        if (fabs(cross) < seps) sgn = 0;
        double dot = vx0 * vx1 + vy0 * vy1;
        double dq = q1 - q0;
        if (fabs(dq) == 3) dq = -dq / 3; else if (fabs(dq) == 2) dq = 2 * sgn;
        sumdq += dq;
        if (sgn == 0 && dot <= 0) on = true;
        ax = bx; ay = by; vx0 = vx1; vy0 = vy1; q0 = q1;
    }
    return (sumdq != 0) || on;
}

// |p_poly_dist(x, y, xv, yv)|: unsigned distance from one point to a closed outline (nv points, first == last).
SZ_HD double abs_poly_dist_xy(const double* vx, const double* vy, int nv, double xq, double yq)
{
    const int ns = nv - 1;
    double dpv_min = SZ_INF; int i_dpv = 0;
    for (int k = 0; k < nv; ++k) {
        double d = hypot(vx[k] - xq, vy[k] - yq);
        if (fabs(d) < dpv_min) { dpv_min = fabs(d); i_dpv = k; }
    }
    double cr_min = 0; int i_cr = 0; bool have = false;
    for (int k = 0; k < ns; ++k) {
        double dvx = vx[k + 1] - vx[k], dvy = vy[k + 1] - vy[k];
        double vds = hypot(dvx, dvy);
        double ct = dvx / vds, st = dvy / vds;
        double p1rx = ct * vx[k] + st * vy[k];
        double p1ry = -st * vx[k] + ct * vy[k];
        double r = (xq * ct + yq * st) - p1rx;
        double cr = (xq * (-st) + yq * ct) - p1ry;
        if (r > 0 && r < vds) { double a = fabs(cr); if (!have || a < cr_min) { cr_min = a; i_cr = k; have = true; } }
    }
    bool is_vertex = !have || ((i_cr != i_dpv) && (cr_min - dpv_min) > 0);
    return is_vertex ? dpv_min : cr_min;
}
SZ_HD bool outline_ok_for_poly_dist_xy(const double* vx, const double* vy, int nv)   // p_poly_dist.m:166-179
{
    const int ns = nv - 1;
    if (nv < 3) return false;
    double s = 0, last = 0;
    for (int k = 0; k < ns; ++k) {
        double vds = hypot(vx[k + 1] - vx[k], vy[k + 1] - vy[k]);
        if (vds < 10 * SZ_EPS) return false;
        if (k + 1 < ns) s += vds; else last = vds;
    }
    return !((s - last) < 10 * SZ_EPS);
}
// the same against the closed outline c1 of the workspace (n1 points, first == last); the reference would raise for
// repeated vertices / a flat polygon (outline_ok)
template <class W>
SZ_HD double abs_poly_dist(const W& w, double xq, double yq) { return abs_poly_dist_xy(w.c1x, w.c1y, w.n1, xq, yq); }
template <class W>
SZ_HD bool outline_ok_for_poly_dist(const W& w) { return outline_ok_for_poly_dist_xy(w.c1x, w.c1y, w.n1); }

// ------------------------------------------------------------------------------------------------
// The general contact-direction branch (floe_interactions.m:117-137) done by the 32 lanes of a warp for ONE of them.
// The branch is rare (0.7 % of the overlap regions of the packed Voronoi field: InterX found a number of points other than two
// near the region's vertices) and long: for every edge of the region one inpolygon() of a probe point against the region
// and one p_poly_dist() of the edge's midpoint against floe 1's outline -- about 6,000 instructions that a single lane used
// to run while the other 31, and with the CTA-wide votes of class C the other 15 warps, waited: leaving the branch out
// (a timing experiment with wrong results) made class C 0.6 ms = 8 % faster.  Here lane src's region and outline are dealt
// out over the lanes (vertex l to lane l), every lane evaluates its own edge of the region for inpolygon and its own vertex
// and segment of the outline for p_poly_dist, and the lanes combine: the winding sum is a sum of small integers (exact in
// any order), the two minima of p_poly_dist are "first index of the smallest value", which a (value, index) reduction
// reproduces.  Every term is the expression the one-lane code evaluates (in_region, abs_poly_dist_xy above), so the
// result is bit-identical; a NaN anywhere makes the function return false and the lane runs the one-lane code.
// Needs nr <= 32 region vertices and n1 <= 32 outline points (the caller checks), all 32 lanes converged.
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void warp_argmin_first(double& v, int& i)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, d); const int oi = __shfl_xor_sync(0xffffffffu, i, d);
        if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
    }
}
__device__ __forceinline__ bool coop_general_direction(int src, int lane, const i64* RX, const i64* RY, int nr_, const double* c1x, const double* c1y, int n1_,
                                                       double force_factor_, double on_edge_tol, double& ofx, double& ofy, double& odl)
{
    const unsigned FULL = 0xffffffffu;
    const int NR = __shfl_sync(FULL, nr_, src), N1 = __shfl_sync(FULL, n1_, src);
    const double ff = __shfl_sync(FULL, force_factor_, src);
    // deal out: lane l keeps region vertex l (metres, as the one-lane code converts it at every use) and outline point l
    double rx = 0, ry = 0, ox = 0, oy = 0;
    const int nmax = NR > N1 ? NR : N1;
    for (int j = 0; j < nmax; ++j) {
        double a = 0, b = 0, c = 0, d = 0;
        if (lane == src) {
            if (j < NR) { a = (double)RX[j] / SZ_SCALE; b = (double)RY[j] / SZ_SCALE; }
            if (j < N1) { c = c1x[j]; d = c1y[j]; }
        }
        a = __shfl_sync(FULL, a, src); b = __shfl_sync(FULL, b, src); c = __shfl_sync(FULL, c, src); d = __shfl_sync(FULL, d, src);
        if (lane == j) { rx = a; ry = b; ox = c; oy = d; }
    }
    // neighbours: next region vertex (closed ring), next outline point
    const int ln = (lane + 1 >= NR) ? 0 : lane + 1;
    const double rxn = __shfl_sync(FULL, rx, ln), ryn = __shfl_sync(FULL, ry, ln);
    const int lo = (lane + 1 > 31) ? 31 : lane + 1;
    const double oxn = __shfl_sync(FULL, ox, lo), oyn = __shfl_sync(FULL, oy, lo);
    // bounding box of the region (inpolygon's quick reject)
    double xmin = lane < NR ? rx : SZ_INF, xmax = lane < NR ? rx : -SZ_INF, ymin = lane < NR ? ry : SZ_INF, ymax = lane < NR ? ry : -SZ_INF;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        xmin = fmin(xmin, __shfl_xor_sync(FULL, xmin, d)); xmax = fmax(xmax, __shfl_xor_sync(FULL, xmax, d));
        ymin = fmin(ymin, __shfl_xor_sync(FULL, ymin, d)); ymax = fmax(ymax, __shfl_xor_sync(FULL, ymax, d));
    }
    // this lane's segment of the outline in p_poly_dist's rotated frame (does not depend on the query point)
    const int ns = N1 - 1;
    double vds = 0, ct = 0, st = 0, p1rx = 0, p1ry = 0;
    if (lane < ns) {
        const double dvx = oxn - ox, dvy = oyn - oy;
        vds = hypot(dvx, dvy);
        ct = dvx / vds; st = dvy / vds;
        p1rx = ct * ox + st * oy;
        p1ry = -st * ox + ct * oy;
    }
    bool bad = false;
    double sx = 0, sy = 0, sb = 0; int non = 0;
    for (int e = 0; e < NR; ++e) {
        const int e1 = (e + 1 == NR) ? 0 : e + 1;
        const double xa = __shfl_sync(FULL, rx, e), ya = __shfl_sync(FULL, ry, e), xb = __shfl_sync(FULL, rx, e1), yb = __shfl_sync(FULL, ry, e1);
        const double xgh = xb - xa, ygh = yb - ya, xm = (xb + xa) / 2, ym = (yb + ya) / 2;
        const double b = sqrt(xgh * xgh + ygh * ygh);
        double nx = -ygh / b, ny = xgh / b;
        const double xt = xm + nx / 100, yt = ym + ny / 100;
        // inpolygon(xt, yt, region): lane m evaluates edge m -> m + 1
        bool inside = false;
        if (xt >= xmin && xt <= xmax && yt >= ymin && yt <= ymax) {
            int dqi = 0; bool on = false;
            if (lane < NR) {
                const double vx0 = rx - xt, vy0 = ry - yt, vx1 = rxn - xt, vy1 = ryn - yt;
                const bool px0 = vx0 > 0, py0 = vy0 > 0, px1 = vx1 > 0, py1 = vy1 > 0;
                const double q0 = (double)((!px0 && py0) + 2 * (!px0 && !py0) + 3 * (px0 && !py0));
                const double q1 = (double)((!px1 && py1) + 2 * (!px1 && !py1) + 3 * (px1 && !py1));
                const double avx = fabs(0.5 * (rx + rxn)), avy = fabs(0.5 * (ry + ryn));
                double sf = avx > avy ? avx : avy; const double pr = avx * avy; if (pr > sf) sf = pr;
                const double seps = sf * SZ_EPS * 3;
                const double cross = vx0 * vy1 - vx1 * vy0;
                double sgn = (double)((cross > 0) - (cross < 0));
                if (fabs(cross) < seps) sgn = 0;
                const double dot = vx0 * vx1 + vy0 * vy1;
                double dq = q1 - q0;
                if (fabs(dq) == 3) dq = -dq / 3; else if (fabs(dq) == 2) dq = 2 * sgn;
                dqi = (int)dq;                                  // -2 .. 2
                on = (sgn == 0 && dot <= 0);
                if (!(cross == cross) || !(dot == dot)) bad = true;
            }
            const int sum = __reduce_add_sync(FULL, dqi);
            inside = (sum != 0) || __any_sync(FULL, on);
        }
        if (!inside) { nx = -nx; ny = -ny; }
        // |p_poly_dist(xm, ym, outline)|: lane k evaluates vertex k and segment k -> k + 1
        double dv = SZ_INF; int iv = 0x7fffffff;
        if (lane < N1) {
            const double a = fabs(hypot(ox - xm, oy - ym));
            if (a < SZ_INF) { dv = a; iv = lane; }
            if (!(a == a)) bad = true;
        }
        warp_argmin_first(dv, iv);
        if (iv == 0x7fffffff) iv = 0;
        double dc = SZ_INF; int ic = 0x7fffffff; bool have = false;
        if (lane < ns) {
            const double r = (xm * ct + ym * st) - p1rx;
            const double cr = (xm * (-st) + ym * ct) - p1ry;
            if (r > 0 && r < vds) { have = true; dc = fabs(cr); ic = lane; if (!(cr == cr)) bad = true; }
        }
        have = __any_sync(FULL, have);
        warp_argmin_first(dc, ic);
        const bool is_vertex = !have || ((ic != iv) && (dc - dv) > 0);
        const double dist = is_vertex ? dv : dc;
        if (dist < on_edge_tol) { sx += (-ff * b) * nx; sy += (-ff * b) * ny; sb += b; ++non; }
    }
    if (__any_sync(FULL, bad)) return false;
    ofx = 0; ofy = 0; odl = 0;
    if (non < NR && non > 0) {
        const double nrm = sqrt(sx * sx + sy * sy);
        ofx = sx / nrm; ofy = sy / nrm; odl = sb / (double)non;
    }
    return true;
}
#endif

// normal + tangential force of one overlap region (floe_interactions.m:167-187): r = Fx Fy Px Py overlap
SZ_HD void force_row(const Body& f1, const Body& f2, const Params& P, double G, double mu, double force_factor,
                     double fdx, double fdy, double dl, double Ak, double pcx, double pcy, double* r)
{
    const double fx = fdx * Ak * force_factor, fy = fdy * Ak * force_factor;  // :167
    // tangential (:170-183)
    const double v1x = f1.Ui + f1.ksi * (pcx - f1.Xi), v1y = f1.Vi + f1.ksi * (pcy - f1.Yi);
    const double v2x = f2.Ui + f2.ksi * (pcx - f2.Xi), v2y = f2.Vi + f2.ksi * (pcy - f2.Yi);
    const double vtx = v1x - v2x, vty = v1y - v2y;
    const double vn = sqrt(vtx * vtx + vty * vty);
    double dtx = 0, dty = 0;
    if (!((fabs(vtx) > fabs(vty) ? fabs(vtx) : fabs(vty)) == 0)) { dtx = vtx / vn; dty = vty / vn; }
    const double dotv = dtx * vtx + dty * vty;
    const double coef = -dotv * dl * G * vn;
    double ftx = coef * dtx * P.dt, fty = coef * dty * P.dt;
    const double fnorm = sqrt(fx * fx + fy * fy);
    if (sqrt(ftx * ftx + fty * fty) > mu * fnorm) { ftx = -mu * fnorm * dtx; fty = -mu * fnorm * dty; }
    r[0] = fx + ftx; r[1] = fy + fty; r[2] = pcx; r[3] = pcy; r[4] = Ak;
}

// ------------------------------------------------------------------------------------------------
// The force law.  On entry w.c1*/w.c2* hold the two world outlines exactly as the reference builds
// them (floe_interactions.m:25, floe_interactions_all.m:105; for the wall c2 = hole vertices).
// rows[r*5 + {0..4}] = Fx, Fy, Px, Py, overlap  for r < res.n_rows.
//
// Phase machine (one loop iteration = at most one Clipper execution per lane):
//   PH_CLIP1  clip #1 (:29/:34)        -> regions, areas, merge test, InterX (:43-84); first region
//   PH_CLIP2  clip #2 (:152-155)       -> the re-clip after the 1 m nudge along force_dir
//   PH_CLIP3  clip #3.. (:158)         -> one per new region: does it meet region k?  (toggles the sign)
// After the last clip of a region its force row is written (:167-187) and the next region's contact
// direction (:96-150) is prepared.
//
// FAST = true is the convex fast path (class C): clip #1 of a strictly convex pair runs in the specialised sweep of
// sz_convex.cuh, and the sign test must be decided by convex_sign_test; whenever either declines, the pair ends with
// PS_BAIL and the caller re-runs it with FAST = false.  Everything between the clips is the same code.
// Experiment (off by default, NOT measured yet): class C split into two kernels so that each one's per-thread working set is
// about half as large -- `convex_sweep_only` runs clip #1 alone and leaves the intersection polygon in a strided global
// buffer, `pair_force_impl<.., MODE = 2>` picks it up and runs the rest of the force law.  Item `item` of a buffer with
// `stride` items: st[item] = vertex count (0 = empty intersection), -1 = the sweep declined the pair or the polygon does not
// fit `cap` vertices; vertex k at x[k * stride + item] (lanes of a warp write neighbouring words).
struct ConvexHandoff { int* st; i64* x; i64* y; int stride; int cap; int item; };

template <class C, bool FAST, class W, int MODE = 0>
SZ_HD void pair_force_impl(W& w, const Body& f1, const Body& f2, bool boundary, const Params& P, PairResult& res, double* rows, bool valid, PairHints hints,
                           const ConvexHandoff* ho = nullptr)
{
    enum { PH_CLIP1 = 0, PH_CLIP2 = 1, PH_CLIP3 = 2, PH_DONE = 3, PH_NEXT = 4 };
    res.status = PS_OK; res.n_rows = 0; res.overlap_state = 0;
    int phase = valid ? PH_CLIP1 : PH_DONE;
    const bool convex_pair = hints.convex;
    double force_factor = 0, overlap = 0, amin = 0;
    const double G = P.modulus / (2 * (1 + P.nu)), mu = P.mu;
    const int method = boundary ? 0 : 1;
    if (valid) {
        double h1 = f1.h; const double h2 = f2.h;
        double r1 = sqrt(f1.area); const double r2 = sqrt(f2.area);
        force_factor = P.modulus * (h1 * h2) / (h1 * r2 + h2 * r1);                  // :12
        if (boundary) force_factor = P.modulus * h1 / r1;                             // :13-14
        else if (r1 > P.big_floe_r || r2 > P.big_floe_r) {                            // :15-19
            r1 = r1 < r2 ? r1 : r2; h1 = h1 < h2 ? h1 : h2;
            force_factor = P.modulus * h1 / r1;
        }
    }
    int k = 0, ii = 0, n_rows = 0, nr = 0;
    const i64* RX = w.rax; const i64* RY = w.ray;
    double fdx = 0, fdy = 0, dl = 0, pcx = 0, pcy = 0, Ak = 0;
    bool outline_checked = false, outline_ok = true;

    // class C: clip #1 in the four-edge sweep of sz_convex.cuh, one scanbeam per iteration with the lanes kept together
    int fast_clip1 = PS_BAIL;
    if constexpr (FAST && MODE == 2) {
        // clip #1 was swept by convex_sweep_only: fetch its polygon
        if (valid && convex_pair && !boundary && hints.no1 >= 3 && hints.no2 >= 3) {
            const int n = ho->st[ho->item];
            if (n >= 0 && n <= C::RV) {
                for (int t = 0; t < n; ++t) { w.rax[t] = ho->x[(size_t)t * ho->stride + ho->item]; w.ray[t] = ho->y[(size_t)t * ho->stride + ho->item]; }
                fast_clip1 = PS_OK; w.ra_off[0] = 0; w.ra_off[1] = n; w.ra_n = n > 0 ? 1 : 0;
            }
        }
        SZ_LANE_SYNC();
    }
    if constexpr (FAST && MODE != 2) {
        bool run = false;
        const bool go = valid && convex_pair && !boundary && hints.no1 >= 3 && hints.no2 >= 3;
        ClipInput subj, clip;
        subj.x = w.c1x; subj.y = w.c1y; subj.dx = 0; subj.dy = 0; subj.ix = subj.iy = 0; subj.n = hints.no1; subj.ring = 0; subj.rot = hints.rot1;
        clip.x = w.c2x; clip.y = w.c2y; clip.dx = 0; clip.dy = 0; clip.ix = clip.iy = 0; clip.n = hints.no2; clip.ring = 0; clip.rot = hints.rot2;
        static_assert(C::NP >= 2 * C::NV, "the InterX buffers must hold both int64 outlines");
        szcvx::ConvexSweep<C::NV> cs;
        const szcvx::SweepMem mem{w.svx, w.svy, w.rbx, w.rby, C::RV};       // outlines in the InterX buffers, deque in the clip #2 buffers
        if (go) { cs.load_ring(mem, 0, subj, subj.n); cs.load_ring(mem, 1, clip, clip.n); }
        SZ_LANE_SYNC();
        if (go) run = cs.begin(mem);
#ifndef SZ_C_STEPS_PER_SYNC
#define SZ_C_STEPS_PER_SYNC 1
#endif
        for (;;) {
#if defined(SZ_C_SWEEP_WARP) && defined(__CUDA_ARCH__)
            if (!__any_sync(0xffffffffu, run)) break;        // experiment: warps sweep independently, the CTA meets again below
#else
            if (!SZ_WARP_ANY(run)) break;
#endif
            for (int u = 0; u < SZ_C_STEPS_PER_SYNC; ++u) { if (run) run = cs.step(mem); SZ_LANE_SYNC(); }
        }
        if (go) {
            int n_out = 0;
            if (cs.finish(mem, w.rax, w.ray, C::RV, n_out) == szcvx::CV_OK) { fast_clip1 = PS_OK; w.ra_off[0] = 0; w.ra_off[1] = n_out; w.ra_n = n_out > 0 ? 1 : 0; }
        }
        SZ_LANE_SYNC();
    }
    // The loop body is a SEQUENCE of predicated blocks with a lane re-convergence point after each one (no early
    // `continue`): the lanes of a warp resolve different pairs, and a block that one lane leaves early must not make
    // the others run the rest of the body one lane at a time (first profile: InterX ran with 1.2 of 32 lanes active).
    for (;;) {
#if defined(SZ_FAST_MAIN_WARP) && defined(__CUDA_ARCH__)
        if (!(FAST ? __any_sync(0xffffffffu, phase != PH_DONE) : SZ_WARP_ANY(phase != PH_DONE))) break;
#else
        if (!SZ_WARP_ANY(phase != PH_DONE)) break;
#endif
        // ---- the clip this lane needs now
        ClipInput subj, clip;
        subj.x = w.c1x; subj.y = w.c1y; subj.dx = 0; subj.dy = 0; subj.ix = subj.iy = 0; subj.n = w.n1; subj.ring = 0; subj.rot = 0;
        clip.x = w.c2x; clip.y = w.c2y; clip.dx = 0; clip.dy = 0; clip.ix = clip.iy = 0; clip.n = w.n2; clip.ring = 0; clip.rot = 0;
        if (convex_pair && !boundary && hints.no1 >= 3 && hints.no2 >= 3) { subj.n = hints.no1; subj.rot = hints.rot1; clip.n = hints.no2; clip.rot = hints.rot2; }
        int m_now = method;
        if (phase == PH_CLIP2) { subj.dx = fdx; subj.dy = fdy; }
        else if (phase == PH_CLIP3) {
            subj.ring = 1; subj.ix = w.rbx + w.rb_off[ii]; subj.iy = w.rby + w.rb_off[ii]; subj.n = w.rb_off[ii + 1] - w.rb_off[ii]; subj.rot = 0;
            clip.ring = 1; clip.ix = RX; clip.iy = RY; clip.n = nr; clip.rot = 0;
            m_now = 1;
        }
        int st = PS_OK;
        if constexpr (FAST) {
            if (phase == PH_CLIP1) st = fast_clip1;                                   // swept before the loop
            else if (phase == PH_CLIP2 || phase == PH_CLIP3) st = PS_BAIL;            // the sign test was not certified
        } else st = run_sweep(w.eng, phase != PH_DONE && phase != PH_NEXT, m_now, subj, clip);
        if (phase != PH_DONE && phase != PH_NEXT && st != PS_OK) { res.status = st; phase = PH_DONE; }
        const int ph = phase;           // the phase whose clip just ran
        bool next_region = (phase == PH_NEXT);       // prepare the contact direction of region k
        bool finish_region = false;     // clips of region k are done: write its row
        bool advance = false;           // look at the next new region of the sign test
        SZ_LANE_SYNC();

        // ---- after clip #1 (:29-84)
        bool a1 = (ph == PH_CLIP1) && phase != PH_DONE;
        if constexpr (!FAST) {
            if (a1) {
                RegionSink<C> sink(w.rax, w.ray, w.ra_off);
                w.eng.emit(sink);
                if (sink.overflow) { res.status = PS_CAPACITY; phase = PH_DONE; a1 = false; }
                else w.ra_n = sink.n_paths;
            }
        }
        SZ_LANE_SYNC();
        if (a1) {
            if (boundary && w.ra_n > 0) {                                             // :35-40
                if (ring_polyarea(w.rax, w.ray, w.ra_off[1]) / f1.area > P.wall_frac) overlap = SZ_INF;
            }
            double sum_ar = 0;                                                        // :43-51
            for (int q = 0; q < w.ra_n; ++q) {
                double a, cx, cy;
                ring_area_centroid(w.rax + w.ra_off[q], w.ray + w.ra_off[q], w.ra_off[q + 1] - w.ra_off[q], a, cx, cy);
                w.ar[q] = a; sum_ar += a;
            }
            {   // merge test (:54-60)
                bool guard = P.periodic != 0;
                if (!guard && P.has_box) {
                    double xmx = w.c1x[0], xmn = w.c1x[0], ymx = w.c1y[0], ymn = w.c1y[0];
                    for (int i = 1; i < w.n1; ++i) {
                        if (w.c1x[i] > xmx) xmx = w.c1x[i];
                        if (w.c1x[i] < xmn) xmn = w.c1x[i];
                        if (w.c1y[i] > ymx) ymx = w.c1y[i];
                        if (w.c1y[i] < ymn) ymn = w.c1y[i];
                    }
                    guard = (xmx < P.bxmax && xmn > P.bxmin && ymx < P.bymax && ymn > P.bymin) || f2.area < P.domain_area_frac * P.barea;
                }
                if (guard) {
                    if (sum_ar / f1.area > P.merge_frac) overlap = SZ_INF;
                    else if (sum_ar / f2.area > P.merge_frac) overlap = -SZ_INF;
                }
            }
            // close the outlines when their ends are more than close_gap apart (:62-67); capacity reserved by the caller
            {
                double gx = w.c1x[0] - w.c1x[w.n1 - 1], gy = w.c1y[0] - w.c1y[w.n1 - 1];
                if (sqrt(gx * gx + gy * gy) > P.close_gap) { w.c1x[w.n1] = w.c1x[0]; w.c1y[w.n1] = w.c1y[0]; ++w.n1; }
                gx = w.c2x[0] - w.c2x[w.n2 - 1]; gy = w.c2y[0] - w.c2y[w.n2 - 1];
                if (sqrt(gx * gx + gy * gy) > P.close_gap) { w.c2x[w.n2] = w.c2x[0]; w.c2y[w.n2] = w.c2y[0]; ++w.n2; }
            }
        }
        SZ_LANE_SYNC();
        if (a1) {
            if (!interx<C>(w)) { res.status = PS_CAPACITY; phase = PH_DONE; a1 = false; }      // :70
        }
        SZ_LANE_SYNC();
        if (a1) {
            res.overlap_state = overlap;
            if (w.np < 2 || overlap == SZ_INF || overlap == -SZ_INF || w.ra_n == 0) phase = PH_DONE;      // :71-74 zero force
            else {
                res.overlap_state = 0;
                const int N1 = w.n1 - 1, N2 = w.n2 - 1;
                amin = (double)(N1 < N2 ? N1 : N2) * P.amin_per_vertex;               // :79
                k = -1; next_region = true;
            }
        }
        if constexpr (!FAST) {
        // ---- after clip #2 (:152-155)
        if (ph == PH_CLIP2 && phase != PH_DONE) {
            RegionSink<C> sink(w.rbx, w.rby, w.rb_off);
            w.eng.emit(sink);
            if (sink.overflow) { res.status = PS_CAPACITY; phase = PH_DONE; }
            else {
                w.rb_n = sink.n_paths;
                ii = 0;
                advance = true;
            }
        }
        // ---- after clip #3 (:158-164): only "empty or not" matters
        if (ph == PH_CLIP3 && phase != PH_DONE) {
            CountSink cs; cs.n = 0;
            w.eng.emit(cs);
            if (cs.n > 0) {
                const double anew = ring_polyarea(w.rbx + w.rb_off[ii], w.rby + w.rb_off[ii], w.rb_off[ii + 1] - w.rb_off[ii]);   // :160
                if (anew / Ak - 1 > 0) { fdx = -fdx; fdy = -fdy; }                     // :161-163
            }
            ++ii;
            advance = true;
        }
        }
        SZ_LANE_SYNC();
        // ---- next new region: answered by the certificate when one point is well inside both rings, else clip #3
        if (advance) {
            phase = PH_CLIP3;
            const double cX = pcx * SZ_SCALE, cY = pcy * SZ_SCALE;      // centroid of region k
            while (ii < w.rb_n) {
                const i64* NX = w.rbx + w.rb_off[ii]; const i64* NY = w.rby + w.rb_off[ii];
                const int nn = w.rb_off[ii + 1] - w.rb_off[ii];
                if (!(Ak != 0 && certify_inside(RX, RY, nr, cX, cY) && certify_inside(NX, NY, nn, cX, cY))) break;
                const double anew = ring_polyarea(NX, NY, nn);                         // :160
                if (anew / Ak - 1 > 0) { fdx = -fdx; fdy = -fdy; }                     // :161-163
                ++ii;
            }
            if (ii >= w.rb_n) finish_region = true;
        }
        SZ_LANE_SYNC();

        // ---- the row of region k (:167-187)
        if (finish_region) {
            force_row(f1, f2, P, G, mu, force_factor, fdx, fdy, dl, Ak, pcx, pcy, rows + (size_t)n_rows * 5);
            ++n_rows;
            next_region = true;
        }
        SZ_LANE_SYNC();

        // ---- contact direction of the next region with Ar >= Amin (:83,92-150)
        if (next_region) {
            ++k;
            while (k < w.ra_n && w.ar[k] < amin) ++k;
            if (k >= w.ra_n) { phase = PH_DONE; next_region = false; }
            else if (n_rows >= C::ROWS) { res.status = PS_CAPACITY; phase = PH_DONE; next_region = false; }
        }
        int m = 0; double cx = 0, cy = 0, p0x = 0, p0y = 0, p1x = 0, p1y = 0;
        if (next_region) {
            RX = w.rax + w.ra_off[k]; RY = w.ray + w.ra_off[k];
            nr = w.ra_off[k + 1] - w.ra_off[k];
            Ak = w.ar[k];
            double a_unused;
            ring_area_centroid(RX, RY, nr, a_unused, cx, cy);                         // :96-97
        }
        SZ_LANE_SYNC();
        if (next_region) {
            // dsearchn + dist<1 (:98-100)
            for (int q = 0; q < w.np; ++q) {
                double best = SZ_INF; int bi = 0;
                for (int v = 0; v < nr; ++v) {
                    double dx = (double)RX[v] / SZ_SCALE - w.px[q], dy = (double)RY[v] / SZ_SCALE - w.py[q];
                    double d2 = dx * dx + dy * dy;
                    if (d2 < best) { best = d2; bi = v; }
                }
                if (sqrt(best) < P.vertex_match_tol) {
                    if (m == 0) { p0x = (double)RX[bi] / SZ_SCALE; p0y = (double)RY[bi] / SZ_SCALE; }
                    else if (m == 1) { p1x = (double)RX[bi] / SZ_SCALE; p1y = (double)RY[bi] / SZ_SCALE; }
                    ++m;
                }
            }
        }
        SZ_LANE_SYNC();
        // the general branch (:117-137) of the lanes that need it, one lane at a time with the whole warp working on it
        bool coop_done = false; double cfx = 0, cfy = 0, cdl = 0;
#if defined(__CUDA_ARCH__) && !defined(SZ_NO_COOP_GENERAL)
        {
            bool ask = next_region && Ak != 0 && m != 2 && m != 0 && nr <= 32 && w.n1 <= 32;
            if (ask) {
                if (!outline_checked) { outline_ok = outline_ok_for_poly_dist(w); outline_checked = true; }
                ask = outline_ok;
            }
            unsigned req = __ballot_sync(0xffffffffu, ask);
            const int lane = threadIdx.x & 31;
            while (req) {
                const int src = __ffs((int)req) - 1; req &= req - 1;
                double gx = 0, gy = 0, gl = 0;
                const bool good = coop_general_direction(src, lane, RX, RY, nr, w.c1x, w.c1y, w.n1, force_factor, P.on_edge_tol, gx, gy, gl);
                if (lane == src && good) { coop_done = true; cfx = gx; cfy = gy; cdl = gl; }
            }
        }
#endif
        if (next_region) {
            fdx = 0; fdy = 0; dl = 0; pcx = cx; pcy = cy;
            if (Ak == 0) { pcx = 0; pcy = 0; }                                        // :103-106
            else if (m == 2) {                                                        // :107-112
                double xgh = p1x - p0x, ygh = p1y - p0y;
                double b = sqrt(xgh * xgh + ygh * ygh);
                fdx = -ygh / b; fdy = xgh / b; dl = b;
            } else if (m != 0 && coop_done) { fdx = cfx; fdy = cfy; dl = cdl; }
            else if (m != 0) {
                // general branch (:117-137), streamed edge by edge
                if (!outline_checked) { outline_ok = outline_ok_for_poly_dist(w); outline_checked = true; }
                if (!outline_ok) { res.status = PS_BAD_POLY; phase = PH_DONE; next_region = false; }
                else {
                    double xmin = SZ_INF, xmax = -SZ_INF, ymin = SZ_INF, ymax = -SZ_INF;
                    for (int v = 0; v < nr; ++v) {
                        double x = (double)RX[v] / SZ_SCALE, y = (double)RY[v] / SZ_SCALE;
                        if (x < xmin) xmin = x;
                        if (x > xmax) xmax = x;
                        if (y < ymin) ymin = y;
                        if (y > ymax) ymax = y;
                    }
                    double sx = 0, sy = 0, sb = 0; int non = 0;
                    for (int e = 0; e < nr; ++e) {
                        int e1 = (e + 1 == nr) ? 0 : e + 1;
                        double xa = (double)RX[e] / SZ_SCALE, ya = (double)RY[e] / SZ_SCALE, xb = (double)RX[e1] / SZ_SCALE, yb = (double)RY[e1] / SZ_SCALE;
                        double xgh = xb - xa, ygh = yb - ya, xm = (xb + xa) / 2, ym = (yb + ya) / 2;
                        double b = sqrt(xgh * xgh + ygh * ygh);
                        double nx = -ygh / b, ny = xgh / b;
                        double xt = xm + nx / 100, yt = ym + ny / 100;
                        if (!in_region(xt, yt, RX, RY, nr, xmin, xmax, ymin, ymax)) { nx = -nx; ny = -ny; }
                        double d = abs_poly_dist(w, xm, ym);
                        if (d < P.on_edge_tol) {
                            sx += (-force_factor * b) * nx; sy += (-force_factor * b) * ny; sb += b; ++non;
                        }
                    }
                    if (non < nr && non > 0) {
                        double nrm = sqrt(sx * sx + sy * sy);
                        fdx = sx / nrm; fdy = sy / nrm; dl = sb / (double)non;
                    }
                }
            }
            if (next_region) {
                if (dl < P.dl_min) { fdx = 0; fdy = 0; }                              // :141-142
                phase = PH_CLIP2;                                                     // sign test (:151-165)
            }
        }
        SZ_LANE_SYNC();
        // ---- convex pairs: the sign test decided with margin, without clips #2/#3 (else the reference's clips run)
        bool fast_done = false;
        if (next_region && convex_pair && !boundary && Ak != 0 && (fdx != 0 || fdy != 0)) {
            const int dec = convex_sign_test<C>(w, fdx, fdy, RX, RY, nr, Ak, pcx, pcy);
            if (dec != 0) { if (dec > 0) { fdx = -fdx; fdy = -fdy; } fast_done = true; }
        }
        SZ_LANE_SYNC();
        if (fast_done) {
            // the row of region k, then straight on to the next region in the next iteration
            force_row(f1, f2, P, G, mu, force_factor, fdx, fdy, dl, Ak, pcx, pcy, rows + (size_t)n_rows * 5);
            ++n_rows;
            phase = PH_NEXT;
        }
    }
    if (res.status != PS_OK) { res.n_rows = 0; return; }
    // caller rule (floe_interactions_all.m:135): rows are kept only when some force component is non-zero
    double sabs = 0;
    for (int r = 0; r < n_rows; ++r) sabs += fabs(rows[r * 5]) + fabs(rows[r * 5 + 1]);
    res.n_rows = (sabs != 0) ? n_rows : 0;
}
// First half of the split class C (see ConvexHandoff): clip #1 of a strictly convex floe-floe pair by the four-edge sweep, the
// same statements as the FAST block of pair_force_impl, with the outlines read where they lie (subj / clip may point at
// global memory: (x[i] + dx) * 2^32 is the arithmetic of polyclip.m:66 either way) and the result handed over.
// Storage: 2 * NV int64 for the outlines, dcap entries for the output deque, dcap for the result ring.
template <int NV>
SZ_HD void convex_sweep_only(bool go, const ClipInput& subj, const ClipInput& clip, i64* svx, i64* svy, i64* dqx, i64* dqy, i64* ox, i64* oy, int dcap, const ConvexHandoff& ho)
{
    szcvx::ConvexSweep<NV> cs;
    const szcvx::SweepMem mem{svx, svy, dqx, dqy, dcap};
    bool run = false;
    if (go) { cs.load_ring(mem, 0, subj, subj.n); cs.load_ring(mem, 1, clip, clip.n); }
    SZ_LANE_SYNC();
    if (go) run = cs.begin(mem);
    for (;;) {
        if (!SZ_WARP_ANY(run)) break;
        if (run) run = cs.step(mem);
        SZ_LANE_SYNC();
    }
    if (go) {
        int n_out = 0, st = -1;
        if (cs.finish(mem, ox, oy, dcap, n_out) == szcvx::CV_OK && n_out <= ho.cap) {
            st = n_out;
            for (int t = 0; t < n_out; ++t) { ho.x[(size_t)t * ho.stride + ho.item] = ox[t]; ho.y[(size_t)t * ho.stride + ho.item] = oy[t]; }
        }
        ho.st[ho.item] = st;
    }
    SZ_LANE_SYNC();
}
// second half: everything after clip #1
template <class C>
SZ_HD void pair_force_convex_after_sweep(WorkspaceLite<C>& w, const Body& f1, const Body& f2, const Params& P, PairResult& res, double* rows, bool valid, PairHints hints, const ConvexHandoff& ho)
{
    pair_force_impl<C, true, WorkspaceLite<C>, 2>(w, f1, f2, false, P, res, rows, valid, hints, &ho);
}

template <class C>
SZ_HD void pair_force(Workspace<C>& w, const Body& f1, const Body& f2, bool boundary, const Params& P, PairResult& res, double* rows, bool valid = true, PairHints hints = PairHints{false, 0, 0, 0, 0})
{
    pair_force_impl<C, false>(w, f1, f2, boundary, P, res, rows, valid, hints);
}
// class C: strictly convex floe-floe pairs; res.status == PS_BAIL sends the pair to pair_force()
template <class C>
SZ_HD void pair_force_convex(WorkspaceLite<C>& w, const Body& f1, const Body& f2, const Params& P, PairResult& res, double* rows, bool valid, PairHints hints)
{
    pair_force_impl<C, true>(w, f1, f2, false, P, res, rows, valid, hints);
}

}  // namespace szpf
