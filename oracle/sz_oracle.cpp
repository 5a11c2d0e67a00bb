// sz_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-timestep
// contact path, used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm
// as the CHECKER.  Nothing in the product (subzero_b200/) may include, link or call this file.
//
// It restates, function by function, the MATLAB sources (which cannot run here: no MATLAB/Octave)
//   floe_interactions_all.m:9-285          -> floe_interactions_all()
//   collisions/floe_interactions.m:1-199   -> floe_interactions()
//   collisions/InterX.m:46-81              -> InterX()
//   polyclip.m:63-73                       -> polyclip()
//   polygon_operations/p_poly_dist.m:111-288 -> p_poly_dist()
//   polygon_operations/inpolygon.m:64-224  -> inpolygon()
//   calc_trajectory.m:9-13, calc_collisionNum.m:3-6
// and calls the UNMODIFIED reference Clipper 6.4.2 (oracle/_ref/libclipper_ref.so, built from
// /root/reference/private/clipper.cpp by oracle/Makefile) for every polygon clip.
//
// Pinning: the Clipper layer is the reference itself; polyshape area/centroid are pinned to the
// 462 known answers in test/test_conservation/FloeShapes.mat (tests/golden/floe_shapes.npz) and
// the clipper_test.m square case.  The force law, InterX, broad phase, ghosts, torque and stress
// have NO numeric pin in the reference ("parity unpinned"): fidelity rests on the
// line-by-line citations below.  MATLAB built-ins without source are restated per SURVEY.md App. C.
//
// Build: -O2 -ffp-contract=off, no -march=native (FMA changes Clipper's output).
#include "../include/subzero_b200.h"
#include "sz_oracle.h"
#include <vector>
#include <array>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <atomic>
#include <string>
#include <stdexcept>

extern "C" int szref_clip(const int64_t* sx, const int64_t* sy, int ns, const int64_t* cx, const int64_t* cy, int nc,
                          int method, int64_t* out_x, int64_t* out_y, int out_cap, int* out_off, int off_cap);

namespace {

typedef std::vector<double> vec;
const double INF = std::numeric_limits<double>::infinity();
const double EPS = 2.220446049250313e-16;   // MATLAB eps

struct Curve { vec x, y; };                 // one polygon / polyline, MATLAB column vectors
struct IPoly { std::vector<int64_t> x, y; };

// ---- MATLAB int64(): round to nearest, ties away from zero, saturating, NaN -> 0 (polyclip.m:66)
int64_t matlab_int64(double v)
{
    if (std::isnan(v)) return 0;
    if (v >= 9223372036854775807.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    double t = std::trunc(v), f = v - t;
    int64_t r = (int64_t)t;
    if (f >= 0.5) r += 1; else if (f <= -0.5) r -= 1;
    return r;
}

struct ClipperError : std::runtime_error { ClipperError() : std::runtime_error("Clipper Error.") {} };

// ---- polyclip.m:63-73 + private/mexclipper.cpp:204-305.  method: 0 dif, 1 int (polyclip.m:52-58).
// Returns the paths both as doubles (what MATLAB sees) and as Clipper's integers (parity/debug).
void polyclip(const Curve& p1, const Curve& p2, int method, std::vector<Curve>& out, std::vector<IPoly>* iout = nullptr)
{
    const double scale = 4294967296.0;   // 2^32
    std::vector<int64_t> sx(p1.x.size()), sy(p1.x.size()), cx(p2.x.size()), cy(p2.x.size());
    for (size_t i = 0; i < p1.x.size(); ++i) { sx[i] = matlab_int64(p1.x[i] * scale); sy[i] = matlab_int64(p1.y[i] * scale); }
    for (size_t i = 0; i < p2.x.size(); ++i) { cx[i] = matlab_int64(p2.x[i] * scale); cy[i] = matlab_int64(p2.y[i] * scale); }
    size_t cap = 4 * (sx.size() + cx.size()) + 64;
    std::vector<int64_t> ox, oy; std::vector<int> off;
    int np;
    for (;;) {
        ox.resize(cap); oy.resize(cap); off.resize(cap);
        np = szref_clip(sx.data(), sy.data(), (int)sx.size(), cx.data(), cy.data(), (int)cx.size(), method,
                        ox.data(), oy.data(), (int)cap, off.data(), (int)cap);
        if (np == -2) { cap *= 4; continue; }
        break;
    }
    if (np < 0) throw ClipperError();
    out.clear(); if (iout) iout->clear();
    for (int k = 0; k < np; ++k) {
        Curve c; IPoly ip;
        for (int v = off[k]; v < off[k + 1]; ++v) {
            // mexclipper.cpp:75-76 stores the int64 in a double; polyclip.m:67 divides by scale
            c.x.push_back((double)ox[v] / scale); c.y.push_back((double)oy[v] / scale);
            ip.x.push_back(ox[v]); ip.y.push_back(oy[v]);
        }
        out.push_back(c); if (iout) iout->push_back(ip);
    }
}

// ---- area(polyshape(X,Y)) / centroid(polyshape(X,Y)): shoelace on coordinates relative to vertex 0
// (pinned by the 462 BoundaryInfo records of FloeShapes.mat to 1.5e-15 / 3.4e-15 sqrt(A))
void polyshape_area_centroid(const vec& X, const vec& Y, double& area, double& cx, double& cy)
{
    const size_t n = X.size();
    if (n < 3) { area = 0; cx = cy = std::nan(""); return; }
    const double x0 = X[0], y0 = Y[0];
    double a2 = 0, sx = 0, sy = 0;
    for (size_t i = 0; i < n; ++i) {
        size_t j = (i + 1 == n) ? 0 : i + 1;
        double xi = X[i] - x0, yi = Y[i] - y0, xj = X[j] - x0, yj = Y[j] - y0;
        double c = xi * yj - xj * yi;
        a2 += c; sx += (xi + xj) * c; sy += (yi + yj) * c;
    }
    area = std::fabs(a2) / 2;
    cx = x0 + sx / (3 * a2);
    cy = y0 + sy / (3 * a2);
}

// ---- polyarea(x,y): abs(sum((x([2:end 1])-x).*(y([2:end 1])+y))/2)   (floe_interactions.m:37,160)
double polyarea(const vec& x, const vec& y)
{
    const size_t n = x.size(); double s = 0;
    for (size_t i = 0; i < n; ++i) { size_t j = (i + 1 == n) ? 0 : i + 1; s += (x[j] - x[i]) * (y[j] + y[i]); }
    return std::fabs(s / 2);
}

// ---- collisions/InterX.m:46-81, two-curve form (hF = @le).  Returns P as rows (x,y), sorted unique.
void InterX(const Curve& L1, const Curve& L2, std::vector<std::array<double, 2>>& P)
{
    P.clear();
    const int n1 = (int)L1.x.size() - 1, n2 = (int)L2.x.size() - 1;   // segments
    if (n1 < 1 || n2 < 1) return;
    vec dx1(n1), dy1(n1), dx2(n2), dy2(n2), S1(n1), S2(n2);
    for (int i = 0; i < n1; ++i) { dx1[i] = L1.x[i + 1] - L1.x[i]; dy1[i] = L1.y[i + 1] - L1.y[i]; }
    for (int j = 0; j < n2; ++j) { dx2[j] = L2.x[j + 1] - L2.x[j]; dy2[j] = L2.y[j + 1] - L2.y[j]; }
    for (int i = 0; i < n1; ++i) S1[i] = dx1[i] * L1.y[i] - dy1[i] * L1.x[i];      // :60
    for (int j = 0; j < n2; ++j) S2[j] = dx2[j] * L2.y[j] - dy2[j] * L2.x[j];      // :61
    // find(C1 & C2) is column-major: j outer, i inner (:67); order is erased by unique() anyway
    for (int j = 0; j < n2; ++j) for (int i = 0; i < n1; ++i) {
        // C1(i,j): D(dx1.*y2 - dy1.*x2, S1) (:63, :79-81)
        double a0 = dx1[i] * L2.y[j] - dy1[i] * L2.x[j];
        double a1 = dx1[i] * L2.y[j + 1] - dy1[i] * L2.x[j + 1];
        bool C1 = ((a0 - S1[i]) * (a1 - S1[i])) <= 0;
        // C2(i,j): D((y1.*dx2 - x1.*dy2)', S2')' (:64)
        double b0 = L1.y[i] * dx2[j] - L1.x[i] * dy2[j];
        double b1 = L1.y[i + 1] * dx2[j] - L1.x[i + 1] * dy2[j];
        bool C2 = ((b0 - S2[j]) * (b1 - S2[j])) <= 0;
        if (!(C1 && C2)) continue;
        double L = dy2[j] * dx1[i] - dy1[i] * dx2[j];                              // :71
        if (L == 0) continue;                                                      // :72
        double px = (dx2[j] * S1[i] - dx1[i] * S2[j]) / L;                         // :75-76
        double py = (dy2[j] * S1[i] - dy1[i] * S2[j]) / L;
        P.push_back({px, py});
    }
    // unique(...,'rows'): ascending by x then y, exact duplicates removed (NaN rows, if any, stay distinct)
    std::sort(P.begin(), P.end(), [](const std::array<double, 2>& a, const std::array<double, 2>& b) {
        if (a[0] < b[0]) return true; if (a[0] > b[0]) return false; return a[1] < b[1]; });
    P.erase(std::unique(P.begin(), P.end(), [](const std::array<double, 2>& a, const std::array<double, 2>& b) {
        return a[0] == b[0] && a[1] == b[1]; }), P.end());
}

// ---- polygon_operations/inpolygon.m:64-224 (the file shadows MATLAB's builtin through paths.m:6).
// Restated algorithm: bounding-box prefilter, loop closing, quadrant-change winding number with the
// per-edge on-boundary tolerance 3*eps*max(|xm|,|ym|,|xm*ym|) of the edge midpoint.
void inpolygon(const vec& px, const vec& py, vec xv, vec yv, std::vector<char>& in)
{
    const size_t np = px.size();
    in.assign(np, 0);
    if (xv.empty()) return;
    double xmin = *std::min_element(xv.begin(), xv.end()), xmax = *std::max_element(xv.begin(), xv.end());
    double ymin = *std::min_element(yv.begin(), yv.end()), ymax = *std::max_element(yv.begin(), yv.end());
    if (xv.size() >= 3 && (xv.front() != xv.back() || yv.front() != yv.back())) { xv.push_back(xv[0]); yv.push_back(yv[0]); }  // close_loops :226-236
    const size_t Nv = xv.size();
    if (Nv < 2) return;
    vec scaledEps(Nv - 1);
    for (size_t m = 0; m + 1 < Nv; ++m) {
        double avx = std::fabs(0.5 * (xv[m] + xv[m + 1])), avy = std::fabs(0.5 * (yv[m] + yv[m + 1]));
        double sf = std::max(avx, avy); sf = std::max(sf, avx * avy);
        scaledEps[m] = sf * EPS * 3;
    }
    for (size_t p = 0; p < np; ++p) {
        const double x = px[p], y = py[p];
        if (!(x >= xmin && x <= xmax && y >= ymin && y <= ymax)) continue;   // mask (:71)
        double sumdq = 0; bool on = false;
        double vx0 = xv[0] - x, vy0 = yv[0] - y;
        auto quad = [](double vx, double vy) { bool posX = vx > 0, posY = vy > 0; return (double)((!posX && posY) + 2 * (!posX && !posY) + 3 * (posX && !posY)); };
        double q0 = quad(vx0, vy0);
        for (size_t m = 0; m + 1 < Nv; ++m) {
            double vx1 = xv[m + 1] - x, vy1 = yv[m + 1] - y;
            double q1 = quad(vx1, vy1);
            double cross = vx0 * vy1 - vx1 * vy0;
            double sgn = (cross > 0) - (cross < 0);
            if (std::fabs(cross) < scaledEps[m]) sgn = 0;
            double dot = vx0 * vx1 + vy0 * vy1;
            double dq = q1 - q0;
            if (std::fabs(dq) == 3) dq = -dq / 3;
            else if (std::fabs(dq) == 2) dq = 2 * sgn;
            sumdq += dq;
            if (sgn == 0 && dot <= 0) on = true;
            vx0 = vx1; vy0 = vy1; q0 = q1;
        }
        in[p] = (sumdq != 0) || on;
    }
}

// ---- polygon_operations/p_poly_dist.m:111-288, four-argument / one-output form (closed polygon,
// distance negative inside).  The caller only uses abs(d) (floe_interactions.m:127).
// Throws like the reference on repeated vertices / degenerate polygons (:166-179).
struct PolyDistError : std::runtime_error { PolyDistError(const char* m) : std::runtime_error(m) {} };
void p_poly_dist(const vec& xp, const vec& yp, vec xv, vec yv, vec& d_min)
{
    size_t nv = xv.size(); const size_t np = xp.size();
    if (nv < 3) throw PolyDistError("Polygon must have at least 3 vertices");
    if (xv[0] != xv[nv - 1] || yv[0] != yv[nv - 1]) { xv.push_back(xv[0]); yv.push_back(yv[0]); nv++; }   // :135-141
    const size_t ns = nv - 1;
    vec vds(ns), ct(ns), st(ns), p1rx(ns), p1ry(ns);
    for (size_t k = 0; k < ns; ++k) {
        double dvx = xv[k + 1] - xv[k], dvy = yv[k + 1] - yv[k];
        vds[k] = std::hypot(dvx, dvy);                                            // :163
        if (vds[k] < 10 * EPS) throw PolyDistError("Points of the polyline are identical");   // :166-169
        ct[k] = dvx / vds[k]; st[k] = dvy / vds[k];                               // :184-185
        p1rx[k] = ct[k] * xv[k] + st[k] * yv[k];                                  // :193-194
        p1ry[k] = -st[k] * xv[k] + ct[k] * yv[k];
    }
    { double s = 0; for (size_t k = 0; k + 1 < ns; ++k) s += vds[k];             // :174-179  s(end-1) - vds(end)
      if ((s - vds[ns - 1]) < 10 * EPS) throw PolyDistError("Polygon vertices should not lie on a straight line"); }
    d_min.assign(np, 0);
    for (size_t j = 0; j < np; ++j) {
        // distances to vertices (:151-155): first minimum
        double dpv_min = INF; size_t I_dpv = 0;
        for (size_t k = 0; k < nv; ++k) {
            double d = std::hypot(xv[k] - xp[j], yv[k] - yp[j]);
            if (std::fabs(d) < dpv_min) { dpv_min = std::fabs(d); I_dpv = k; }
        }
        // projections in each segment's rotated frame (:205-256)
        double cr_min = std::nan(""); size_t I_cr = 0;
        for (size_t k = 0; k < ns; ++k) {
            double r = (xp[j] * ct[k] + yp[j] * st[k]) - p1rx[k];                 // Pp*Cer21 - P1r(:,1)
            double cr = (xp[j] * (-st[k]) + yp[j] * ct[k]) - p1ry[k];             // Pp*Cer22 - P1r(:,2)
            if (r > 0 && r < vds[k]) {                                            // :256
                double a = std::fabs(cr);
                if (std::isnan(cr_min) || a < cr_min) { cr_min = a; I_cr = k; }   // min(abs(B),[],2) ignoring NaN (:264)
            }
        }
        bool is_vertex = std::isnan(cr_min) || ((I_cr != I_dpv) && (cr_min - dpv_min) > 0);   // :271-277
        d_min[j] = is_vertex ? dpv_min : cr_min;
    }
    std::vector<char> in;                                                         // :285-288
    inpolygon(xp, yp, xv, yv, in);
    for (size_t j = 0; j < np; ++j) if (in[j]) d_min[j] = -d_min[j];
}

// ---- one element of the (extended) Floe array, hot-path fields only (initialize_floe_values.m:12-52)
struct Partner {                       // Floe(i).potentialInteractions(k)   floe_interactions_all.m:104-112
    double floeNum;                    // 1-based index, or Inf for the wall
    Curve c; double Ui, Vi, h, area, Xi, Yi, ksi; bool is_boundary;
};
typedef std::array<double, 7> Row;
struct Floe {
    Curve c_alpha; double Xi, Yi, rmax, h, area, Ui, Vi, ksi; int alive;
    std::vector<Row> interactions; double OverlapArea; double cf[2]; double ct;
    std::vector<Partner> potentialInteractions;
    bool has_pi_from_before;           // never set: the field is rmfield-ed at :507 every call
};

struct PairDebug {                     // per candidate pair, for parity checks
    int i, j; double overlap_state; int n_regions; int status;
    std::vector<IPoly> clip1;
};

struct ForceOut { std::vector<std::array<double, 2>> force, pcontact; vec overlap; bool overlap_is_scalar; };

// ---- collisions/floe_interactions.m:1-199
void floe_interactions(const Floe& floe1, const Partner& floe2, const SzParams& P, const Curve& c2_boundary,
                       ForceOut& out, std::vector<IPoly>* clip1_dbg)
{
    const double Modulus = P.modulus, dt = P.dt;
    double h1 = floe1.h, h2 = floe2.h;
    double r1 = std::sqrt(floe1.area), r2 = std::sqrt(floe2.area);
    double Force_factor = Modulus * (h1 * h2) / (h1 * r2 + h2 * r1);               // :12
    double overlap = 0;
    if (floe2.is_boundary) Force_factor = Modulus * h1 / r1;                        // :13-14
    else if (r1 > P.big_floe_r || r2 > P.big_floe_r) {                              // :15-19
        r1 = std::min(r1, r2); h1 = std::min(h1, h2); Force_factor = Modulus * h1 / r1;
    }
    const double nu = P.nu, G = Modulus / (2 * (1 + nu)), mu = P.mu;                // :20-21

    Curve c1;                                                                       // :25
    for (size_t k = 0; k < floe1.c_alpha.x.size(); ++k) { c1.x.push_back(floe1.c_alpha.x[k] + floe1.Xi); c1.y.push_back(floe1.c_alpha.y[k] + floe1.Yi); }
    Curve c2 = floe2.c;                                                             // :27 / :31-32
    const bool boundary = floe2.is_boundary;
    std::vector<Curve> R;                                                           // Xi, Yi cell arrays
    polyclip(c1, c2, boundary ? 0 : 1, R, clip1_dbg);                               // :29 / :34
    if (boundary && !R.empty()) {                                                   // :35-40
        if (polyarea(R[0].x, R[0].y) / floe1.area > P.wall_frac) overlap = INF;
    }
    vec Ar;                                                                         // :43-51
    if (R.empty()) Ar.push_back(0);
    else for (auto& r : R) { double a, cx, cy; polyshape_area_centroid(r.x, r.y, a, cx, cy); Ar.push_back(a); }

    // :54-60 merge test
    {
        double c1xmax = *std::max_element(c1.x.begin(), c1.x.end()), c1xmin = *std::min_element(c1.x.begin(), c1.x.end());
        double c1ymax = *std::max_element(c1.y.begin(), c1.y.end()), c1ymin = *std::min_element(c1.y.begin(), c1.y.end());
        double bxmax = -INF, bxmin = INF, bymax = -INF, bymin = INF, barea = 0;
        if (!c2_boundary.x.empty()) {
            bxmax = *std::max_element(c2_boundary.x.begin(), c2_boundary.x.end()); bxmin = *std::min_element(c2_boundary.x.begin(), c2_boundary.x.end());
            bymax = *std::max_element(c2_boundary.y.begin(), c2_boundary.y.end()); bymin = *std::min_element(c2_boundary.y.begin(), c2_boundary.y.end());
            double cx, cy; polyshape_area_centroid(c2_boundary.x, c2_boundary.y, barea, cx, cy);   // area(polyshape(c2_boundary'))
        }
        bool guard = (c1xmax < bxmax && c1xmin > bxmin && c1ymax < bymax && c1ymin > bymin) ||
                     floe2.area < P.domain_area_frac * barea || P.periodic;
        if (guard) {
            double s = 0; for (double a : Ar) s += a;
            if (s / floe1.area > P.merge_frac) overlap = INF;
            else if (s / floe2.area > P.merge_frac) overlap = -INF;
        }
    }
    // :62-67 close the outlines if needed
    { size_t n = c1.x.size(); if (std::sqrt((c1.x[0] - c1.x[n - 1]) * (c1.x[0] - c1.x[n - 1]) + (c1.y[0] - c1.y[n - 1]) * (c1.y[0] - c1.y[n - 1])) > P.close_gap) { c1.x.push_back(c1.x[0]); c1.y.push_back(c1.y[0]); } }
    { size_t n = c2.x.size(); if (std::sqrt((c2.x[0] - c2.x[n - 1]) * (c2.x[0] - c2.x[n - 1]) + (c2.y[0] - c2.y[n - 1]) * (c2.y[0] - c2.y[n - 1])) > P.close_gap) { c2.x.push_back(c2.x[0]); c2.y.push_back(c2.y[0]); } }

    std::vector<std::array<double, 2>> Px;
    InterX(c1, c2, Px);                                                             // :70
    out.force.clear(); out.pcontact.clear(); out.overlap.clear();
    if (Px.empty() || Px.size() < 2 || std::isinf(overlap) || R.empty()) {          // :71-74
        out.force.push_back({0, 0}); out.pcontact.push_back({0, 0});
        out.overlap.push_back(overlap); out.overlap_is_scalar = true;
        return;
    }
    const int N1 = (int)c1.x.size() - 1, N2 = (int)c2.x.size() - 1;                  // :78
    const double Amin = std::min(N1, N2) * P.amin_per_vertex;                        // :79  min([N1,N2])*100/1.75
    { std::vector<Curve> R2; vec Ar2;                                                // :83
      for (size_t k = 0; k < R.size(); ++k) if (!(Ar[k] < Amin)) { R2.push_back(R[k]); Ar2.push_back(Ar[k]); }
      R.swap(R2); Ar.swap(Ar2); }
    const int N_contact = (int)R.size();
    out.overlap_is_scalar = (N_contact == 0);
    if (N_contact == 0) { out.overlap.push_back(0); return; }                        // force_1 = zeros(0,2), overlap stays 0

    for (int k = 0; k < N_contact; ++k) {                                            // :92-190
        const vec& X = R[k].x; const vec& Y = R[k].y;
        double a_unused, cx, cy; polyshape_area_centroid(X, Y, a_unused, cx, cy);    // :96-97
        // dsearchn([X Y],P') (:98): nearest region vertex per crossing point, first index on ties
        std::vector<std::array<double, 2>> p;
        for (auto& q : Px) {
            double best = INF; size_t bi = 0;
            for (size_t v = 0; v < X.size(); ++v) {
                double dx = X[v] - q[0], dy = Y[v] - q[1]; double d2 = dx * dx + dy * dy;
                if (d2 < best) { best = d2; bi = v; }
            }
            if (std::sqrt(best) < P.vertex_match_tol) p.push_back({X[bi], Y[bi]});   // :99
        }
        const int m = (int)p.size();
        double fdx = 0, fdy = 0, dl = 0, pcx = 0, pcy = 0;
        if (Ar[k] == 0) { fdx = fdy = 0; pcx = pcy = 0; dl = 0; }                    // :103-106
        else if (m == 2) {                                                           // :107-112
            pcx = cx; pcy = cy;
            double xgh = p[1][0] - p[0][0], ygh = p[1][1] - p[0][1];
            double b = std::sqrt(xgh * xgh + ygh * ygh);
            fdx = -ygh / b; fdy = xgh / b; dl = b;
        } else if (m == 0) { fdx = fdy = 0; pcx = cx; pcy = cy; dl = 0; }            // :113-116
        else {                                                                       // :117-137
            const size_t nr = X.size();
            vec xv(X), yv(Y); xv.push_back(X[0]); yv.push_back(Y[0]);
            vec xgh(nr), ygh(nr), xm(nr), ym(nr), b(nr), nx(nr), ny(nr), xt(nr), yt(nr);
            for (size_t e = 0; e < nr; ++e) {
                xgh[e] = xv[e + 1] - xv[e]; xm[e] = (xv[e + 1] + xv[e]) / 2;
                ygh[e] = yv[e + 1] - yv[e]; ym[e] = (yv[e + 1] + yv[e]) / 2;
                b[e] = std::sqrt(xgh[e] * xgh[e] + ygh[e] * ygh[e]);
                nx[e] = -ygh[e] / b[e]; ny[e] = xgh[e] / b[e];
                xt[e] = xm[e] + nx[e] / 100; yt[e] = ym[e] + ny[e] / 100;
            }
            std::vector<char> in; inpolygon(xt, yt, xv, yv, in);                     // :123
            vec Fnx(nr), Fny(nr);
            for (size_t e = 0; e < nr; ++e) {
                if (!in[e]) { nx[e] = -nx[e]; ny[e] = -ny[e]; }                      // :124
                Fnx[e] = (-Force_factor * b[e]) * nx[e]; Fny[e] = (-Force_factor * b[e]) * ny[e];   // :125
            }
            vec d_min1;
            p_poly_dist(xm, ym, c1.x, c1.y, d_min1);                                 // :126
            size_t non = 0; for (size_t e = 0; e < nr; ++e) if (std::fabs(d_min1[e]) < P.on_edge_tol) ++non;   // :127
            if (non < nr && non > 0) {                                               // :128-131
                double sx = 0, sy = 0, sb = 0;
                for (size_t e = 0; e < nr; ++e) if (std::fabs(d_min1[e]) < P.on_edge_tol) { sx += Fnx[e]; sy += Fny[e]; sb += b[e]; }
                double nrm = std::sqrt(sx * sx + sy * sy);
                fdx = sx / nrm; fdy = sy / nrm; dl = sb / (double)non;
            } else { fdx = fdy = 0; dl = 0; }
            pcx = cx; pcy = cy;                                                      // :136
        }
        // :140-150 direction of force
        Curve c1new;
        if (dl < P.dl_min) { fdx = fdy = 0; c1new = c1; }
        else { c1new = c1; for (auto& v : c1new.x) v += fdx; for (auto& v : c1new.y) v += fdy; }
        std::vector<Curve> Rnew;
        polyclip(c1new, c2, boundary ? 0 : 1, Rnew);                                 // :151-155
        Curve reg; reg.x = X; reg.y = Y;
        for (auto& rn : Rnew) {                                                      // :156-165
            std::vector<Curve> Xn;
            polyclip(rn, reg, 1, Xn);
            if (!Xn.empty()) {
                double Anew = polyarea(rn.x, rn.y);
                if (Anew / Ar[k] - 1 > 0) { fdx = -fdx; fdy = -fdy; }
            }
        }
        double fx = fdx * Ar[k] * Force_factor, fy = fdy * Ar[k] * Force_factor;     // :167
        // :170-183 tangential force
        double v1x = floe1.Ui + floe1.ksi * (pcx - floe1.Xi), v1y = floe1.Vi + floe1.ksi * (pcy - floe1.Yi);
        double v2x = floe2.Ui + floe2.ksi * (pcx - floe2.Xi), v2y = floe2.Vi + floe2.ksi * (pcy - floe2.Yi);
        double vtx = v1x - v2x, vty = v1y - v2y;
        double dtx, dty;
        double vn = std::sqrt(vtx * vtx + vty * vty);
        if (std::max(std::fabs(vtx), std::fabs(vty)) == 0) { dtx = dty = 0; } else { dtx = vtx / vn; dty = vty / vn; }
        double dotv = dtx * vtx + dty * vty;
        double coef = -dotv * dl * G * vn;
        double ftx = coef * dtx * dt, fty = coef * dty * dt;                         // :178
        double fnorm = std::sqrt(fx * fx + fy * fy);
        if (std::sqrt(ftx * ftx + fty * fty) > mu * fnorm) { ftx = -mu * fnorm * dtx; fty = -mu * fnorm * dty; }   // :180-183
        out.force.push_back({fx + ftx, fy + fty});                                   // :186
        out.overlap.push_back(Ar[k]);                                                // :187
        out.pcontact.push_back({pcx, pcy});
    }
}

}  // namespace

// ================================================================================================
struct SzoResult {
    int n0 = 0, n = 0;
    std::vector<int> parent, floe_num; vec gx, gy;           // ghosts
    std::vector<PairDebug> pairs;
    std::vector<int64_t> row_off; vec rows;
    vec fx, fy, torque, overlap_area, stress, xi, yi; std::vector<uint8_t> alive; std::vector<int> kill, transfer;
    double collision_count = 0; int n_clipper_fail = 0; long n_pairs_force = 0;
    std::string error;
};

namespace {

// floe_interactions_all.m:9-285
void floe_interactions_all(const SzParams& P, const SzFloesSoA& in, const SzBoundary* bnd, int nthreads, int broad_mode, SzoResult& res)
{
    const double Lx = P.Lx, Ly = P.Ly; const int Nb = P.Nb;
    std::vector<Floe> F(in.n);
    for (int i = 0; i < in.n; ++i) {
        Floe& f = F[i];
        for (int v = in.voff[i]; v < in.voff[i + 1]; ++v) { f.c_alpha.x.push_back(in.vx[v]); f.c_alpha.y.push_back(in.vy[v]); }
        f.Xi = in.x[i]; f.Yi = in.y[i]; f.rmax = in.rmax[i]; f.h = in.h[i]; f.area = in.area[i];
        f.Ui = in.u[i]; f.Vi = in.v[i]; f.ksi = in.ksi[i]; f.alive = in.alive[i];
        f.OverlapArea = 0; f.cf[0] = f.cf[1] = 0; f.ct = 0;
    }
    const int N0 = (int)F.size();
    std::vector<int> FloeNums(N0); for (int i = 0; i < N0; ++i) FloeNums[i] = i + 1;   // :14
    std::vector<int> parent;
    auto sgn = [](double v) { return (double)((v > 0) - (v < 0)); };
    if (P.periodic) {                                                                  // :18-66
        std::vector<Floe> ghostX;
        for (int i = 0; i < N0; ++i) {
            double m = -INF;   // max() skips NaN; max of nothing / all-NaN compares false
            for (size_t v = 0; v < F[i].c_alpha.x.size(); ++v) { double a = std::fabs(F[i].c_alpha.x[v] + F[i].Xi); if (a > m) m = a; }
            if (F[i].alive && m > Lx) {                                         // :31
                Floe g = F[i]; g.Xi = F[i].Xi - 2 * Lx * sgn(F[i].Xi);                 // :34
                ghostX.push_back(g); FloeNums.push_back(-std::abs(FloeNums[i])); parent.push_back(i + 1);
            }
        }
        F.insert(F.end(), ghostX.begin(), ghostX.end());
        const int N1 = (int)F.size();
        std::vector<Floe> ghostY;
        for (int i = 0; i < N1; ++i) {
            double m = -INF;
            for (size_t v = 0; v < F[i].c_alpha.y.size(); ++v) { double a = std::fabs(F[i].c_alpha.y[v] + F[i].Yi); if (a > m) m = a; }
            if (F[i].alive && m > Ly) {                                         // :52
                Floe g = F[i]; g.Yi = F[i].Yi - 2 * Ly * sgn(F[i].Yi);                 // :55
                ghostY.push_back(g); FloeNums.push_back(-std::abs(FloeNums[i])); parent.push_back(i + 1);
            }
        }
        F.insert(F.end(), ghostY.begin(), ghostY.end());
    }
    const int N = (int)F.size();
    vec x(N), y(N), rmax(N); std::vector<int> alive(N);
    for (int i = 0; i < N; ++i) { x[i] = F[i].Xi; y[i] = F[i].Yi; rmax[i] = F[i].rmax; alive[i] = F[i].alive; }

    // ---- broad phase :76-120 (strictly sequential in i because of `mems`)
    // broad_mode 0: the literal O(N^2) double loop.  broad_mode 1: the same predicate evaluated only on
    // candidates from a uniform cell grid (cell >= 2*max(rmax)), partners kept in ascending j -- identical
    // output, used for N beyond ~3e4 (tests/ check the two modes against each other).
    double cell = 0, gx0 = 0, gy0 = 0; int gnx = 0, gny = 0; std::vector<int> cell_start, cell_items;
    if (broad_mode == 1) {
        double rm = 0; for (int i = 0; i < N; ++i) if (rmax[i] > rm) rm = rmax[i];
        cell = 2 * rm; if (!(cell > 0)) cell = 1;
        double xmn = INF, xmx = -INF, ymn = INF, ymx = -INF;
        for (int i = 0; i < N; ++i) if (!std::isnan(x[i]) && !std::isnan(y[i])) { xmn = std::min(xmn, x[i]); xmx = std::max(xmx, x[i]); ymn = std::min(ymn, y[i]); ymx = std::max(ymx, y[i]); }
        gx0 = xmn; gy0 = ymn; gnx = (int)((xmx - xmn) / cell) + 1; gny = (int)((ymx - ymn) / cell) + 1;
        std::vector<int> cnt(gnx * gny + 1, 0), cid(N, -1);
        for (int i = 0; i < N; ++i) if (!std::isnan(x[i]) && !std::isnan(y[i])) { cid[i] = (int)((y[i] - gy0) / cell) * gnx + (int)((x[i] - gx0) / cell); cnt[cid[i] + 1]++; }
        for (int c = 0; c < gnx * gny; ++c) cnt[c + 1] += cnt[c];
        cell_start = cnt; cell_items.resize(N); std::vector<int> pos(cnt.begin(), cnt.end() - 1);
        for (int i = 0; i < N; ++i) if (cid[i] >= 0) cell_items[pos[cid[i]]++] = i;
    }
    const double minL2 = std::min(2 * Lx, 2 * Ly);
    for (int i = Nb; i < N; ++i) {                                                    // i = 1+Nb:N (0-based here)
        Floe& fi = F[i];
        fi.interactions.clear(); fi.OverlapArea = 0; fi.potentialInteractions.clear(); fi.cf[0] = fi.cf[1] = 0; fi.ct = 0;
        std::vector<int> mems;                                                        // :93-99
        if (FloeNums[i] < 0) {
            int num = std::abs(FloeNums[i]);
            for (auto& pi : F[num - 1].potentialInteractions) if (pi.floeNum > Nb) mems.push_back((int)pi.floeNum);      // (opt-in topography partners take no part in the de-dup)
        }
        if (!(alive[i] && !std::isnan(x[i]) && P.collision)) continue;               // :101
        std::vector<int> cand;
        const bool opt_in = P.pair_with_boundary_floes != 0;      // NOT the reference: topography floes j <= Nb as partners of i > Nb
        if (broad_mode == 0) { if (opt_in) for (int j = 0; j < Nb && j < i; ++j) cand.push_back(j); for (int j = i + 1; j < N; ++j) cand.push_back(j); }
        else if (!std::isnan(y[i])) {
            int cxi = (int)((x[i] - gx0) / cell), cyi = (int)((y[i] - gy0) / cell);
            for (int dy = -1; dy <= 1; ++dy) for (int dx = -1; dx <= 1; ++dx) {
                int cx = cxi + dx, cy = cyi + dy; if (cx < 0 || cy < 0 || cx >= gnx || cy >= gny) continue;
                int c = cy * gnx + cx;
                for (int t = cell_start[c]; t < cell_start[c + 1]; ++t) if (cell_items[t] > i || (opt_in && cell_items[t] < Nb)) cand.push_back(cell_items[t]);
            }
            std::sort(cand.begin(), cand.end());
        }
        for (int j : cand) {                                                          // :102-117
            if (!alive[j]) continue;
            double ddx = x[i] - x[j], ddy = y[i] - y[j];
            if (!(std::sqrt(ddx * ddx + ddy * ddy) < (rmax[i] + rmax[j]))) continue;
            bool member = std::find(mems.begin(), mems.end(), std::abs(FloeNums[j])) != mems.end();
            if (!(!member || 2 * (rmax[i] + rmax[j]) > minL2)) continue;
            Partner p; p.floeNum = j + 1; p.is_boundary = false;
            for (size_t v = 0; v < F[j].c_alpha.x.size(); ++v) { p.c.x.push_back(F[j].c_alpha.x[v] + x[j]); p.c.y.push_back(F[j].c_alpha.y[v] + y[j]); }
            p.Ui = F[j].Ui; p.Vi = F[j].Vi; p.h = F[j].h; p.area = F[j].area; p.Xi = x[j]; p.Yi = y[j]; p.ksi = F[j].ksi;
            fi.potentialInteractions.push_back(p);
            if (j >= Nb) mems.push_back(FloeNums[j]);                                 // :113 (signed)
        }
    }

    // pair bookkeeping for parity output (ascending (i,j))
    std::vector<size_t> pair_base(N + 1, 0);
    for (int i = 0; i < N; ++i) pair_base[i + 1] = pair_base[i] + F[i].potentialInteractions.size();
    res.pairs.resize(pair_base[N]);

    Curve c2_boundary;
    if (bnd && bnd->box_n > 0) { c2_boundary.x.assign(bnd->box_x, bnd->box_x + bnd->box_n); c2_boundary.y.assign(bnd->box_y, bnd->box_y + bnd->box_n); }
    Partner floebound; floebound.is_boundary = true; floebound.floeNum = INF;
    if (bnd) { floebound.c.x.assign(bnd->x, bnd->x + bnd->n); floebound.c.y.assign(bnd->y, bnd->y + bnd->n);
               floebound.area = bnd->area; floebound.h = bnd->h; floebound.Xi = bnd->xi; floebound.Yi = bnd->yi; floebound.Ui = bnd->u; floebound.Vi = bnd->v; floebound.ksi = bnd->ksi; }

    // ---- pair loop :123-174 (parfor over i)
    std::vector<int> kill(N0, 0), transfer(N0, 0);
    std::vector<int> kill_i(N, 0), transfer_i(N, 0);          // per-i results (i may be a ghost: only i<=N0 entries are used)
    std::vector<int> alive_out(N); for (int i = 0; i < N; ++i) alive_out[i] = F[i].alive;
    std::atomic<int> next(Nb), nfail(0);
    std::atomic<long> npf(0);
    auto worker = [&]() {
        ForceOut fo;
        for (;;) {
            int i = next.fetch_add(1); if (i >= N) break;
            Floe& fi = F[i];
            for (size_t k = 0; k < fi.potentialInteractions.size(); ++k) {
                const Partner& pk = fi.potentialInteractions[k];
                PairDebug& pd = res.pairs[pair_base[i] + k];
                pd.i = i + 1; pd.j = (int)pk.floeNum; pd.status = 0; pd.overlap_state = 0; pd.n_regions = 0;
                try { floe_interactions(fi, pk, P, c2_boundary, fo, P.want_clip_polys ? &pd.clip1 : nullptr); }
                catch (ClipperError&) { pd.status = SZ_ERR_CLIPPER; nfail++; continue; }
                catch (PolyDistError&) { pd.status = SZ_ERR_ARG; nfail++; continue; }   // reference: error() aborts the parfor
                double sabs = 0; for (auto& f : fo.force) sabs += std::fabs(f[0]) + std::fabs(f[1]);
                if (fo.overlap_is_scalar) pd.overlap_state = fo.overlap[0];
                if (sabs != 0) {                                                       // :135-137
                    double so = 0;
                    for (size_t r = 0; r < fo.force.size(); ++r) {
                        fi.interactions.push_back({pk.floeNum, fo.force[r][0], fo.force[r][1], fo.pcontact[r][0], fo.pcontact[r][1], 0.0, fo.overlap[r]});
                        so += fo.overlap[r];
                    }
                    fi.OverlapArea = so + fi.OverlapArea;
                    pd.n_regions = (int)fo.force.size(); npf++;
                } else if (fo.overlap_is_scalar && std::isinf(fo.overlap[0]) && i + 1 > Nb && pk.floeNum > Nb) {   // :138-145 (an opt-in topography partner raises no kill / transfer)
                    if (i + 1 <= N0 && fo.overlap[0] > 0) { kill_i[i] = i + 1; transfer_i[i] = (int)pk.floeNum; }
                    else if (pk.floeNum <= N0) kill_i[i] = (int)pk.floeNum;
                }
            }
            if (!P.periodic && bnd) {                                                  // :150-172
                try {
                    floe_interactions(fi, floebound, P, c2_boundary, fo, nullptr);
                    vec qx(1, x[i]), qy(1, y[i]); std::vector<char> in;
                    inpolygon(qx, qy, c2_boundary.x, c2_boundary.y, in);              // :152
                    if (!in[0]) alive_out[i] = 0;
                    double sabs = 0; for (auto& f : fo.force) sabs += std::fabs(f[0]) + std::fabs(f[1]);
                    if (sabs != 0) {
                        double so = 0;
                        for (size_t r = 0; r < fo.force.size(); ++r) {
                            double fbx = fo.force[r][0], fby = fo.force[r][1];
                            if (std::fabs(fo.pcontact[r][1]) == Ly) fbx = 0;           // :160-162
                            if (std::fabs(fo.pcontact[r][0]) == Lx) fby = 0;           // :163-165
                            fi.interactions.push_back({INF, fbx, fby, fo.pcontact[r][0], fo.pcontact[r][1], 0.0, fo.overlap[r]});
                            so += fo.overlap[r];
                        }
                        fi.OverlapArea = so + fi.OverlapArea;
                    }
                } catch (std::runtime_error&) { nfail++; }
            }
        }
    };
    { std::vector<std::thread> th; for (int t = 1; t < nthreads; ++t) th.emplace_back(worker); worker(); for (auto& t : th) t.join(); }
    res.n_clipper_fail = nfail; res.n_pairs_force = npf;
    // kill(i)=..., transfer(i)=... are written at index i of 1xN0 arrays; MATLAB would grow the array for a ghost i>N0
    // with kill(i)=floeNum.  Entries beyond N0 are never read back as parents (:175-179 loops over length(kill)):
    // we keep the first N0 entries plus the growth, as MATLAB does.
    std::vector<int> killv(kill_i), transv(transfer_i);
    { int len = N0; for (int i = N0; i < N; ++i) if (kill_i[i] != 0) len = i + 1; killv.resize(len); transv.resize(std::max(len, N0)); }
    for (int i = 0; i < (int)killv.size(); ++i)                                       // :175-179
        if (std::abs(killv[i] - (i + 1)) > 0 && killv[i] > 0) { if ((int)transv.size() < killv[i]) transv.resize(killv[i], 0); transv[killv[i] - 1] = i + 1; }
    for (int i = 0; i < N0; ++i) { kill[i] = killv[i]; transfer[i] = i < (int)transv.size() ? transv[i] : 0; }

    // ---- mirror :187-214 (serial)
    for (int i = 0; i < N; ++i) {
        if (F[i].interactions.empty()) continue;
        std::vector<Row> a = F[i].interactions;                                       // snapshot
        for (auto& r : a) {
            if (r[0] <= N && r[0] > i + 1) {
                Floe& fj = F[(int)r[0] - 1];
                fj.interactions.push_back({(double)(i + 1), -r[1], -r[2], r[3], r[4], 0.0, r[6]});
                fj.OverlapArea = fj.OverlapArea + r[6];
            }
        }
    }
    // ---- torques: ghosts first, folded into parents :218-246
    auto torque_and_sums = [&](int i, bool add_existing) {
        Floe& f = F[i];
        if (f.interactions.empty()) return;
        double sfx = 0, sfy = 0, st = 0;
        for (auto& r : f.interactions) {
            double rx = r[3] - x[i], ry = r[4] - y[i];
            r[5] = rx * r[2] - ry * r[1];                                             // cross([P-r 0],[F 0])(3)
            sfx += r[1]; sfy += r[2]; st += r[5];
        }
        if (add_existing) { f.cf[0] = sfx + f.cf[0]; f.cf[1] = sfy + f.cf[1]; f.ct = st + f.ct; }   // :262-263
        else { f.cf[0] = sfx; f.cf[1] = sfy; f.ct = st; }                                            // :234-235
    };
    if (P.periodic) {
        for (int i = N0; i < N; ++i) torque_and_sums(i, false);
        for (size_t k = 0; k < parent.size(); ++k) {                                  // :242-245, serial and in order
            Floe& p = F[parent[k] - 1]; Floe& g = F[N0 + k];
            p.cf[0] = p.cf[0] + g.cf[0]; p.cf[1] = p.cf[1] + g.cf[1]; p.ct = p.ct + g.ct;
        }
    }
    res.xi.resize(N0); res.yi.resize(N0); res.stress.assign((size_t)N0 * 4, 0.0);
    for (int i = 0; i < N0; ++i) { res.xi[i] = F[i].Xi; res.yi[i] = F[i].Yi; }
    for (int i = Nb; i < N0; ++i) {                                                   // :249-285
        torque_and_sums(i, true);
        if (P.periodic) {                                                             // :267-277
            if (std::fabs(F[i].Xi) > Lx) F[i].Xi = F[i].Xi - 2 * Lx * sgn(F[i].Xi);
            if (std::fabs(F[i].Yi) > Ly) F[i].Yi = F[i].Yi - 2 * Ly * sgn(F[i].Yi);
        }
        res.xi[i] = F[i].Xi; res.yi[i] = F[i].Yi;
        if (alive_out[i] && !F[i].interactions.empty()) {                             // calc_trajectory.m:9-13 (r = post-wrap centroid)
            double rx = F[i].Xi, ry = F[i].Yi;
            double s11 = 0, s12 = 0, s21 = 0, s22 = 0, t11 = 0, t12 = 0, t21 = 0, t22 = 0;
            for (auto& r : F[i].interactions) {
                s11 += (r[3] - rx) * r[1]; s12 += (r[4] - ry) * r[1]; s21 += (r[3] - rx) * r[2]; s22 += (r[4] - ry) * r[2];
                t11 += r[1] * (r[3] - rx); t12 += r[2] * (r[3] - rx); t21 += r[1] * (r[4] - ry); t22 += r[2] * (r[4] - ry);
            }
            double k = 1 / (2 * F[i].area * F[i].h);
            res.stress[(size_t)i * 4 + 0] = k * (s11 + t11); res.stress[(size_t)i * 4 + 1] = k * (s12 + t12);
            res.stress[(size_t)i * 4 + 2] = k * (s21 + t21); res.stress[(size_t)i * 4 + 3] = k * (s22 + t22);
        }
    }

    // ---- pack results
    res.n0 = N0; res.n = N;
    res.parent = parent; res.floe_num.assign(FloeNums.begin() + N0, FloeNums.end());
    res.gx.resize(N - N0); res.gy.resize(N - N0);
    for (int i = N0; i < N; ++i) { res.gx[i - N0] = F[i].Xi; res.gy[i - N0] = F[i].Yi; }
    res.row_off.assign(N + 1, 0);
    for (int i = 0; i < N; ++i) res.row_off[i + 1] = res.row_off[i] + (int64_t)F[i].interactions.size();
    res.rows.resize((size_t)res.row_off[N] * 7);
    for (int i = 0; i < N; ++i) for (size_t r = 0; r < F[i].interactions.size(); ++r)
        std::memcpy(&res.rows[((size_t)res.row_off[i] + r) * 7], F[i].interactions[r].data(), 7 * sizeof(double));
    res.fx.resize(N0); res.fy.resize(N0); res.torque.resize(N0); res.overlap_area.resize(N0); res.alive.resize(N0);
    double nfin = 0, ninf = 0;
    for (int i = 0; i < N0; ++i) {
        res.fx[i] = F[i].cf[0]; res.fy[i] = F[i].cf[1]; res.torque[i] = F[i].ct; res.overlap_area[i] = F[i].OverlapArea; res.alive[i] = (uint8_t)alive_out[i];
        for (auto& r : F[i].interactions) { if (r[0] < INF) nfin += 1; else if (r[0] == INF) ninf += 1; }
    }
    res.collision_count = nfin / 2 + ninf;                                            // calc_collisionNum.m:6
    res.kill = kill; res.transfer = transfer;
}

}  // namespace

// ================================================================================================ C interface
extern "C" {

SzoResult* szo_contact_step(const SzParams* prm, const SzFloesSoA* floes, const SzBoundary* bnd, int nthreads, int broad_mode)
{
    SzoResult* r = new SzoResult;
    try { floe_interactions_all(*prm, *floes, bnd, nthreads < 1 ? 1 : nthreads, broad_mode, *r); }
    catch (std::exception& e) { r->error = e.what(); }
    return r;
}
void szo_free(SzoResult* r) { delete r; }
const char* szo_error(const SzoResult* r) { return r->error.c_str(); }
void szo_summary(const SzoResult* r, SzSummary* s)
{
    std::memset(s, 0, sizeof(*s));
    s->n0 = r->n0; s->n = r->n; s->n_pairs = (int64_t)r->pairs.size(); s->n_pairs_force = r->n_pairs_force;
    s->n_rows = r->row_off.empty() ? 0 : r->row_off.back(); s->collision_count = r->collision_count; s->n_clipper_fail = r->n_clipper_fail;
    int64_t np = 0, nv = 0; for (auto& p : r->pairs) { np += (int64_t)p.clip1.size(); for (auto& c : p.clip1) nv += (int64_t)c.x.size(); }
    s->n_clip_paths = np; s->n_clip_verts = nv;
}
void szo_get_floe_outputs(const SzoResult* r, double* fx, double* fy, double* torque, double* overlap_area, double* stress,
                          double* xi, double* yi, uint8_t* alive, int32_t* kill, int32_t* transfer)
{
    const size_t n = (size_t)r->n0;
    if (fx) std::memcpy(fx, r->fx.data(), n * 8); if (fy) std::memcpy(fy, r->fy.data(), n * 8);
    if (torque) std::memcpy(torque, r->torque.data(), n * 8); if (overlap_area) std::memcpy(overlap_area, r->overlap_area.data(), n * 8);
    if (stress) std::memcpy(stress, r->stress.data(), n * 32);
    if (xi) std::memcpy(xi, r->xi.data(), n * 8); if (yi) std::memcpy(yi, r->yi.data(), n * 8);
    if (alive) std::memcpy(alive, r->alive.data(), n);
    if (kill) std::memcpy(kill, r->kill.data(), n * 4); if (transfer) std::memcpy(transfer, r->transfer.data(), n * 4);
}
void szo_get_ghosts(const SzoResult* r, int32_t* parent, int32_t* floe_num, double* gx, double* gy)
{
    const size_t g = (size_t)(r->n - r->n0);
    if (parent) std::memcpy(parent, r->parent.data(), g * 4); if (floe_num) std::memcpy(floe_num, r->floe_num.data(), g * 4);
    if (gx) std::memcpy(gx, r->gx.data(), g * 8); if (gy) std::memcpy(gy, r->gy.data(), g * 8);
}
void szo_get_pairs(const SzoResult* r, int32_t* pi, int32_t* pj, double* overlap_state, int32_t* n_regions, int32_t* status)
{
    for (size_t k = 0; k < r->pairs.size(); ++k) {
        if (pi) pi[k] = r->pairs[k].i; if (pj) pj[k] = r->pairs[k].j;
        if (overlap_state) overlap_state[k] = r->pairs[k].overlap_state;
        if (n_regions) n_regions[k] = r->pairs[k].n_regions; if (status) status[k] = r->pairs[k].status;
    }
}
void szo_get_rows(const SzoResult* r, int64_t* row_off, double* rows)
{
    if (row_off) std::memcpy(row_off, r->row_off.data(), r->row_off.size() * 8);
    if (rows) std::memcpy(rows, r->rows.data(), r->rows.size() * 8);
}
void szo_get_clip_polys(const SzoResult* r, int64_t* pair_path_off, int64_t* path_vert_off, int64_t* x, int64_t* y)
{
    int64_t np = 0, nv = 0; pair_path_off[0] = 0; path_vert_off[0] = 0;
    for (size_t k = 0; k < r->pairs.size(); ++k) {
        for (auto& c : r->pairs[k].clip1) {
            for (size_t v = 0; v < c.x.size(); ++v) { x[nv] = c.x[v]; y[nv] = c.y[v]; ++nv; }
            path_vert_off[++np] = nv;
        }
        pair_path_off[k + 1] = np;
    }
}

// ---- small entry points used by the tests to pin individual restatements
int szo_polyclip(const double* x1, const double* y1, int n1, const double* x2, const double* y2, int n2, int method,
                 double* ox, double* oy, int cap, int* off, int off_cap)
{
    Curve a, b; a.x.assign(x1, x1 + n1); a.y.assign(y1, y1 + n1); b.x.assign(x2, x2 + n2); b.y.assign(y2, y2 + n2);
    std::vector<Curve> out;
    try { polyclip(a, b, method, out); } catch (ClipperError&) { return -1; }
    if ((int)out.size() + 1 > off_cap) return -2;
    int pos = 0; off[0] = 0;
    for (size_t k = 0; k < out.size(); ++k) {
        if (pos + (int)out[k].x.size() > cap) return -2;
        for (size_t v = 0; v < out[k].x.size(); ++v) { ox[pos] = out[k].x[v]; oy[pos] = out[k].y[v]; ++pos; }
        off[k + 1] = pos;
    }
    return (int)out.size();
}
void szo_polyshape_area_centroid(const double* x, const double* y, int n, double* out3)
{
    vec X(x, x + n), Y(y, y + n); polyshape_area_centroid(X, Y, out3[0], out3[1], out3[2]);
}
double szo_polyarea(const double* x, const double* y, int n) { vec X(x, x + n), Y(y, y + n); return polyarea(X, Y); }
int szo_interx(const double* x1, const double* y1, int n1, const double* x2, const double* y2, int n2, double* out, int cap)
{
    Curve a, b; a.x.assign(x1, x1 + n1); a.y.assign(y1, y1 + n1); b.x.assign(x2, x2 + n2); b.y.assign(y2, y2 + n2);
    std::vector<std::array<double, 2>> P; InterX(a, b, P);
    if ((int)P.size() > cap) return -2;
    for (size_t k = 0; k < P.size(); ++k) { out[2 * k] = P[k][0]; out[2 * k + 1] = P[k][1]; }
    return (int)P.size();
}
void szo_inpolygon(const double* px, const double* py, int np, const double* xv, const double* yv, int nv, uint8_t* in)
{
    vec a(px, px + np), b(py, py + np), c(xv, xv + nv), d(yv, yv + nv); std::vector<char> r;
    inpolygon(a, b, c, d, r); for (int i = 0; i < np; ++i) in[i] = (uint8_t)r[i];
}
int szo_p_poly_dist(const double* px, const double* py, int np, const double* xv, const double* yv, int nv, double* d)
{
    vec a(px, px + np), b(py, py + np), c(xv, xv + nv), e(yv, yv + nv), r;
    try { p_poly_dist(a, b, c, e, r); } catch (PolyDistError&) { return -1; }
    for (int i = 0; i < np; ++i) d[i] = r[i];
    return 0;
}
int64_t szo_matlab_int64(double v) { return matlab_int64(v); }
// one call of collisions/floe_interactions.m.  c1 = floe1.c_alpha (closed) + centroid is formed inside, as in :25.
// body = {h, area, Xi, Yi, Ui, Vi, ksi}.  rows_out[r*5..] = Fx Fy Px Py overlap.  Returns the number of rows the CALLER
// would append (floe_interactions_all.m:135), or <0: -3 Clipper error, -1 p_poly_dist error.
int szo_floe_interactions(const SzParams* prm, const double* cax, const double* cay, int n1, const double* body1,
                          const double* c2x, const double* c2y, int n2, const double* body2, int is_boundary,
                          const double* boxx, const double* boxy, int nbox,
                          double* rows_out, int rows_cap, double* overlap_state)
{
    Floe f1; f1.c_alpha.x.assign(cax, cax + n1); f1.c_alpha.y.assign(cay, cay + n1);
    f1.h = body1[0]; f1.area = body1[1]; f1.Xi = body1[2]; f1.Yi = body1[3]; f1.Ui = body1[4]; f1.Vi = body1[5]; f1.ksi = body1[6];
    Partner p2; p2.c.x.assign(c2x, c2x + n2); p2.c.y.assign(c2y, c2y + n2); p2.is_boundary = is_boundary != 0; p2.floeNum = 2;
    p2.h = body2[0]; p2.area = body2[1]; p2.Xi = body2[2]; p2.Yi = body2[3]; p2.Ui = body2[4]; p2.Vi = body2[5]; p2.ksi = body2[6];
    Curve box; if (nbox > 0) { box.x.assign(boxx, boxx + nbox); box.y.assign(boxy, boxy + nbox); }
    ForceOut fo;
    try { floe_interactions(f1, p2, *prm, box, fo, nullptr); }
    catch (ClipperError&) { return -3; } catch (PolyDistError&) { return -1; }
    *overlap_state = fo.overlap_is_scalar ? fo.overlap[0] : 0.0;
    double sabs = 0; for (auto& f : fo.force) sabs += std::fabs(f[0]) + std::fabs(f[1]);
    if (!(sabs != 0)) return 0;
    int n = (int)fo.force.size(); if (n > rows_cap) return -2;
    for (int r = 0; r < n; ++r) { rows_out[r * 5] = fo.force[r][0]; rows_out[r * 5 + 1] = fo.force[r][1]; rows_out[r * 5 + 2] = fo.pcontact[r][0]; rows_out[r * 5 + 3] = fo.pcontact[r][1]; rows_out[r * 5 + 4] = fo.overlap[r]; }
    return n;
}
int szo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---- calc_trajectory.m, the branch the contact-loop benchmark exercises: doInt.flag = false with
// FxOA/FyOA/torqueOA carried over, no ocean/wind evaluation.  One call = the loop floe_interactions_all.m:279-284 over
// the floes of the input list.  Arrays are per floe; c0/c_alpha pools share voff.  sacked[i] = 1 when the reference
// returns floe = [] (:89,117); unsupported[i] = 1 when the reference would evaluate the ocean (h < 0.1, :94,121).
void szo_calc_trajectory(int n, double dt, double HFo, double xo_min, double xo_max, double yo_min, double yo_max, int nz,
                         const double* cfx, const double* cfy, const double* ctq, const double* stress_now, const uint8_t* has_rows,
                         const double* area, double* x, double* y, double* u, double* v, double* ksi, double* h, uint8_t* alive,
                         double* mass, double* inertia, double* alpha, double* dXi_p, double* dYi_p, double* dUi_p, double* dVi_p,
                         double* dalpha_p, double* dksi_p, const double* FxOA, const double* FyOA, const double* torqueOA,
                         const int32_t* voff, const double* c0x, const double* c0y, double* cax, double* cay,
                         double* stress_h, int32_t* stress_count, double* stress_out, uint8_t* sacked, uint8_t* unsupported)
{
    auto sgn = [](double a) { return (double)((a > 0) - (a < 0)); };
    for (int i = 0; i < n; ++i) {
        sacked[i] = 0; unsupported[i] = 0;
        if (!alive[i]) continue;                                                   // floe_interactions_all.m:280
        double ext_fx = cfx[i], ext_fy = cfy[i], ext_t = ctq[i];                    // :3-4
        // a sacked floe comes back as [] and the caller keeps the OLD struct (floe_interactions_all.m:282): remember it
        const double mass0 = mass[i], inertia0 = inertia[i], h0 = h[i]; const uint8_t alive0 = alive[i]; const int32_t sc0 = stress_count[i];
        // :9-29 stress history ring (StressCount is 1-based)
        if (stress_count[i] > nz) stress_count[i] = 1;
        double* slot = stress_h + ((size_t)i * nz + (stress_count[i] - 1)) * 4;
        double slot0[4]; for (int k = 0; k < 4; ++k) slot0[k] = slot[k];
        for (int k = 0; k < 4; ++k) slot[k] = has_rows[i] ? stress_now[(size_t)i * 4 + k] : 0.0;
        stress_count[i] += 1;
        double st_new[4];
        for (int k = 0; k < 4; ++k) { double sm = 0; for (int z = 0; z < nz; ++z) sm += stress_h[((size_t)i * nz + z) * 4 + k]; st_new[k] = sm / nz; }   // mean(StressH,3)
        // :36-41
        if (h[i] > 10) h[i] = 10; else if (mass[i] < 100) { mass[i] = 1e3; alive[i] = 0; }
        while (std::max(std::fabs(ext_fx), std::fabs(ext_fy)) > mass[i] / (5 * dt)) { ext_fx = ext_fx / 10; ext_fy = ext_fy / 10; ext_t = ext_t / 10; }   // :42-46
        // :67-80 thermodynamic growth
        const double floe_area = area[i];
        double floe_mass = mass[i], hh = h[i], floe_inertia = inertia[i];
        const double dh = HFo * dt / hh;
        floe_mass = (hh - dh) / hh * floe_mass; mass[i] = floe_mass;
        floe_inertia = (hh - dh) / hh * floe_inertia; inertia[i] = floe_inertia;
        h[i] = hh - dh;
        bool sack = std::isnan(x[i]);                                              // :89
        if (!sack && h[i] < 0.1) { unsupported[i] = 1; continue; }                 // :94 would re-evaluate the ocean forcing
        if (!sack) {
            double cmaxx = -INF, cminx = INF, cmaxy = -INF, cminy = INF;
            for (int t = voff[i]; t < voff[i + 1]; ++t) { cmaxx = std::max(cmaxx, cax[t]); cminx = std::min(cminx, cax[t]); cmaxy = std::max(cmaxy, cay[t]); cminy = std::min(cminy, cay[t]); }
            sack = (cmaxx + x[i] > xo_max || cminx + x[i] < xo_min || cmaxy + y[i] > yo_max || cminy + y[i] < yo_min);   // :116-117
        }
        if (sack) {
            sacked[i] = 1; mass[i] = mass0; inertia[i] = inertia0; h[i] = h0; alive[i] = alive0; stress_count[i] = sc0;
            for (int k = 0; k < 4; ++k) slot[k] = slot0[k];
            continue;
        }
        for (int k = 0; k < 4; ++k) stress_out[(size_t)i * 4 + k] = st_new[k];
        if (alive[i] != 1) continue;                                               // :118
        // :174-177 positions, AB2
        const double dx = 1.5 * dt * u[i] - 0.5 * dt * dXi_p[i], dy = 1.5 * dt * v[i] - 0.5 * dt * dYi_p[i];
        x[i] = x[i] + dx; dXi_p[i] = u[i];
        y[i] = y[i] + dy; dYi_p[i] = v[i];
        alpha[i] = alpha[i] + 1.5 * dt * ksi[i] - 0.5 * dt * dalpha_p[i]; dalpha_p[i] = ksi[i];
        // :181-208 velocities with the 0.5 h / dt limiter
        const double ax0 = (FxOA[i] * floe_area + ext_fx), ay0 = (FyOA[i] * floe_area + ext_fy);
        double dU = ax0 / floe_mass, dV = ay0 / floe_mass;
        bool have_frac = false; double frac = 0;
        const double lim = 0.5 * h[i];
        if (std::fabs(dt * dU) > lim && std::fabs(dt * dV) > lim) {
            dU = sgn(dU) * 0.5 * h[i] / dt; dV = sgn(dV) * 0.5 * h[i] / dt;
            const double f1 = dU / ax0 * floe_mass, f2 = dV / ay0 * floe_mass;
            frac = std::min(f1, f2); have_frac = true;
            dU = ax0 / floe_mass; dV = ay0 / floe_mass; dU = frac * dU; dV = frac * dV;
        } else if (std::fabs(dt * dU) > lim && std::fabs(dt * dV) < lim) {
            dU = sgn(dU) * 0.5 * h[i] / dt;
            frac = dU / ax0 * floe_mass; have_frac = true;
            dU = ax0 / floe_mass; dV = ay0 / floe_mass; dU = frac * dU; dV = frac * dV;
        } else if (std::fabs(dt * dU) < lim && std::fabs(dt * dV) > lim) {
            dV = sgn(dV) * 0.5 * h[i] / dt;
            frac = dV / ay0 * floe_mass; have_frac = true;
            dU = ax0 / floe_mass; dV = ay0 / floe_mass; dU = frac * dU; dV = frac * dV;
        }
        u[i] = u[i] + 1.5 * dt * dU - 0.5 * dt * dUi_p[i];
        v[i] = v[i] + 1.5 * dt * dV - 0.5 * dt * dVi_p[i];
        dUi_p[i] = dU; dVi_p[i] = dV;
        // :210-219 angular velocity, clamped to 1e-5
        double dksi = (torqueOA[i] * floe_area + ext_t) / floe_inertia;
        if (have_frac) dksi = frac * dksi;
        double k2 = ksi[i] + 1.5 * dt * dksi - 0.5 * dt * dksi_p[i];
        if (std::fabs(k2) > 1e-5) k2 = sgn(k2) * 1e-5;
        ksi[i] = k2; dksi_p[i] = dksi;
        // :221-222 rotate the outline
        const double ca = std::cos(alpha[i]), sa = std::sin(alpha[i]);
        for (int t = voff[i]; t < voff[i + 1]; ++t) { cax[t] = ca * c0x[t] + (-sa) * c0y[t]; cay[t] = sa * c0x[t] + ca * c0y[t]; }
    }
}

// interp2(X, Y, V, xq, yq) of MATLAB, 'linear' (calc_trajectory.m:135-138): X [nx], Y [ny] ascending grid vectors,
// V(iy, ix) stored column-major like MATLAB (iy + ix*ny); NaN outside the grid.  MATLAB's own kernel is not in the
// reference; this is the textbook bilinear form (parity with MATLAB's rounding unpinned, well inside 1e-9).
static double interp2_linear(const double* X, int nx, const double* Y, int ny, const double* V, double xq, double yq)
{
    if (!(xq >= X[0] && xq <= X[nx - 1] && yq >= Y[0] && yq <= Y[ny - 1])) return INF - INF;
    int ix = (int)(std::upper_bound(X, X + nx, xq) - X) - 1; if (ix > nx - 2) ix = nx - 2; if (ix < 0) ix = 0;
    int iy = (int)(std::upper_bound(Y, Y + ny, yq) - Y) - 1; if (iy > ny - 2) iy = ny - 2; if (iy < 0) iy = 0;
    const double t = (xq - X[ix]) / (X[ix + 1] - X[ix]), sy = (yq - Y[iy]) / (Y[iy + 1] - Y[iy]);
    const double v00 = V[iy + (size_t)ix * ny], v10 = V[iy + (size_t)(ix + 1) * ny], v01 = V[iy + 1 + (size_t)ix * ny], v11 = V[iy + 1 + (size_t)(ix + 1) * ny];
    return (v00 * (1 - t) + v10 * t) * (1 - sy) + (v01 * (1 - t) + v11 * t) * sy;
}

// calc_trajectory.m:94-166: the ocean / atmosphere forcing of every floe the reference would evaluate it for in this
// call -- doInt.flag (do_int), or floe.h < 0.1 after the thermodynamic thinning -- from the state BEFORE the
// Adams-Bashforth update.  It repeats the bounds and thinning of :36-41,67-80 to obtain floe_mass and the sacking tests
// of :89,116-117 (a sacked floe keeps its old FxOA).  x, y, alive are the contact step's per-floe outputs (wrapped
// centroid, alive after the wall test), like in szo_calc_trajectory.  PX, PY, PA [n * npts]: Floe.X, Floe.Y, Floe.A
// (Monte-Carlo points of initialize_floe_values.m:31-33, body frame).  no_points[i] = 1 when sum(A) == 0: the reference
// would draw new random points (:100-111), which cannot be reproduced; FxOA is left untouched for that floe.
void szo_ocean_forcing(int n, double dt, double HFo, double xo_min, double xo_max, double yo_min, double yo_max, int do_int,
                       const uint8_t* alive, const double* x, const double* y, const double* u, const double* v, const double* ksi,
                       const double* h, const double* mass, const double* area, const double* alpha,
                       const int32_t* voff, const double* cax, const double* cay,
                       int npts, const double* PX, const double* PY, const uint8_t* PA,
                       int nx, int ny, const double* Xo, const double* Yo, const double* Uocn, const double* Vocn, const double* Uwinds, const double* Vwinds,
                       double fc, double turn_angle, double rho0, double Cd, double rho_air, double Cd_atm,
                       double* FxOA, double* FyOA, double* torqueOA, uint8_t* evaluated, uint8_t* no_points)
{
    for (int i = 0; i < n; ++i) {
        evaluated[i] = 0; no_points[i] = 0;
        if (!alive[i]) continue;                                                   // floe_interactions_all.m:280
        double hh = h[i], m = mass[i]; uint8_t al = alive[i];
        if (hh > 10) hh = 10; else if (m < 100) { m = 1e3; al = 0; }               // :36-41
        const double dh = HFo * dt / hh;                                           // :75-79
        const double floe_mass = (hh - dh) / hh * m, h_new = hh - dh, floe_area = area[i];
        if (std::isnan(x[i])) continue;                                            // :89
        if (!(do_int || h_new < 0.1)) continue;                                    // :94
        double cmaxx = -INF, cminx = INF, cmaxy = -INF, cminy = INF;
        for (int t = voff[i]; t < voff[i + 1]; ++t) { cmaxx = std::max(cmaxx, cax[t]); cminx = std::min(cminx, cax[t]); cmaxy = std::max(cmaxy, cay[t]); cminy = std::min(cminy, cay[t]); }
        if (cmaxx + x[i] > xo_max || cminx + x[i] < xo_min || cmaxy + y[i] > yo_max || cminy + y[i] < yo_min) continue;   // :116-117
        if (al != 1) continue;                                                     // :118
        const double* px = PX + (size_t)i * npts; const double* py = PY + (size_t)i * npts; const uint8_t* pa = PA + (size_t)i * npts;
        int cnt = 0; for (int k = 0; k < npts; ++k) cnt += pa[k] != 0;
        if (cnt == 0) { no_points[i] = 1; continue; }                              // :100-111 draws new random points
        const double ca = std::cos(alpha[i]), sa = std::sin(alpha[i]);
        const double Xi = x[i], Yi = y[i], Ui = u[i], Vi = v[i], K = ksi[i];
        // winds averaged over the floe (:140)
        double su = 0, sv = 0;
        for (int k = 0; k < npts; ++k) if (pa[k]) {
            const double xr = ca * px[k] + (-sa) * py[k], yr = sa * px[k] + ca * py[k];                                  // :97
            su += interp2_linear(Xo, nx, Yo, ny, Uwinds, xr + Xi, yr + Yi); sv += interp2_linear(Xo, nx, Yo, ny, Vwinds, xr + Xi, yr + Yi);
        }
        const double U10 = su / cnt, V10 = sv / cnt;
        const double Fx_atm = rho_air * Cd_atm * std::sqrt(U10 * U10 + V10 * V10) * U10, Fy_atm = rho_air * Cd_atm * std::sqrt(U10 * U10 + V10 * V10) * V10;   // :141-142
        const double mfa = floe_mass / floe_area;
        double sfx = 0, sfy = 0, stq = 0;
        for (int k = 0; k < npts; ++k) if (pa[k]) {
            const double xr = ca * px[k] + (-sa) * py[k], yr = sa * px[k] + ca * py[k];
            const double theta = std::atan2(yr, xr), rho = std::hypot(xr, yr);                                           // cart2pol :124
            const double Uice = Ui - rho * K * std::sin(theta), Vice = Vi + rho * K * std::cos(theta);                   // :127-128
            const double uo = interp2_linear(Xo, nx, Yo, ny, Uocn, xr + Xi, yr + Yi), vo = interp2_linear(Xo, nx, Yo, ny, Vocn, xr + Xi, yr + Yi);
            const double fxp = -mfa * fc * vo, fyp = +mfa * fc * uo;                                                     // :144-145
            const double du = uo - Uice, dv = vo - Vice;                                                                 // :147
            const double sp = std::sqrt(du * du + dv * dv);
            const double tx = rho0 * Cd * sp * (std::cos(turn_angle) * du - std::sin(turn_angle) * dv);                  // :149-150
            const double ty = rho0 * Cd * sp * (std::sin(turn_angle) * du + std::cos(turn_angle) * dv);
            double Fx = tx + Fx_atm + fxp, Fy = ty + Fy_atm + fyp;                                                       // :152-153
            const double tq = (-Fx * std::sin(theta) + Fy * std::cos(theta)) * rho;                                      // :157
            Fx = Fx + mfa * fc * Vi; Fy = Fy - mfa * fc * Ui;                                                            // :160-161
            sfx += Fx; sfy += Fy; stq += tq;
        }
        FxOA[i] = sfx / cnt; FyOA[i] = sfy / cnt; torqueOA[i] = stq / cnt;                                               // :164-166
        evaluated[i] = 1;
    }
}

// calc_trajectory.m:224-234 (doInt.flag): strain rate of every floe that went through the update branch (alive == 1, not
// sacked) from its UPDATED outline and velocities; the others keep their old floe.strain.
void szo_floe_strain(int n, const uint8_t* alive, const uint8_t* sacked, const double* area, const double* u, const double* v, const double* ksi,
                     const int32_t* voff, const double* cax, const double* cay, double* strain)
{
    for (int i = 0; i < n; ++i) {
        if (alive[i] != 1 || sacked[i]) continue;
        const int o = voff[i], m = voff[i + 1] - o;
        if (m < 1) continue;
        double sxu = 0, syu = 0, sxv = 0, syv = 0;      // sum(diff([Uice Uice(1)]) .* diff([c_alpha(2,:) c_alpha(2,1)])) ...
        auto vel = [&](int t, double& U, double& V) {
            const double theta = std::atan2(cay[o + t], cax[o + t]), rho = std::hypot(cax[o + t], cay[o + t]);
            U = u[i] - rho * ksi[i] * std::sin(theta); V = v[i] + rho * ksi[i] * std::cos(theta);
        };
        for (int t = 0; t < m; ++t) {
            const int t1 = (t + 1 == m) ? 0 : t + 1;
            double U0, V0, U1, V1; vel(t, U0, V0); vel(t1, U1, V1);
            const double dxc = cax[o + t1] - cax[o + t], dyc = cay[o + t1] - cay[o + t];
            sxu += (U1 - U0) * dyc; syu += (U1 - U0) * dxc; sxv += (V1 - V0) * dyc; syv += (V1 - V0) * dxc;
        }
        const double du_dx = 0.5 * sxu / area[i], du_dy = 0.5 * syu / area[i], dv_dx = 0.5 * sxv / area[i], dv_dy = 0.5 * syv / area[i];
        // 1/2*([du_dx du_dy; dv_dx dv_dy] + [du_dx dv_dx; du_dy dv_dy]), stored row-major like the stress
        strain[(size_t)i * 4 + 0] = 0.5 * (du_dx + du_dx); strain[(size_t)i * 4 + 1] = 0.5 * (du_dy + dv_dx);
        strain[(size_t)i * 4 + 2] = 0.5 * (dv_dx + du_dy); strain[(size_t)i * 4 + 3] = 0.5 * (dv_dy + dv_dy);
    }
}

// Physical_Processes/fracture_floe.m:12-52 -- the "slight permanent deformation" a floe receives before it is fractured,
// the deterministic consumer of the contact rows (the Voronoi split that follows draws random points).  For every
// selected floe (idx, 1-based position in the floe list the contact step ran on = Floe0): the row with the largest
// overlap among its floe-floe rows (:17-22); if the partner is an original floe (:26), clip the two outlines ('int', :29),
// take the first region, its polyshape centroid and the distance from the centroid to the region's outline
// (p_poly_dist, :34-35), push the partner by half that distance along the contact force (:36-39), subtract it from the
// floe ('dif', :40) and, if more than 90 % of the floe's area is left, make that region the floe's new outline about its
// new centroid (:41-50).  Outputs per selected floe: changed, Xi, Yi, area and, for changed floes, the new c_alpha
// (open ring in Clipper's order) in the pools (vert_off [count + 1]).  Returns the vertices needed, or -1 on "Clipper Error.",
// -3 when p_poly_dist would raise.
int szo_fracture_deform(const SzFloesSoA* f, const int64_t* row_off, const double* rows, int count, const int32_t* idx,
                        uint8_t* changed, double* xi, double* yi, double* area, int64_t* vert_off, double* cx, double* cy, int64_t vcap)
{
    const int N0 = f->n;
    int64_t pos = 0; vert_off[0] = 0;
    auto world = [&](int i, Curve& c) {
        const int o = f->voff[i], n = f->voff[i + 1] - o;
        c.x.resize(n); c.y.resize(n);
        for (int t = 0; t < n; ++t) { c.x[t] = f->vx[o + t] + f->x[i]; c.y[t] = f->vy[o + t] + f->y[i]; }
    };
    for (int q = 0; q < count; ++q) {
        const int i = idx[q] - 1;
        changed[q] = 0; xi[q] = f->x[i]; yi[q] = f->y[i]; area[q] = f->area[i]; vert_off[q + 1] = pos;
        // a = floe.interactions without the wall rows; [~,k] = max(a(:,7)) (first maximum)
        int k = -1; double best = 0;
        for (int64_t r = row_off[i]; r < row_off[i + 1]; ++r) {
            const double* a = rows + r * 7;
            if (std::isinf(a[0])) continue;
            if (k < 0 || a[6] > best) { best = a[6]; k = (int)r; }
        }
        if (k < 0) continue;
        const double* a = rows + (int64_t)k * 7;
        const double clip = a[0];
        if (!(clip < N0 + 1)) continue;                                            // :26 ghost partners are skipped
        const int j = (int)clip - 1;
        Curve c1, c2; world(i, c1); world(j, c2);
        std::vector<Curve> out;
        try { polyclip(c1, c2, 1, out); } catch (ClipperError&) { return -1; }    // :29
        if (out.empty()) continue;                                                 // :32
        double ar, xm, ym;
        polyshape_area_centroid(out[0].x, out[0].y, ar, xm, ym);                   // :34
        vec d;
        try { p_poly_dist(vec{xm}, vec{ym}, out[0].x, out[0].y, d); } catch (PolyDistError&) { return -3; }   // :35
        const double F = std::sqrt(a[1] * a[1] + a[2] * a[2]);                     // vecnorm(a(k,2:3))
        const double xs = a[1] * std::fabs(d[0]) / 2 / F, ys = a[2] * std::fabs(d[0]) / 2 / F;   // :37-38
        for (size_t t = 0; t < c2.x.size(); ++t) { c2.x[t] = c2.x[t] + xs; c2.y[t] = c2.y[t] + ys; }   // :39
        try { polyclip(c1, c2, 0, out); } catch (ClipperError&) { return -1; }    // :40
        if (out.empty()) continue;
        const double Anew = polyarea(out[0].x, out[0].y);                          // :43
        if (!(Anew / f->area[i] > 0.9)) continue;                                  // :44
        polyshape_area_centroid(out[0].x, out[0].y, ar, xm, ym);                   // :45
        changed[q] = 1; xi[q] = xm; yi[q] = ym; area[q] = Anew;                    // :46-47
        for (size_t t = 0; t < out[0].x.size(); ++t) {
            if (pos < vcap) { cx[pos] = out[0].x[t] - xm; cy[pos] = out[0].y[t] - ym; }   // :48
            ++pos;
        }
        vert_off[q + 1] = pos;
    }
    return (int)pos;
}

// calc_eulerian_data.m:1-192: mass-weighted coarse-grid averages of the floe state.  Restated as
// written, including its quirks: dead floes are dropped first (:7-8); boundary floes (Nb > 0, :11-25) are not supported
// here (returns -1); for PERIODIC domains an x-ghost is appended for every floe with a vertex beyond +-Lx (:39-48), and
// the y pass (:56-65) tests the polygon left over from the LAST iteration of the x loop for every floe, so either every
// floe of the list (x-ghosts included) gets a y-ghost or none does; grid rows run from ymax down (:72 fliplr); a cell is
// processed when the masses of its candidate floes (centre distance < rmax + cell half-diagonal, :113-119) sum to > 0;
// Aover = area(intersect(box, poly)) -- here Clipper's intersection + the polyshape area of every returned region,
// geometrically the same set (MATLAB's own polygon kernel is not in the reference: parity unpinned, ~1e-12 relative);
// sums run over ascending floe index.  Per-floe inputs: [n]; stress, strain: [n][4] row-major.  Outputs: 18 arrays of
// Ny*Nx doubles, element (jj, ii) at jj*Nx + ii (jj = 0 is the TOP row), in the order
//   u v du dv stress stressxx stressyx stressxy stressyy strainux strainvx strainuy strainvy c Over Mtot area h
int szo_calc_eulerian_data(const SzFloesSoA* f, const double* mass, const double* overlap_area, const double* dUi_p, const double* dVi_p,
                           const double* stress, const double* strain, int Nx, int Ny, int Nb,
                           double xmin, double xmax, double ymin, double ymax, int periodic, double* out)
{
    if (Nb != 0) return -1;
    struct F { Curve c; double Xi, Yi, rmax, mass, area, over, U, V, H, dU, dV, S[4], E[4]; };
    auto nz = [](double v) { return std::isnan(v) ? 0.0 : v; };
    std::vector<F> fl;
    for (int i = 0; i < f->n; ++i) {
        if (!f->alive[i]) continue;                                                // :7-8
        F g; const int o = f->voff[i], m = f->voff[i + 1] - o;
        g.c.x.assign(f->vx + o, f->vx + o + m); g.c.y.assign(f->vy + o, f->vy + o + m);
        g.Xi = f->x[i]; g.Yi = f->y[i]; g.rmax = f->rmax[i]; g.mass = nz(mass[i]); g.area = nz(f->area[i]); g.over = overlap_area[i];
        g.U = nz(f->u[i]); g.V = nz(f->v[i]); g.H = nz(f->h[i]); g.dU = nz(dUi_p[i]); g.dV = nz(dVi_p[i]);       // :100-111
        for (int k = 0; k < 4; ++k) { g.S[k] = stress[(size_t)i * 4 + k]; g.E[k] = strain[(size_t)i * 4 + k]; }
        fl.push_back(g);
    }
    const double Lx = xmax, Ly = ymax;                                             // :28-29 max(c2_boundary)
    auto sgn = [](double a) { return (double)((a > 0) - (a < 0)); };
    if (periodic && !fl.empty()) {
        const size_t n1 = fl.size();
        double last_max_abs_y = 0;
        for (size_t i = 0; i < n1; ++i) {                                          // :39-48
            double mx = 0, my = 0;
            for (size_t t = 0; t < fl[i].c.x.size(); ++t) { mx = std::max(mx, std::fabs(fl[i].c.x[t] + fl[i].Xi)); my = std::max(my, std::fabs(fl[i].c.y[t] + fl[i].Yi)); }
            last_max_abs_y = my;                                                   // `poly` keeps the last floe's polygon
            if (mx > Lx) { F g = fl[i]; g.Xi = fl[i].Xi - 2 * Lx * sgn(fl[i].Xi); fl.push_back(g); }
        }
        const size_t n2 = fl.size();
        if (last_max_abs_y > Ly)                                                   // :56-65 the stale-polygon test
            for (size_t i = 0; i < n2; ++i) { F g = fl[i]; g.Yi = fl[i].Yi - 2 * Ly * sgn(fl[i].Yi); fl.push_back(g); }
    }
    // :70-80 grid (colon: a + k*d), rows from the top
    std::vector<double> xe(Nx + 1), ye(Ny + 1);
    for (int k = 0; k <= Nx; ++k) xe[k] = xmin + k * ((xmax - xmin) / Nx);
    for (int k = 0; k <= Ny; ++k) ye[k] = ymin + k * ((ymax - ymin) / Ny);
    xe[Nx] = xmax; ye[Ny] = ymax;
    std::reverse(ye.begin(), ye.end());
    const double dx = std::fabs(xe[1] - xe[0]), dy = std::fabs(ye[1] - ye[0]);
    const double r_max = std::sqrt((dx / 2) * (dx / 2) + (dy / 2) * (dy / 2));
    const size_t cells = (size_t)Nx * Ny;
    for (size_t k = 0; k < 18 * cells; ++k) out[k] = 0;
    enum { O_U, O_V, O_DU, O_DV, O_STRESS, O_SXX, O_SYX, O_SXY, O_SYY, O_EUX, O_EVX, O_EUY, O_EVY, O_C, O_OVER, O_MTOT, O_AREA, O_H };
    for (int ii = 0; ii < Nx; ++ii) for (int jj = 0; jj < Ny; ++jj) {
        const double xc = 0.5 * (xe[ii] + xe[ii + 1]), yc = 0.5 * (ye[jj] + ye[jj + 1]);
        std::vector<int> live; double M0 = 0;
        for (size_t q = 0; q < fl.size(); ++q) {
            const double pint = std::sqrt((xc - fl[q].Xi) * (xc - fl[q].Xi) + (yc - fl[q].Yi) * (yc - fl[q].Yi)) - (fl[q].rmax + r_max);
            if (pint < 0) { live.push_back((int)q); M0 += fl[q].mass; }
        }
        if (!(M0 > 0)) continue;                                                   // :131
        Curve box; box.x = {xe[ii], xe[ii], xe[ii + 1], xe[ii + 1], xe[ii]}; box.y = {ye[jj], ye[jj + 1], ye[jj + 1], ye[jj], ye[jj]};   // :134
        const double abox = dx * dy;
        std::vector<int> nums; std::vector<double> aover;
        for (int q : live) {
            Curve c; c.x.resize(fl[q].c.x.size()); c.y.resize(c.x.size());
            for (size_t t = 0; t < c.x.size(); ++t) { c.x[t] = fl[q].c.x[t] + fl[q].Xi; c.y[t] = fl[q].c.y[t] + fl[q].Yi; }
            std::vector<Curve> reg;
            try { polyclip(box, c, 1, reg); } catch (ClipperError&) { return -2; }
            double a = 0;
            for (auto& r : reg) { double ar, cx, cy; polyshape_area_centroid(r.x, r.y, ar, cx, cy); a += ar; }
            if (a != 0) { nums.push_back(q); aover.push_back(a); }                 // :146-147
        }
        double Mtot = 0, Atot = 0;
        for (size_t k = 0; k < nums.size(); ++k) { Mtot += fl[nums[k]].mass * aover[k] / fl[nums[k]].area; Atot += aover[k]; }   // :149-150
        const size_t e = (size_t)jj * Nx + ii;
        out[O_C * cells + e] = Atot / abox;                                        // :151
        if (!(Mtot > 0)) continue;
        auto wsum = [&](auto get) { double s2 = 0; for (size_t k = 0; k < nums.size(); ++k) s2 += get(fl[nums[k]]) * fl[nums[k]].mass * aover[k] / fl[nums[k]].area; return s2 / Mtot; };
        double so = 0; for (int q : nums) so += fl[q].over;
        out[O_OVER * cells + e] = so / (double)aover.size();                       // :153
        out[O_MTOT * cells + e] = Mtot; out[O_AREA * cells + e] = Atot;
        out[O_H * cells + e] = wsum([](const F& g) { return g.H; });
        out[O_U * cells + e] = wsum([](const F& g) { return g.U; }); out[O_V * cells + e] = wsum([](const F& g) { return g.V; });
        out[O_DU * cells + e] = wsum([](const F& g) { return g.dU; }); out[O_DV * cells + e] = wsum([](const F& g) { return g.dV; });
        const double sxx = wsum([](const F& g) { return g.S[0]; }), syx = wsum([](const F& g) { return g.S[1]; });   // Stress(1,1), Stress(1,2)
        const double sxy = wsum([](const F& g) { return g.S[2]; }), syy = wsum([](const F& g) { return g.S[3]; });   // Stress(2,1), Stress(2,2)
        out[O_SXX * cells + e] = sxx; out[O_SYX * cells + e] = syx; out[O_SXY * cells + e] = sxy; out[O_SYY * cells + e] = syy;
        out[O_EUX * cells + e] = wsum([](const F& g) { return g.E[0]; }); out[O_EVX * cells + e] = wsum([](const F& g) { return g.E[1]; });
        out[O_EUY * cells + e] = wsum([](const F& g) { return g.E[2]; }); out[O_EVY * cells + e] = wsum([](const F& g) { return g.E[3]; });
        // max(eig([sxx syx; sxy syy])) (:170); the stress tensor is symmetric, its eigenvalues real
        const double tr = sxx + syy, det = sxx * syy - syx * sxy, disc = tr * tr / 4 - det;
        double lam = tr / 2 + std::sqrt(disc > 0 ? disc : 0);
        if (std::fabs(lam) > 1e8) lam = 0;                                         // :171-173
        out[O_STRESS * cells + e] = lam;
    }
    return 0;
}

// Physical_Processes/corners.m:10-99, the deterministic half of the corner-grinding rule: which
// vertices of a floe are "in contact" (da).  The other half, break1 = rand > angle/Anorm (:72), and frac_corner stay with
// the host.  As written: the floe list Floe0 is extended with periodic images whether or not the run is periodic
// (:13-48: x images of floes with a vertex beyond +-Lx, then y images over the extended list); N0 = its length (:51);
// for every selected floe (idx, 1-based positions in the list; the reference skips the first Nb of the SELECTION, :54)
// with contact rows: the vertex nearest to each contact point whose partner number is <= N0 (dsearchn, :74), every
// vertex inside the outline of a partner (inpolygon counts the boundary, :78-82; partners are looked up in the extended
// list rebuilt from the CURRENT positions), and, when the floe touches the wall (an Inf partner), every vertex outside
// c2_boundary (:83-86).  Vertices are those of polyshape(c_alpha'): taken here as c_alpha without its closing duplicate,
// in the stored order (polyshape keeps the order of a clockwise simple outline; it would also drop exactly collinear
// vertices, which these outlines do not have -- parity with polyshape's clean-up unpinned).
// Output: da for the vertices of selected floe q at da[da_off[q] .. da_off[q+1]); returns the total, or -1 if vcap is too small.
int szo_corner_eligibility(const SzFloesSoA* f, const int64_t* row_off, const double* rows, int count, const int32_t* idx, int Nb,
                           double Lx, double Ly, const double* boxx, const double* boxy, int nbox,
                           int64_t* da_off, uint8_t* da, int64_t vcap)
{
    struct G { int src; double Xi, Yi; };
    auto sgn = [](double a) { return (double)((a > 0) - (a < 0)); };
    auto max_abs = [&](int src, double X, double Y, bool ycoord) {
        double m = 0;
        for (int t = f->voff[src]; t < f->voff[src + 1]; ++t) m = std::max(m, std::fabs(ycoord ? f->vy[t] + Y : f->vx[t] + X));
        return m;
    };
    std::vector<G> ext;
    for (int i = 0; i < f->n; ++i) ext.push_back({i, f->x[i], f->y[i]});
    for (int i = 0; i < f->n; ++i)
        if (f->alive[i] && max_abs(i, f->x[i], f->y[i], false) > Lx) ext.push_back({i, f->x[i] - 2 * Lx * sgn(f->x[i]), f->y[i]});            // :21-30
    const size_t n1 = ext.size();
    for (size_t i = 0; i < n1; ++i)
        if (f->alive[ext[i].src] && max_abs(ext[i].src, ext[i].Xi, ext[i].Yi, true) > Ly) ext.push_back({ext[i].src, ext[i].Xi, ext[i].Yi - 2 * Ly * sgn(ext[i].Yi)});   // :38-46
    const double N0 = (double)ext.size();
    int64_t pos = 0; da_off[0] = 0;
    for (int q = 0; q < count; ++q) {
        const int i = idx[q] - 1;
        int nv = f->voff[i + 1] - f->voff[i];
        const int o = f->voff[i];
        if (nv > 1 && f->vx[o] == f->vx[o + nv - 1] && f->vy[o] == f->vy[o + nv - 1]) --nv;                                                  // polyshape drops the closing vertex
        if (pos + nv > vcap) return -1;
        for (int t = 0; t < nv; ++t) da[pos + t] = 0;
        da_off[q + 1] = pos + nv;
        if (q + 1 < 1 + Nb || row_off[i + 1] == row_off[i]) { pos += nv; continue; }                                                          // :54-55
        vec vxw(nv), vyw(nv);
        for (int t = 0; t < nv; ++t) { vxw[t] = f->vx[o + t] + f->x[i]; vyw[t] = f->vy[o + t] + f->y[i]; }
        std::vector<int> in(nv, 0);
        bool bnd = false;
        for (int64_t r = row_off[i]; r < row_off[i + 1]; ++r) {
            const double* a = rows + r * 7;
            if (std::isinf(a[0])) { bnd = true; continue; }
            if (!(a[0] <= N0)) continue;
            // break2 = dsearchn(polytrue.Vertices, [Xi Yi]) (:74): first nearest vertex
            int best = 0; double bd = INF;
            for (int t = 0; t < nv; ++t) { const double d = (vxw[t] - a[3]) * (vxw[t] - a[3]) + (vyw[t] - a[4]) * (vyw[t] - a[4]); if (d < bd) { bd = d; best = t; } }
            da[pos + best] = 1;
            // inpolygon of the floe's vertices in the partner's outline (:78-82)
            const G& g = ext[(size_t)a[0] - 1];
            vec cx, cy;
            for (int t = f->voff[g.src]; t < f->voff[g.src + 1]; ++t) { cx.push_back(f->vx[t] + g.Xi); cy.push_back(f->vy[t] + g.Yi); }
            std::vector<char> inn;
            inpolygon(vxw, vyw, cx, cy, inn);
            for (int t = 0; t < nv; ++t) in[t] += inn[t] ? 1 : 0;
        }
        if (bnd) {                                                                                                                            // :83-86
            vec bx(boxx, boxx + nbox), by(boxy, boxy + nbox);
            std::vector<char> inn;
            inpolygon(vxw, vyw, bx, by, inn);
            for (int t = 0; t < nv; ++t) in[t] += inn[t] ? 0 : 1;
        }
        for (int t = 0; t < nv; ++t) if (in[t] > 0) da[pos + t] = 1;                                                                          // :87
        pos += nv;
    }
    return (int)pos;
}

// ------------------------------------------------------------------------------------------------
// SURVEY.md 8f row f4, second half: the bounding-radius pair searches of weld.m and FloeSimplify.m.
//
// szo_weld_search -- Physical_Processes/weld.m:25-81 restated literally.  The first Nb (boundary) floes are cut off (:25), the
// rest are binned on an Nx x Ny grid (:31-36: Binx = fix((Xi-min(x))/(max(x)-min(x))*Nx+1), the same for y; xmin..ymax are
// min(x), max(x), min(y), max(y) of the colon vectors of :31-32 as the caller evaluated them), bin `count` = (i-1)*Ny + j holds
// the floes with Binx == i and Biny == j in their original order (:40-48).  Inside a bin floe i records every floe j of the
// SAME bin with  alive(j) && d > 1 && d < rmax(i)+rmax(j),  d = sqrt((Xi(i)-Xi(j))^2 + (Yi(i)-Yi(j))^2), ascending j (:66-79).
// Output, per floe q of the cut list (q = 0 .. n-Nb-1): bin[q] = count (1-based) or 0 when the floe is in no bin;
// partner[off[q] .. off[q+1]) = partners as 1-BASED positions in the cut list, ascending (their bin-local numbers floeNum are
// the ranks of those positions within the bin).  Returns the number of partners, or -1 when cap is too small.
int64_t szo_weld_search(const SzFloesSoA* f, int Nb, int Nx, int Ny, double xmin, double xmax, double ymin, double ymax,
                        int32_t* bin, int64_t* off, int32_t* partner, int64_t cap)
{
    const int n = f->n - Nb;
    if (n <= 0) { if (off) off[0] = 0; return 0; }
    std::vector<int> bx(n), by(n);
    auto fixd = [](double v) { return std::trunc(v); };
    for (int q = 0; q < n; ++q) {
        const double X = f->x[Nb + q], Y = f->y[Nb + q];
        const double a = fixd((X - xmin) / (xmax - xmin) * Nx + 1), b = fixd((Y - ymin) / (ymax - ymin) * Ny + 1);      // :35-36
        bx[q] = (a >= 1 && a <= Nx) ? (int)a : 0; by[q] = (b >= 1 && b <= Ny) ? (int)b : 0;                                // NaN compares false
        bin[q] = (bx[q] && by[q]) ? (bx[q] - 1) * Ny + by[q] : 0;                                                          // :40-48
    }
    std::vector<std::vector<int>> members((size_t)Nx * Ny + 1);
    for (int q = 0; q < n; ++q) if (bin[q]) members[bin[q]].push_back(q);
    int64_t pos = 0; off[0] = 0;
    for (int q = 0; q < n; ++q) {
        if (bin[q]) {
            const double Xi = f->x[Nb + q], Yi = f->y[Nb + q], ri = f->rmax[Nb + q];
            for (int j : members[bin[q]]) {                                                                                // :66-67, j ascending
                const double dx = Xi - f->x[Nb + j], dy = Yi - f->y[Nb + j];
                const double d = std::sqrt(dx * dx + dy * dy);
                if (f->alive[Nb + j] && d > 1 && d < (ri + f->rmax[Nb + j])) { if (pos >= cap) return -1; partner[pos++] = j + 1; }
            }
        }
        off[q + 1] = pos;
    }
    return pos;
}
// szo_simplify_search -- polygon_operations/FloeSimplify.m:13-31: the floes Floe0(j) that may overlap the floe about to be
// simplified:  alive(j) && d > 1 && d < floe.rmax + rmax(j),  ascending j over the whole list.  idx: `count` floe numbers
// (1-based) whose own record in `f` is the query floe (Subzero.m:176-186 passes Floe(ii) and the list).
int64_t szo_simplify_search(const SzFloesSoA* f, int count, const int32_t* idx, int64_t* off, int32_t* partner, int64_t cap)
{
    int64_t pos = 0; off[0] = 0;
    for (int q = 0; q < count; ++q) {
        const int i = idx[q] - 1;
        for (int j = 0; j < f->n; ++j) {
            const double dx = f->x[i] - f->x[j], dy = f->y[i] - f->y[j];
            const double d = std::sqrt(dx * dx + dy * dy);
            if (f->alive[j] && d > 1 && d < (f->rmax[i] + f->rmax[j])) { if (pos >= cap) return -1; partner[pos++] = j + 1; }
        }
        off[q + 1] = pos;
    }
    return pos;
}

}  // extern "C"
