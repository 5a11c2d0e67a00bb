// sz_euler.cuh -- calc_eulerian_data.m:68-187: mass-weighted coarse-grid averages of the floe state.
//
// The reference loops over the cells of an Nx x Ny grid; for each cell: the floes whose centre is closer than
// rmax + half the cell diagonal (:113-119), the area each of them shares with the cell (:134-147, intersect(box, poly):
// here the Clipper-exact intersection + the polyshape area of every returned region), and sums weighted by
// mass * shared area / floe area over those floes in ascending list order (:149-187).
//
// Decomposition on the device (all pieces below are host/device code, so the CPU-side parity test runs the same source):
//   items      one (cell, floe) pair per candidate, generated floe by floe (a floe only meets the few cells around it);
//   item_area  one Clipper sweep per item, all items in parallel (size classes like the narrow phase);
//   sort       a stable radix sort of the items by cell: inside a cell they stay in ascending floe order;
//   cell_reduce  one thread per cell adds its items up sequentially, i.e. in the reference's order (bit-identical sums).
#pragma once
#include "sz_pairforce.cuh"

namespace szeul {

using szclip::i64;
using szclip::P64;

// output planes, each Ny x Nx, element (jj, ii) at jj*Nx + ii, jj = 0 is the TOP row (:72 fliplr)
enum { O_U, O_V, O_DU, O_DV, O_STRESS, O_SXX, O_SYX, O_SXY, O_SYY, O_EUX, O_EVX, O_EUY, O_EVY, O_C, O_OVER, O_MTOT, O_AREA, O_H, N_OUT };

struct Grid {       // :70-80, edges by MATLAB's colon (a + k*d), the last one pinned to the upper bound
    int Nx, Ny; double xmin, xmax, ymin, ymax;
    SZ_HD double xe(int k) const { return k == Nx ? xmax : xmin + k * ((xmax - xmin) / Nx); }
    SZ_HD double ye_up(int k) const { return k == Ny ? ymax : ymin + k * ((ymax - ymin) / Ny); }
    SZ_HD double ye(int j) const { return ye_up(Ny - j); }                    // rows run from ymax down
    SZ_HD double dx() const { return fabs(xe(1) - xe(0)); }
    SZ_HD double dy() const { return fabs(ye(1) - ye(0)); }
    SZ_HD double r_max() const { const double a = dx() / 2, b = dy() / 2; return sqrt(a * a + b * b); }
    SZ_HD double xc(int ii) const { return 0.5 * (xe(ii) + xe(ii + 1)); }
    SZ_HD double yc(int jj) const { return 0.5 * (ye(jj) + ye(jj + 1)); }
};

struct EulerArgs {
    Grid g;
    // the floe list of :7-65 (dead floes dropped, periodic images appended): entry q = source floe + centroid
    int n_list; const int* lsrc; const double* lx; const double* ly;
    // per source floe
    const double* rmax; const double* area; const double* h; const double* u; const double* v;
    const double* mass; const double* over; const double* dU; const double* dV; const double* stress; const double* strain;   // stress, strain: [n][4]
    const int* voff; const double* vx; const double* vy;
    // items
    int n_items; const int* item_cell; const int* item_q; double* item_area; int* item_status;
    const int* sorted;          // item numbers ordered by (cell, q)
    const int* cell_off;        // [Nx*Ny + 1] into `sorted`
    double* out;                // N_OUT planes
};

SZ_HD double nz(double v) { return (v != v) ? 0.0 : v; }                      // :100-111 `x(isnan(x)) = 0`

// candidate test of :113-119 for list entry q and cell (ii, jj)
SZ_HD bool is_candidate(const EulerArgs& a, int q, int ii, int jj)
{
    const double xc = a.g.xc(ii), yc = a.g.yc(jj);
    const double pint = sqrt((xc - a.lx[q]) * (xc - a.lx[q]) + (yc - a.ly[q]) * (yc - a.ly[q])) - (a.rmax[a.lsrc[q]] + a.g.r_max());
    return pint < 0;
}
// a conservative index box around the cells entry q can be a candidate of (one cell of slack on every side); empty when the
// entry can be no candidate at all (NaN centre or radius)
SZ_HD bool cell_range(const EulerArgs& a, int q, int& i0, int& i1, int& j0, int& j1)
{
    const double X = a.lx[q], Y = a.ly[q], R = a.rmax[a.lsrc[q]] + a.g.r_max();
    if (X != X || Y != Y || R != R) return false;
    const double dx = a.g.dx(), dy = a.g.dy();
    double fi0 = floor((X - R - a.g.xmin) / dx) - 1, fi1 = floor((X + R - a.g.xmin) / dx) + 1;
    double fj0 = floor((a.g.ymax - (Y + R)) / dy) - 1, fj1 = floor((a.g.ymax - (Y - R)) / dy) + 1;
    if (!(fi0 > 0)) fi0 = 0;                      // also catches NaN (0/0 cells) and -Inf
    if (!(fj0 > 0)) fj0 = 0;
    if (!(fi1 < a.g.Nx - 1)) fi1 = a.g.Nx - 1;
    if (!(fj1 < a.g.Ny - 1)) fj1 = a.g.Ny - 1;
    if (fi0 > a.g.Nx) fi0 = a.g.Nx;               // far outside the grid (or an infinite centre): keep the casts defined, the range empty
    if (fj0 > a.g.Ny) fj0 = a.g.Ny;
    if (fi1 < -1) fi1 = -1;
    if (fj1 < -1) fj1 = -1;
    i0 = (int)fi0; i1 = (int)fi1; j0 = (int)fj0; j1 = (int)fj1;
    return i0 <= i1 && j0 <= j1;
}

// area(polyshape) of every emitted path, added up in emission order: the shoelace about vertex 0 of ring_area_centroid,
// streamed (the first and the closing term of that loop are zero)
struct AreaSink {
    double total, x0, y0, xi, yi, a2; int k, cnt;
    SZ_HD void reset() { total = 0; cnt = 0; k = 0; a2 = 0; x0 = y0 = xi = yi = 0; }
    SZ_HD void finish() { if (cnt >= 3) total += fabs(a2) / 2; cnt = 0; }
    SZ_HD void begin_path(int c) { finish(); cnt = c; k = 0; a2 = 0; }
    SZ_HD void point(P64 p)
    {
        const double X = (double)p.x / SZ_SCALE, Y = (double)p.y / SZ_SCALE;
        if (k == 0) { x0 = X; y0 = Y; xi = 0; yi = 0; }
        else { const double xj = X - x0, yj = Y - y0; const double c = xi * yj - xj * yi; a2 += c; xi = xj; yi = yj; }
        ++k;
    }
};
struct BoxGetter {      // :134 box = [xe(ii) xe(ii) xe(ii+1) xe(ii+1) xe(ii); ye(jj) ye(jj+1) ye(jj+1) ye(jj) ye(jj)], packed like polyclip.m:66
    double x[5], y[5];
    SZ_HD P64 operator()(int i) const { P64 p; p.x = szpf::matlab_int64(x[i] * SZ_SCALE); p.y = szpf::matlab_int64(y[i] * SZ_SCALE); return p; }
};
struct OutlineGetter {
    const double* vx; const double* vy; double X, Y;
    SZ_HD P64 operator()(int i) const { P64 p; p.x = szpf::matlab_int64((vx[i] + X) * SZ_SCALE); p.y = szpf::matlab_int64((vy[i] + Y) * SZ_SCALE); return p; }
};

// Aover of one item (:140-147): polyclip(box, floe outline, 'int') and the area of what comes back.
// Returns PS_OK / PS_CAPACITY (arena too small: rerun in a larger class) / PS_CLIPPER_FAIL.
template <class E>
SZ_HD int item_area(E& eng, const EulerArgs& a, int item, double& area_out)
{
    const int cell = a.item_cell[item], q = a.item_q[item];
    const int jj = cell / a.g.Nx, ii = cell - jj * a.g.Nx;
    BoxGetter box;
    box.x[0] = a.g.xe(ii); box.x[1] = a.g.xe(ii); box.x[2] = a.g.xe(ii + 1); box.x[3] = a.g.xe(ii + 1); box.x[4] = a.g.xe(ii);
    box.y[0] = a.g.ye(jj); box.y[1] = a.g.ye(jj + 1); box.y[2] = a.g.ye(jj + 1); box.y[3] = a.g.ye(jj); box.y[4] = a.g.ye(jj);
    const int s = a.lsrc[q], o = a.voff[s];
    OutlineGetter fl; fl.vx = a.vx + o; fl.vy = a.vy + o; fl.X = a.lx[q]; fl.Y = a.ly[q];
    area_out = 0;
    eng.begin(1);
    eng.add_path(box, 5, 0);
    eng.add_path(fl, a.voff[s + 1] - o, 1);
    const int st = eng.execute();
    if (st == szclip::ST_OVERFLOW) return szpf::PS_CAPACITY;
    if (st != szclip::ST_OK) return szpf::PS_CLIPPER_FAIL;
    AreaSink sink; sink.reset();
    eng.emit(sink);
    sink.finish();
    area_out = sink.total;
    return szpf::PS_OK;
}

// ---- one cell (:131-187): its items in ascending list order.
// What one item adds to the sums of its cell: Mtot (:149), Atot (:150), the OverlapArea sum and the item count of :153, the 13
// weighted sums of :155-169, and the candidate-mass sum M0 of :131.  Returns false for an item without shared area
// (:146-147), which contributes to M0 only.
enum { T_MTOT = 0, T_ATOT = 1, T_OVER = 2, T_CNT = 3, T_VAL = 4 /* h u v du dv sxx syx sxy syy eux evx euy evy */, T_M0 = 17, N_TERMS = 18 };
SZ_HD bool item_terms(const EulerArgs& a, int item, double* t)
{
    const int s = a.lsrc[a.item_q[item]];
    const double m = nz(a.mass[s]);
    t[T_M0] = m;
    const double ar = a.item_area[item];
    if (!(ar != 0)) return false;
    const double A = nz(a.area[s]);
    t[T_MTOT] = m * ar / A; t[T_ATOT] = ar; t[T_OVER] = a.over[s]; t[T_CNT] = 1.0;
    t[T_VAL + 0] = nz(a.h[s]) * m * ar / A;
    t[T_VAL + 1] = nz(a.u[s]) * m * ar / A; t[T_VAL + 2] = nz(a.v[s]) * m * ar / A;
    t[T_VAL + 3] = nz(a.dU[s]) * m * ar / A; t[T_VAL + 4] = nz(a.dV[s]) * m * ar / A;
    for (int j = 0; j < 4; ++j) { t[T_VAL + 5 + j] = a.stress[(size_t)s * 4 + j] * m * ar / A; t[T_VAL + 9 + j] = a.strain[(size_t)s * 4 + j] * m * ar / A; }
    return true;
}
// the cell's outputs from its sums S[N_TERMS]
SZ_HD void cell_finalize(const EulerArgs& a, int cell, const double* S)
{
    const size_t cells = (size_t)a.g.Nx * (size_t)a.g.Ny, e = (size_t)cell;
    if (!(S[T_M0] > 0)) return;                                               // :131
    const double Mtot = S[T_MTOT], Atot = S[T_ATOT];
    a.out[O_C * cells + e] = Atot / (a.g.dx() * a.g.dy());                    // :151
    if (!(Mtot > 0)) return;
    a.out[O_OVER * cells + e] = S[T_OVER] / S[T_CNT];                         // :153
    a.out[O_MTOT * cells + e] = Mtot; a.out[O_AREA * cells + e] = Atot;
    a.out[O_H * cells + e] = S[T_VAL + 0] / Mtot;
    a.out[O_U * cells + e] = S[T_VAL + 1] / Mtot; a.out[O_V * cells + e] = S[T_VAL + 2] / Mtot;
    a.out[O_DU * cells + e] = S[T_VAL + 3] / Mtot; a.out[O_DV * cells + e] = S[T_VAL + 4] / Mtot;
    const double sxx = S[T_VAL + 5] / Mtot, syx = S[T_VAL + 6] / Mtot, sxy = S[T_VAL + 7] / Mtot, syy = S[T_VAL + 8] / Mtot;
    a.out[O_SXX * cells + e] = sxx; a.out[O_SYX * cells + e] = syx; a.out[O_SXY * cells + e] = sxy; a.out[O_SYY * cells + e] = syy;
    a.out[O_EUX * cells + e] = S[T_VAL + 9] / Mtot; a.out[O_EVX * cells + e] = S[T_VAL + 10] / Mtot;
    a.out[O_EUY * cells + e] = S[T_VAL + 11] / Mtot; a.out[O_EVY * cells + e] = S[T_VAL + 12] / Mtot;
    // max(eig([sxx syx; sxy syy])) (:170), eigenvalues of the symmetric stress tensor
    const double tr = sxx + syy, det = sxx * syy - syx * sxy, disc = tr * tr / 4 - det;
    double lam = tr / 2 + sqrt(disc > 0 ? disc : 0);
    if (fabs(lam) > 1e8) lam = 0;                                             // :171-173
    a.out[O_STRESS * cells + e] = lam;
}
// sequential form: one caller adds the cell's items up in order (the host test; the device's fallback kernel).  The device's
// default is a warp per cell that computes 32 items' terms at a time and adds them in this same order (euler_cell_warp_kernel).
SZ_HD void cell_reduce(const EulerArgs& a, int cell)
{
    double S[N_TERMS], t[N_TERMS];
    for (int j = 0; j < N_TERMS; ++j) S[j] = 0;
    for (int k = a.cell_off[cell]; k < a.cell_off[cell + 1]; ++k) {
        for (int j = 0; j < N_TERMS; ++j) t[j] = 0;
        const bool valid = item_terms(a, a.sorted[k], t);
        S[T_M0] += t[T_M0];
        if (valid) for (int j = 0; j < T_M0; ++j) S[j] += t[j];
    }
    cell_finalize(a, cell, S);
}

}  // namespace szeul
