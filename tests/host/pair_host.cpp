// Host build of the product's pair force law (subzero_b200/csrc/sz_pairforce.cuh) for CPU-side
// parity tests against the oracle (test infrastructure; the product never runs this on the CPU).
#include "../../subzero_b200/csrc/sz_pairforce.cuh"
#include "../../subzero_b200/csrc/sz_corners.cuh"
#include "../../subzero_b200/csrc/sz_euler.cuh"
#include "../../include/subzero_b200.h"
#include <vector>
#include <memory>
#include <cstring>
#include <algorithm>

using namespace szpf;
typedef szclip::ClipCaps<2600, 1300, 10000, 2600, 10000, 2600, 512, 3900> BigClip;
typedef PairCaps<BigClip, 1300, 5200, 64, 6000, 32> BigPair;
typedef szclip::ClipCaps<32, 16, 96, 32, 64, 32, 16, 48> SmallClip;
typedef PairCaps<SmallClip, 20, 64, 6, 40, 4> SmallPair;

static void fill_params(const SzParams* p, const double* boxx, const double* boxy, int nbox, Params& P)
{
    P.Lx = p->Lx; P.Ly = p->Ly; P.modulus = p->modulus; P.dt = p->dt; P.nu = p->nu; P.mu = p->mu; P.merge_frac = p->merge_frac;
    P.wall_frac = p->wall_frac; P.amin_per_vertex = p->amin_per_vertex; P.vertex_match_tol = p->vertex_match_tol;
    P.on_edge_tol = p->on_edge_tol; P.dl_min = p->dl_min; P.close_gap = p->close_gap; P.big_floe_r = p->big_floe_r;
    P.domain_area_frac = p->domain_area_frac; P.Nb = p->Nb; P.periodic = p->periodic; P.collision = p->collision;
    P.has_box = nbox > 0;
    if (nbox > 0) {
        P.bxmin = P.bxmax = boxx[0]; P.bymin = P.bymax = boxy[0];
        for (int i = 1; i < nbox; ++i) { if (boxx[i] < P.bxmin) P.bxmin = boxx[i]; if (boxx[i] > P.bxmax) P.bxmax = boxx[i]; if (boxy[i] < P.bymin) P.bymin = boxy[i]; if (boxy[i] > P.bymax) P.bymax = boxy[i]; }
        // area(polyshape(c2_boundary')) with the same vertex-0-relative shoelace
        double a2 = 0; for (int i = 0; i < nbox; ++i) { int j = (i + 1) % nbox; double xi = boxx[i] - boxx[0], yi = boxy[i] - boxy[0], xj = boxx[j] - boxx[0], yj = boxy[j] - boxy[0]; a2 += xi * yj - xj * yi; }
        P.barea = fabs(a2) / 2;
    } else { P.bxmin = P.bxmax = P.bymin = P.bymax = P.barea = 0; }
}

template <class CAPS>
static int run(const SzParams* prm, const double* cax, const double* cay, int n1, const double* body1,
               const double* c2x, const double* c2y, int n2, const double* body2, int is_boundary,
               const double* boxx, const double* boxy, int nbox, double* rows_out, int rows_cap, double* overlap_state, bool fast = false, bool split = false)
{
    if (n1 + 1 > CAPS::NV || n2 + 1 > CAPS::NV) return PS_CAPACITY;
    std::unique_ptr<Workspace<CAPS>> w(new Workspace<CAPS>);
    Params P; fill_params(prm, boxx, boxy, nbox, P);
    Body b1{body1[0], body1[1], body1[2], body1[3], body1[4], body1[5], body1[6]};
    Body b2{body2[0], body2[1], body2[2], body2[3], body2[4], body2[5], body2[6]};
    w->n1 = n1; w->n2 = n2;
    for (int i = 0; i < n1; ++i) { w->c1x[i] = cax[i] + b1.Xi; w->c1y[i] = cay[i] + b1.Yi; }
    for (int i = 0; i < n2; ++i) { w->c2x[i] = c2x[i]; w->c2y[i] = c2y[i]; }
    PairResult res; std::vector<double> rows(CAPS::ROWS * 5);
    // convexity of both outlines in Clipper's coordinates, as the device computes it per floe (ext_prep_kernel)
    struct G { const double* x; const double* y; P64 operator()(int i) const { P64 p; p.x = matlab_int64(x[i] * SZ_SCALE); p.y = matlab_int64(y[i] * SZ_SCALE); return p; } };
    auto open_n = [](const double* x, const double* y, int n) { while (n > 1 && x[n - 1] == x[0] && y[n - 1] == y[0]) --n; return n; };
    const bool convex = !is_boundary && ring_is_strictly_convex(G{w->c1x, w->c1y}, open_n(w->c1x, w->c1y, n1)) && ring_is_strictly_convex(G{w->c2x, w->c2y}, open_n(w->c2x, w->c2y, n2));
    PairHints hints{convex, 0, 0, 0, 0};
    if (convex) {
        hints.no1 = open_n(w->c1x, w->c1y, n1); hints.no2 = open_n(w->c2x, w->c2y, n2);
        hints.rot1 = ring_bottom_vertex(G{w->c1x, w->c1y}, hints.no1); hints.rot2 = ring_bottom_vertex(G{w->c2x, w->c2y}, hints.no2);
    }
    if (split) {
        // class C split in two (experiment): clip #1 alone, reading the outline where it lies (relative outline + centroid, like
        // the device's sweep kernel reads the vertex pool), its polygon handed over through a strided buffer; then the rest
        std::unique_ptr<WorkspaceLite<CAPS>> wl(new WorkspaceLite<CAPS>);
        wl->n1 = n1; wl->n2 = n2;
        for (int i = 0; i < n1; ++i) { wl->c1x[i] = w->c1x[i]; wl->c1y[i] = w->c1y[i]; }
        for (int i = 0; i < n2; ++i) { wl->c2x[i] = w->c2x[i]; wl->c2y[i] = w->c2y[i]; }
        const int stride = 3, item = 1, cap = 16;
        std::vector<int> hst(stride, -5); std::vector<i64> hx((size_t)cap * stride, 0), hy((size_t)cap * stride, 0);
        ConvexHandoff ho{hst.data(), hx.data(), hy.data(), stride, cap, item};
        ClipInput subj, clip;
        subj.x = cax; subj.y = cay; subj.dx = b1.Xi; subj.dy = b1.Yi; subj.ix = subj.iy = 0; subj.n = hints.no1; subj.ring = 0; subj.rot = hints.rot1;
        clip.x = c2x; clip.y = c2y; clip.dx = 0; clip.dy = 0; clip.ix = clip.iy = 0; clip.n = hints.no2; clip.ring = 0; clip.rot = hints.rot2;
        const bool go = convex && hints.no1 >= 3 && hints.no2 >= 3;
        std::vector<i64> svx(2 * CAPS::NV), svy(2 * CAPS::NV), dqx(CAPS::RV), dqy(CAPS::RV), ox(CAPS::RV), oy(CAPS::RV);
        convex_sweep_only<CAPS::NV>(go, subj, clip, svx.data(), svy.data(), dqx.data(), dqy.data(), ox.data(), oy.data(), CAPS::RV, ho);
        if (go && hst[item] == -5) return -9;                 // the sweep must have answered
        if (hst[0] != -5 || hst[2] != -5) return -9;          // ... in its own slot only
        pair_force_convex_after_sweep<CAPS>(*wl, b1, b2, P, res, rows.data(), true, hints, ho);
    } else if (fast) {
        // class C as the device runs it: the convex fast path on the engine-less workspace; PS_BAIL = declined
        std::unique_ptr<WorkspaceLite<CAPS>> wl(new WorkspaceLite<CAPS>);
        wl->n1 = n1; wl->n2 = n2;
        for (int i = 0; i < n1; ++i) { wl->c1x[i] = w->c1x[i]; wl->c1y[i] = w->c1y[i]; }
        for (int i = 0; i < n2; ++i) { wl->c2x[i] = w->c2x[i]; wl->c2y[i] = w->c2y[i]; }
        pair_force_convex<CAPS>(*wl, b1, b2, P, res, rows.data(), true, hints);
    } else pair_force(*w, b1, b2, is_boundary != 0, P, res, rows.data(), true, hints);
    if (res.status != PS_OK) return res.status;
    *overlap_state = res.overlap_state;
    if (res.n_rows > rows_cap) return -2;
    for (int i = 0; i < res.n_rows * 5; ++i) rows_out[i] = rows[i];
    return res.n_rows;
}

extern "C" int szport_floe_interactions(const SzParams* prm, const double* cax, const double* cay, int n1, const double* body1,
                                        const double* c2x, const double* c2y, int n2, const double* body2, int is_boundary,
                                        const double* boxx, const double* boxy, int nbox,
                                        double* rows_out, int rows_cap, double* overlap_state, int small_class)
{
    if (small_class == 3) return run<SmallPair>(prm, cax, cay, n1, body1, c2x, c2y, n2, body2, is_boundary, boxx, boxy, nbox, rows_out, rows_cap, overlap_state, true, true);
    if (small_class == 2) return run<SmallPair>(prm, cax, cay, n1, body1, c2x, c2y, n2, body2, is_boundary, boxx, boxy, nbox, rows_out, rows_cap, overlap_state, true);
    if (small_class) return run<SmallPair>(prm, cax, cay, n1, body1, c2x, c2y, n2, body2, is_boundary, boxx, boxy, nbox, rows_out, rows_cap, overlap_state);
    return run<BigPair>(prm, cax, cay, n1, body1, c2x, c2y, n2, body2, is_boundary, boxx, boxy, nbox, rows_out, rows_cap, overlap_state);
}

// corners.m:10-88, the contact mask: the product's per-floe core (sz_corners.cuh, one lane per group on the host) over the
// periodic list built here the way the device kernels build it (originals, x images, y images over the extended list).
extern "C" int szport_corner_mask(const SzFloesSoA* f, const int64_t* row_off64, const double* rows, int count, const int32_t* idx, int nb_skip,
                                  double Lx, double Ly, const double* boxx, const double* boxy, int nbox, int64_t* da_off, uint8_t* da, int64_t vcap)
{
    const int n0 = f->n;
    std::vector<double> ex(f->x, f->x + n0), ey(f->y, f->y + n0); std::vector<int> esrc(n0); std::vector<uint8_t> ealive(f->alive, f->alive + n0);
    for (int i = 0; i < n0; ++i) esrc[i] = i;
    auto sgn = [](double v) { return (double)((v > 0) - (v < 0)); };
    auto pass = [&](int axis, double L) {
        const size_t n = ex.size();
        for (size_t i = 0; i < n; ++i) {
            if (!ealive[i]) continue;
            const int s = esrc[i]; double m = -SZ_INF;
            for (int t = f->voff[s]; t < f->voff[s + 1]; ++t) { const double a = fabs((axis ? f->vy[t] : f->vx[t]) + (axis ? ey[i] : ex[i])); if (a > m) m = a; }
            if (!(m > L)) continue;
            if (axis == 0) { ex.push_back(ex[i] - 2 * L * sgn(ex[i])); ey.push_back(ey[i]); }
            else { ex.push_back(ex[i]); ey.push_back(ey[i] - 2 * L * sgn(ey[i])); }
            esrc.push_back(esrc[i]); ealive.push_back(ealive[i]);
        }
    };
    pass(0, Lx); pass(1, Ly);
    const int n_ext = (int)ex.size();
    std::vector<int> row_off(n0 + 1), off(count + 1, 0);
    for (int i = 0; i <= n0; ++i) row_off[i] = (int)row_off64[i];
    szcorn::CornerArgs a; memset(&a, 0, sizeof(a));
    a.count = count; a.idx = idx; a.nb_skip = nb_skip; a.x = f->x; a.y = f->y; a.voff = f->voff; a.vx = f->vx; a.vy = f->vy;
    a.row_off = row_off.data(); a.rows = rows; a.cex = ex.data(); a.cey = ey.data(); a.cesrc = esrc.data(); a.n_ext = &n_ext;
    a.boxx = boxx; a.boxy = boxy; a.nbox = nbox;
    for (int q = 0; q < count; ++q) off[q + 1] = off[q] + szcorn::open_count(a, idx[q] - 1);
    if (off[count] > vcap) return -1;
    a.da_off = off.data(); a.da = da;
    for (int q = 0; q < count; ++q) szcorn::corner_mask_floe<1>(a, q, 0, 1u);
    for (int q = 0; q <= count; ++q) da_off[q] = off[q];
    return off[count];
}

// calc_eulerian_data.m: the product's item / reduction code (sz_euler.cuh) over the list and the items built here the way
// the device kernels build them (list: alive floes, x images, the stale-polygon y pass; items floe by floe, stable sort
// by cell).  Same signature as the checker's entry point.  Returns 0, -1 (Nb != 0), -2 (Clipper failure), -3 (cell_range
// missed a candidate: a bug), -4 (capacity).
extern "C" int szport_calc_eulerian_data(const SzFloesSoA* f, const double* mass, const double* overlap_area, const double* dUi_p, const double* dVi_p,
                                         const double* stress, const double* strain, int Nx, int Ny, int Nb,
                                         double xmin, double xmax, double ymin, double ymax, int periodic, double* out)
{
    if (Nb != 0) return -1;
    std::vector<int> lsrc; std::vector<double> lx, ly;
    for (int i = 0; i < f->n; ++i) if (f->alive[i]) { lsrc.push_back(i); lx.push_back(f->x[i]); ly.push_back(f->y[i]); }
    const double Lx = xmax, Ly = ymax;
    auto sgn = [](double v) { return (double)((v > 0) - (v < 0)); };
    if (periodic && !lsrc.empty()) {
        const size_t n1 = lsrc.size();
        double last_my = 0;
        for (size_t i = 0; i < n1; ++i) {
            const int s = lsrc[i]; double mx = 0, my = 0;
            for (int t = f->voff[s]; t < f->voff[s + 1]; ++t) { mx = std::max(mx, fabs(f->vx[t] + lx[i])); my = std::max(my, fabs(f->vy[t] + ly[i])); }
            last_my = my;
            if (mx > Lx) { lsrc.push_back(s); lx.push_back(lx[i] - 2 * Lx * sgn(lx[i])); ly.push_back(ly[i]); }
        }
        const size_t n2 = lsrc.size();
        if (last_my > Ly) for (size_t i = 0; i < n2; ++i) { lsrc.push_back(lsrc[i]); lx.push_back(lx[i]); ly.push_back(ly[i] - 2 * Ly * sgn(ly[i])); }
    }
    const size_t cells = (size_t)Nx * Ny;
    for (size_t k = 0; k < szeul::N_OUT * cells; ++k) out[k] = 0;
    szeul::EulerArgs a; memset(&a, 0, sizeof(a));
    a.g.Nx = Nx; a.g.Ny = Ny; a.g.xmin = xmin; a.g.xmax = xmax; a.g.ymin = ymin; a.g.ymax = ymax;
    a.n_list = (int)lsrc.size(); a.lsrc = lsrc.data(); a.lx = lx.data(); a.ly = ly.data();
    a.rmax = f->rmax; a.area = f->area; a.h = f->h; a.u = f->u; a.v = f->v;
    a.mass = mass; a.over = overlap_area; a.dU = dUi_p; a.dV = dVi_p; a.stress = stress; a.strain = strain;
    a.voff = f->voff; a.vx = f->vx; a.vy = f->vy; a.out = out;
    std::vector<int> item_cell, item_q;
    for (int q = 0; q < a.n_list; ++q) {
        int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
        const bool any = szeul::cell_range(a, q, i0, i1, j0, j1);
        for (int jj = 0; jj < Ny; ++jj) for (int ii = 0; ii < Nx; ++ii) {
            if (!szeul::is_candidate(a, q, ii, jj)) continue;
            if (!any || ii < i0 || ii > i1 || jj < j0 || jj > j1) return -3;
        }
        if (!any) continue;
        for (int jj = j0; jj <= j1; ++jj) for (int ii = i0; ii <= i1; ++ii) if (szeul::is_candidate(a, q, ii, jj)) { item_cell.push_back(jj * Nx + ii); item_q.push_back(q); }
    }
    const int n_items = (int)item_cell.size();
    std::vector<double> item_area(n_items, 0.0); std::vector<int> item_status(n_items, 0), sorted(n_items), cell_off(cells + 1, 0);
    a.n_items = n_items; a.item_cell = item_cell.data(); a.item_q = item_q.data(); a.item_area = item_area.data(); a.item_status = item_status.data();
    std::unique_ptr<Workspace<BigPair>> w(new Workspace<BigPair>);
    for (int k = 0; k < n_items; ++k) {
        const int st = szeul::item_area(w->eng, a, k, item_area[k]);
        if (st == PS_CLIPPER_FAIL) return -2;
        if (st != PS_OK) return -4;
    }
    for (int k = 0; k < n_items; ++k) sorted[k] = k;
    std::stable_sort(sorted.begin(), sorted.end(), [&](int p, int q) { return item_cell[p] < item_cell[q]; });
    for (int k = 0; k < n_items; ++k) ++cell_off[item_cell[k] + 1];
    for (size_t c = 0; c < cells; ++c) cell_off[c + 1] += cell_off[c];
    a.sorted = sorted.data(); a.cell_off = cell_off.data();
    for (size_t c = 0; c < cells; ++c) szeul::cell_reduce(a, (int)c);
    return 0;
}

// the classifier's certificate for outlines of any shape (sz_apart.cuh); bounding boxes computed here like ext_prep_kernel's
#include "../../subzero_b200/csrc/sz_apart.cuh"
extern "C" int szport_rings_apart(const double* ax, const double* ay, int na, double AX, double AY, const double* bx, const double* by, int nb, double BX, double BY)
{
    auto box = [](const double* x, const double* y, int n, double X, double Y, double* o) {
        o[0] = o[2] = 1e300; o[1] = o[3] = -1e300;
        for (int i = 0; i < n; ++i) { const double u = x[i] + X, v = y[i] + Y; o[0] = u < o[0] ? u : o[0]; o[1] = u > o[1] ? u : o[1]; o[2] = v < o[2] ? v : o[2]; o[3] = v > o[3] ? v : o[3]; }
    };
    double A[4], B[4]; box(ax, ay, na, AX, AY, A); box(bx, by, nb, BX, BY, B);
    return szapart::rings_apart(ax, ay, na, AX, AY, A[0], A[1], A[2], A[3], bx, by, nb, BX, BY, B[0], B[1], B[2], B[3]) ? 1 : 0;
}
