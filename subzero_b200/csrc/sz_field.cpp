// sz_field.cpp -- synthetic input of BASELINE.json configs[4] ("10k-1M packed Voronoi polygon floes in
// a doubly-periodic domain"): host-side generator, no CUDA.  SURVEY.md 8(d) fixes the recipe:
//   domain [-L,L]^2 with L = 0.5*sqrt(N*A0); N uniform random sites (std::mt19937_64, given seed);
//   periodic Voronoi cells; every cell scaled by (1+eps) about its centroid so neighbours overlap in
//   strips; h = 0.25 (Subzero.m:37); Ui,Vi ~ U(-0.1,0.1); ksi ~ U(-1e-5,1e-5);
//   Modulus = 1.5e3*(mean(sqrt(A)) + min(sqrt(A))) (Subzero.m:77).
// Each floe is laid out like Initialize_Model/initialize_floe_values.m:12-52 does it: centroid and
// area by the polyshape formulas, c_alpha = outline about the centroid, CLOSED (first vertex
// repeated, :17) and clockwise like the reference's fixture shapes, rmax = farthest vertex (:21).
#include "../../include/subzero_b200.h"
#include <vector>
#include <random>
#include <cmath>
#include <cstring>
#include <thread>
#include <algorithm>
#include <string>

#ifdef SZ_FIELD_STANDALONE
// libsz_field.so: the generator without the CUDA library around it (bench.py's CPU arm); errors go to stderr
#include <cstdio>
#include <cstdarg>
void sz_set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap); fputc('\n', stderr); }
#else
void sz_set_error(const char* fmt, ...);   // sz_contact.cu
#endif

struct SzField {
    std::vector<double> x, y, rmax, h, area, u, v, ksi, vx, vy;
    std::vector<uint8_t> alive;
    std::vector<int32_t> voff;
};

namespace {

struct Pt { double x, y; };

// clip convex polygon `poly` by the half-plane {p : (p - m).d <= 0}
void clip_halfplane(std::vector<Pt>& poly, std::vector<Pt>& tmp, Pt m, Pt d)
{
    tmp.clear();
    const size_t n = poly.size();
    if (n == 0) return;
    double sprev = (poly[n - 1].x - m.x) * d.x + (poly[n - 1].y - m.y) * d.y;
    for (size_t i = 0; i < n; ++i) {
        const Pt& a = poly[(i + n - 1) % n]; const Pt& b = poly[i];
        double s = (b.x - m.x) * d.x + (b.y - m.y) * d.y;
        if ((sprev <= 0) != (s <= 0)) {
            double t = sprev / (sprev - s);
            tmp.push_back({a.x + t * (b.x - a.x), a.y + t * (b.y - a.y)});
        }
        if (s <= 0) tmp.push_back(b);
        sprev = s;
    }
    poly.swap(tmp);
}

inline double u01(std::mt19937_64& g) { return (double)(g() >> 11) * (1.0 / 9007199254740992.0); }

}  // namespace

extern "C" int sz_field_voronoi(SzField** out, int32_t n, uint64_t seed, double mean_area, double inflate, SzParams* prm_out)
{
    if (!out || n < 4 || !(mean_area > 0)) { sz_set_error("sz_field_voronoi: bad arguments"); return SZ_ERR_ARG; }
    const double L = 0.5 * std::sqrt((double)n * mean_area);
    std::mt19937_64 gen(seed);
    std::vector<Pt> site(n);
    for (int i = 0; i < n; ++i) { site[i].x = (2 * u01(gen) - 1) * L; site[i].y = (2 * u01(gen) - 1) * L; }
    std::vector<double> ru(n), rv(n), rk(n);
    for (int i = 0; i < n; ++i) { ru[i] = -0.1 + 0.2 * u01(gen); rv[i] = -0.1 + 0.2 * u01(gen); rk[i] = -1e-5 + 2e-5 * u01(gen); }

    // bucket grid, ~2 sites per bucket
    int g = std::max(1, (int)std::floor(std::sqrt(n / 2.0)));
    const double cs = 2 * L / g;
    std::vector<int> start(g * g + 1, 0), items(n), cid(n);
    for (int i = 0; i < n; ++i) {
        int cx = std::min(g - 1, std::max(0, (int)((site[i].x + L) / cs))), cy = std::min(g - 1, std::max(0, (int)((site[i].y + L) / cs)));
        cid[i] = cy * g + cx; start[cid[i] + 1]++;
    }
    for (int c = 0; c < g * g; ++c) start[c + 1] += start[c];
    { std::vector<int> pos(start.begin(), start.end() - 1); for (int i = 0; i < n; ++i) items[pos[cid[i]]++] = i; }

    std::vector<std::vector<Pt>> cells(n);
    auto work = [&](int t0, int t1) {
        std::vector<Pt> poly, tmp;
        for (int i = t0; i < t1; ++i) {
            const Pt s = site[i];
            const double R = 2 * L;   // starting box: the whole (periodic) plane tile around the site
            poly = {{s.x - R, s.y - R}, {s.x + R, s.y - R}, {s.x + R, s.y + R}, {s.x - R, s.y + R}};
            const int cx = cid[i] % g, cy = cid[i] / g;
            for (int ring = 0; ring <= g; ++ring) {
                if (ring > 0) {
                    double md = 0; for (auto& p : poly) md = std::max(md, std::hypot(p.x - s.x, p.y - s.y));
                    if (2 * md < (ring - 1) * cs) break;   // nothing farther away can cut the cell
                }
                for (int dy = -ring; dy <= ring; ++dy) for (int dx = -ring; dx <= ring; ++dx) {
                    if (std::max(std::abs(dx), std::abs(dy)) != ring) continue;
                    int bx = cx + dx, by = cy + dy; double ox = 0, oy = 0;
                    while (bx < 0) { bx += g; ox -= 2 * L; } while (bx >= g) { bx -= g; ox += 2 * L; }
                    while (by < 0) { by += g; oy -= 2 * L; } while (by >= g) { by -= g; oy += 2 * L; }
                    const int c = by * g + bx;
                    for (int t = start[c]; t < start[c + 1]; ++t) {
                        const int j = items[t];
                        if (j == i && ox == 0 && oy == 0) continue;
                        Pt q{site[j].x + ox, site[j].y + oy};
                        Pt m{0.5 * (s.x + q.x), 0.5 * (s.y + q.y)}, d{q.x - s.x, q.y - s.y};
                        clip_halfplane(poly, tmp, m, d);
                    }
                }
            }
            cells[i] = poly;
        }
    };
    {
        int nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
        std::vector<std::thread> th; int chunk = (n + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) { int a = t * chunk, b = std::min(n, a + chunk); if (a < b) th.emplace_back(work, a, b); }
        for (auto& t : th) t.join();
    }

    SzField* f = new SzField;
    f->x.resize(n); f->y.resize(n); f->rmax.resize(n); f->h.assign(n, 0.25); f->area.resize(n);
    f->u = ru; f->v = rv; f->ksi = rk; f->alive.assign(n, 1); f->voff.assign(n + 1, 0);
    double sum_sqrt = 0, min_sqrt = 1e300;
    for (int i = 0; i < n; ++i) {
        std::vector<Pt>& c = cells[i];
        // drop numerically repeated vertices
        std::vector<Pt> q;
        for (auto& p : c) if (q.empty() || std::hypot(p.x - q.back().x, p.y - q.back().y) > 1e-6) q.push_back(p);
        while (q.size() > 1 && std::hypot(q.front().x - q.back().x, q.front().y - q.back().y) <= 1e-6) q.pop_back();
        // polyshape-style area / centroid (vertex-0-relative shoelace)
        const size_t m = q.size();
        double a2 = 0, sx = 0, sy = 0;
        for (size_t k = 0; k < m; ++k) {
            size_t k1 = (k + 1) % m;
            double xi = q[k].x - q[0].x, yi = q[k].y - q[0].y, xj = q[k1].x - q[0].x, yj = q[k1].y - q[0].y;
            double cr = xi * yj - xj * yi; a2 += cr; sx += (xi + xj) * cr; sy += (yi + yj) * cr;
        }
        const double cxx = q[0].x + sx / (3 * a2), cyy = q[0].y + sy / (3 * a2);
        if (a2 > 0) std::reverse(q.begin(), q.end());   // clockwise, like FloeShapes.mat
        const double s = 1 + inflate;
        f->x[i] = cxx; f->y[i] = cyy;
        f->area[i] = std::fabs(a2) / 2 * s * s;
        double r2 = 0;
        for (size_t k = 0; k <= m; ++k) {
            const Pt& p = q[k % m];
            double ax = s * (p.x - cxx), ay = s * (p.y - cyy);
            f->vx.push_back(ax); f->vy.push_back(ay);
            r2 = std::max(r2, ax * ax + ay * ay);
        }
        f->rmax[i] = std::sqrt(r2);
        f->voff[i + 1] = (int32_t)f->vx.size();
        double sq = std::sqrt(f->area[i]); sum_sqrt += sq; min_sqrt = std::min(min_sqrt, sq);
    }
    if (prm_out) {
        prm_out->Lx = L; prm_out->Ly = L;
        prm_out->modulus = 1.5e3 * (sum_sqrt / n + min_sqrt);   // Subzero.m:77
    }
    *out = f;
    return SZ_OK;
}

extern "C" int sz_field_view(const SzField* f, SzFloesSoA* v)
{
    if (!f || !v) return SZ_ERR_ARG;
    v->n = (int32_t)f->x.size(); v->nverts = (int64_t)f->vx.size();
    v->x = f->x.data(); v->y = f->y.data(); v->rmax = f->rmax.data(); v->h = f->h.data(); v->area = f->area.data();
    v->u = f->u.data(); v->v = f->v.data(); v->ksi = f->ksi.data(); v->alive = f->alive.data();
    v->voff = f->voff.data(); v->vx = f->vx.data(); v->vy = f->vy.data();
    return SZ_OK;
}

extern "C" void sz_field_free(SzField* f) { delete f; }
