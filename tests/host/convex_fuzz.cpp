// Host-side fuzz harness (test infrastructure): the convex-pair sweep of subzero_b200/csrc/sz_convex.cuh against the
// UNMODIFIED reference Clipper 6.4.2 (oracle/_ref/libclipper_ref.so), vertex for vertex.  Every case the sweep
// accepts (CV_OK) must reproduce the reference's Paths exactly; cases it declines (CV_BAIL) are counted per input
// family -- the product re-runs those with the general sweep.
//   build: see tests/host/Makefile     run: ./convex_fuzz [cases] [seed]
#include "../../subzero_b200/csrc/sz_convex.cuh"
#include "../../subzero_b200/csrc/sz_pairforce.cuh"
#include <vector>
#include <random>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>

extern "C" int szref_clip(const int64_t* sx, const int64_t* sy, int ns, const int64_t* cx, const int64_t* cy, int nc,
                          int method, int64_t* out_x, int64_t* out_y, int out_cap, int* out_off, int off_cap);
using namespace szclip;
typedef std::vector<P64> Poly;
struct VecGet { const Poly* v; int rot; P64 operator()(int i) const { int k = i + rot; if (k >= (int)v->size()) k -= (int)v->size(); return (*v)[k]; } };

static std::mt19937_64 rng;
static double urand() { return std::uniform_real_distribution<double>(0, 1)(rng); }
static int irand(int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(rng); }

static Poly hull(std::vector<P64> p)   // Andrew's monotone chain, strict (no collinear points), counter-clockwise
{
    std::sort(p.begin(), p.end(), [](const P64& a, const P64& b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
    p.erase(std::unique(p.begin(), p.end(), [](const P64& a, const P64& b) { return a == b; }), p.end());
    const int n = (int)p.size();
    if (n < 3) return Poly();
    Poly h(2 * n); int k = 0;
    auto cross = [](const P64& o, const P64& a, const P64& b) { return (__int128)(a.x - o.x) * (b.y - o.y) - (__int128)(a.y - o.y) * (b.x - o.x); };
    for (int i = 0; i < n; ++i) { while (k >= 2 && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k; h[k++] = p[i]; }
    for (int i = n - 2, t = k + 1; i >= 0; --i) { while (k >= t && cross(h[k - 2], h[k - 1], p[i]) <= 0) --k; h[k++] = p[i]; }
    h.resize(k - 1);
    if (h.size() < 3) return Poly();
    return h;
}
static Poly random_convex(double cx, double cy, double r, int npts, double scale)
{
    std::vector<P64> p(npts);
    for (auto& q : p) { double a = 6.283185307179586 * urand(), rr = r * std::sqrt(urand()); q.x = szpf::matlab_int64((cx + rr * std::cos(a)) * scale); q.y = szpf::matlab_int64((cy + rr * std::sin(a)) * scale); }
    return hull(p);
}
// clip a convex polygon (doubles) by the half-plane a*x + b*y <= c
static void halfplane(std::vector<double>& X, std::vector<double>& Y, double a, double b, double c)
{
    std::vector<double> ox, oy; const int n = (int)X.size();
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        const double di = a * X[i] + b * Y[i] - c, dj = a * X[j] + b * Y[j] - c;
        if (di <= 0) { ox.push_back(X[i]); oy.push_back(Y[i]); }
        if ((di < 0 && dj > 0) || (di > 0 && dj < 0)) { const double t = di / (di - dj); ox.push_back(X[i] + t * (X[j] - X[i])); oy.push_back(Y[i] + t * (Y[j] - Y[i])); }
    }
    X.swap(ox); Y.swap(oy);
}
// a small Voronoi diagram in a box: cells inflated about their centroid like the benchmark field
static void voronoi_cells(int nsites, double ox, double oy, double size, double inflate, bool cw, std::vector<Poly>& cells, std::vector<std::pair<double, double>>& sites)
{
    sites.clear(); cells.clear();
    const double tilt = 0.2 + urand(), tilt_c = std::cos(tilt), tilt_s = std::sin(tilt);
    for (int i = 0; i < nsites; ++i) sites.push_back({ox + size * urand(), oy + size * urand()});
    for (int i = 0; i < nsites; ++i) {
        // a tilted bounding square, so that boundary cells get no axis-parallel edges
        const double mx = ox + 0.5 * size, my = oy + 0.5 * size, hx = 1.5 * size * tilt_c, hy = 1.5 * size * tilt_s;
        std::vector<double> X = {mx - hx + hy, mx + hx + hy, mx + hx - hy, mx - hx - hy}, Y = {my - hy - hx, my + hy - hx, my + hy + hx, my - hy + hx};
        for (int j = 0; j < nsites && X.size() >= 3; ++j) if (j != i) {
            const double a = sites[j].first - sites[i].first, b = sites[j].second - sites[i].second;
            const double c = 0.5 * (sites[j].first * sites[j].first + sites[j].second * sites[j].second - sites[i].first * sites[i].first - sites[i].second * sites[i].second);
            halfplane(X, Y, a, b, c);
        }
        Poly p;
        if (X.size() >= 3) {
            double A = 0, cx = 0, cy = 0;
            for (size_t k = 0; k < X.size(); ++k) { size_t m = (k + 1) % X.size(); double cr = X[k] * Y[m] - X[m] * Y[k]; A += cr; cx += (X[k] + X[m]) * cr; cy += (Y[k] + Y[m]) * cr; }
            cx /= 3 * A; cy /= 3 * A;
            for (size_t k = 0; k < X.size(); ++k) {
                P64 q; q.x = szpf::matlab_int64((cx + inflate * (X[k] - cx)) * 4294967296.0); q.y = szpf::matlab_int64((cy + inflate * (Y[k] - cy)) * 4294967296.0);
                p.push_back(q);
            }
            if (cw) std::reverse(p.begin(), p.end());
        }
        cells.push_back(p);
    }
}

static long g_why[32];
static long g_ok = 0, g_bail = 0, g_nonempty = 0, g_skipped = 0, g_bad = 0;
struct Fam { const char* name; long ok, bail, nonempty; };
static Fam fams[8] = {{"voronoi", 0, 0, 0}, {"voronoi-exact", 0, 0, 0}, {"random-hull", 0, 0, 0}, {"grid-hull", 0, 0, 0}, {"grid-hull-scaled", 0, 0, 0}, {"nudged-copy", 0, 0, 0}, {"far-offset", 0, 0, 0}, {"tiny-overlap", 0, 0, 0}};

static bool check(const Poly& s_in, const Poly& c_in, int fam, long caseno)
{
    // what the product does: strict convexity in Clipper coordinates, open ring rotated to its bottom vertex
    VecGet gs0{&s_in, 0}, gc0{&c_in, 0};
    if (s_in.size() < 3 || c_in.size() < 3 || !szpf::ring_is_strictly_convex(gs0, (int)s_in.size()) || !szpf::ring_is_strictly_convex(gc0, (int)c_in.size())) { ++g_skipped; return true; }
    VecGet gs{&s_in, szpf::ring_bottom_vertex(gs0, (int)s_in.size())}, gc{&c_in, szpf::ring_bottom_vertex(gc0, (int)c_in.size())};
    static szcvx::ConvexSweep<64> sw;
    static i64 ringx[128], ringy[128];
    i64 wx[64], wy[64], ox[64], oy[64]; int nout = 0;
    const szcvx::SweepMem mem{ringx, ringy, wx, wy, 64};
    const int st = sw.run(mem, gs, (int)s_in.size(), gc, (int)c_in.size(), ox, oy, 64, nout);
    if (st != szcvx::CV_OK) { ++g_bail; ++fams[fam].bail; ++g_why[sw.why & 31]; return true; }
    ++g_ok; ++fams[fam].ok;
    // the reference sees the path as stored (any start vertex, optionally closed)
    static std::vector<int64_t> rx(1 << 12), ry(1 << 12); static std::vector<int> ro(64);
    std::vector<int64_t> sx, sy, cx, cy;
    for (auto& p : s_in) { sx.push_back(p.x); sy.push_back(p.y); }
    for (auto& p : c_in) { cx.push_back(p.x); cy.push_back(p.y); }
    if (caseno % 3 == 0) { sx.push_back(sx[0]); sy.push_back(sy[0]); cx.push_back(cx[0]); cy.push_back(cy[0]); }
    const int nr = szref_clip(sx.data(), sy.data(), (int)sx.size(), cx.data(), cy.data(), (int)cx.size(), 1, rx.data(), ry.data(), 1 << 12, ro.data(), 64);
    bool ok = (nr == (nout > 0 ? 1 : 0));
    if (ok && nr == 1) ok = (ro[1] == nout) && memcmp(rx.data(), ox, 8 * nout) == 0 && memcmp(ry.data(), oy, 8 * nout) == 0;
    if (nr > 0) { ++g_nonempty; ++fams[fam].nonempty; }
    if (!ok) {
        ++g_bad;
        fprintf(stderr, "MISMATCH case %ld family %s: ref paths=%d, sweep n=%d\nsubj:", caseno, fams[fam].name, nr, nout);
        for (auto& p : s_in) fprintf(stderr, " (%lld,%lld)", p.x, p.y);
        fprintf(stderr, "\nclip:"); for (auto& p : c_in) fprintf(stderr, " (%lld,%lld)", p.x, p.y);
        fprintf(stderr, "\n");
        for (int k = 0; k < nr; ++k) { fprintf(stderr, " ref[%d]:", k); for (int v = ro[k]; v < ro[k + 1]; ++v) fprintf(stderr, " (%lld,%lld)", (long long)rx[v], (long long)ry[v]); fprintf(stderr, "\n"); }
        fprintf(stderr, " sweep:"); for (int v = 0; v < nout; ++v) fprintf(stderr, " (%lld,%lld)", (long long)ox[v], (long long)oy[v]); fprintf(stderr, "\n");
    }
    return ok;
}
static void rotate_random(Poly& p) { if (p.size() > 1) std::rotate(p.begin(), p.begin() + irand(0, (int)p.size() - 1), p.end()); }

int main(int argc, char** argv)
{
    long cases = argc > 1 ? atol(argv[1]) : 300000;
    unsigned long long seed = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
    rng.seed(seed);
    const double S = 4294967296.0;
    {   // polyclip.m:66 int64(): the one-conversion form against the general one, ties and range ends included
        long bad = 0;
        for (long k = 0; k < 4000000 && bad < 5; ++k) {
            double v;
            switch (k % 6) {
                case 0: v = (urand() - 0.5) * 9.1e15; break;
                case 1: v = (double)irand(-2000000, 2000000) + 0.5; break;
                case 2: v = std::ldexp((double)irand(-(1 << 30), 1 << 30), irand(-40, 22)); break;
                case 3: v = std::nextafter((double)irand(-1000, 1000) + 0.5, (k & 8) ? 1e300 : -1e300); break;
                case 4: v = (urand() - 0.5) * 2.0; break;
                default: v = (k & 16) ? 2251799813685248.0 - (double)irand(0, 4) * 0.5 : -2251799813685248.0 + (double)irand(0, 4) * 0.5; break;
            }
            if (szpf::matlab_int64(v) != szpf::matlab_int64_general(v)) { fprintf(stderr, "matlab_int64 mismatch at %.17g\n", v); ++bad; }
        }
        const double specials[] = {0.0, -0.0, 0.5, -0.5, 1.5, -1.5, 2.5, -2.5, 0.49999999999999994, -0.49999999999999994, 1e300, -1e300, 9.3e18, -9.3e18, std::nan("")};
        for (double v : specials) if (szpf::matlab_int64(v) != szpf::matlab_int64_general(v)) { fprintf(stderr, "matlab_int64 mismatch at %.17g\n", v); ++bad; }
        if (bad) return 3;
    }
    long t = 0, round = 0;
    while (t < cases && g_bad < 5) {
        const int fam = (int)(round++ % 8);
        if (fam <= 1) {
            // Voronoi neighbourhood: every pair of cells whose inflated outlines may touch
            std::vector<Poly> cells; std::vector<std::pair<double, double>> sites;
            const double size = 2000.0 * std::sqrt(12.0);
            const double off = (t % 16 < 8) ? 0.0 : (urand() - 0.5) * 1.8e6;
            voronoi_cells(12, off, -off * 0.7, size, fam == 0 ? 1.02 : 1.0, (t / 8) % 2 == 0, cells, sites);
            for (size_t i = 0; i < cells.size() && g_bad < 5; ++i) for (size_t j = i + 1; j < cells.size() && g_bad < 5; ++j) {
                Poly a = cells[i], b = cells[j]; rotate_random(a); rotate_random(b);
                check(a, b, fam, t); ++t;
            }
            continue;
        }
        Poly s, c;
        switch (fam) {
            case 2: { double d = 1800 * urand(), a = 6.28 * urand(), ox = (urand() - 0.5) * 2e6, oy = (urand() - 0.5) * 2e6;
                      s = random_convex(ox, oy, 1000, irand(3, 14), S); c = random_convex(ox + d * cos(a), oy + d * sin(a), 1000, irand(3, 14), S); break; }
            case 3: { std::vector<P64> p(irand(3, 9)), q(irand(3, 9)); for (auto& v : p) { v.x = irand(0, 9); v.y = irand(0, 9); } for (auto& v : q) { v.x = irand(0, 9); v.y = irand(0, 9); }
                      s = hull(p); c = hull(q); break; }
            case 4: { std::vector<P64> p(irand(3, 9)), q(irand(3, 9)); for (auto& v : p) { v.x = (i64)irand(0, 12) << 32; v.y = (i64)irand(0, 12) << 32; } for (auto& v : q) { v.x = (i64)irand(0, 12) << 32; v.y = (i64)irand(0, 12) << 32; }
                      s = hull(p); c = hull(q); break; }
            case 5: { s = random_convex(0, 0, 1500, irand(3, 12), S); c = s; if (t & 8) std::reverse(c.begin(), c.end());
                      const i64 dx = irand(-3, 3) * ((t & 16) ? (i64)1 : ((i64)1 << 31)), dy = irand(-3, 3) * ((t & 32) ? (i64)1 : ((i64)1 << 31));
                      for (auto& p : c) { p.x += dx; p.y += dy; } break; }
            case 6: { double ox = (urand() < 0.5 ? -1 : 1) * (9e5 + 1e5 * urand()), oy = (urand() - 0.5) * 2e6, d = 1500 * urand(), a = 6.28 * urand();
                      s = random_convex(ox, oy, 1200, irand(4, 19), S); c = random_convex(ox + d * cos(a), oy + d * sin(a), 900, irand(4, 19), S); break; }
            default: { // overlaps of a few grid units up to a millimetre
                      s = random_convex(0, 0, 1000, irand(3, 10), S); c = random_convex(0, 0, 1000, irand(3, 10), S);
                      i64 smax = -((i64)1 << 62), cmin = (i64)1 << 62; for (auto& p : s) smax = std::max(smax, p.x); for (auto& p : c) cmin = std::min(cmin, p.x);
                      const i64 ov = (t & 8) ? irand(0, 40) : (i64)(urand() * 4e6); for (auto& p : c) p.x += smax - cmin - ov; break; }
        }
        if (!s.empty() && !c.empty()) {
            if (t & 64) std::reverse(s.begin(), s.end());
            if (t & 128) std::reverse(c.begin(), c.end());
            rotate_random(s); rotate_random(c);
            check(s, c, fam, t);
        }
        ++t;
    }
    printf("cases=%ld accepted=%ld (nonempty %ld) bailed=%ld skipped(non-convex)=%ld mismatches=%ld\n", t, g_ok, g_nonempty, g_bail, g_skipped, g_bad);
    for (auto& f : fams) printf("  %-18s accepted %8ld  nonempty %8ld  bailed %8ld (%.3f%%)\n", f.name, f.ok, f.nonempty, f.bail, 100.0 * f.bail / std::max(1L, f.ok + f.bail));
    printf("  bail reasons:"); for (int i = 0; i < 32; ++i) if (g_why[i]) printf(" #%d:%ld", i, g_why[i]); printf("\n");
    return g_bad ? 1 : 0;
}
