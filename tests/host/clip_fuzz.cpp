// Host-side fuzz harness (test infrastructure): compiles the product's clip engine
// (subzero_b200/csrc/sz_clip.cuh) with g++ and compares it, polygon for polygon and vertex for
// vertex, with the UNMODIFIED reference Clipper 6.4.2 (oracle/_ref/libclipper_ref.so).
// Also checks the std::sort replica against libstdc++'s std::sort on tie-heavy inputs.
//   build: see tests/host/Makefile     run: ./clip_fuzz [cases] [seed]
#include "../../subzero_b200/csrc/sz_clip.cuh"
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
typedef ClipCaps<2048, 1024, 8192, 2048, 8192, 2048, 512, 3072> BigCaps;
typedef ClipEngine<BigCaps> BigEngine;

struct VecGetter { const std::vector<P64>* v; P64 operator()(int i) const { return (*v)[i]; } };
struct VecSink {
    std::vector<std::vector<P64>> paths;
    void begin_path(int) { paths.emplace_back(); }
    void point(P64 p) { paths.back().push_back(p); }
};

long g_stat_joins = 0; int g_stat_maxop = 0, g_stat_maxor = 0, g_stat_maxjn = 0, g_stat_maxsb = 0;
extern "C" int szport_clip(const int64_t* sx, const int64_t* sy, int ns, const int64_t* cx, const int64_t* cy, int nc,
                           int method, int64_t* out_x, int64_t* out_y, int out_cap, int* out_off, int off_cap)
{
    static thread_local BigEngine* eng = new BigEngine;
    std::vector<P64> s(ns), c(nc);
    for (int i = 0; i < ns; ++i) { s[i].x = sx[i]; s[i].y = sy[i]; }
    for (int i = 0; i < nc; ++i) { c[i].x = cx[i]; c[i].y = cy[i]; }
    eng->begin(method);
    VecGetter gs{&s}, gc{&c};
    eng->add_path(gs, ns, 0);
    eng->add_path(gc, nc, 1);
    int st = eng->execute();
    g_stat_joins += eng->n_jn > 0; g_stat_maxop = std::max(g_stat_maxop, eng->n_op); g_stat_maxor = std::max(g_stat_maxor, eng->n_or);
    g_stat_maxjn = std::max(g_stat_maxjn, eng->n_jn); g_stat_maxsb = std::max(g_stat_maxsb, eng->n_sb);
    if (st != ST_OK) return st == ST_CLIPPER_FAIL ? -1 : (st == ST_OVERFLOW ? -2 : -3);
    VecSink sink;
    int np = eng->emit(sink);
    if (np + 1 > off_cap) return -2;
    int pos = 0; out_off[0] = 0;
    for (int k = 0; k < np; ++k) {
        for (auto& p : sink.paths[k]) { if (pos >= out_cap) return -2; out_x[pos] = p.x; out_y[pos] = p.y; ++pos; }
        out_off[k + 1] = pos;
    }
    return np;
}

#ifndef SZ_NO_MAIN
static std::mt19937_64 rng;
static double urand() { return std::uniform_real_distribution<double>(0, 1)(rng); }
static int irand(int lo, int hi) { return std::uniform_int_distribution<int>(lo, hi)(rng); }

typedef std::vector<P64> Poly;
static Poly star(double cx, double cy, double r, int n, double jitter, double scale)
{
    Poly p(n);
    double ph = urand() * 6.283185307179586;
    for (int i = 0; i < n; ++i) {
        double a = ph - 6.283185307179586 * i / n;   // clockwise like the fixture
        double rr = r * (1 - jitter + jitter * urand());
        p[i].x = llround((cx + rr * cos(a)) * scale);
        p[i].y = llround((cy + rr * sin(a)) * scale);
    }
    return p;
}
static Poly grid_poly(int n, int span)   // random vertices on a tiny integer grid: many degeneracies
{
    Poly p(n);
    for (int i = 0; i < n; ++i) { p[i].x = irand(0, span); p[i].y = irand(0, span); }
    return p;
}
static Poly rect(i64 x0, i64 y0, i64 x1, i64 y1, bool cw)
{
    Poly p = {{x0, y0}, {x1, y0}, {x1, y1}, {x0, y1}};
    if (cw) std::reverse(p.begin(), p.end());
    return p;
}
static Poly ortho(int n, int span, i64 mul)   // rectilinear polygon: horizontals + verticals everywhere
{
    Poly p; i64 x = irand(0, span), y = irand(0, span);
    for (int i = 0; i < n; ++i) {
        p.push_back({x * mul, y * mul});
        if (i & 1) x = irand(0, span); else y = irand(0, span);
    }
    return p;
}

static long g_nonempty = 0, g_multi = 0, g_fail = 0;
static bool compare(const Poly& s, const Poly& c, int method, long caseno, const char* kind)
{
    static std::vector<int64_t> ox(1 << 16), oy(1 << 16), px(1 << 16), py(1 << 16);
    static std::vector<int> oo(4096), po(4096);
    std::vector<int64_t> sx(s.size()), sy(s.size()), cx(c.size()), cy(c.size());
    for (size_t i = 0; i < s.size(); ++i) { sx[i] = s[i].x; sy[i] = s[i].y; }
    for (size_t i = 0; i < c.size(); ++i) { cx[i] = c[i].x; cy[i] = c[i].y; }
    int nr = szref_clip(sx.data(), sy.data(), (int)s.size(), cx.data(), cy.data(), (int)c.size(), method, ox.data(), oy.data(), 1 << 16, oo.data(), 4096);
    int np = szport_clip(sx.data(), sy.data(), (int)s.size(), cx.data(), cy.data(), (int)c.size(), method, px.data(), py.data(), 1 << 16, po.data(), 4096);
    bool ok = (nr == np);
    g_nonempty += nr > 0; g_multi += nr > 1; g_fail += nr < 0;
    if (ok && nr > 0) {
        ok = memcmp(oo.data(), po.data(), sizeof(int) * (nr + 1)) == 0;
        if (ok) ok = memcmp(ox.data(), px.data(), sizeof(int64_t) * oo[nr]) == 0 && memcmp(oy.data(), py.data(), sizeof(int64_t) * oo[nr]) == 0;
    }
    if (!ok) {
        fprintf(stderr, "MISMATCH case %ld kind %s method %d: ref=%d port=%d\n", caseno, kind, method, nr, np);
        fprintf(stderr, "subj:"); for (auto& p : s) fprintf(stderr, " (%lld,%lld)", p.x, p.y); fprintf(stderr, "\nclip:");
        for (auto& p : c) fprintf(stderr, " (%lld,%lld)", p.x, p.y);
        fprintf(stderr, "\n");
        for (int k = 0; k < nr; ++k) { fprintf(stderr, " ref[%d]:", k); for (int v = oo[k]; v < oo[k + 1]; ++v) fprintf(stderr, " (%lld,%lld)", (long long)ox[v], (long long)oy[v]); fprintf(stderr, "\n"); }
        for (int k = 0; k < np; ++k) { fprintf(stderr, " port[%d]:", k); for (int v = po[k]; v < po[k + 1]; ++v) fprintf(stderr, " (%lld,%lld)", (long long)px[v], (long long)py[v]); fprintf(stderr, "\n"); }
    }
    return ok;
}

struct Tagged { long long key; int id; };
static bool sort_selftest(int rounds)
{
    for (int r = 0; r < rounds; ++r) {
        int n = (r % 7 == 0) ? irand(0, 3000) : irand(0, 80);
        int nkeys = irand(1, std::max(1, n / (1 + irand(0, 6))));
        std::vector<Tagged> a(n);
        int mode = irand(0, 3);
        for (int i = 0; i < n; ++i) {
            long long k = irand(0, nkeys);
            if (mode == 1) k = i / (1 + nkeys % 5);           // ascending runs
            if (mode == 2) k = (n - i) / (1 + nkeys % 5);     // descending runs
            a[i] = {k, i};
        }
        std::vector<Tagged> b = a;
        std::sort(a.begin(), a.end(), [](const Tagged& x, const Tagged& y) { return y.key < x.key; });
        stl_sort(b.data(), n, [](const Tagged& x, const Tagged& y) { return y.key < x.key; });
        for (int i = 0; i < n; ++i) if (a[i].id != b[i].id) { fprintf(stderr, "sort replica mismatch n=%d at %d\n", n, i); return false; }
    }
    return true;
}

int main(int argc, char** argv)
{
    long cases = argc > 1 ? atol(argv[1]) : 200000;
    unsigned long long seed = argc > 2 ? strtoull(argv[2], 0, 10) : 1;
    rng.seed(seed);
    if (!sort_selftest(4000)) return 2;
    const double S = 4294967296.0;
    long bad = 0;
    for (long t = 0; t < cases; ++t) {
        int kind = (int)(t % 10);
        Poly s, c; const char* name = "";
        int method = (t / 10) % 4 == 3 ? 0 : 1;            // mostly intersection, 1 in 4 difference
        if ((t / 40) % 8 == 7) method = 2 + (t & 1);       // a few xor/union
        switch (kind) {
            case 0: { name = "hex-pair"; double d = 1000 + 1500 * urand(); double a = 6.28 * urand();
                      double ox = (urand() - 0.5) * 2e6, oy = (urand() - 0.5) * 2e6;
                      s = star(ox, oy, 1000, irand(3, 9), 0.3, S); c = star(ox + d * cos(a), oy + d * sin(a), 1000, irand(3, 9), 0.3, S); break; }
            case 1: { name = "concave-pair"; double d = 3000 * urand(); double a = 6.28 * urand();
                      s = star(0, 0, 2000, irand(8, 60), 0.7, S); c = star(d * cos(a), d * sin(a), 2000, irand(8, 60), 0.7, S); break; }
            case 2: { name = "grid-small"; s = grid_poly(irand(3, 8), 6); c = grid_poly(irand(3, 8), 6); break; }
            case 3: { name = "grid-mid"; s = grid_poly(irand(3, 14), 12); c = grid_poly(irand(3, 14), 12); break; }
            case 4: { name = "ortho"; s = ortho(2 * irand(2, 8), 8, 1); c = ortho(2 * irand(2, 8), 8, 1); break; }
            case 5: { name = "rects"; i64 m = (t & 64) ? (i64)1 << 32 : 1;
                      s = rect(irand(0, 5) * m, irand(0, 5) * m, irand(6, 10) * m, irand(6, 10) * m, t & 16);
                      c = rect(irand(0, 8) * m, irand(0, 8) * m, irand(9, 14) * m, irand(9, 14) * m, t & 32); break; }
            case 6: { name = "shared-edge"; // two stars sharing vertices exactly, one nudged by a few units
                      s = star(0, 0, 1500, irand(4, 12), 0.4, S); c = s; std::reverse(c.begin(), c.end());
                      i64 dx = irand(-3, 3) * ((t & 128) ? (i64)1 : ((i64)1 << 31)), dy = irand(-3, 3) * ((t & 256) ? (i64)1 : ((i64)1 << 31));
                      for (auto& p : c) { p.x += dx; p.y += dy; } if (t & 512) { c.erase(c.begin() + irand(0, (int)c.size() - 1)); } break; }
            case 7: { name = "big-concave"; double d = 4000 * urand(); s = star(0, 0, 3000, irand(100, 400), 0.8, S); c = star(d, 0.3 * d, 3000, irand(100, 400), 0.8, S); break; }
            case 8: { name = "grid-scaled"; s = grid_poly(irand(3, 10), 8); c = grid_poly(irand(3, 10), 8);
                      for (auto& p : s) { p.x <<= 32; p.y <<= 32; } for (auto& p : c) { p.x <<= 32; p.y <<= 32; } break; }
            default: { name = "ortho-vs-star"; s = ortho(2 * irand(2, 10), 10, (i64)1 << 32); c = star(5, 5, 4, irand(3, 12), 0.5, S); break; }
        }
        if (t % 97 == 0 && !s.empty()) s.push_back(s[0]);    // closed input (first vertex repeated), like c_alpha
        if (t % 89 == 0 && c.size() > 2) c.insert(c.begin() + 1, c[1]);   // duplicate vertex
        if (!compare(s, c, method, t, name)) { if (++bad > 5) break; }
    }
    printf("cases=%ld mismatches=%ld nonempty=%ld multi=%ld ref_fail=%ld with_joins=%ld max_op=%d max_or=%d max_jn=%d\n", cases, bad, g_nonempty, g_multi, g_fail, g_stat_joins, g_stat_maxop, g_stat_maxor, g_stat_maxjn);
    return bad ? 1 : 0;
}
#endif
