// sz_convex.cuh -- Clipper 6.4.2's scan-beam sweep specialised to ONE call pattern: the intersection
// (ctIntersection, even-odd) of two STRICTLY CONVEX closed paths without horizontal edges.
//
// This is not a different algorithm: it executes the reference's own sequence of steps
// (private/clipper.cpp ExecuteInternal :1560-1621 -- InsertLocalMinimaIntoAEL :1978, ProcessIntersections
// :2827 with BuildIntersectList :2856 / FixupIntersectionOrder :2934, IntersectEdges :2106,
// ProcessEdgesAtTopOfScanbeam :3009 with DoMaxima :2957, AddLocalMinPoly :1841, AddLocalMaxPoly :1884, AddOutPt
// :2463, the orientation fix :1594-1600, FixupOutPolygon :3143, BuildResult :3199) with the same FP64
// TopX / IntersectPoint / Round arithmetic (sz_clip.cuh fp::), but on the tiny state such an input needs:
//   * a convex path has ONE local minimum and two bounds, so the active edge list holds at most four edges
//     (left/right bound of the subject, left/right bound of the clip): four records + a 4-entry order array
//     instead of index-linked edge arenas; edges are materialised from the outline when a bound advances;
//   * the pending scanbeams are exactly {Top.Y of the active edges} U {Y of the pending local minima};
//   * the intersection of two convex paths is one output record built by front/back insertions: a deque.
// Whatever falls outside this model makes run() return CV_BAIL *before* anything is output, and the caller
// re-runs the pair with the general sweep (sz_clip.cuh): horizontal edges, a second output record, any join
// (clipper.cpp:1942 AddJoin call sites), an edge of the other path passing between the two bounds of a local
// minimum (:2059-2076), FixupIntersectionOrder / DoMaxima failures, vertices FixupOutPolygon would remove,
// buffer exhaustion.  tests/host/convex_fuzz.cpp checks it vertex for vertex against the UNMODIFIED reference
// Clipper (generic, near-degenerate, shared-edge and grid inputs) and reports the bail rate.
#pragma once
#include "sz_clip.cuh"

namespace szcvx {

using szclip::i64;
using szclip::P64;
namespace fp = szclip::fp;

enum { CV_OK = 0, CV_BAIL = 1 };
enum { NILE = -1 };

struct Act {                 // the current edge of one bound (clipper.cpp:66-84 TEdge, reduced like szclip::Edge)
    P64 bot, top;
    i64 curx, cury;
    double dx;
    short vi;                // ring index of `top`
    signed char step;        // +1: the bound walks the ring forwards, -1: backwards
    signed char side;        // 1 left, 2 right
    signed char out;         // OutIdx: 0 or -1
    signed char wc2;         // even-odd parity of the other path to the left of this edge
    signed char last;        // NextInLML == NULL (top is the path's top vertex)
    signed char pad;
};
struct INode { P64 pt; int e1, e2; };

// G: callable (int i) -> P64, vertex i of the open ring, i in [0, n), vertex 0 = bottom vertex (largest Y, then
// smallest X).  Edge ids: 2*poly + {0 left bound, 1 right bound}; poly 0 = subject, 1 = clip.
template <class G>
struct ConvexSweep {
    const G* g[2];
    int n[2];
    Act a[4];
    int ord[4], na;          // AEL, left to right
    int lm_ord[2], cur_lm; i64 lm_y[2];
    INode il[6]; int n_il;
    // output record: deque d[lo..hi] in (dx_, dy_), front = d[lo] (OutRec.Pts), back = d[hi] (Pts->Prev)
    i64* dqx; i64* dqy; int dcap, lo, hi, n_or;
    bool bail; int why;
    SZ_HD void set_bail(int r) { if (!bail) { bail = true; why = r; } }

    SZ_HD P64 vert(int p, int i) const { return (*g[p])(i); }
    SZ_HD int pos_of(int e) const { for (int k = 0; k < na; ++k) if (ord[k] == e) return k; return -1; }
    SZ_HD int pael(int e) const { const int k = pos_of(e); return k > 0 ? ord[k - 1] : NILE; }
    SZ_HD int nael(int e) const { const int k = pos_of(e); return (k >= 0 && k + 1 < na) ? ord[k + 1] : NILE; }

    SZ_HD i64 top_x(const Act& e, i64 y) const     // clipper.cpp:615-619
    {
        return (y == e.top.y) ? e.top.x : e.bot.x + fp::round_half(fp::mul(e.dx, fp::cvt(y - e.bot.y)));
    }
    // the edge of bound `e` that follows vertex `from` (ring index) -- SetDx :591-596, InitEdge2 :729-742
    SZ_HD void load_edge(int e, int from)
    {
        Act& r = a[e];
        const int p = e >> 1, nn = n[p];
        int to = from + r.step; if (to >= nn) to -= nn; else if (to < 0) to += nn;
        r.bot = vert(p, from); r.top = vert(p, to); r.vi = (short)to;
        if (r.top.y >= r.bot.y) { set_bail(1); return; }                 // horizontal (or not a bound of a convex path)
        r.dx = fp::div(fp::cvt(r.top.x - r.bot.x), fp::cvt(r.top.y - r.bot.y));
        int nx = to + r.step; if (nx >= nn) nx -= nn; else if (nx < 0) nx += nn;
        const i64 ny = vert(p, nx).y;
        if (ny == r.top.y) { set_bail(2); return; }                      // horizontal edge at the top of this one
        r.last = (signed char)(ny > r.top.y);
        r.curx = r.bot.x; r.cury = r.bot.y;
    }
    SZ_HD bool inserts_before(const Act& e1, const Act& e2) const   // E2InsertsBeforeE1 :3278-3287
    {
        if (e2.curx == e1.curx) {
            if (e2.top.y > e1.top.y) return e2.top.x < top_x(e1, e2.top.y);
            else return e1.top.x > top_x(e2, e1.top.y);
        }
        return e2.curx < e1.curx;
    }
    SZ_HD void insert_at(int k, int e) { for (int t = na; t > k; --t) ord[t] = ord[t - 1]; ord[k] = e; ++na; }
    SZ_HD void insert_into_ael(int e, int start)   // :3319-3345
    {
        if (na == 0) insert_at(0, e);
        else if (start == NILE && inserts_before(a[ord[0]], a[e])) insert_at(0, e);
        else {
            int k = (start == NILE) ? 0 : pos_of(start);
            while (k + 1 < na && !inserts_before(a[ord[k + 1]], a[e])) ++k;
            insert_at(k + 1, e);
        }
    }
    SZ_HD void delete_from_ael(int e)
    {
        const int k = pos_of(e);
        if (k < 0) return;
        for (int t = k; t + 1 < na; ++t) ord[t] = ord[t + 1];
        --na;
    }
    SZ_HD void swap_in_ael(int e1, int e2)
    {
        const int k1 = pos_of(e1), k2 = pos_of(e2);
        if (k1 < 0 || k2 < 0) return;
        ord[k1] = e2; ord[k2] = e1;
    }

    // ---- output record
    SZ_HD void add_out_pt(int e, P64 pt)   // :2463-2499
    {
        Act& r = a[e];
        if (r.out < 0) {
            if (n_or != 0) { set_bail(3); return; }       // a second OutRec: outside the model
            n_or = 1; lo = hi = dcap / 2; dqx[lo] = pt.x; dqy[lo] = pt.y;
            r.out = 0;                                    // SetHoleState :2301-2324: no other output yet -> not a hole
            return;
        }
        if (r.side == 1) {
            if (pt.x == dqx[lo] && pt.y == dqy[lo]) return;
            if (lo == 0) { set_bail(4); return; }
            --lo; dqx[lo] = pt.x; dqy[lo] = pt.y;
        } else {
            if (pt.x == dqx[hi] && pt.y == dqy[hi]) return;
            if (hi + 1 >= dcap) { set_bail(5); return; }
            ++hi; dqx[hi] = pt.x; dqy[hi] = pt.y;
        }
    }
    SZ_HD void add_local_min_poly(int e1, int e2, P64 pt)   // :1841-1881 (the join test needs an existing OutRec: unreachable)
    {
        if (a[e1].dx > a[e2].dx) { add_out_pt(e1, pt); a[e2].out = a[e1].out; a[e1].side = 1; a[e2].side = 2; }
        else { add_out_pt(e2, pt); a[e1].out = a[e2].out; a[e1].side = 2; a[e2].side = 1; }
    }
    SZ_HD void add_local_max_poly(int e1, int e2, P64 pt)   // :1884-1897 (one OutRec: the indices are equal)
    {
        add_out_pt(e1, pt);
        a[e1].out = -1; a[e2].out = -1;
    }
    SZ_HD void swap_sides_idx(int e1, int e2)
    {
        signed char s = a[e1].side; a[e1].side = a[e2].side; a[e2].side = s;
        signed char o = a[e1].out; a[e1].out = a[e2].out; a[e2].out = o;
    }
    SZ_HD void intersect_edges(int e1, int e2, P64 pt)   // :2106-2298, closed even-odd paths, ctIntersection
    {
        const bool c1 = a[e1].out >= 0, c2 = a[e2].out >= 0;
        const bool same = (e1 >> 1) == (e2 >> 1);
        if (!same) { a[e1].wc2 ^= 1; a[e2].wc2 ^= 1; }
        if (c1 && c2) {
            if (!same) add_local_max_poly(e1, e2, pt);
            else { add_out_pt(e1, pt); add_out_pt(e2, pt); swap_sides_idx(e1, e2); }
        } else if (c1) { add_out_pt(e1, pt); swap_sides_idx(e1, e2); }
        else if (c2) { add_out_pt(e2, pt); swap_sides_idx(e1, e2); }
        else {
            if (!same) add_local_min_poly(e1, e2, pt);
            else if (a[e1].wc2 > 0 && a[e2].wc2 > 0) add_local_min_poly(e1, e2, pt);
        }
    }

    // ---- InsertLocalMinimaIntoAEL :1978-2077
    SZ_HD void insert_local_minima(i64 bot_y)
    {
        while (cur_lm < 2 && lm_y[lm_ord[cur_lm]] == bot_y && !bail) {
            const int p = lm_ord[cur_lm++];
            const int lb = 2 * p, rb = 2 * p + 1;
            insert_into_ael(lb, NILE);
            insert_into_ael(rb, lb);
            {   // SetWindingCount :1624-1722 reduced to the other path's parity
                const int k = pos_of(lb);
                int q = k - 1;
                while (q >= 0 && (ord[q] >> 1) != p) --q;
                signed char w = (q < 0) ? 0 : a[ord[q]].wc2;
                if (((k - 1 - q) & 1) != 0) w ^= 1;
                a[lb].wc2 = w; a[rb].wc2 = w;
            }
            if (a[lb].wc2 != 0) add_local_min_poly(lb, rb, a[lb].bot);
            if (bail) return;
            const int pl = pael(lb);
            if (a[lb].out >= 0 && pl != NILE) {
                const Act& q = a[pl]; const Act& l = a[lb];
                P64 lc; lc.x = l.curx; lc.y = l.cury;
                if (q.curx == l.bot.x && q.out >= 0 && szclip::slopes_eq4(q.bot, q.top, lc, l.top)) { set_bail(6); return; }   // AddJoin :2046-2055
            }
            if (nael(lb) != rb) { set_bail(7); return; }                                                                     // :2057-2076
        }
    }

    // ---- IntersectPoint :622-689
    SZ_HD P64 intersect_point(const Act& e1, const Act& e2) const
    {
        P64 ip;
        if (e1.dx == e2.dx) { ip.y = e1.cury; ip.x = top_x(e1, ip.y); return ip; }
        else if (e1.dx == 0) {
            ip.x = e1.bot.x;
            double b2 = fp::sub(fp::cvt(e2.bot.y), fp::div(fp::cvt(e2.bot.x), e2.dx));
            ip.y = fp::round_half(fp::add(fp::div(fp::cvt(ip.x), e2.dx), b2));
        } else if (e2.dx == 0) {
            ip.x = e2.bot.x;
            double b1 = fp::sub(fp::cvt(e1.bot.y), fp::div(fp::cvt(e1.bot.x), e1.dx));
            ip.y = fp::round_half(fp::add(fp::div(fp::cvt(ip.x), e1.dx), b1));
        } else {
            double b1 = fp::sub(fp::cvt(e1.bot.x), fp::mul(fp::cvt(e1.bot.y), e1.dx));
            double b2 = fp::sub(fp::cvt(e2.bot.x), fp::mul(fp::cvt(e2.bot.y), e2.dx));
            double q = fp::div(fp::sub(b2, b1), fp::sub(e1.dx, e2.dx));
            ip.y = fp::round_half(q);
            if (fabs(e1.dx) < fabs(e2.dx)) ip.x = fp::round_half(fp::add(fp::mul(e1.dx, q), b1));
            else ip.x = fp::round_half(fp::add(fp::mul(e2.dx, q), b2));
        }
        if (ip.y < e1.top.y || ip.y < e2.top.y) {
            ip.y = (e1.top.y > e2.top.y) ? e1.top.y : e2.top.y;
            ip.x = (fabs(e1.dx) < fabs(e2.dx)) ? top_x(e1, ip.y) : top_x(e2, ip.y);
        }
        if (ip.y > e1.cury) {
            ip.y = e1.cury;
            ip.x = (fabs(e1.dx) > fabs(e2.dx)) ? top_x(e2, ip.y) : top_x(e1, ip.y);
        }
        return ip;
    }
    // ---- ProcessIntersections :2827-2845
    SZ_HD void process_intersections(i64 top_y)
    {
        if (na == 0) return;
        int sel[4]; int ns = na;
        for (int k = 0; k < na; ++k) { sel[k] = ord[k]; a[ord[k]].curx = top_x(a[ord[k]], top_y); }
        n_il = 0;
        bool modified;
        do {   // BuildIntersectList :2871-2900: bubble sort; every pass drops its last element
            modified = false;
            for (int k = 0; k + 1 < ns; ++k) {
                const int e = sel[k], en = sel[k + 1];
                if (a[e].curx > a[en].curx) {
                    P64 pt = intersect_point(a[e], a[en]);
                    if (pt.y < top_y) { pt.x = top_x(a[e], top_y); pt.y = top_y; }
                    if (n_il >= 6) { set_bail(8); return; }
                    il[n_il].e1 = e; il[n_il].e2 = en; il[n_il].pt = pt; ++n_il;
                    sel[k] = en; sel[k + 1] = e;
                    modified = true;
                }
            }
            if (ns > 1) --ns; else break;
        } while (modified);
        if (n_il == 0) return;
        if (n_il > 1) {
            // FixupIntersectionOrder :2934-2954: std::sort by Y descending (<= 6 elements: libstdc++ runs a stable
            // insertion sort), then make every intersection one of adjacent edges
            for (int k = 0; k < na; ++k) sel[k] = ord[k];
            for (int i = 1; i < n_il; ++i) {
                INode v = il[i]; int k = i - 1;
                while (k >= 0 && il[k].pt.y < v.pt.y) { il[k + 1] = il[k]; --k; }
                il[k + 1] = v;
            }
            for (int i = 0; i < n_il; ++i) {
                int j = i;
                while (j < n_il && !sel_adjacent(sel, il[j])) ++j;
                if (j == n_il) { set_bail(9); return; }
                if (j != i) { INode t = il[i]; il[i] = il[j]; il[j] = t; }
                sel_swap(sel, il[i].e1, il[i].e2);
            }
        }
        for (int i = 0; i < n_il && !bail; ++i) {          // ProcessIntersectList :2906-2918
            intersect_edges(il[i].e1, il[i].e2, il[i].pt);
            swap_in_ael(il[i].e1, il[i].e2);
        }
        n_il = 0;
    }
    SZ_HD bool sel_adjacent(const int* sel, const INode& nd) const
    {
        int k1 = -1, k2 = -1;
        for (int k = 0; k < na; ++k) { if (sel[k] == nd.e1) k1 = k; if (sel[k] == nd.e2) k2 = k; }
        return k1 - k2 == 1 || k2 - k1 == 1;
    }
    SZ_HD void sel_swap(int* sel, int e1, int e2) const
    {
        int k1 = -1, k2 = -1;
        for (int k = 0; k < na; ++k) { if (sel[k] == e1) k1 = k; if (sel[k] == e2) k2 = k; }
        sel[k1] = e2; sel[k2] = e1;
    }

    // ---- ProcessEdgesAtTopOfScanbeam :3009-3113
    SZ_HD void do_maxima(int e)   // :2957-3006
    {
        const int mp = e ^ 1;
        // GetMaximaPairEx :2548-2555: the other bound of the path ends at the same top vertex and is active
        if (!(pos_of(mp) >= 0 && a[mp].last && a[mp].top.x == a[e].top.x && a[mp].top.y == a[e].top.y)) { set_bail(10); return; }
        int en = nael(e);
        while (en != NILE && en != mp && !bail) {
            intersect_edges(e, en, a[e].top);
            swap_in_ael(e, en);
            en = nael(e);
        }
        if (bail) return;
        if (a[e].out == -1 && a[mp].out == -1) { delete_from_ael(e); delete_from_ael(mp); }
        else if (a[e].out >= 0 && a[mp].out >= 0) { add_local_max_poly(e, mp, a[e].top); delete_from_ael(e); delete_from_ael(mp); }
        else set_bail(11);      // "DoMaxima error": let the general sweep report it
    }
    SZ_HD void process_edges_at_top(i64 top_y)
    {
        int i = 0;
        while (i < na && !bail) {
            const int e = ord[i];
            if (a[e].top.y == top_y && a[e].last) do_maxima(e);       // e and its pair leave the AEL; position i now holds the next edge
            else { a[e].curx = top_x(a[e], top_y); a[e].cury = top_y; ++i; }
        }
        // 4. promote intermediate vertices
        for (i = 0; i < na && !bail; ++i) {
            const int e = ord[i];
            Act& r = a[e];
            if (r.top.y == top_y && !r.last) {
                const bool o = r.out >= 0;
                if (o) add_out_pt(e, r.top);
                load_edge(e, r.vi);                                   // UpdateEdgeIntoAEL :1442-1462 (out, side, wc2 carry over)
                if (bail) return;
                if (o) {
                    const int ep = (i > 0) ? ord[i - 1] : NILE, en = (i + 1 < na) ? ord[i + 1] : NILE;
                    P64 rc; rc.x = r.curx; rc.y = r.cury;
                    if (ep != NILE) {
                        const Act& q = a[ep]; P64 qc; qc.x = q.curx; qc.y = q.cury;
                        if (q.curx == r.bot.x && q.cury == r.bot.y && q.out >= 0 && q.cury > q.top.y && szclip::slopes_eq4(rc, r.top, qc, q.top)) { set_bail(12); return; }
                    }
                    if (en != NILE) {
                        const Act& q = a[en]; P64 qc; qc.x = q.curx; qc.y = q.cury;
                        if (q.curx == r.bot.x && q.cury == r.bot.y && q.out >= 0 && q.cury > q.top.y && szclip::slopes_eq4(rc, r.top, qc, q.top)) { set_bail(13); return; }
                    }
                }
            }
        }
    }

    // Runs the whole clip.  On CV_OK the solution ring (BuildResult order) is in (ox, oy)[0, n_out), n_out = 0 when
    // the intersection is empty.  (wx, wy)[0, wcap) is scratch for the output record.
    SZ_HD int run(const G& subj, int n1, const G& clip, int n2, i64* wx, i64* wy, int wcap, i64* ox, i64* oy, int ocap, int& n_out)
    {
        n_out = 0;
        why = 17; if (n1 < 3 || n2 < 3) return CV_BAIL;
        g[0] = &subj; g[1] = &clip; n[0] = n1; n[1] = n2;
        dqx = wx; dqy = wy; dcap = wcap; lo = hi = 0; n_or = 0; na = 0; n_il = 0; bail = false; why = 0; cur_lm = 0;
        for (int p = 0; p < 2; ++p) {
            // the two bounds of the path's single local minimum (AddPath :1172-1219)
            Act& f = a[2 * p]; Act& b = a[2 * p + 1];
            f.step = 1; b.step = -1;
            f.side = b.side = 0; f.out = b.out = -1; f.wc2 = b.wc2 = 0; f.pad = b.pad = 0;
            load_edge(2 * p, 0); load_edge(2 * p + 1, 0);
            if (bail) return CV_BAIL;
            // e = forward edge, e.prev = backward edge: left bound = the one with the larger Dx (:1192-1203)
            if (f.dx < b.dx) { Act t = f; f = b; b = t; }
            f.side = 1; b.side = 2;
            lm_y[p] = f.bot.y;
        }
        // Reset :1247-1276: minima sorted by Y descending (std::sort of two elements is stable)
        if (lm_y[1] > lm_y[0]) { lm_ord[0] = 1; lm_ord[1] = 0; } else { lm_ord[0] = 0; lm_ord[1] = 1; }
        i64 bot_y = lm_y[lm_ord[0]];
        insert_local_minima(bot_y);
        while (!bail) {
            // PopScanbeam: the largest pending Y = tops of the active edges and pending local minima
            if (na == 0 && cur_lm >= 2) break;
            bool have = false; i64 top_y = 0;
            for (int k = 0; k < na; ++k) { const i64 y = a[ord[k]].top.y; if (!have || y > top_y) { top_y = y; have = true; } }
            if (cur_lm < 2) { const i64 y = lm_y[lm_ord[cur_lm]]; if (!have || y > top_y) { top_y = y; have = true; } }
            process_intersections(top_y);
            if (bail) break;
            process_edges_at_top(top_y);
            if (bail) break;
            insert_local_minima(top_y);
        }
        if (bail) return CV_BAIL;
        if (n_or == 0) return CV_OK;
        const int m = hi - lo + 1;
        // ring in Next order from Pts: r(t) = d[lo + t].  Area :406-416
        double ar = 0;
        for (int t = 0; t < m; ++t) {
            const int c = lo + t, pv = (t == 0) ? hi : c - 1;
            ar = fp::add(ar, fp::mul(fp::cvt(dqx[pv] + dqx[c]), fp::cvt(dqy[pv] - dqy[c])));
        }
        ar = fp::mul(ar, 0.5);
        const bool rev = !(ar > 0);        // :1594-1600: not a hole, so reversed unless the area is positive
        // FixupOutPolygon :3143-3181 would remove duplicate / collinear points: outside the model
        if (m < 3) { why = 14; return CV_BAIL; }
        for (int t = 0; t < m; ++t) {
            const int c = lo + t, pv = (t == 0) ? hi : c - 1, nx = (t == m - 1) ? lo : c + 1;
            P64 A, B, C; A.x = dqx[pv]; A.y = dqy[pv]; B.x = dqx[c]; B.y = dqy[c]; C.x = dqx[nx]; C.y = dqy[nx];
            if (B == C || B == A || szclip::slopes_eq3(A, B, C)) { why = 15; return CV_BAIL; }
        }
        if (m > ocap) { why = 16; return CV_BAIL; }
        // BuildResult :3199-3217: start at Pts->Prev, walk Prev
        if (!rev) { for (int t = 0; t < m; ++t) { ox[t] = dqx[hi - t]; oy[t] = dqy[hi - t]; } }
        else { for (int t = 0; t < m; ++t) { const int c = (t == m - 1) ? lo : lo + 1 + t; ox[t] = dqx[c]; oy[t] = dqy[c]; } }
        n_out = m;
        return CV_OK;
    }
};

}  // namespace szcvx
