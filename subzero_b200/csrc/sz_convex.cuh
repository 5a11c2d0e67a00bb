// ------------------------------------------------------------------------------------------------
// DERIVATIVE WORK NOTICE.  This file is a function-by-function re-implementation, for fixed-capacity GPU data structures,
// of the Clipper library 6.4.2 -- Copyright Angus Johnson 2010-2017, http://www.angusj.com -- which the reference vendors as
// private/clipper.cpp / clipper.hpp (an extension of Bala Vatti's clipping algorithm, CACM 35(7), 1992).  Clipper is
// distributed under the Boost Software License, Version 1.0; that licence and the attribution are reproduced in the
// NOTICE file at the root of this repository (http://www.boost.org/LICENSE_1_0.txt).
// ------------------------------------------------------------------------------------------------
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
#include <type_traits>

namespace szcvx {

using szclip::i64;
using szclip::P64;
namespace fp = szclip::fp;

enum { CV_OK = 0, CV_BAIL = 1 };
enum { NILE = -1 };

struct INode { P64 pt; int e1, e2; };

// Four-entry state arrays.  To keep them in registers on the device an entry chosen at run time has to be read through a
// select chain and written through predicated moves (every subscript is a literal; a single run-time subscript
// would send the whole array to local memory).  Measured on B200 this is SLOWER than leaving the arrays in L1-cached local
// memory (12.0 vs 10.4 ms at 1M floes: the select chains cost more issue slots than the loads they save), so plain
// subscripts are the default and SZ_CVX_REG_ARRAYS is the experiment switch.
#if defined(__CUDA_ARCH__) && defined(SZ_CVX_REG_ARRAYS)
template <class T> SZ_HD T rd4(const T (&a)[4], int e) { return e == 0 ? a[0] : (e == 1 ? a[1] : (e == 2 ? a[2] : a[3])); }
template <class T> SZ_HD void wr4(T (&a)[4], int e, T v) { if (e == 0) a[0] = v; if (e == 1) a[1] = v; if (e == 2) a[2] = v; if (e == 3) a[3] = v; }
#else
template <class T> SZ_HD T rd4(const T (&a)[4], int e) { return a[e]; }
template <class T> SZ_HD void wr4(T (&a)[4], int e, T v) { a[e] = v; }
#endif
#if defined(__CUDA_ARCH__)
#define SZ_UNROLL4 _Pragma("unroll")
#else
#define SZ_UNROLL4
#endif

// Experiment switch SZ_C_SMEM_EDGES=<mask> (device only, off by default, NOT measured yet): the chosen 8-byte fields of the
// four edge records live in dynamic shared memory instead of local memory -- bit 0 botx, 1 boty, 2 topx, 3 topy, 4 curx,
// 5 dx.  Entry e of field slot f of thread t sits at ((4 f + e) * blockDim + t) * 8: any mix of run-time subscripts across
// a warp is bank-conflict free, and the accessor is stateless (it names the extern array itself) so the accesses compile
// to LDS/STS.  Cost: popcount(mask) * 32 B per thread (all six: 96 KB per 512-thread CTA, 192 KB per SM at two CTAs, which
// leaves 64 KB of L1 for the rest of the local-memory working set).  The launcher (sz_narrow_C.cu) sizes the allocation
// with smem_edge_bytes().
#if defined(SZ_C_SMEM_EDGES)
#define SZ_C_SMEM_MASK (SZ_C_SMEM_EDGES)
#else
#define SZ_C_SMEM_MASK 0
#endif
constexpr int smem_edge_slot(int field) { int s = 0; for (int k = 0; k < field; ++k) s += (SZ_C_SMEM_MASK >> k) & 1; return s; }
constexpr int smem_edge_fields() { return smem_edge_slot(6); }
constexpr size_t smem_edge_bytes(int threads_per_block) { return (size_t)smem_edge_fields() * 4 * 8 * (size_t)threads_per_block; }
#if defined(__CUDA_ARCH__) && SZ_C_SMEM_MASK != 0
template <class T, int FIELD> struct SmemArr4 {
    static_assert(sizeof(T) == 8, "8-byte fields only");
    enum { SLOT = smem_edge_slot(FIELD) };
    __device__ __forceinline__ T& operator[](int e) const
    {
        extern __shared__ __align__(16) unsigned char sz_c_edge_smem[];
        return reinterpret_cast<T*>(sz_c_edge_smem)[(SLOT * 4 + e) * blockDim.x + threadIdx.x];
    }
};
template <class T, int F> SZ_HD T rd4(const SmemArr4<T, F>& a, int e) { return a[e]; }
template <class T, int F> SZ_HD void wr4(const SmemArr4<T, F>& a, int e, T v) { a[e] = v; }
#define SZ_EDGE_FIELD(T, name, field) typename std::conditional<((SZ_C_SMEM_MASK >> (field)) & 1) != 0, SmemArr4<T, field>, T[4]>::type name
#else
#define SZ_EDGE_FIELD(T, name, field) T name[4]
#endif

// The sweep's two pieces of bulk storage, lent by the caller: the int64 outlines (ring p at [p*NV, p*NV + n[p]) of vxs/vys)
// and the output record's deque (dqx/dqy, dcap entries).  It is handed down the (inlined) call chain by reference instead of
// living in the sweep object: pointers read back from a struct in local memory lose their address space and every access
// through them becomes a generic LD/ST instead of LDL/STL.
struct SweepMem { i64* vxs; i64* vys; i64* dqx; i64* dqy; int dcap; };

// Edge ids: 2*slot + {0 left bound, 1 right bound}; slot 0 is the path whose local minimum is popped first (larger
// bottom Y; the subject on a tie, like the stable sort of Reset :1251), slot 1 the other one.  Flags and orders live
// in registers (bit id of a mask; the AEL is four packed nibbles).
template <int NV>
struct ConvexSweep {
    // the two open rings in Clipper coordinates (SweepMem), vertex 0 = bottom vertex (largest Y, then smallest X); [0] subject, [1] clip
    int n[2];
    typedef const SweepMem& M;
    SZ_HD static P64 vat(const SweepMem& m, int p, int i) { P64 q; q.x = m.vxs[p * NV + i]; q.y = m.vys[p * NV + i]; return q; }
    int sw;                                   // slot q holds ring q ^ sw
    // current edge of every bound (clipper.cpp:66-84 TEdge, reduced like szclip::Edge)
    SZ_EDGE_FIELD(i64, botx, 0); SZ_EDGE_FIELD(i64, boty, 1); SZ_EDGE_FIELD(i64, topx, 2); SZ_EDGE_FIELD(i64, topy, 3); SZ_EDGE_FIELD(i64, curx, 4);
    SZ_EDGE_FIELD(double, dx, 5); int vi[4];      // vi: ring index of `top`
    i64 cur_y;                                // Curr.Y of every active edge = bottom of the current scanbeam
    unsigned ordp; int na;                    // AEL left to right: nibble k of ordp is the edge id at position k
    unsigned act, f_right, f_out, f_wc2, f_last, f_back;    // bit id: in the AEL / Side == right / OutIdx >= 0 /
                                              // other path's parity / NextInLML == NULL / the bound walks the ring backwards
    i64 lm_y[2]; int cur_lm;                  // the two local minima (slot order = Y descending)
    INode il[6]; int n_il;
    // output record: deque d[lo..hi], front = d[lo] (OutRec.Pts), back = d[hi] (Pts->Prev); both ends cached
    int lo, hi, n_or; P64 fr, bk;
    bool bail; int why;
    SZ_HD void set_bail(int r) { if (!bail) { bail = true; why = r; } }

    SZ_HD int ord_at(int k) const { return (int)((ordp >> (4 * k)) & 15u); }
    // position of edge e in the AEL, -1 when absent: nibble k of ordp equals e exactly when nibble k of x is zero, and
    // (x - 0x1111) & ~x & 0x8888 marks zero nibbles (its lowest mark is always a true one; an id occurs once among the first na)
    SZ_HD int pos_of(int e) const
    {
        const unsigned x = (ordp ^ ((unsigned)e * 0x1111u)) & 0xffffu;
        const unsigned z = (x - 0x1111u) & ~x & 0x8888u;
        if (z == 0) return -1;
#if defined(__CUDA_ARCH__)
        const int k = (__ffs((int)z) - 1) >> 2;
#else
        const int k = __builtin_ctz(z) >> 2;
#endif
        return k < na ? k : -1;
    }
    SZ_HD int pael(int e) const { const int k = pos_of(e); return k > 0 ? ord_at(k - 1) : NILE; }
    SZ_HD int nael(int e) const { const int k = pos_of(e); return (k >= 0 && k + 1 < na) ? ord_at(k + 1) : NILE; }
    SZ_HD bool bit(unsigned m, int e) const { return ((m >> e) & 1u) != 0; }
    SZ_HD P64 bot(int e) const { P64 p; p.x = rd4(botx, e); p.y = rd4(boty, e); return p; }
    SZ_HD P64 top(int e) const { P64 p; p.x = rd4(topx, e); p.y = rd4(topy, e); return p; }
    SZ_HD P64 cur(int e) const { P64 p; p.x = rd4(curx, e); p.y = cur_y; return p; }

    SZ_HD static i64 top_x_of(i64 bx, i64 by, i64 tx, i64 ty, double d, i64 y)     // clipper.cpp:615-619
    {
        return (y == ty) ? tx : bx + fp::round_half(fp::mul(d, fp::cvt(y - by)));
    }
    SZ_HD i64 top_x(int e, i64 y) const { return top_x_of(rd4(botx, e), rd4(boty, e), rd4(topx, e), rd4(topy, e), rd4(dx, e), y); }
    // the edge of bound `e` that follows vertex `from` (ring index) -- SetDx :591-596, InitEdge2 :729-742
    SZ_HD void load_edge(M m, int e, int from)
    {
        const int p = (e >> 1) ^ sw, nn = n[p], st = bit(f_back, e) ? -1 : 1;
        int to = from + st; if (to >= nn) to -= nn; else if (to < 0) to += nn;
        const P64 pb = vat(m, p, from), pt = vat(m, p, to);
        const i64 bx = pb.x, by = pb.y, tx = pt.x, ty = pt.y;
        wr4(botx, e, bx); wr4(boty, e, by); wr4(topx, e, tx); wr4(topy, e, ty); wr4(vi, e, to); wr4(curx, e, bx);
        if (ty >= by) { set_bail(1); return; }                           // horizontal (or not a bound of a convex path)
        wr4(dx, e, fp::div(fp::cvt(tx - bx), fp::cvt(ty - by)));
        int nx = to + st; if (nx >= nn) nx -= nn; else if (nx < 0) nx += nn;
        const i64 ny = vat(m, p, nx).y;
        if (ny == ty) { set_bail(2); return; }                           // horizontal edge at the top of this one
        if (ny > ty) f_last |= 1u << e; else f_last &= ~(1u << e);
    }
    SZ_HD bool inserts_before(int e1, int e2) const   // E2InsertsBeforeE1 :3278-3287
    {
        const i64 c1 = rd4(curx, e1), c2 = rd4(curx, e2);
        if (c2 == c1) {
            const i64 t1y = rd4(topy, e1), t2y = rd4(topy, e2);
            if (t2y > t1y) return rd4(topx, e2) < top_x(e1, t2y);
            else return rd4(topx, e1) > top_x(e2, t1y);
        }
        return c2 < c1;
    }
    SZ_HD void insert_at(int k, int e)
    {
        const unsigned lowm = (1u << (4 * k)) - 1u;
        ordp = (ordp & lowm) | ((unsigned)e << (4 * k)) | ((ordp & ~lowm) << 4);
        ++na; act |= 1u << e;
    }
    SZ_HD void insert_into_ael(int e, int start)   // :3319-3345
    {
        if (na == 0) insert_at(0, e);
        else if (start == NILE && inserts_before(ord_at(0), e)) insert_at(0, e);
        else {
            int k = (start == NILE) ? 0 : pos_of(start);
            while (k + 1 < na && !inserts_before(ord_at(k + 1), e)) ++k;
            insert_at(k + 1, e);
        }
    }
    SZ_HD void delete_from_ael(int e)
    {
        const int k = pos_of(e);
        if (k < 0) return;
        const unsigned lowm = (1u << (4 * k)) - 1u;
        ordp = (ordp & lowm) | ((ordp >> 4) & ~lowm);
        --na; act &= ~(1u << e);
    }
    SZ_HD static unsigned swap_nibbles(unsigned v, int k1, int k2)
    {
        const unsigned a = (v >> (4 * k1)) & 15u, b = (v >> (4 * k2)) & 15u;
        v &= ~((15u << (4 * k1)) | (15u << (4 * k2)));
        return v | (b << (4 * k1)) | (a << (4 * k2));
    }
    SZ_HD void swap_in_ael(int e1, int e2)
    {
        const int k1 = pos_of(e1), k2 = pos_of(e2);
        if (k1 < 0 || k2 < 0) return;
        ordp = swap_nibbles(ordp, k1, k2);
    }

    // ---- output record
    SZ_HD void add_out_pt(M m, int e, P64 pt)   // :2463-2499
    {
        if (!bit(f_out, e)) {
            if (n_or != 0) { set_bail(3); return; }       // a second OutRec: outside the model
            n_or = 1; lo = hi = m.dcap / 2; m.dqx[lo] = pt.x; m.dqy[lo] = pt.y; fr = pt; bk = pt;
            f_out |= 1u << e;                             // SetHoleState :2301-2324: no other output yet -> not a hole
            return;
        }
        if (!bit(f_right, e)) {
            if (pt == fr) return;
            if (lo == 0) { set_bail(4); return; }
            --lo; m.dqx[lo] = pt.x; m.dqy[lo] = pt.y; fr = pt;
        } else {
            if (pt == bk) return;
            if (hi + 1 >= m.dcap) { set_bail(5); return; }
            ++hi; m.dqx[hi] = pt.x; m.dqy[hi] = pt.y; bk = pt;
        }
    }
    SZ_HD void set_flag(unsigned& m, int e, bool v) { m = (m & ~(1u << e)) | ((v ? 1u : 0u) << e); }
    SZ_HD void add_local_min_poly(M m, int e1, int e2, P64 pt)   // :1841-1881 (the join test needs an existing OutRec: unreachable)
    {
        if (rd4(dx, e1) > rd4(dx, e2)) { add_out_pt(m, e1, pt); set_flag(f_out, e2, bit(f_out, e1)); set_flag(f_right, e1, false); set_flag(f_right, e2, true); }
        else { add_out_pt(m, e2, pt); set_flag(f_out, e1, bit(f_out, e2)); set_flag(f_right, e1, true); set_flag(f_right, e2, false); }
    }
    SZ_HD void add_local_max_poly(M m, int e1, int e2, P64 pt)   // :1884-1897 (one OutRec: the indices are equal)
    {
        add_out_pt(m, e1, pt);
        f_out &= ~((1u << e1) | (1u << e2));
    }
    SZ_HD void swap_sides_idx(int e1, int e2)
    {
        const bool s1 = bit(f_right, e1), s2 = bit(f_right, e2), o1 = bit(f_out, e1), o2 = bit(f_out, e2);
        set_flag(f_right, e1, s2); set_flag(f_right, e2, s1); set_flag(f_out, e1, o2); set_flag(f_out, e2, o1);
    }
    SZ_HD void intersect_edges(M m, int e1, int e2, P64 pt)   // :2106-2298, closed even-odd paths, ctIntersection
    {
        const bool c1 = bit(f_out, e1), c2 = bit(f_out, e2);
        const bool same = (e1 >> 1) == (e2 >> 1);
        if (!same) f_wc2 ^= (1u << e1) | (1u << e2);
        if (c1 && c2) {
            if (!same) add_local_max_poly(m, e1, e2, pt);
            else { add_out_pt(m, e1, pt); add_out_pt(m, e2, pt); swap_sides_idx(e1, e2); }
        } else if (c1) { add_out_pt(m, e1, pt); swap_sides_idx(e1, e2); }
        else if (c2) { add_out_pt(m, e2, pt); swap_sides_idx(e1, e2); }
        else {
            if (!same) add_local_min_poly(m, e1, e2, pt);
            else if (bit(f_wc2, e1) && bit(f_wc2, e2)) add_local_min_poly(m, e1, e2, pt);
        }
    }

    // ---- InsertLocalMinimaIntoAEL :1978-2077
    SZ_HD void insert_local_minima(M m, i64 bot_y)
    {
        while (cur_lm < 2 && lm_y[cur_lm] == bot_y && !bail) {
            const int p = cur_lm++;
            const int lb = 2 * p, rb = 2 * p + 1;
            wr4(curx, lb, rd4(botx, lb)); wr4(curx, rb, rd4(botx, rb));      // Curr = Bot (Reset :1262-1274)
            insert_into_ael(lb, NILE);
            insert_into_ael(rb, lb);
            {   // SetWindingCount :1624-1722 reduced to the other path's parity
                const int k = pos_of(lb);
                int q = k - 1;
                while (q >= 0 && (ord_at(q) >> 1) != p) --q;
                bool w = (q < 0) ? false : bit(f_wc2, ord_at(q));
                if (((k - 1 - q) & 1) != 0) w = !w;
                set_flag(f_wc2, lb, w); set_flag(f_wc2, rb, w);
            }
            if (bit(f_wc2, lb)) add_local_min_poly(m, lb, rb, bot(lb));
            if (bail) return;
            const int pl = pael(lb);
            if (bit(f_out, lb) && pl != NILE) {
                if (rd4(curx, pl) == rd4(botx, lb) && bit(f_out, pl) && szclip::slopes_eq4(bot(pl), top(pl), cur(lb), top(lb))) { set_bail(6); return; }   // AddJoin :2046-2055
            }
            if (nael(lb) != rb) { set_bail(7); return; }                                                                     // :2057-2076
        }
    }

    // ---- IntersectPoint :622-689 (cur_y is still the bottom of the scanbeam being processed)
    SZ_HD P64 intersect_point(int a, int b) const
    {
        P64 ip;
        const double d1 = rd4(dx, a), d2 = rd4(dx, b);
        const i64 bxa = rd4(botx, a), bya = rd4(boty, a), txa = rd4(topx, a), tya = rd4(topy, a);
        const i64 bxb = rd4(botx, b), byb = rd4(boty, b), txb = rd4(topx, b), tyb = rd4(topy, b);
        if (d1 == d2) { ip.y = cur_y; ip.x = top_x_of(bxa, bya, txa, tya, d1, ip.y); return ip; }
        else if (d1 == 0) {
            ip.x = bxa;
            double b2 = fp::sub(fp::cvt(byb), fp::div(fp::cvt(bxb), d2));
            ip.y = fp::round_half(fp::add(fp::div(fp::cvt(ip.x), d2), b2));
        } else if (d2 == 0) {
            ip.x = bxb;
            double b1 = fp::sub(fp::cvt(bya), fp::div(fp::cvt(bxa), d1));
            ip.y = fp::round_half(fp::add(fp::div(fp::cvt(ip.x), d1), b1));
        } else {
            double b1 = fp::sub(fp::cvt(bxa), fp::mul(fp::cvt(bya), d1));
            double b2 = fp::sub(fp::cvt(bxb), fp::mul(fp::cvt(byb), d2));
            double q = fp::div(fp::sub(b2, b1), fp::sub(d1, d2));
            ip.y = fp::round_half(q);
            if (fabs(d1) < fabs(d2)) ip.x = fp::round_half(fp::add(fp::mul(d1, q), b1));
            else ip.x = fp::round_half(fp::add(fp::mul(d2, q), b2));
        }
        if (ip.y < tya || ip.y < tyb) {
            ip.y = (tya > tyb) ? tya : tyb;
            ip.x = (fabs(d1) < fabs(d2)) ? top_x_of(bxa, bya, txa, tya, d1, ip.y) : top_x_of(bxb, byb, txb, tyb, d2, ip.y);
        }
        if (ip.y > cur_y) {
            ip.y = cur_y;
            ip.x = (fabs(d1) > fabs(d2)) ? top_x_of(bxb, byb, txb, tyb, d2, ip.y) : top_x_of(bxa, bya, txa, tya, d1, ip.y);
        }
        return ip;
    }
    // ---- ProcessIntersections :2827-2845 when some neighbours change places in this scanbeam.
    // gt: bit (4a + b) set when Curr.X of edge a > Curr.X of edge b at the top of the scanbeam.
    SZ_HD void intersections(M m, i64 top_y, unsigned gt)
    {
        unsigned sel = ordp; int ns = na;
        n_il = 0;
        bool modified;
        do {   // BuildIntersectList :2871-2900: bubble sort; every pass drops its last element
            modified = false;
            for (int k = 0; k + 1 < ns; ++k) {
                const int e = (int)((sel >> (4 * k)) & 15u), en = (int)((sel >> (4 * k + 4)) & 15u);
                if ((gt >> (4 * e + en)) & 1u) {
                    P64 pt = intersect_point(e, en);
                    if (pt.y < top_y) { pt.x = top_x(e, top_y); pt.y = top_y; }
                    if (n_il >= 6) { set_bail(8); return; }
                    il[n_il].e1 = e; il[n_il].e2 = en; il[n_il].pt = pt; ++n_il;
                    sel = swap_nibbles(sel, k, k + 1);
                    modified = true;
                }
            }
            if (ns > 1) --ns; else break;
        } while (modified);
        if (n_il > 1) {
            // FixupIntersectionOrder :2934-2954: std::sort by Y descending (<= 6 elements: libstdc++ runs a stable
            // insertion sort), then make every intersection one of adjacent edges
            sel = ordp;
            for (int i = 1; i < n_il; ++i) {
                INode v = il[i]; int k = i - 1;
                while (k >= 0 && il[k].pt.y < v.pt.y) { il[k + 1] = il[k]; --k; }
                il[k + 1] = v;
            }
            for (int i = 0; i < n_il; ++i) {
                int j = i, k1 = 0, k2 = 0;
                for (; j < n_il; ++j) {
                    k1 = k2 = -9;
                    for (int k = 0; k < na; ++k) { const int id = (int)((sel >> (4 * k)) & 15u); if (id == il[j].e1) k1 = k; if (id == il[j].e2) k2 = k; }
                    if (k1 - k2 == 1 || k2 - k1 == 1) break;
                }
                if (j == n_il) { set_bail(9); return; }
                if (j != i) { INode t = il[i]; il[i] = il[j]; il[j] = t; }
                sel = swap_nibbles(sel, k1, k2);
            }
        }
        for (int i = 0; i < n_il && !bail; ++i) {          // ProcessIntersectList :2906-2918
            intersect_edges(m, il[i].e1, il[i].e2, il[i].pt);
            swap_in_ael(il[i].e1, il[i].e2);
        }
        n_il = 0;
    }

    SZ_HD void do_maxima(M m, int e)   // :2957-3006
    {
        const int mp = e ^ 1;
        // GetMaximaPairEx :2548-2555: the other bound of the path ends at the same top vertex and is active
        const P64 te = top(e), tm = top(mp);
        if (!(bit(act, mp) && bit(f_last, mp) && tm == te)) { set_bail(10); return; }
        int en = nael(e);
        while (en != NILE && en != mp && !bail) {
            intersect_edges(m, e, en, te);
            swap_in_ael(e, en);
            en = nael(e);
        }
        if (bail) return;
        if (!bit(f_out, e) && !bit(f_out, mp)) { delete_from_ael(e); delete_from_ael(mp); }
        else if (bit(f_out, e) && bit(f_out, mp)) { add_local_max_poly(m, e, mp, te); delete_from_ael(e); delete_from_ael(mp); }
        else set_bail(11);      // "DoMaxima error": let the general sweep report it
    }

    // ---- stepwise interface (the caller keeps the lanes of a warp / CTA together between scanbeams)
    // before begin(): fill vx/vy/n (load_ring).  (wx, wy)[0, wcap) is scratch for the output record.
    template <class G> SZ_HD void load_ring(M m, int p, const G& get, int cnt)
    {
        n[p] = cnt;
        for (int i = 0; i < cnt && i < NV; ++i) { const P64 q = get(i); m.vxs[p * NV + i] = q.x; m.vys[p * NV + i] = q.y; }
    }
    SZ_HD bool begin(M m)
    {
        lo = hi = 0; n_or = 0; na = 0; n_il = 0; bail = false; why = 0; cur_lm = 0;
        ordp = 0; act = f_right = f_out = f_wc2 = f_last = f_back = 0; cur_y = 0; fr.x = fr.y = bk.x = bk.y = 0; sw = 0;
        SZ_UNROLL4
        for (int id = 0; id < 4; ++id) { botx[id] = boty[id] = topx[id] = topy[id] = curx[id] = 0; dx[id] = 0; vi[id] = 0; }
        if (n[0] < 3 || n[1] < 3 || n[0] > NV || n[1] > NV) { set_bail(17); return false; }
        // Reset :1247-1276: minima sorted by Y descending (std::sort of two elements is stable: the subject on a tie)
        sw = (vat(m, 1, 0).y > vat(m, 0, 0).y) ? 1 : 0;
        SZ_UNROLL4
        for (int q = 0; q < 2; ++q) {
            // the two bounds of the path's single local minimum (AddPath :1172-1219): e = forward edge, e.prev = backward
            // edge; the left bound is the one with the larger Dx (:1192-1203)
            const int l = 2 * q, r = 2 * q + 1;
            f_back |= 1u << r;
            load_edge(m, l, 0); load_edge(m, r, 0);
            if (bail) return false;
            if (dx[l] < dx[r]) {
                f_back ^= (1u << l) | (1u << r);
                i64 t; double d; int u;
                t = topx[l]; topx[l] = topx[r]; topx[r] = t; t = topy[l]; topy[l] = topy[r]; topy[r] = t;
                d = dx[l]; dx[l] = dx[r]; dx[r] = d; u = vi[l]; vi[l] = vi[r]; vi[r] = u;
                const bool ll = bit(f_last, l), lr = bit(f_last, r); set_flag(f_last, l, lr); set_flag(f_last, r, ll);
            }
            f_right |= 1u << r;
            lm_y[q] = boty[l];
        }
        cur_y = lm_y[0];
        insert_local_minima(m, cur_y);
        if (bail) return false;
        // Scanbeams above which only the first path is active (its two bounds, not contributing: nothing can be
        // output and nothing is inserted) are walked here, outside the lane-synchronous loop: per scanbeam the sweep
        // only promotes the bound(s) whose top is reached (:3086-3111).  The one other thing it could do -- swap the two
        // bounds because their rounded TopX values invert (within a grid unit of the path's top) -- is outside the model.
        if (cur_lm == 1) {
            const i64 y2 = lm_y[1];
            for (;;) {
                const i64 ty = topy[0] > topy[1] ? topy[0] : topy[1];
                if (ty <= y2) break;
                if (top_x(0, ty) > top_x(1, ty)) { set_bail(18); return false; }
                cur_y = ty;
                if ((bit(f_last, 0) && topy[0] == ty) || (bit(f_last, 1) && topy[1] == ty)) {
                    // the first path ends below the bottom of the second: DoMaxima removes both bounds, the rest of the
                    // sweep sees the second path alone and outputs nothing
                    na = 0; act = 0; cur_lm = 2; return false;
                }
                if (topy[0] == ty) load_edge(m, 0, vi[0]);
                if (topy[1] == ty && !bail) load_edge(m, 1, vi[1]);
                if (bail) return false;
            }
        }
        return true;
    }
    // one scanbeam: PopScanbeam, ProcessIntersections, ProcessEdgesAtTopOfScanbeam, InsertLocalMinimaIntoAEL.
    // Returns false when the sweep is over (or bailed).
    SZ_HD bool step(M m)
    {
        if (bail) return false;
        // PopScanbeam: the largest pending Y = tops of the active edges and pending local minima
        bool have = false; i64 top_y = 0;
        SZ_UNROLL4
        for (int id = 0; id < 4; ++id) if (bit(act, id)) { const i64 y = topy[id]; if (!have || y > top_y) { top_y = y; have = true; } }
        if (cur_lm < 2) { const i64 y = lm_y[1]; if (!have || y > top_y) { top_y = y; have = true; } }
        if (!have) return false;
        // BuildIntersectList :2863-2868: Curr.X of every active edge at the top of the scanbeam
        SZ_UNROLL4
        for (int id = 0; id < 4; ++id) curx[id] = bit(act, id) ? top_x(id, top_y) : 0;
        // Curr.X in AEL order; positions behind the list compare as "in order"
        const i64 big = 0x7fffffffffffffffLL;
        const i64 c0 = na > 0 ? rd4(curx, ord_at(0)) : big, c1 = na > 1 ? rd4(curx, ord_at(1)) : big, c2 = na > 2 ? rd4(curx, ord_at(2)) : big, c3 = na > 3 ? rd4(curx, ord_at(3)) : big;
        const bool i01 = c0 > c1, i12 = c1 > c2, i23 = c2 > c3;
        const bool inv = i01 || i12 || i23;
        // The usual scanbeam with an intersection has exactly ONE: two neighbours change places and fit between their outer
        // neighbours afterwards.  BuildIntersectList's bubble sort (:2871-2900) then swaps that one pair in its first pass and
        // nothing in its second, the list has one node, FixupIntersectionOrder (:2934) is not entered: the node is processed here
        // directly, without the 12-comparison order matrix and the sort (they were 9 % of class C's instructions at 5 of 32 lanes).
        bool simple = false; int ks = 0;
#if !defined(SZ_CVX_NO_SIMPLE_IL)
        if (((int)i01 + (int)i12 + (int)i23) == 1) {
            ks = i01 ? 0 : (i12 ? 1 : 2);
            const i64 lft = i01 ? -big - 1 : (i12 ? c0 : c1), a = i01 ? c0 : (i12 ? c1 : c2), b = i01 ? c1 : (i12 ? c2 : c3), rgt = i01 ? c2 : (i12 ? c3 : big);
            simple = lft <= b && a <= rgt;
        }
#endif
        if (simple) {
            const int e = ord_at(ks), en = ord_at(ks + 1);
            P64 pt = intersect_point(e, en);
            if (pt.y < top_y) { pt.x = top_x(e, top_y); pt.y = top_y; }
            intersect_edges(m, e, en, pt);                       // ProcessIntersectList :2906-2918
            ordp = swap_nibbles(ordp, ks, ks + 1);
            if (bail) return false;
        } else if (inv) {
            unsigned gt = 0;
            SZ_UNROLL4
            for (int a = 0; a < 4; ++a) {
                SZ_UNROLL4
                for (int b = 0; b < 4; ++b) if (a != b && bit(act, a) && bit(act, b) && curx[a] > curx[b]) gt |= 1u << (4 * a + b);
            }
            intersections(m, top_y, gt);
            if (bail) return false;
        }
        cur_y = top_y;
        // ProcessEdgesAtTopOfScanbeam :3009-3113.  Curr of the edges that stay is already (TopX, topY).
        unsigned at_top = 0;
        SZ_UNROLL4
        for (int id = 0; id < 4; ++id) if (bit(act, id) && topy[id] == top_y) at_top |= 1u << id;
        if (at_top & f_last) {
            int i = 0;
            while (i < na && !bail) {
                const int e = ord_at(i);
                if (bit(at_top & f_last, e)) do_maxima(m, e);      // e and its pair leave the AEL; position i now holds the next edge
                else ++i;
            }
            if (bail) return false;
        }
        // 4. promote intermediate vertices, in AEL order
        unsigned prom = at_top & ~f_last & act;
        while (prom != 0 && !bail) {
            int k = 0;
            while (!bit(prom, ord_at(k))) ++k;
            const int e = ord_at(k);
            prom &= ~(1u << e);
            const bool o = bit(f_out, e);
            if (o) add_out_pt(m, e, top(e));
            load_edge(m, e, rd4(vi, e));                             // UpdateEdgeIntoAEL :1442-1462 (out, side, wc2 carry over)
            if (bail) return false;
            if (o) {
                const int ep = (k > 0) ? ord_at(k - 1) : NILE, en = (k + 1 < na) ? ord_at(k + 1) : NILE;
                const P64 be = bot(e);
                if (ep != NILE && rd4(curx, ep) == be.x && cur_y == be.y && bit(f_out, ep) && cur_y > rd4(topy, ep) && szclip::slopes_eq4(cur(e), top(e), cur(ep), top(ep))) { set_bail(12); return false; }
                if (en != NILE && rd4(curx, en) == be.x && cur_y == be.y && bit(f_out, en) && cur_y > rd4(topy, en) && szclip::slopes_eq4(cur(e), top(e), cur(en), top(en))) { set_bail(13); return false; }
            }
        }
        insert_local_minima(m, top_y);
        if (bail) return false;
        // with both minima inserted and one path gone, the other path's bounds are outside it for good: no edge can
        // contribute again, whatever the remaining scanbeams hold
        if (cur_lm >= 2 && ((act & 3u) == 0 || (act & 12u) == 0)) return false;
        return true;
    }
    // On CV_OK the solution ring (BuildResult order) is in (ox, oy)[0, n_out), n_out = 0 when the intersection is empty.
    SZ_HD int finish(M m, i64* ox, i64* oy, int ocap, int& n_out)
    {
        n_out = 0;
        if (bail) return CV_BAIL;
        if (n_or == 0) return CV_OK;
        const int cnt = hi - lo + 1;
        // ring in Next order from Pts: r(t) = d[lo + t].  Area :406-416
        // (both loops carry the previous vertices over in registers: one load of each deque entry per loop)
        double ar = 0;
        {
            P64 pv; pv.x = m.dqx[hi]; pv.y = m.dqy[hi];
            for (int t = 0; t < cnt; ++t) {
                P64 c; c.x = m.dqx[lo + t]; c.y = m.dqy[lo + t];
                ar = fp::add(ar, fp::mul(fp::cvt(pv.x + c.x), fp::cvt(pv.y - c.y)));
                pv = c;
            }
        }
        ar = fp::mul(ar, 0.5);
        const bool rev = !(ar > 0);        // :1594-1600: not a hole, so reversed unless the area is positive
        // FixupOutPolygon :3143-3181 would remove duplicate / collinear points: outside the model
        if (cnt < 3) { why = 14; return CV_BAIL; }
        {
            P64 A, B; A.x = m.dqx[hi]; A.y = m.dqy[hi]; B.x = m.dqx[lo]; B.y = m.dqy[lo];
            for (int t = 0; t < cnt; ++t) {
                const int nx = (t == cnt - 1) ? lo : lo + t + 1;
                P64 C; C.x = m.dqx[nx]; C.y = m.dqy[nx];
                if (B == C || B == A || szclip::slopes_eq3(A, B, C)) { why = 15; return CV_BAIL; }
                A = B; B = C;
            }
        }
        if (cnt > ocap) { why = 16; return CV_BAIL; }
        // BuildResult :3199-3217: start at Pts->Prev, walk Prev
        if (!rev) { for (int t = 0; t < cnt; ++t) { ox[t] = m.dqx[hi - t]; oy[t] = m.dqy[hi - t]; } }
        else { for (int t = 0; t < cnt; ++t) { const int c = (t == cnt - 1) ? lo : lo + 1 + t; ox[t] = m.dqx[c]; oy[t] = m.dqy[c]; } }
        n_out = cnt;
        return CV_OK;
    }
    // the whole clip for a single caller
    template <class G>
    SZ_HD int run(M m, const G& subj, int n1, const G& clip, int n2, i64* ox, i64* oy, int ocap, int& n_out)
    {
        load_ring(m, 0, subj, n1); load_ring(m, 1, clip, n2);
        bool go = begin(m);
        while (go) go = step(m);
        return finish(m, ox, oy, ocap, n_out);
    }
};

}  // namespace szcvx
