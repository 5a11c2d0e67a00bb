// ------------------------------------------------------------------------------------------------
// DERIVATIVE WORK NOTICE.  This file is a function-by-function re-implementation, for fixed-capacity GPU data structures,
// of the Clipper library 6.4.2 -- Copyright Angus Johnson 2010-2017, http://www.angusj.com -- which the reference vendors as
// private/clipper.cpp / clipper.hpp (an extension of Bala Vatti's clipping algorithm, CACM 35(7), 1992).  Clipper is
// distributed under the Boost Software License, Version 1.0; that licence and the attribution are reproduced in the
// NOTICE file at the root of this repository (http://www.boost.org/LICENSE_1_0.txt).
// ------------------------------------------------------------------------------------------------
// sz_clip.cuh -- fixed-capacity, pointer-free Vatti scan-beam polygon clipper whose output is
// bit-identical to Clipper 6.4.2 for the only call pattern SubZero's contact loop uses:
//   one closed subject path, one closed clip path, even-odd fill on both, Paths output,
//   default options (no PreserveCollinear / StrictlySimple / ReverseSolution / PolyTree).
// (collisions/floe_interactions.m:29,34,152-158 -> polyclip.m:73 -> private/mexclipper.cpp:291-298
//  -> private/clipper.cpp:1508 Execute.)
//
// It is written for one CUDA thread per clip: every container of the reference (heap-allocated
// TEdge arrays, OutPt rings, std::vector/priority_queue/list) becomes an index-linked array in a
// caller-provided arena that can live in registers/local memory, shared memory or HBM scratch.
// There is no allocation, recursion or exception; capacity exhaustion and Clipper's own failure
// paths are reported through `status`.  Because the call pattern is closed + even-odd only, the
// winding bookkeeping collapses to one parity bit per edge (see wc2 below), and the open-path,
// Skip-edge, maxima-list and PolyTree branches of the reference are unreachable and absent.
//
// What must match the reference exactly, and where it lives here:
//   * 128-bit slope comparisons         -> slopes_eq*()          (clipper.cpp:354-378, 541-575)
//   * FP64 Dx / TopX / IntersectPoint   -> fp:: helpers, never contracted to FMA
//                                                                 (clipper.cpp:136-140, 591-596, 615-689)
//   * libstdc++ std::sort tie order     -> stl_sort()            (clipper.cpp:1251, 2940)
//   * AEL/SEL/OutPt link manipulation   -> ClipEngine methods, each citing its reference lines
//
// The same header is compiled by g++ for the host-side fuzz harness in tests/ (which checks it
// against the unmodified reference Clipper) and by nvcc for the sm_100a narrow-phase kernels.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define SZ_HD __host__ __device__ __forceinline__
#define SZ_HDN __host__ __device__ __noinline__
#else
#define SZ_HD inline
#define SZ_HDN inline
#endif
// medium-size engine methods: force-inlined by default (measured faster on B200: fewer spills at call
// boundaries), __noinline__ when SZ_COMPACT_CODE is defined (3.4x smaller SASS)
#if defined(__CUDACC__) && defined(SZ_COMPACT_CODE)
#define SZ_HDM __host__ __device__ __noinline__
#else
#define SZ_HDM SZ_HD
#endif

namespace szclip {

typedef long long i64;
typedef unsigned long long u64;
typedef short idx_t;              // link type inside an arena (capacities stay < 32767)
static const idx_t NIL = -1;

struct P64 { i64 x, y; };
SZ_HD bool operator==(const P64& a, const P64& b) { return a.x == b.x && a.y == b.y; }
SZ_HD bool operator!=(const P64& a, const P64& b) { return a.x != b.x || a.y != b.y; }

enum ClipOp { OP_DIFFERENCE = 0, OP_INTERSECTION = 1, OP_XOR = 2, OP_UNION = 3 };  // polyclip.m:52-58 numbering
enum Status {
    ST_OK = 0,
    ST_CLIPPER_FAIL = -1,   // reference Execute() would return false / throw ("Clipper Error.")
    ST_OVERFLOW = -2,       // an arena capacity was exceeded: rerun in a larger size class
    ST_RANGE = -3           // coordinate beyond hiRange (clipper.cpp:896-908)
};

// ---------------------------------------------------------------------------------------------
// FP64 helpers: every operation individually rounded (no FMA), as in the reference binary.
// ---------------------------------------------------------------------------------------------
namespace fp {
#if defined(__CUDA_ARCH__)
SZ_HD double mul(double a, double b) { return __dmul_rn(a, b); }
SZ_HD double add(double a, double b) { return __dadd_rn(a, b); }
SZ_HD double sub(double a, double b) { return __dsub_rn(a, b); }
SZ_HD double div(double a, double b) { return __ddiv_rn(a, b); }
SZ_HD double cvt(i64 v) { return __ll2double_rn(v); }
#else
SZ_HD double mul(double a, double b) { return a * b; }   // host TU is built with -ffp-contract=off
SZ_HD double add(double a, double b) { return a + b; }
SZ_HD double sub(double a, double b) { return a - b; }
SZ_HD double div(double a, double b) { return a / b; }
SZ_HD double cvt(i64 v) { return (double)v; }
#endif
// clipper.cpp:136-140: add +-0.5 then truncate toward zero (not llround)
SZ_HD i64 round_half(double v) { return (v < 0) ? (i64)sub(v, 0.5) : (i64)add(v, 0.5); }
}  // namespace fp

#define SZ_HORIZONTAL (-1.0E+40)
static const i64 SZ_HI_RANGE = 0x3FFFFFFFFFFFFFFFLL;

// exact signed 128-bit product comparison  a*b == c*d   (clipper.cpp:354-378 Int128Mul + operator==)
SZ_HD bool prod_eq(i64 a, i64 b, i64 c, i64 d)
{
#if defined(__CUDA_ARCH__)
    return ((u64)a * (u64)b == (u64)c * (u64)d) && (__mul64hi(a, b) == __mul64hi(c, d));
#else
    return (__int128)a * (__int128)b == (__int128)c * (__int128)d;
#endif
}
// clipper.cpp:554-563 (three points) and :566-575 (four points); full-range form always
SZ_HD bool slopes_eq3(P64 p1, P64 p2, P64 p3) { return prod_eq(p1.y - p2.y, p2.x - p3.x, p1.x - p2.x, p2.y - p3.y); }
SZ_HD bool slopes_eq4(P64 p1, P64 p2, P64 p3, P64 p4) { return prod_eq(p1.y - p2.y, p3.x - p4.x, p1.x - p2.x, p3.y - p4.y); }

// clipper.cpp:584-588
SZ_HD double dx_of(P64 a, P64 b)
{
    return (a.y == b.y) ? SZ_HORIZONTAL : fp::div(fp::cvt(b.x - a.x), fp::cvt(b.y - a.y));
}

// ---------------------------------------------------------------------------------------------
// libstdc++ std::sort, reproduced step for step (introsort: median-of-3 quicksort down to 16
// elements, heapsort when the depth limit hits, then one insertion pass) so that elements with
// equal keys end up in the same order as in the reference binary.  `less(a,b)` is the comparator.
// ---------------------------------------------------------------------------------------------
template <class T, class Less>
struct StlSort {
    T* a; Less less;
    SZ_HD StlSort(T* a_, Less l) : a(a_), less(l) {}
    SZ_HD void swp(int i, int j) { T t = a[i]; a[i] = a[j]; a[j] = t; }
    SZ_HD void linear_insert(int last)   // __unguarded_linear_insert
    {
        T val = a[last]; int next = last - 1;
        while (less(val, a[next])) { a[last] = a[next]; last = next; --next; }
        a[last] = val;
    }
    SZ_HD void insertion(int first, int last)   // __insertion_sort
    {
        if (first == last) return;
        for (int i = first + 1; i != last; ++i) {
            if (less(a[i], a[first])) { T val = a[i]; for (int k = i; k > first; --k) a[k] = a[k - 1]; a[first] = val; }
            else linear_insert(i);
        }
    }
    SZ_HD void push_heap(int first, int hole, int top, T val)
    {
        int parent = (hole - 1) / 2;
        while (hole > top && less(a[first + parent], val)) { a[first + hole] = a[first + parent]; hole = parent; parent = (hole - 1) / 2; }
        a[first + hole] = val;
    }
    SZ_HD void adjust_heap(int first, int hole, int len, T val)
    {
        const int top = hole; int child = hole;
        while (child < (len - 1) / 2) {
            child = 2 * (child + 1);
            if (less(a[first + child], a[first + (child - 1)])) child--;
            a[first + hole] = a[first + child]; hole = child;
        }
        if ((len & 1) == 0 && child == (len - 2) / 2) {
            child = 2 * (child + 1);
            a[first + hole] = a[first + (child - 1)]; hole = child - 1;
        }
        push_heap(first, hole, top, val);
    }
    SZ_HD void heap_sort(int first, int last)   // __partial_sort(first,last,last)
    {
        int len = last - first;
        if (len >= 2) {
            int parent = (len - 2) / 2;
            for (;;) { T v = a[first + parent]; adjust_heap(first, parent, len, v); if (parent == 0) break; parent--; }
        }
        while (last - first > 1) {
            --last;
            T v = a[last]; a[last] = a[first];
            adjust_heap(first, 0, last - first, v);
        }
    }
    SZ_HD int partition_pivot(int first, int last)
    {
        int mid = first + (last - first) / 2;
        int A = first + 1, B = mid, C = last - 1;   // __move_median_to_first(first, A, B, C)
        if (less(a[A], a[B])) {
            if (less(a[B], a[C])) swp(first, B);
            else if (less(a[A], a[C])) swp(first, C);
            else swp(first, A);
        } else if (less(a[A], a[C])) swp(first, A);
        else if (less(a[B], a[C])) swp(first, C);
        else swp(first, B);
        int f = first + 1, l = last;               // __unguarded_partition(first+1, last, first)
        for (;;) {
            while (less(a[f], a[first])) ++f;
            --l;
            while (less(a[first], a[l])) --l;
            if (!(f < l)) return f;
            swp(f, l);
            ++f;
        }
    }
    SZ_HDM void sort(int n)
    {
        if (n <= 0) return;
        if (n > 16) {
            int lg = 0; for (int t = n; t > 1; t >>= 1) ++lg;
            // explicit stack replaces the recursion on the right-hand part
            int stk_first[32], stk_last[32], stk_depth[32]; int sp = 0;
            stk_first[0] = 0; stk_last[0] = n; stk_depth[0] = lg * 2; sp = 1;
            while (sp > 0) {
                --sp;
                int first = stk_first[sp], last = stk_last[sp], depth = stk_depth[sp];
                // the reference recurses into [cut,last) BEFORE continuing with [first,cut); the two
                // ranges are disjoint, so processing order does not change the outcome.
                while (last - first > 16) {
                    if (depth == 0) { heap_sort(first, last); break; }
                    --depth;
                    int cut = partition_pivot(first, last);
                    if (sp < 32) { stk_first[sp] = cut; stk_last[sp] = last; stk_depth[sp] = depth; ++sp; }
                    last = cut;
                }
            }
            insertion(0, 16);
            for (int i = 16; i != n; ++i) linear_insert(i);
        } else {
            insertion(0, n);
        }
    }
};
template <class T, class Less> SZ_HD void stl_sort(T* a, int n, Less less) { StlSort<T, Less> s(a, less); s.sort(n); }

// ---------------------------------------------------------------------------------------------
// Arena records
// ---------------------------------------------------------------------------------------------
struct Edge {                 // clipper.cpp:66-84 TEdge, with winding state reduced to one bit
    P64 bot, cur, top;
    double dx;
    idx_t next, prev;         // polygon ring
    idx_t nlml;               // next edge of the bound (NextInLML)
    idx_t nael, pael;         // active edge list
    idx_t nsel, psel;         // sorted edge list (also the pending-horizontal stack)
    idx_t out;                // OutIdx: >=0 output record, -1 unassigned
    signed char poly;         // 0 subject, 1 clip
    signed char side;         // 1 left, 2 right
    signed char wc2;          // even-odd parity of the OTHER polygon type to the left of this edge
    signed char pad;
};
struct OutPt { P64 pt; idx_t rec, next, prev, pad; };                 // clipper.cpp:112-117
struct OutRec { idx_t idx, first_left, pts, bottom; bool hole; };      // clipper.cpp:102-110
struct LocMin { i64 y; idx_t left, right; };                           // clipper.cpp:92-96
struct INode { P64 pt; idx_t e1, e2; };                                // clipper.cpp:86-90
struct Join { P64 off; idx_t op1, op2; };                              // clipper.cpp:119-123

template <int E_, int LM_, int OP_, int OR_, int IN_, int J_, int GJ_, int SB_>
struct ClipCaps { enum { E = E_, LM = LM_, OP = OP_, OR = OR_, IN = IN_, J = J_, GJ = GJ_, SB = SB_ }; };

template <class C>
struct ClipEngine {
    Edge   ed[C::E];
    OutPt  op[C::OP];
    OutRec orec[C::OR];
    LocMin lm[C::LM];
    INode  il[C::IN];
    Join   jn[C::J];
    Join   gj[C::GJ];
    i64    sb[C::SB];          // scanbeam Ys (see insert_scanbeam)
    int n_ed, n_op, n_or, n_lm, n_il, n_jn, n_gj, n_sb;   // n_sb: scanbeam insertions of this sweep
    int sb_lo, sb_hi;
    int cur_lm;
    idx_t ael, sel;            // m_ActiveEdges, m_SortedEdges
    int clip_op;
    int status;

    // ---------------------------------------------------------------- setup
    SZ_HD void begin(int op_)
    {
        n_ed = n_op = n_or = n_lm = n_il = n_jn = n_gj = n_sb = 0;
        cur_lm = 0; ael = sel = NIL; clip_op = op_; status = ST_OK; sb_lo = sb_hi = C::SB;
    }
    SZ_HD bool is_horz(idx_t e) const { return ed[e].dx == SZ_HORIZONTAL; }
    SZ_HD void fail(int st) { if (status == ST_OK) status = st; }

    // clipper.cpp:615-619
    SZ_HD i64 top_x(idx_t e, i64 y) const
    {
        const Edge& g = ed[e];
        return (y == g.top.y) ? g.top.x : g.bot.x + fp::round_half(fp::mul(g.dx, fp::cvt(y - g.bot.y)));
    }
    SZ_HD void reverse_horizontal(idx_t e) { i64 t = ed[e].top.x; ed[e].top.x = ed[e].bot.x; ed[e].bot.x = t; }   // :756-765

    // clipper.cpp:911-925
    SZ_HDM idx_t find_next_loc_min(idx_t e)
    {
        for (;;) {
            while (ed[e].bot != ed[ed[e].prev].bot || ed[e].cur == ed[e].top) e = ed[e].next;
            if (!is_horz(e) && !is_horz(ed[e].prev)) break;
            while (is_horz(ed[e].prev)) e = ed[e].prev;
            idx_t e2 = e;
            while (is_horz(e)) e = ed[e].next;
            if (ed[e].top.y == ed[ed[e].prev].bot.y) continue;   // just an intermediate horizontal
            if (ed[ed[e2].prev].bot.x < ed[e].bot.x) e = e2;
            break;
        }
        return e;
    }

    // clipper.cpp:928-1042 without the Skip-edge (open path) branches
    SZ_HDM idx_t process_bound(idx_t e, bool fwd)
    {
        idx_t result = e, horz;
        if (is_horz(e)) {
            idx_t es = fwd ? ed[e].prev : ed[e].next;
            if (is_horz(es)) {
                if (ed[es].bot.x != ed[e].bot.x && ed[es].top.x != ed[e].bot.x) reverse_horizontal(e);
            } else if (ed[es].bot.x != ed[e].bot.x) reverse_horizontal(e);
        }
        const idx_t estart = e;
        if (fwd) {
            while (ed[result].top.y == ed[ed[result].next].bot.y) result = ed[result].next;
            if (is_horz(result)) {
                // at the top of a bound, horizontals are added to the bound only when the
                // preceding edge attaches to the horizontal's left vertex
                horz = result;
                while (is_horz(ed[horz].prev)) horz = ed[horz].prev;
                if (ed[ed[horz].prev].top.x > ed[ed[result].next].top.x) result = ed[horz].prev;
            }
            while (e != result) {
                ed[e].nlml = ed[e].next;
                if (is_horz(e) && e != estart && ed[e].bot.x != ed[ed[e].prev].top.x) reverse_horizontal(e);
                e = ed[e].next;
            }
            if (is_horz(e) && e != estart && ed[e].bot.x != ed[ed[e].prev].top.x) reverse_horizontal(e);
            result = ed[result].next;
        } else {
            while (ed[result].top.y == ed[ed[result].prev].bot.y) result = ed[result].prev;
            if (is_horz(result)) {
                horz = result;
                while (is_horz(ed[horz].next)) horz = ed[horz].next;
                if (ed[ed[horz].next].top.x == ed[ed[result].prev].top.x ||
                    ed[ed[horz].next].top.x > ed[ed[result].prev].top.x) result = ed[horz].next;
            }
            while (e != result) {
                ed[e].nlml = ed[e].prev;
                if (is_horz(e) && e != estart && ed[e].bot.x != ed[ed[e].next].top.x) reverse_horizontal(e);
                e = ed[e].prev;
            }
            if (is_horz(e) && e != estart && ed[e].bot.x != ed[ed[e].next].top.x) reverse_horizontal(e);
            result = ed[result].prev;
        }
        return result;
    }

    // clipper.cpp:1045-1221 AddPath(pg, polyType, Closed=true).  `get(i)` returns vertex i of the
    // caller's path (n vertices).  Returns false when the reference would (degenerate path).
    template <class Getter>
    SZ_HDM bool add_path(const Getter& get, int n, int poly_type)
    {
        int hi = n - 1;
        if (hi < 0) return false;
        const P64 p0 = get(0);
        while (hi > 0 && get(hi) == p0) --hi;
        while (hi > 0 && get(hi) == get(hi - 1)) --hi;
        if (hi < 2) return false;
        if (n_ed + hi + 1 > C::E) { fail(ST_OVERFLOW); return false; }

        // 1. ring of edges, Curr = vertex
        const int base = n_ed;
        for (int i = 0; i <= hi; ++i) {
            Edge& g = ed[base + i];
            P64 p = get(i);
            if (p.x > SZ_HI_RANGE || p.y > SZ_HI_RANGE || -p.x > SZ_HI_RANGE || -p.y > SZ_HI_RANGE) { fail(ST_RANGE); return false; }
            g.cur = p; g.bot.x = g.bot.y = g.top.x = g.top.y = 0; g.dx = 0;
            g.next = (idx_t)(base + (i == hi ? 0 : i + 1));
            g.prev = (idx_t)(base + (i == 0 ? hi : i - 1));
            g.nlml = g.nael = g.pael = g.nsel = g.psel = NIL;
            g.out = -1; g.poly = (signed char)poly_type; g.side = 0; g.wc2 = 0; g.pad = 0;
        }
        n_ed += hi + 1;      // removed edges keep their slots, like the reference's edge array
        idx_t estart = (idx_t)base;

        // 2. remove duplicate vertices and collinear edges (:1085-1117)
        idx_t e = estart, loop_stop = estart;
        for (;;) {
            if (ed[e].cur == ed[ed[e].next].cur) {
                if (e == ed[e].next) break;
                if (e == estart) estart = ed[e].next;
                e = remove_edge(e);
                loop_stop = e;
                continue;
            }
            if (ed[e].prev == ed[e].next) break;   // only two vertices
            if (slopes_eq3(ed[ed[e].prev].cur, ed[e].cur, ed[ed[e].next].cur)) {
                if (e == estart) estart = ed[e].next;
                e = remove_edge(e);
                e = ed[e].prev;
                loop_stop = e;
                continue;
            }
            e = ed[e].next;
            if (e == loop_stop) break;
        }
        if (ed[e].prev == ed[e].next) return false;

        // 3. second stage of edge initialisation (:729-742, :1131-1139)
        bool flat = true;
        e = estart;
        do {
            Edge& g = ed[e];
            const P64 nc = ed[g.next].cur;
            if (g.cur.y >= nc.y) { g.bot = g.cur; g.top = nc; } else { g.top = g.cur; g.bot = nc; }
            const i64 dy = g.top.y - g.bot.y;                                  // SetDx :591-596
            g.dx = (dy == 0) ? SZ_HORIZONTAL : fp::div(fp::cvt(g.top.x - g.bot.x), fp::cvt(dy));
            e = g.next;
            if (flat && ed[e].cur.y != ed[estart].cur.y) flat = false;
        } while (e != estart);
        if (flat) return false;   // totally flat closed path (:1145-1151)

        // 4. bounds -> local minima (:1172-1219)
        idx_t emin = NIL;
        if (ed[ed[e].prev].bot == ed[ed[e].prev].top) e = ed[e].next;
        for (;;) {
            e = find_next_loc_min(e);
            if (e == emin) break;
            else if (emin == NIL) emin = e;
            if (n_lm >= C::LM) { fail(ST_OVERFLOW); return false; }
            LocMin m; m.y = ed[e].bot.y;
            bool left_fwd;
            if (ed[e].dx < ed[ed[e].prev].dx) { m.left = ed[e].prev; m.right = e; left_fwd = false; }
            else { m.left = e; m.right = ed[e].prev; left_fwd = true; }
            e = process_bound(m.left, left_fwd);
            idx_t e2 = process_bound(m.right, !left_fwd);
            lm[n_lm++] = m;
            if (!left_fwd) e = e2;
        }
        return true;
    }
    SZ_HD idx_t remove_edge(idx_t e)   // :745-753
    {
        ed[ed[e].prev].next = ed[e].next;
        ed[ed[e].next].prev = ed[e].prev;
        idx_t r = ed[e].next;
        ed[e].prev = NIL;
        return r;
    }

    // ---------------------------------------------------------------- scanbeam (:1335-1348)
    // The reference keeps the pending scanbeam Ys in a priority_queue and pops duplicates; here they are the sorted,
    // duplicate-free run sb[sb_lo, sb_hi), anchored at the TOP of the buffer: the sweep pops the largest Y from sb_hi and
    // almost every insertion is a new smallest Y, which lands at sb_lo - 1 without moving anything.  The run only moves
    // down, so the capacity bounds the number of insertions of one sweep (<= edges + local minima).
    SZ_HDM void insert_scanbeam(i64 y)
    {
        int k = sb_lo;
        while (k < sb_hi && sb[k] < y) ++k;
        if (k < sb_hi && sb[k] == y) return;
        if (sb_lo == 0) { fail(ST_OVERFLOW); return; }
        for (int t = sb_lo; t < k; ++t) sb[t - 1] = sb[t];
        sb[k - 1] = y; --sb_lo; ++n_sb;
    }
    SZ_HD bool pop_scanbeam(i64& y) { if (sb_lo == sb_hi) return false; y = sb[--sb_hi]; return true; }

    // ---------------------------------------------------------------- AEL / SEL plumbing
    SZ_HDM void delete_from_ael(idx_t e)   // :1367-1377
    {
        idx_t p = ed[e].pael, n = ed[e].nael;
        if (p == NIL && n == NIL && e != ael) return;
        if (p != NIL) ed[p].nael = n; else ael = n;
        if (n != NIL) ed[n].pael = p;
        ed[e].nael = NIL; ed[e].pael = NIL;
    }
    SZ_HDM void delete_from_sel(idx_t e)   // :2080-2090
    {
        idx_t p = ed[e].psel, n = ed[e].nsel;
        if (p == NIL && n == NIL && e != sel) return;
        if (p != NIL) ed[p].nsel = n; else sel = n;
        if (n != NIL) ed[n].psel = p;
        ed[e].nsel = NIL; ed[e].psel = NIL;
    }
    SZ_HDM void swap_in_ael(idx_t a, idx_t b)   // :1395-1439
    {
        if (ed[a].nael == ed[a].pael || ed[b].nael == ed[b].pael) return;
        if (ed[a].nael == b) {
            idx_t n = ed[b].nael; if (n != NIL) ed[n].pael = a;
            idx_t p = ed[a].pael; if (p != NIL) ed[p].nael = b;
            ed[b].pael = p; ed[b].nael = a; ed[a].pael = b; ed[a].nael = n;
        } else if (ed[b].nael == a) {
            idx_t n = ed[a].nael; if (n != NIL) ed[n].pael = b;
            idx_t p = ed[b].pael; if (p != NIL) ed[p].nael = a;
            ed[a].pael = p; ed[a].nael = b; ed[b].pael = a; ed[b].nael = n;
        } else {
            idx_t n = ed[a].nael, p = ed[a].pael;
            ed[a].nael = ed[b].nael; if (ed[a].nael != NIL) ed[ed[a].nael].pael = a;
            ed[a].pael = ed[b].pael; if (ed[a].pael != NIL) ed[ed[a].pael].nael = a;
            ed[b].nael = n; if (n != NIL) ed[n].pael = b;
            ed[b].pael = p; if (p != NIL) ed[p].nael = b;
        }
        if (ed[a].pael == NIL) ael = a; else if (ed[b].pael == NIL) ael = b;
    }
    SZ_HDM void swap_in_sel(idx_t a, idx_t b)   // :2558-2601
    {
        if (ed[a].nsel == NIL && ed[a].psel == NIL) return;
        if (ed[b].nsel == NIL && ed[b].psel == NIL) return;
        if (ed[a].nsel == b) {
            idx_t n = ed[b].nsel; if (n != NIL) ed[n].psel = a;
            idx_t p = ed[a].psel; if (p != NIL) ed[p].nsel = b;
            ed[b].psel = p; ed[b].nsel = a; ed[a].psel = b; ed[a].nsel = n;
        } else if (ed[b].nsel == a) {
            idx_t n = ed[a].nsel; if (n != NIL) ed[n].psel = b;
            idx_t p = ed[b].psel; if (p != NIL) ed[p].nsel = a;
            ed[a].psel = p; ed[a].nsel = b; ed[b].psel = a; ed[b].nsel = n;
        } else {
            idx_t n = ed[a].nsel, p = ed[a].psel;
            ed[a].nsel = ed[b].nsel; if (ed[a].nsel != NIL) ed[ed[a].nsel].psel = a;
            ed[a].psel = ed[b].psel; if (ed[a].psel != NIL) ed[ed[a].psel].nsel = a;
            ed[b].nsel = n; if (n != NIL) ed[n].psel = b;
            ed[b].psel = p; if (p != NIL) ed[p].nsel = b;
        }
        if (ed[a].psel == NIL) sel = a; else if (ed[b].psel == NIL) sel = b;
    }
    SZ_HDM void update_edge_into_ael(idx_t& e)   // :1442-1462
    {
        const idx_t nx = ed[e].nlml;
        if (nx == NIL) { fail(ST_CLIPPER_FAIL); return; }
        ed[nx].out = ed[e].out;
        idx_t p = ed[e].pael, n = ed[e].nael;
        if (p != NIL) ed[p].nael = nx; else ael = nx;
        if (n != NIL) ed[n].pael = nx;
        ed[nx].side = ed[e].side;
        ed[nx].wc2 = ed[e].wc2;
        e = nx;
        ed[e].cur = ed[e].bot;
        ed[e].pael = p; ed[e].nael = n;
        if (!is_horz(e)) insert_scanbeam(ed[e].top.y);
    }
    SZ_HD bool inserts_before(idx_t e1, idx_t e2)   // E2InsertsBeforeE1 :3278-3287
    {
        if (ed[e2].cur.x == ed[e1].cur.x) {
            if (ed[e2].top.y > ed[e1].top.y) return ed[e2].top.x < top_x(e1, ed[e2].top.y);
            else return ed[e1].top.x > top_x(e2, ed[e1].top.y);
        }
        return ed[e2].cur.x < ed[e1].cur.x;
    }
    SZ_HDM void insert_into_ael(idx_t e, idx_t start)   // :3319-3345
    {
        if (ael == NIL) { ed[e].pael = NIL; ed[e].nael = NIL; ael = e; }
        else if (start == NIL && inserts_before(ael, e)) { ed[e].pael = NIL; ed[e].nael = ael; ed[ael].pael = e; ael = e; }
        else {
            if (start == NIL) start = ael;
            while (ed[start].nael != NIL && !inserts_before(ed[start].nael, e)) start = ed[start].nael;
            ed[e].nael = ed[start].nael;
            if (ed[start].nael != NIL) ed[ed[start].nael].pael = e;
            ed[e].pael = start;
            ed[start].nael = e;
        }
    }
    SZ_HD void add_edge_to_sel(idx_t e)   // :1900-1917
    {
        if (sel == NIL) { sel = e; ed[e].psel = NIL; ed[e].nsel = NIL; }
        else { ed[e].nsel = sel; ed[e].psel = NIL; ed[sel].psel = e; sel = e; }
    }

    // ---------------------------------------------------------------- winding (even-odd, closed)
    // clipper.cpp:1624-1722 reduced: WindDelta is always +-1 and |WindCnt| always 1, so only the
    // opposite-type parity (WindCnt2 in {0,1}) carries information.
    SZ_HD void set_winding(idx_t edge)
    {
        idx_t e = ed[edge].pael;
        while (e != NIL && ed[e].poly != ed[edge].poly) e = ed[e].pael;
        signed char w;
        if (e == NIL) { w = 0; e = ael; }
        else { w = ed[e].wc2; e = ed[e].nael; }
        while (e != edge) { w ^= 1; e = ed[e].nael; }
        ed[edge].wc2 = w;
    }
    SZ_HD bool is_contributing(idx_t e) const   // :1741-1838 with pft == pft2 == EvenOdd
    {
        switch (clip_op) {
            case OP_INTERSECTION: return ed[e].wc2 != 0;
            case OP_UNION:        return ed[e].wc2 == 0;
            case OP_DIFFERENCE:   return ed[e].poly == 0 ? (ed[e].wc2 == 0) : (ed[e].wc2 != 0);
            default:              return true;   // xor
        }
    }

    // ---------------------------------------------------------------- output records
    SZ_HDM void set_hole_state(idx_t e, idx_t r)   // :2301-2324
    {
        idx_t e2 = ed[e].pael, tmp = NIL;
        while (e2 != NIL) {
            if (ed[e2].out >= 0) {
                if (tmp == NIL) tmp = e2;
                else if (ed[tmp].out == ed[e2].out) tmp = NIL;
            }
            e2 = ed[e2].pael;
        }
        if (tmp == NIL) { orec[r].first_left = NIL; orec[r].hole = false; }
        else { orec[r].first_left = ed[tmp].out; orec[r].hole = !orec[ed[tmp].out].hole; }
    }
    SZ_HD idx_t create_outrec()   // :1380-1392 (caller checked capacity)
    {
        idx_t r = (idx_t)n_or++;
        orec[r].idx = r; orec[r].hole = false; orec[r].first_left = NIL; orec[r].pts = NIL; orec[r].bottom = NIL;
        return r;
    }
    // :2463-2499.  On arena exhaustion: flag the overflow and change nothing structurally.
    SZ_HDM idx_t add_out_pt(idx_t e, P64 pt)
    {
        if (ed[e].out < 0) {
            if (n_or >= C::OR || n_op >= C::OP) { fail(ST_OVERFLOW); return 0; }
            idx_t r = create_outrec();
            idx_t q = (idx_t)n_op++;
            orec[r].pts = q;
            op[q].rec = r; op[q].pt = pt; op[q].next = q; op[q].prev = q;
            set_hole_state(e, r);
            ed[e].out = r;
            return q;
        }
        const idx_t r = ed[e].out;
        const idx_t o = orec[r].pts;
        const bool front = (ed[e].side == 1);
        if (front && pt == op[o].pt) return o;
        else if (!front && pt == op[op[o].prev].pt) return op[o].prev;
        if (n_op >= C::OP) { fail(ST_OVERFLOW); return o; }
        idx_t q = (idx_t)n_op++;
        op[q].rec = r; op[q].pt = pt;
        op[q].next = o; op[q].prev = op[o].prev;
        op[op[q].prev].next = q; op[o].prev = q;
        if (front) orec[r].pts = q;
        return q;
    }
    SZ_HD idx_t last_out_pt(idx_t e) const   // :2502-2509
    {
        idx_t r = ed[e].out;
        return (ed[e].side == 1) ? orec[r].pts : op[orec[r].pts].prev;
    }
    SZ_HD void add_join(idx_t o1, idx_t o2, P64 off)   // :1942-1949
    {
        if (n_jn >= C::J) { fail(ST_OVERFLOW); return; }
        jn[n_jn].op1 = o1; jn[n_jn].op2 = o2; jn[n_jn].off = off; ++n_jn;
    }
    SZ_HD void add_ghost_join(idx_t o, P64 off)   // :1968-1975
    {
        if (n_gj >= C::GJ) { fail(ST_OVERFLOW); return; }
        gj[n_gj].op1 = o; gj[n_gj].op2 = NIL; gj[n_gj].off = off; ++n_gj;
    }
    SZ_HD static bool horz_segments_overlap(i64 a1, i64 a2, i64 b1, i64 b2)   // :872-877
    {
        if (a1 > a2) { i64 t = a1; a1 = a2; a2 = t; }
        if (b1 > b2) { i64 t = b1; b1 = b2; b2 = t; }
        return (a1 < b2) && (b1 < a2);
    }

    // :1841-1881
    SZ_HDM idx_t add_local_min_poly(idx_t e1, idx_t e2, P64 pt)
    {
        idx_t result, e, prev_e;
        if (is_horz(e2) || ed[e1].dx > ed[e2].dx) {
            result = add_out_pt(e1, pt);
            ed[e2].out = ed[e1].out;
            ed[e1].side = 1; ed[e2].side = 2;
            e = e1;
            prev_e = (ed[e].pael == e2) ? ed[e2].pael : ed[e].pael;
        } else {
            result = add_out_pt(e2, pt);
            ed[e1].out = ed[e2].out;
            ed[e1].side = 2; ed[e2].side = 1;
            e = e2;
            prev_e = (ed[e].pael == e1) ? ed[e1].pael : ed[e].pael;
        }
        if (prev_e != NIL && ed[prev_e].out >= 0 && ed[prev_e].top.y < pt.y && ed[e].top.y < pt.y) {
            i64 xp = top_x(prev_e, pt.y), xe = top_x(e, pt.y);
            P64 a; a.x = xp; a.y = pt.y; P64 b; b.x = xe; b.y = pt.y;
            if (xp == xe && slopes_eq4(a, ed[prev_e].top, b, ed[e].top)) {
                idx_t o = add_out_pt(prev_e, pt);
                add_join(result, o, ed[e].top);
            }
        }
        return result;
    }
    SZ_HD void reverse_links(idx_t pp)   // :692-703
    {
        if (pp == NIL) return;
        idx_t p1 = pp;
        do { idx_t p2 = op[p1].next; op[p1].next = op[p1].prev; op[p1].prev = p2; p1 = p2; } while (p1 != pp);
    }
    SZ_HD double ring_area(idx_t o) const   // :406-416
    {
        if (o == NIL) return 0;
        const idx_t start = o; double a = 0;
        do {
            const P64 pp = op[op[o].prev].pt, pc = op[o].pt;
            a = fp::add(a, fp::mul(fp::cvt(pp.x + pc.x), fp::cvt(pp.y - pc.y)));
            o = op[o].next;
        } while (o != start);
        return fp::mul(a, 0.5);
    }
    SZ_HDM bool first_is_bottom_pt(idx_t b1, idx_t b2) const   // :798-819
    {
        idx_t p = op[b1].prev;
        while (op[p].pt == op[b1].pt && p != b1) p = op[p].prev;
        double dx1p = fabs(dx_of(op[b1].pt, op[p].pt));
        p = op[b1].next;
        while (op[p].pt == op[b1].pt && p != b1) p = op[p].next;
        double dx1n = fabs(dx_of(op[b1].pt, op[p].pt));
        p = op[b2].prev;
        while (op[p].pt == op[b2].pt && p != b2) p = op[p].prev;
        double dx2p = fabs(dx_of(op[b2].pt, op[p].pt));
        p = op[b2].next;
        while (op[p].pt == op[b2].pt && p != b2) p = op[p].next;
        double dx2n = fabs(dx_of(op[b2].pt, op[p].pt));
        const double mx1 = dx1p < dx1n ? dx1n : dx1p, mn1 = dx1n < dx1p ? dx1n : dx1p;   // std::max / std::min
        const double mx2 = dx2p < dx2n ? dx2n : dx2p, mn2 = dx2n < dx2p ? dx2n : dx2p;
        if (mx1 == mx2 && mn1 == mn2) return ring_area(b1) > 0;
        return (dx1p >= dx2p && dx1p >= dx2n) || (dx1n >= dx2p && dx1n >= dx2n);
    }
    SZ_HDM idx_t get_bottom_pt(idx_t pp) const   // :822-857
    {
        idx_t dups = NIL;
        idx_t p = op[pp].next;
        while (p != pp) {
            if (op[p].pt.y > op[pp].pt.y) { pp = p; dups = NIL; }
            else if (op[p].pt.y == op[pp].pt.y && op[p].pt.x <= op[pp].pt.x) {
                if (op[p].pt.x < op[pp].pt.x) { dups = NIL; pp = p; }
                else if (op[p].next != pp && op[p].prev != pp) dups = p;
            }
            p = op[p].next;
        }
        if (dups != NIL) {
            while (dups != p) {
                if (!first_is_bottom_pt(p, dups)) pp = dups;
                dups = op[dups].next;
                while (op[dups].pt != op[pp].pt) dups = op[dups].next;
            }
        }
        return pp;
    }
    SZ_HDM idx_t lowermost_rec(idx_t r1, idx_t r2)   // :2327-2344
    {
        if (orec[r1].bottom == NIL) orec[r1].bottom = get_bottom_pt(orec[r1].pts);
        if (orec[r2].bottom == NIL) orec[r2].bottom = get_bottom_pt(orec[r2].pts);
        const idx_t o1 = orec[r1].bottom, o2 = orec[r2].bottom;
        if (op[o1].pt.y > op[o2].pt.y) return r1;
        else if (op[o1].pt.y < op[o2].pt.y) return r2;
        else if (op[o1].pt.x < op[o2].pt.x) return r1;
        else if (op[o1].pt.x > op[o2].pt.x) return r2;
        else if (op[o1].next == o1) return r2;
        else if (op[o2].next == o2) return r1;
        else if (first_is_bottom_pt(o1, o2)) return r1;
        else return r2;
    }
    SZ_HD bool rec1_right_of_rec2(idx_t r1, idx_t r2) const   // :2347-2355
    {
        do { r1 = orec[r1].first_left; if (r1 == r2) return true; } while (r1 != NIL);
        return false;
    }
    SZ_HD idx_t get_outrec(idx_t i) const   // :2358-2364
    {
        idx_t r = i;
        while (r != orec[r].idx) r = orec[r].idx;
        return r;
    }
    // :2367-2460
    SZ_HDM void append_polygon(idx_t e1, idx_t e2)
    {
        const idx_t r1 = ed[e1].out, r2 = ed[e2].out;
        idx_t hole_rec;
        if (rec1_right_of_rec2(r1, r2)) hole_rec = r2;
        else if (rec1_right_of_rec2(r2, r1)) hole_rec = r1;
        else hole_rec = lowermost_rec(r1, r2);

        const idx_t p1l = orec[r1].pts, p1r = op[p1l].prev;
        const idx_t p2l = orec[r2].pts, p2r = op[p2l].prev;
        if (ed[e1].side == 1) {
            if (ed[e2].side == 1) {   // z y x a b c
                reverse_links(p2l);
                op[p2l].next = p1l; op[p1l].prev = p2l;
                op[p1r].next = p2r; op[p2r].prev = p1r;
                orec[r1].pts = p2r;
            } else {                  // x y z a b c
                op[p2r].next = p1l; op[p1l].prev = p2r;
                op[p2l].prev = p1r; op[p1r].next = p2l;
                orec[r1].pts = p2l;
            }
        } else {
            if (ed[e2].side == 2) {   // a b c z y x
                reverse_links(p2l);
                op[p1r].next = p2r; op[p2r].prev = p1r;
                op[p2l].next = p1l; op[p1l].prev = p2l;
            } else {                  // a b c x y z
                op[p1r].next = p2l; op[p2l].prev = p1r;
                op[p1l].prev = p2r; op[p2r].next = p1l;
            }
        }
        orec[r1].bottom = NIL;
        if (hole_rec == r2) {
            if (orec[r2].first_left != r1) orec[r1].first_left = orec[r2].first_left;
            orec[r1].hole = orec[r2].hole;
        }
        orec[r2].pts = NIL; orec[r2].bottom = NIL; orec[r2].first_left = r1;

        const idx_t ok_idx = ed[e1].out, obsolete = ed[e2].out;
        ed[e1].out = -1; ed[e2].out = -1;
        for (idx_t e = ael; e != NIL; e = ed[e].nael) {
            if (ed[e].out == obsolete) { ed[e].out = ok_idx; ed[e].side = ed[e1].side; break; }
        }
        orec[r2].idx = orec[r1].idx;
    }
    SZ_HDM void add_local_max_poly(idx_t e1, idx_t e2, P64 pt)   // :1884-1897
    {
        add_out_pt(e1, pt);
        if (ed[e1].out == ed[e2].out) { ed[e1].out = -1; ed[e2].out = -1; }
        else if (ed[e1].out < ed[e2].out) append_polygon(e1, e2);
        else append_polygon(e2, e1);
    }

    // :2106-2298 for closed even-odd paths (|WindCnt| == 1 on every edge, see header note)
    SZ_HDM void intersect_edges(idx_t e1, idx_t e2, P64 pt)
    {
        const bool c1 = ed[e1].out >= 0, c2 = ed[e2].out >= 0;
        const bool same = ed[e1].poly == ed[e2].poly;
        if (!same) { ed[e1].wc2 ^= 1; ed[e2].wc2 ^= 1; }
        if (c1 && c2) {
            if (!same && clip_op != OP_XOR) add_local_max_poly(e1, e2, pt);
            else { add_out_pt(e1, pt); add_out_pt(e2, pt); swap_sides_idx(e1, e2); }
        } else if (c1) { add_out_pt(e1, pt); swap_sides_idx(e1, e2); }
        else if (c2) { add_out_pt(e2, pt); swap_sides_idx(e1, e2); }
        else {
            const int w1 = ed[e1].wc2, w2 = ed[e2].wc2;
            if (!same) add_local_min_poly(e1, e2, pt);
            else switch (clip_op) {
                case OP_INTERSECTION: if (w1 > 0 && w2 > 0) add_local_min_poly(e1, e2, pt); break;
                case OP_UNION:        if (w1 <= 0 && w2 <= 0) add_local_min_poly(e1, e2, pt); break;
                case OP_DIFFERENCE:
                    if ((ed[e1].poly == 1 && w1 > 0 && w2 > 0) || (ed[e1].poly == 0 && w1 <= 0 && w2 <= 0))
                        add_local_min_poly(e1, e2, pt);
                    break;
                default: add_local_min_poly(e1, e2, pt);
            }
        }
    }
    SZ_HD void swap_sides_idx(idx_t e1, idx_t e2)   // SwapSides + SwapPolyIndexes :599-612
    {
        signed char s = ed[e1].side; ed[e1].side = ed[e2].side; ed[e2].side = s;
        idx_t o = ed[e1].out; ed[e1].out = ed[e2].out; ed[e2].out = o;
    }

    // ---------------------------------------------------------------- local minima -> AEL (:1978-2077)
    SZ_HDM void insert_local_minima(i64 bot_y)
    {
        while (cur_lm < n_lm && lm[cur_lm].y == bot_y) {
            const idx_t lb = lm[cur_lm].left, rb = lm[cur_lm].right;
            ++cur_lm;
            idx_t op1 = NIL;
            insert_into_ael(lb, NIL);
            insert_into_ael(rb, lb);
            set_winding(lb);
            ed[rb].wc2 = ed[lb].wc2;
            if (is_contributing(lb)) op1 = add_local_min_poly(lb, rb, ed[lb].bot);
            insert_scanbeam(ed[lb].top.y);
            if (is_horz(rb)) {
                add_edge_to_sel(rb);
                if (ed[rb].nlml != NIL) insert_scanbeam(ed[ed[rb].nlml].top.y);
            } else insert_scanbeam(ed[rb].top.y);

            // if any output polygons share an edge, they'll need joining later
            if (op1 != NIL && is_horz(rb) && n_gj > 0) {
                for (int i = 0; i < n_gj; ++i) {
                    if (horz_segments_overlap(op[gj[i].op1].pt.x, gj[i].off.x, ed[rb].bot.x, ed[rb].top.x))
                        add_join(gj[i].op1, op1, gj[i].off);
                }
            }
            if (ed[lb].out >= 0 && ed[lb].pael != NIL) {
                const idx_t p = ed[lb].pael;
                if (ed[p].cur.x == ed[lb].bot.x && ed[p].out >= 0 &&
                    slopes_eq4(ed[p].bot, ed[p].top, ed[lb].cur, ed[lb].top)) {
                    idx_t op2 = add_out_pt(p, ed[lb].bot);
                    add_join(op1, op2, ed[lb].top);
                }
            }
            if (ed[lb].nael != rb) {
                const idx_t rp = ed[rb].pael;
                if (ed[rb].out >= 0 && ed[rp].out >= 0 &&
                    slopes_eq4(ed[rp].cur, ed[rp].top, ed[rb].cur, ed[rb].top)) {
                    idx_t op2 = add_out_pt(rp, ed[rb].bot);
                    add_join(op1, op2, ed[rb].top);
                }
                idx_t e = ed[lb].nael;
                if (e != NIL) {
                    while (e != rb) {
                        // intersect_edges assumes param1 is to the right of param2 above the point
                        intersect_edges(rb, e, ed[lb].cur);
                        e = ed[e].nael;
                    }
                }
            }
        }
    }

    // ---------------------------------------------------------------- horizontals (:2512-2824)
    SZ_HD idx_t maxima_pair(idx_t e) const   // :2538-2545
    {
        const idx_t n = ed[e].next, p = ed[e].prev;
        if (ed[n].top == ed[e].top && ed[n].nlml == NIL) return n;
        else if (ed[p].top == ed[e].top && ed[p].nlml == NIL) return p;
        return NIL;
    }
    SZ_HD idx_t maxima_pair_ex(idx_t e) const   // :2548-2555
    {
        idx_t r = maxima_pair(e);
        if (r != NIL && (ed[r].nael == ed[r].pael && !is_horz(r))) return NIL;
        return r;
    }
    SZ_HD void horz_direction(idx_t h, bool& l2r, i64& left, i64& right) const   // :2610-2623
    {
        if (ed[h].bot.x < ed[h].top.x) { left = ed[h].bot.x; right = ed[h].top.x; l2r = true; }
        else { left = ed[h].top.x; right = ed[h].bot.x; l2r = false; }
    }
    SZ_HD void joins_with_pending_horizontals(idx_t h, idx_t op1)   // the loop at :2721-2732 / :2774-2785
    {
        for (idx_t nh = sel; nh != NIL; nh = ed[nh].nsel) {
            if (ed[nh].out >= 0 && horz_segments_overlap(ed[h].bot.x, ed[h].top.x, ed[nh].bot.x, ed[nh].top.x)) {
                idx_t op2 = last_out_pt(nh);
                add_join(op2, op1, ed[nh].top);
            }
        }
    }
    SZ_HDN void process_horizontal(idx_t h)
    {
        bool l2r; i64 hl, hr;
        horz_direction(h, l2r, hl, hr);
        idx_t last_h = h, max_pair = NIL;
        while (ed[last_h].nlml != NIL && is_horz(ed[last_h].nlml)) last_h = ed[last_h].nlml;
        if (ed[last_h].nlml == NIL) max_pair = maxima_pair(last_h);
        idx_t op1 = NIL;
        for (;;) {   // loop through consecutive horizontal edges
            const bool is_last = (h == last_h);
            idx_t e = l2r ? ed[h].nael : ed[h].pael;
            while (e != NIL) {
                if ((l2r && ed[e].cur.x > hr) || (!l2r && ed[e].cur.x < hl)) break;
                // also break at the end of an intermediate horizontal (smaller Dx lies to the right above)
                if (ed[e].cur.x == ed[h].top.x && ed[h].nlml != NIL && ed[e].dx < ed[ed[h].nlml].dx) break;
                if (ed[h].out >= 0) {
                    op1 = add_out_pt(h, ed[e].cur);
                    joins_with_pending_horizontals(h, op1);
                    add_ghost_join(op1, ed[h].bot);
                }
                if (e == max_pair && is_last) {
                    if (ed[h].out >= 0) add_local_max_poly(h, max_pair, ed[h].top);
                    delete_from_ael(h);
                    delete_from_ael(max_pair);
                    return;
                }
                P64 pt; pt.x = ed[e].cur.x; pt.y = ed[h].cur.y;
                if (l2r) intersect_edges(h, e, pt); else intersect_edges(e, h, pt);
                idx_t en = l2r ? ed[e].nael : ed[e].pael;
                swap_in_ael(h, e);
                e = en;
            }
            if (ed[h].nlml == NIL || !is_horz(ed[h].nlml)) break;
            update_edge_into_ael(h);
            if (ed[h].out >= 0) add_out_pt(h, ed[h].bot);
            horz_direction(h, l2r, hl, hr);
        }
        if (ed[h].out >= 0 && op1 == NIL) {
            op1 = last_out_pt(h);
            joins_with_pending_horizontals(h, op1);
            add_ghost_join(op1, ed[h].top);
        }
        if (ed[h].nlml != NIL) {
            if (ed[h].out >= 0) {
                op1 = add_out_pt(h, ed[h].top);
                update_edge_into_ael(h);
                // h is no longer horizontal here
                const idx_t ep = ed[h].pael, en = ed[h].nael;
                if (ep != NIL && ed[ep].cur.x == ed[h].bot.x && ed[ep].cur.y == ed[h].bot.y &&
                    ed[ep].out >= 0 && ed[ep].cur.y > ed[ep].top.y && edge_slopes_eq(h, ep)) {
                    idx_t op2 = add_out_pt(ep, ed[h].bot);
                    add_join(op1, op2, ed[h].top);
                } else if (en != NIL && ed[en].cur.x == ed[h].bot.x && ed[en].cur.y == ed[h].bot.y &&
                           ed[en].out >= 0 && ed[en].cur.y > ed[en].top.y && edge_slopes_eq(h, en)) {
                    idx_t op2 = add_out_pt(en, ed[h].bot);
                    add_join(op1, op2, ed[h].top);
                }
            } else update_edge_into_ael(h);
        } else {
            if (ed[h].out >= 0) add_out_pt(h, ed[h].top);
            delete_from_ael(h);
        }
    }
    SZ_HD bool edge_slopes_eq(idx_t a, idx_t b) const   // :541-551
    {
        return prod_eq(ed[a].top.y - ed[a].bot.y, ed[b].top.x - ed[b].bot.x, ed[a].top.x - ed[a].bot.x, ed[b].top.y - ed[b].bot.y);
    }
    SZ_HD void process_horizontals()   // :2512-2517
    {
        while (sel != NIL) { idx_t h = sel; delete_from_sel(h); process_horizontal(h); }
    }

    // ---------------------------------------------------------------- intersections (:622-689, :2827-2954)
    SZ_HDM P64 intersect_point(idx_t a, idx_t b) const
    {
        const Edge& e1 = ed[a]; const Edge& e2 = ed[b];
        P64 ip;
        if (e1.dx == e2.dx) { ip.y = e1.cur.y; ip.x = top_x(a, ip.y); return ip; }
        else if (e1.dx == 0) {
            ip.x = e1.bot.x;
            if (is_horz(b)) ip.y = e2.bot.y;
            else {
                double b2 = fp::sub(fp::cvt(e2.bot.y), fp::div(fp::cvt(e2.bot.x), e2.dx));
                ip.y = fp::round_half(fp::add(fp::div(fp::cvt(ip.x), e2.dx), b2));
            }
        } else if (e2.dx == 0) {
            ip.x = e2.bot.x;
            if (is_horz(a)) ip.y = e1.bot.y;
            else {
                double b1 = fp::sub(fp::cvt(e1.bot.y), fp::div(fp::cvt(e1.bot.x), e1.dx));
                ip.y = fp::round_half(fp::add(fp::div(fp::cvt(ip.x), e1.dx), b1));
            }
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
            ip.x = (fabs(e1.dx) < fabs(e2.dx)) ? top_x(a, ip.y) : top_x(b, ip.y);
        }
        // don't allow ip below the bottom of the scanbeam
        if (ip.y > e1.cur.y) {
            ip.y = e1.cur.y;
            ip.x = (fabs(e1.dx) > fabs(e2.dx)) ? top_x(b, ip.y) : top_x(a, ip.y);   // use the more vertical edge
        }
        return ip;
    }
    SZ_HDN bool build_intersect_list(i64 top_y)   // :2856-2902; false only on arena overflow
    {
        if (ael == NIL) return true;
        idx_t e = ael;
        sel = e;
        while (e != NIL) {
            ed[e].psel = ed[e].pael; ed[e].nsel = ed[e].nael;
            ed[e].cur.x = top_x(e, top_y);
            e = ed[e].nael;
        }
        bool modified;
        do {   // bubble sort
            modified = false;
            e = sel;
            while (ed[e].nsel != NIL) {
                const idx_t en = ed[e].nsel;
                if (ed[e].cur.x > ed[en].cur.x) {
                    P64 pt = intersect_point(e, en);
                    if (pt.y < top_y) { pt.x = top_x(e, top_y); pt.y = top_y; }
                    if (n_il >= C::IN) { fail(ST_OVERFLOW); sel = NIL; return false; }
                    il[n_il].e1 = e; il[n_il].e2 = en; il[n_il].pt = pt; ++n_il;
                    swap_in_sel(e, en);
                    modified = true;
                } else e = en;
            }
            if (ed[e].psel != NIL) ed[ed[e].psel].nsel = NIL; else break;
        } while (modified);
        sel = NIL;
        return true;
    }
    struct INodeLess { SZ_HD bool operator()(const INode& a, const INode& b) const { return b.pt.y < a.pt.y; } };   // :2921-2924
    SZ_HD bool edges_adjacent(const INode& n) const { return ed[n.e1].nsel == n.e2 || ed[n.e1].psel == n.e2; }   // :2927-2931
    SZ_HDM bool fixup_intersection_order()   // :2934-2954
    {
        sel = ael;                                        // CopyAELToSEL :1929-1939
        for (idx_t e = ael; e != NIL; e = ed[e].nael) { ed[e].psel = ed[e].pael; ed[e].nsel = ed[e].nael; }
        stl_sort(il, n_il, INodeLess());
        for (int i = 0; i < n_il; ++i) {
            if (!edges_adjacent(il[i])) {
                int j = i + 1;
                while (j < n_il && !edges_adjacent(il[j])) j++;
                if (j == n_il) return false;
                INode t = il[i]; il[i] = il[j]; il[j] = t;
            }
            swap_in_sel(il[i].e1, il[i].e2);
        }
        return true;
    }
    SZ_HDM bool process_intersections(i64 top_y)   // :2827-2845
    {
        if (ael == NIL) return true;
        n_il = 0;
        if (!build_intersect_list(top_y)) return false;
        if (n_il == 0) return true;
        if (n_il == 1 || fixup_intersection_order()) {
            for (int i = 0; i < n_il; ++i) {              // ProcessIntersectList :2906-2918
                intersect_edges(il[i].e1, il[i].e2, il[i].pt);
                swap_in_ael(il[i].e1, il[i].e2);
            }
            n_il = 0;
        } else return false;
        sel = NIL;
        return true;
    }

    // ---------------------------------------------------------------- top of scanbeam (:2957-3113)
    SZ_HDM void do_maxima(idx_t e)
    {
        const idx_t mp = maxima_pair_ex(e);
        if (mp == NIL) {
            if (ed[e].out >= 0) add_out_pt(e, ed[e].top);
            delete_from_ael(e);
            return;
        }
        idx_t en = ed[e].nael;
        while (en != NIL && en != mp) {
            intersect_edges(e, en, ed[e].top);
            swap_in_ael(e, en);
            en = ed[e].nael;
        }
        if (ed[e].out == -1 && ed[mp].out == -1) { delete_from_ael(e); delete_from_ael(mp); }
        else if (ed[e].out >= 0 && ed[mp].out >= 0) {
            add_local_max_poly(e, mp, ed[e].top);
            delete_from_ael(e); delete_from_ael(mp);
        } else fail(ST_CLIPPER_FAIL);   // "DoMaxima error"
    }
    SZ_HDN void process_edges_at_top(i64 top_y)
    {
        idx_t e = ael;
        while (e != NIL) {
            // 1. maxima are treated as 'bent' horizontal edges, excluding maxima with horizontal partners
            bool is_max = (ed[e].top.y == top_y && ed[e].nlml == NIL);
            if (is_max) { idx_t mp = maxima_pair_ex(e); is_max = (mp == NIL || !is_horz(mp)); }
            if (is_max) {
                idx_t ep = ed[e].pael;
                do_maxima(e);
                if (status == ST_CLIPPER_FAIL) return;
                e = (ep == NIL) ? ael : ed[ep].nael;
            } else {
                // 2. promote horizontal edges, otherwise update cur
                if (ed[e].top.y == top_y && ed[e].nlml != NIL && is_horz(ed[e].nlml)) {
                    update_edge_into_ael(e);
                    if (ed[e].out >= 0) add_out_pt(e, ed[e].bot);
                    add_edge_to_sel(e);
                } else {
                    ed[e].cur.x = top_x(e, top_y);
                    ed[e].cur.y = top_y;
                }
                e = ed[e].nael;
            }
        }
        // 3. horizontals at the top of the scanbeam
        process_horizontals();
        // 4. promote intermediate vertices
        e = ael;
        while (e != NIL) {
            if (ed[e].top.y == top_y && ed[e].nlml != NIL) {
                idx_t o = NIL;
                if (ed[e].out >= 0) o = add_out_pt(e, ed[e].top);
                update_edge_into_ael(e);
                const idx_t ep = ed[e].pael, en = ed[e].nael;
                if (ep != NIL && ed[ep].cur.x == ed[e].bot.x && ed[ep].cur.y == ed[e].bot.y && o != NIL &&
                    ed[ep].out >= 0 && ed[ep].cur.y > ed[ep].top.y &&
                    slopes_eq4(ed[e].cur, ed[e].top, ed[ep].cur, ed[ep].top)) {
                    idx_t o2 = add_out_pt(ep, ed[e].bot);
                    add_join(o, o2, ed[e].top);
                } else if (en != NIL && ed[en].cur.x == ed[e].bot.x && ed[en].cur.y == ed[e].bot.y && o != NIL &&
                           ed[en].out >= 0 && ed[en].cur.y > ed[en].top.y &&
                           slopes_eq4(ed[e].cur, ed[e].top, ed[en].cur, ed[en].top)) {
                    idx_t o2 = add_out_pt(en, ed[e].bot);
                    add_join(o, o2, ed[e].top);
                }
            }
            e = ed[e].nael;
        }
    }

    // ---------------------------------------------------------------- joins (:3290-3763)
    SZ_HD idx_t dup_out_pt(idx_t o, bool after)   // :3348-3368 (capacity pre-checked by join_points)
    {
        idx_t r = (idx_t)n_op++;
        op[r].pt = op[o].pt; op[r].rec = op[o].rec;
        if (after) { op[r].next = op[o].next; op[r].prev = o; op[op[o].next].prev = r; op[o].next = r; }
        else { op[r].prev = op[o].prev; op[r].next = o; op[op[o].prev].next = r; op[o].prev = r; }
        return r;
    }
    SZ_HD static bool get_overlap(i64 a1, i64 a2, i64 b1, i64 b2, i64& left, i64& right)   // :3290-3304
    {
        if (a1 < a2) {
            if (b1 < b2) { left = a1 > b1 ? a1 : b1; right = a2 < b2 ? a2 : b2; }
            else { left = a1 > b2 ? a1 : b2; right = a2 < b1 ? a2 : b1; }
        } else {
            if (b1 < b2) { left = a2 > b1 ? a2 : b1; right = a1 < b2 ? a1 : b2; }
            else { left = a2 > b2 ? a2 : b2; right = a1 < b1 ? a1 : b1; }
        }
        return left < right;
    }
    SZ_HDM bool join_horz(idx_t o1, idx_t o1b, idx_t o2, idx_t o2b, P64 pt, bool discard_left)   // :3371-3455
    {
        const bool d1_l2r = !(op[o1].pt.x > op[o1b].pt.x);
        const bool d2_l2r = !(op[o2].pt.x > op[o2b].pt.x);
        if (d1_l2r == d2_l2r) return false;
        if (d1_l2r) {
            while (op[op[o1].next].pt.x <= pt.x && op[op[o1].next].pt.x >= op[o1].pt.x && op[op[o1].next].pt.y == pt.y) o1 = op[o1].next;
            if (discard_left && op[o1].pt.x != pt.x) o1 = op[o1].next;
            o1b = dup_out_pt(o1, !discard_left);
            if (op[o1b].pt != pt) { o1 = o1b; op[o1].pt = pt; o1b = dup_out_pt(o1, !discard_left); }
        } else {
            while (op[op[o1].next].pt.x >= pt.x && op[op[o1].next].pt.x <= op[o1].pt.x && op[op[o1].next].pt.y == pt.y) o1 = op[o1].next;
            if (!discard_left && op[o1].pt.x != pt.x) o1 = op[o1].next;
            o1b = dup_out_pt(o1, discard_left);
            if (op[o1b].pt != pt) { o1 = o1b; op[o1].pt = pt; o1b = dup_out_pt(o1, discard_left); }
        }
        if (d2_l2r) {
            while (op[op[o2].next].pt.x <= pt.x && op[op[o2].next].pt.x >= op[o2].pt.x && op[op[o2].next].pt.y == pt.y) o2 = op[o2].next;
            if (discard_left && op[o2].pt.x != pt.x) o2 = op[o2].next;
            o2b = dup_out_pt(o2, !discard_left);
            if (op[o2b].pt != pt) { o2 = o2b; op[o2].pt = pt; o2b = dup_out_pt(o2, !discard_left); }
        } else {
            while (op[op[o2].next].pt.x >= pt.x && op[op[o2].next].pt.x <= op[o2].pt.x && op[op[o2].next].pt.y == pt.y) o2 = op[o2].next;
            if (!discard_left && op[o2].pt.x != pt.x) o2 = op[o2].next;
            o2b = dup_out_pt(o2, discard_left);
            if (op[o2b].pt != pt) { o2 = o2b; op[o2].pt = pt; o2b = dup_out_pt(o2, discard_left); }
        }
        if (d1_l2r == discard_left) {
            op[o1].prev = o2; op[o2].next = o1; op[o1b].next = o2b; op[o2b].prev = o1b;
        } else {
            op[o1].next = o2; op[o2].prev = o1; op[o1b].prev = o2b; op[o2b].next = o1b;
        }
        return true;
    }
    SZ_HDN bool join_points(Join& j, idx_t r1, idx_t r2)   // :3458-3614
    {
        if (n_op + 4 > C::OP) { fail(ST_OVERFLOW); return false; }
        idx_t o1 = j.op1, o1b, o2 = j.op2, o2b;
        const bool horizontal = (op[j.op1].pt.y == j.off.y);
        if (horizontal && j.off == op[j.op1].pt && j.off == op[j.op2].pt) {
            // strictly-simple join
            if (r1 != r2) return false;
            o1b = op[j.op1].next;
            while (o1b != o1 && op[o1b].pt == j.off) o1b = op[o1b].next;
            const bool rev1 = op[o1b].pt.y > j.off.y;
            o2b = op[j.op2].next;
            while (o2b != o2 && op[o2b].pt == j.off) o2b = op[o2b].next;
            const bool rev2 = op[o2b].pt.y > j.off.y;
            if (rev1 == rev2) return false;
            if (rev1) {
                o1b = dup_out_pt(o1, false); o2b = dup_out_pt(o2, true);
                op[o1].prev = o2; op[o2].next = o1; op[o1b].next = o2b; op[o2b].prev = o1b;
            } else {
                o1b = dup_out_pt(o1, true); o2b = dup_out_pt(o2, false);
                op[o1].next = o2; op[o2].prev = o1; op[o1b].prev = o2b; op[o2b].next = o1b;
            }
            j.op1 = o1; j.op2 = o1b;
            return true;
        } else if (horizontal) {
            // op1 and op2 may be anywhere along the horizontal edge
            o1b = o1;
            while (op[op[o1].prev].pt.y == op[o1].pt.y && op[o1].prev != o1b && op[o1].prev != o2) o1 = op[o1].prev;
            while (op[op[o1b].next].pt.y == op[o1b].pt.y && op[o1b].next != o1 && op[o1b].next != o2) o1b = op[o1b].next;
            if (op[o1b].next == o1 || op[o1b].next == o2) return false;   // a flat 'polygon'
            o2b = o2;
            while (op[op[o2].prev].pt.y == op[o2].pt.y && op[o2].prev != o2b && op[o2].prev != o1b) o2 = op[o2].prev;
            while (op[op[o2b].next].pt.y == op[o2b].pt.y && op[o2b].next != o2 && op[o2b].next != o1) o2b = op[o2b].next;
            if (op[o2b].next == o2 || op[o2b].next == o1) return false;   // a flat 'polygon'
            i64 left, right;
            if (!get_overlap(op[o1].pt.x, op[o1b].pt.x, op[o2].pt.x, op[o2b].pt.x, left, right)) return false;
            P64 pt; bool discard_left;
            if (op[o1].pt.x >= left && op[o1].pt.x <= right) { pt = op[o1].pt; discard_left = op[o1].pt.x > op[o1b].pt.x; }
            else if (op[o2].pt.x >= left && op[o2].pt.x <= right) { pt = op[o2].pt; discard_left = op[o2].pt.x > op[o2b].pt.x; }
            else if (op[o1b].pt.x >= left && op[o1b].pt.x <= right) { pt = op[o1b].pt; discard_left = op[o1b].pt.x > op[o1].pt.x; }
            else { pt = op[o2b].pt; discard_left = op[o2b].pt.x > op[o2].pt.x; }
            j.op1 = o1; j.op2 = o2;
            return join_horz(o1, o1b, o2, o2b, pt, discard_left);
        } else {
            // non-horizontal: op1.pt.y == op2.pt.y and op1.pt.y > off.y
            o1b = op[o1].next;
            while (op[o1b].pt == op[o1].pt && o1b != o1) o1b = op[o1b].next;
            const bool rev1 = (op[o1b].pt.y > op[o1].pt.y) || !slopes_eq3(op[o1].pt, op[o1b].pt, j.off);
            if (rev1) {
                o1b = op[o1].prev;
                while (op[o1b].pt == op[o1].pt && o1b != o1) o1b = op[o1b].prev;
                if ((op[o1b].pt.y > op[o1].pt.y) || !slopes_eq3(op[o1].pt, op[o1b].pt, j.off)) return false;
            }
            o2b = op[o2].next;
            while (op[o2b].pt == op[o2].pt && o2b != o2) o2b = op[o2b].next;
            const bool rev2 = (op[o2b].pt.y > op[o2].pt.y) || !slopes_eq3(op[o2].pt, op[o2b].pt, j.off);
            if (rev2) {
                o2b = op[o2].prev;
                while (op[o2b].pt == op[o2].pt && o2b != o2) o2b = op[o2b].prev;
                if ((op[o2b].pt.y > op[o2].pt.y) || !slopes_eq3(op[o2].pt, op[o2b].pt, j.off)) return false;
            }
            if (o1b == o1 || o2b == o2 || o1b == o2b || (r1 == r2 && rev1 == rev2)) return false;
            if (rev1) {
                o1b = dup_out_pt(o1, false); o2b = dup_out_pt(o2, true);
                op[o1].prev = o2; op[o2].next = o1; op[o1b].next = o2b; op[o2b].prev = o1b;
            } else {
                o1b = dup_out_pt(o1, true); o2b = dup_out_pt(o2, false);
                op[o1].next = o2; op[o2].prev = o1; op[o1b].prev = o2b; op[o2b].next = o1b;
            }
            j.op1 = o1; j.op2 = o1b;
            return true;
        }
    }
    // :484-523, returns 0 outside, +1 inside, -1 on the boundary
    SZ_HDM int point_in_ring(P64 pt, idx_t o) const
    {
        int result = 0;
        const idx_t start = o;
        for (;;) {
            const P64 a = op[o].pt, b = op[op[o].next].pt;
            if (b.y == pt.y) {
                if (b.x == pt.x || (a.y == pt.y && ((b.x > pt.x) == (a.x < pt.x)))) return -1;
            }
            if ((a.y < pt.y) != (b.y < pt.y)) {
                if (a.x >= pt.x) {
                    if (b.x > pt.x) result = 1 - result;
                    else {
                        double d = fp::sub(fp::mul(fp::cvt(a.x - pt.x), fp::cvt(b.y - pt.y)), fp::mul(fp::cvt(b.x - pt.x), fp::cvt(a.y - pt.y)));
                        if (!d) return -1;
                        if ((d > 0) == (b.y > a.y)) result = 1 - result;
                    }
                } else if (b.x > pt.x) {
                    double d = fp::sub(fp::mul(fp::cvt(a.x - pt.x), fp::cvt(b.y - pt.y)), fp::mul(fp::cvt(b.x - pt.x), fp::cvt(a.y - pt.y)));
                    if (!d) return -1;
                    if ((d > 0) == (b.y > a.y)) result = 1 - result;
                }
            }
            o = op[o].next;
            if (o == start) break;
        }
        return result;
    }
    SZ_HD bool ring2_contains_ring1(idx_t o1, idx_t o2) const   // :526-538
    {
        idx_t o = o1;
        do {
            int res = point_in_ring(op[o].pt, o2);
            if (res >= 0) return res > 0;
            o = op[o].next;
        } while (o != o1);
        return true;
    }
    SZ_HDN void join_common_edges()   // :3679-3763 (no PolyTree => no FixupFirstLefts)
    {
        for (int i = 0; i < n_jn; ++i) {
            if (status != ST_OK) return;
            Join& j = jn[i];
            idx_t r1 = get_outrec(op[j.op1].rec);
            idx_t r2 = get_outrec(op[j.op2].rec);
            if (orec[r1].pts == NIL || orec[r2].pts == NIL) continue;
            idx_t hole_rec;
            if (r1 == r2) hole_rec = r1;
            else if (rec1_right_of_rec2(r1, r2)) hole_rec = r2;
            else if (rec1_right_of_rec2(r2, r1)) hole_rec = r1;
            else hole_rec = lowermost_rec(r1, r2);
            if (!join_points(j, r1, r2)) continue;
            if (r1 == r2) {
                // one polygon was split into two
                if (n_or >= C::OR) { fail(ST_OVERFLOW); return; }
                orec[r1].pts = j.op1; orec[r1].bottom = NIL;
                r2 = create_outrec();
                orec[r2].pts = j.op2;
                { idx_t o = orec[r2].pts; do { op[o].rec = orec[r2].idx; o = op[o].prev; } while (o != orec[r2].pts); }   // UpdateOutPtIdxs
                if (ring2_contains_ring1(orec[r2].pts, orec[r1].pts)) {
                    orec[r2].hole = !orec[r1].hole; orec[r2].first_left = r1;
                    if (orec[r2].hole == (ring_area(orec[r2].pts) > 0)) reverse_links(orec[r2].pts);
                } else if (ring2_contains_ring1(orec[r1].pts, orec[r2].pts)) {
                    orec[r2].hole = orec[r1].hole; orec[r1].hole = !orec[r2].hole;
                    orec[r2].first_left = orec[r1].first_left; orec[r1].first_left = r2;
                    if (orec[r1].hole == (ring_area(orec[r1].pts) > 0)) reverse_links(orec[r1].pts);
                } else {
                    orec[r2].hole = orec[r1].hole; orec[r2].first_left = orec[r1].first_left;
                }
            } else {
                orec[r2].pts = NIL; orec[r2].bottom = NIL; orec[r2].idx = orec[r1].idx;
                orec[r1].hole = orec[hole_rec].hole;
                if (hole_rec == r2) orec[r1].first_left = orec[r2].first_left;
                orec[r2].first_left = r1;
            }
        }
    }
    SZ_HDM void fixup_out_polygon(idx_t r)   // :3143-3181 (PreserveCollinear off)
    {
        idx_t last_ok = NIL;
        orec[r].bottom = NIL;
        idx_t pp = orec[r].pts;
        for (;;) {
            if (op[pp].prev == pp || op[pp].prev == op[pp].next) { orec[r].pts = NIL; return; }
            const P64 a = op[op[pp].prev].pt, b = op[pp].pt, c = op[op[pp].next].pt;
            if (b == c || b == a || slopes_eq3(a, b, c)) {
                last_ok = NIL;
                op[op[pp].prev].next = op[pp].next;
                op[op[pp].next].prev = op[pp].prev;
                pp = op[pp].prev;
            } else if (pp == last_ok) break;
            else { if (last_ok == NIL) last_ok = pp; pp = op[pp].next; }
        }
        orec[r].pts = pp;
    }

    // ---------------------------------------------------------------- Execute (:1508-1525, :1560-1621)
    // The sweep is exposed step by step so that a warp running 32 independent clips can re-converge
    // between the phases of every scanbeam (sz_pairforce.cuh: run_sweep); execute() is the same sequence
    // for a single caller.  After the sweep the solution is left in orec/op (read it with emit()).
    i64 sw_top_y;
    // Reset() :1247-1276 + the first PopScanbeam / InsertLocalMinimaIntoAEL of ExecuteInternal :1569-1571.
    // Returns true while the sweep has scanbeams to process.
    SZ_HDN bool sweep_begin()
    {
        if (status != ST_OK) return false;
        if (n_lm == 0) { status = ST_CLIPPER_FAIL; return false; }    // nothing to process -> PopScanbeam fails
        stl_sort(lm, n_lm, LocMinLess());
        for (int i = 0; i < n_lm; ++i) {
            insert_scanbeam(lm[i].y);
            Edge& l = ed[lm[i].left];  l.cur = l.bot; l.side = 1; l.out = -1;
            Edge& r = ed[lm[i].right]; r.cur = r.bot; r.side = 2; r.out = -1;
        }
        ael = NIL; sel = NIL; cur_lm = 0; sw_top_y = 0;
        i64 bot_y;
        if (!pop_scanbeam(bot_y)) { status = ST_CLIPPER_FAIL; return false; }
        insert_local_minima(bot_y);
        return status == ST_OK;
    }
    SZ_HD bool sweep_next() { return pop_scanbeam(sw_top_y) || cur_lm < n_lm; }     // :1572
    SZ_HDN bool sweep_intersections()                                               // :1574-1576
    {
        process_horizontals();
        n_gj = 0;
        if (!process_intersections(sw_top_y)) { fail(ST_CLIPPER_FAIL); return false; }
        return status == ST_OK;
    }
    SZ_HD bool sweep_top() { process_edges_at_top(sw_top_y); return status == ST_OK; }   // :1577
    SZ_HDN bool sweep_minima() { insert_local_minima(sw_top_y); return status == ST_OK; }   // :1578-1579
    SZ_HDN void sweep_finish()                                                      // :1594-1613
    {
        if (status != ST_OK) return;
        for (int i = 0; i < n_or; ++i) {
            if (orec[i].pts == NIL) continue;
            if (orec[i].hole == (ring_area(orec[i].pts) > 0)) reverse_links(orec[i].pts);
        }
        if (n_jn > 0) join_common_edges();
        if (status != ST_OK) return;
        for (int i = 0; i < n_or; ++i) if (orec[i].pts != NIL) fixup_out_polygon((idx_t)i);
    }
    SZ_HD int execute()
    {
        bool run = sweep_begin();
        while (run && sweep_next()) run = sweep_intersections() && sweep_top() && sweep_minima();
        sweep_finish();
        return status;
    }
    struct LocMinLess { SZ_HD bool operator()(const LocMin& a, const LocMin& b) const { return b.y < a.y; } };   // :125-131

    // ---------------------------------------------------------------- BuildResult (:3199-3217)
    SZ_HD int ring_count(idx_t o) const
    {
        if (o == NIL) return 0;
        int c = 0; idx_t p = o;
        do { ++c; p = op[p].next; } while (p != o);
        return c;
    }
    // Calls sink.begin_path(cnt), then sink.point(P64) cnt times, for every solution path in the
    // reference's output order (OutRec order; each ring starts at Pts->Prev and walks Prev).
    template <class Sink>
    SZ_HD int emit(Sink& sink) const
    {
        int np = 0;
        for (int i = 0; i < n_or; ++i) {
            if (orec[i].pts == NIL) continue;
            idx_t p = op[orec[i].pts].prev;
            int cnt = ring_count(p);
            if (cnt < 2) continue;
            sink.begin_path(cnt);
            for (int k = 0; k < cnt; ++k) { sink.point(op[p].pt); p = op[p].prev; }
            ++np;
        }
        return np;
    }
};

}  // namespace szclip
