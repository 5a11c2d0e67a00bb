// sz_narrow.cuh -- narrow-phase kernels: one CUDA thread resolves one candidate pair
// (collisions/floe_interactions.m, reached from floe_interactions_all.m:125-172) or one stand-alone
// polygon clip (private/mexclipper.cpp:291-298) with the Clipper-exact sweep of sz_clip.cuh.
//
// Size classes.  The sweep needs a per-pair arena whose size depends on the vertex counts, so pairs
// run in three classes; a pair that does not fit (too many vertices, or an arena overflowed during
// the sweep) is appended to the next class's work list and re-run there from scratch:
//   class S  <= 19 vertices per outline     arena in per-thread local memory (interleaved by the
//            (packed Voronoi: 3..13)        hardware, so lanes running the same sweep step touch
//                                           the same sectors), one thread per pair, grid = all pairs
//   class T  <= 47 vertices                 the same with a 35 KB arena (simplified real floes, concave)
//   class M  <= 191 vertices                arena in an HBM scratch slab, persistent threads that
//   class L  <= 1299 vertices               stride over the work list
// Each class is instantiated in its own translation unit (sz_narrow_{S,M,L}.cu) so the three
// ~40 s template instantiations build in parallel; sz_contact.cu calls the extern "C" launchers.
#pragma once
#include "sz_pairforce.cuh"
#include "sz_euler.cuh"
#include <cuda_runtime.h>

namespace sznarrow {

using szpf::Params;
using szpf::Body;
using szclip::i64;
using szclip::P64;

typedef szclip::ClipCaps<40, 20, 112, 32, 64, 32, 16, 64> ClipS;
typedef szpf::PairCaps<ClipS, 20, 64, 6, 40, 4> PairS;
typedef szclip::ClipCaps<96, 48, 320, 64, 192, 64, 32, 144> ClipT;      // class T ("thirty"): simplified real floes
typedef szpf::PairCaps<ClipT, 48, 192, 8, 128, 8> PairT;
typedef szclip::ClipCaps<384, 192, 1536, 256, 1536, 256, 128, 576> ClipM;
typedef szpf::PairCaps<ClipM, 192, 768, 16, 512, 16> PairM;
typedef szclip::ClipCaps<2600, 1300, 10000, 2600, 10000, 2600, 512, 3900> ClipL;
typedef szpf::PairCaps<ClipL, 1300, 5200, 64, 6000, 32> PairL;

// Everything a narrow-phase launch reads and writes.  Indices are 0-based positions in the extended
// floe list (originals, x-ghosts, y-ghosts); `src` maps a position to the original floe whose
// outline and body fields it shares (a ghost is its parent with a shifted centroid,
// floe_interactions_all.m:33-34,54-55).
struct NarrowArgs {
    // extended list
    const double* ex; const double* ey; const int* esrc; const int* egid; const unsigned char* eowned; const unsigned char* econvex; const unsigned char* erot; const unsigned char* eno;
    // per original floe
    const double* h; const double* area; const double* u; const double* v; const double* ksi;
    const int* voff; const double* vx; const double* vy;
    // work: pair k = (pi[k], pj[k]); in wall mode "pair" k is floe first_floe + k against the boundary
    const int* pi; const int* pj;
    int n_work;                    // class S: number of pairs; M/L: unused (list_count is read on the device)
    const int* list; const int* list_count;      // M/L work list (pair indices)
    int* next_list; int* next_count;              // escalation target (NULL in class L)
    // per-pair outputs
    int* status; int* nrows; int* row_start; double* ovl_state;
    // contact-row pool: 5 doubles per row (Fx Fy Px Py overlap), allocated with one atomicAdd per pair
    double* row_pool; int row_cap; int* row_used;
    // optional clip #1 polygons (Clipper coordinates) for bit-exact parity checks
    int want_polys; int* poly_path_start; int* poly_npaths;
    int* path_vstart; int* path_len; int path_cap; int* path_used;
    i64* pvx; i64* pvy; int vert_cap; int* vert_used;
    // wall mode (floe_interactions_all.m:150-172)
    int wall; int first_floe; const double* bx; const double* by; int bn; Body bbody;
    void* scratch; int n_threads;  // M/L arenas: n_threads * sizeof(Workspace<C>)
    int* ho_st; i64* ho_x; i64* ho_y; int ho_stride; int ho_cap;      // experiment: class C split in two kernels (szpf::ConvexHandoff), indexed by work-list position
    Params P;
};

// Stand-alone clip batch (mex gateway semantics): item k clips subject [soff[k],soff[k+1]) with clip
// [coff[k],coff[k+1]) using method[k]; result paths go to the pools.
struct ClipArgs {
    int count; const int* method;
    const i64* soff; const i64* sx; const i64* sy;
    const i64* coff; const i64* cx; const i64* cy;
    const int* list; const int* list_count; int* next_list; int* next_count;
    int* status; int* item_path_start; int* item_npaths;
    int* path_vstart; int* path_len; int path_cap; int* path_used;
    i64* pvx; i64* pvy; int vert_cap; int* vert_used;
    void* scratch; int n_threads;
};

// Physical_Processes/fracture_floe.m:12-52: the deformation of the floes about to be fractured (see fracture_deform_kernel)
struct FractureArgs {
    int count; const int* idx;                  // floe numbers, 1-based positions in the floe list of the last contact step
    int n0;
    const double* x; const double* y; const double* area; const int* voff; const double* vx; const double* vy;
    const int* row_off; const double* rows;     // Floe(i).interactions of that step: [K][7]
    unsigned char* changed; double* oxi; double* oyi; double* oarea; int* vstart; int* vcount; int* status;
    double* pvx; double* pvy; int vert_cap; int* vert_used; int* n_changed;
    void* scratch; int n_threads;
};

struct PtrGetter { const i64* x; const i64* y; SZ_HD P64 operator()(int i) const { P64 p; p.x = x[i]; p.y = y[i]; return p; } };

// sink that stores emitted paths straight into the global pools (space reserved beforehand)
struct PoolSink {
    i64* x; i64* y; int* vstart; int* len; int path; int vert;
    SZ_HD void begin_path(int cnt) { vstart[path] = vert; len[path] = cnt; ++path; }
    SZ_HD void point(P64 p) { x[vert] = p.x; y[vert] = p.y; ++vert; }
};
struct SizeSink { int paths, verts; SZ_HD void begin_path(int c) { ++paths; verts += c; } SZ_HD void point(P64) {} };

#if defined(__CUDACC__)

// All 32 lanes of a warp call this together (szpf::pair_force is warp-synchronous); `valid` says whether
// the lane has a pair.
template <class C, bool FAST, class W, int MODE = 0>
__device__ __forceinline__ void resolve_pair_impl(const NarrowArgs& a, int k, bool valid, W& w, int item = 0)
{
    int i = 0, j = -1;
    Body b1, b2;
    b1.h = b1.area = b1.Xi = b1.Yi = b1.Ui = b1.Vi = b1.ksi = 0; b2 = b1;
    if (valid && a.wall && (a.egid[k] <= a.P.Nb || !a.eowned[k])) valid = false;     // floes below Nb take no part in the pair loop (floe_interactions_all.m:125)
    bool escalate = false;
    szpf::PairHints hints{false, 0, 0, 0, 0};
    if (valid) {
        if (a.wall) { i = a.first_floe + k; b2 = a.bbody; }
        else { i = a.pi[k]; j = a.pj[k]; hints.convex = a.econvex[i] && a.econvex[j]; hints.rot1 = a.erot[i]; hints.no1 = a.eno[i]; hints.rot2 = a.erot[j]; hints.no2 = a.eno[j]; }
        const int si = a.esrc[i];
        const int o1 = a.voff[si], n1 = a.voff[si + 1] - o1;
        int o2 = 0, n2 = a.bn, sj = 0;
        if (!a.wall) { sj = a.esrc[j]; o2 = a.voff[sj]; n2 = a.voff[sj + 1] - o2; }
        // + 1: room for the closing vertex of floe_interactions.m:62-67
        if (n1 + 1 > C::NV || n2 + 1 > C::NV || n1 < 1 || n2 < 1) {
            if (n1 >= 1 && n2 >= 1 && a.next_list) escalate = true;
            else { a.status[k] = (n1 < 1 || n2 < 1) ? szpf::PS_BAD_POLY : szpf::PS_CAPACITY; a.nrows[k] = 0; a.ovl_state[k] = 0; if (a.want_polys) a.poly_npaths[k] = 0; }
            valid = false;
        } else {
            b1.h = a.h[si]; b1.area = a.area[si]; b1.Xi = a.ex[i]; b1.Yi = a.ey[i]; b1.Ui = a.u[si]; b1.Vi = a.v[si]; b1.ksi = a.ksi[si];
            w.n1 = n1; w.n2 = n2;
            for (int t = 0; t < n1; ++t) { w.c1x[t] = a.vx[o1 + t] + b1.Xi; w.c1y[t] = a.vy[o1 + t] + b1.Yi; }       // floe_interactions.m:25
            if (a.wall) { for (int t = 0; t < n2; ++t) { w.c2x[t] = a.bx[t]; w.c2y[t] = a.by[t]; } }                    // :31-32
            else {
                b2.h = a.h[sj]; b2.area = a.area[sj]; b2.Xi = a.ex[j]; b2.Yi = a.ey[j]; b2.Ui = a.u[sj]; b2.Vi = a.v[sj]; b2.ksi = a.ksi[sj];
                for (int t = 0; t < n2; ++t) { w.c2x[t] = a.vx[o2 + t] + b2.Xi; w.c2y[t] = a.vy[o2 + t] + b2.Yi; }    // floe_interactions_all.m:105
            }
        }
    }
    szpf::PairResult res;
    double rows[C::ROWS * 5];
    if constexpr (FAST && MODE == 2) {
        const szpf::ConvexHandoff ho{a.ho_st, a.ho_x, a.ho_y, a.ho_stride, a.ho_cap, item};
        szpf::pair_force_convex_after_sweep<C>(w, b1, b2, a.P, res, rows, valid, hints, ho);
    } else if constexpr (FAST) szpf::pair_force_convex<C>(w, b1, b2, a.P, res, rows, valid, hints);
    else szpf::pair_force(w, b1, b2, a.wall != 0, a.P, res, rows, valid, hints);
    if (valid && (res.status == szpf::PS_CAPACITY || res.status == szpf::PS_BAIL) && a.next_list) { escalate = true; valid = false; }
    if (escalate) { int t = atomicAdd(a.next_count, 1); a.next_list[t] = k; }
    if (!valid) return;
    a.status[k] = res.status; a.ovl_state[k] = res.overlap_state;
    int nr = (res.status == szpf::PS_OK) ? res.n_rows : 0;
    a.nrows[k] = nr;
    if (nr > 0) {
        int s = atomicAdd(a.row_used, nr);
        a.row_start[k] = s;
        if (s + nr <= a.row_cap) for (int t = 0; t < nr * 5; ++t) a.row_pool[(size_t)s * 5 + t] = rows[t];
    }
    if (a.want_polys) {
        int np = (res.status == szpf::PS_OK) ? w.ra_n : 0;
        a.poly_npaths[k] = np;
        if (np > 0) {
            const int nv = w.ra_off[np];
            int ps = atomicAdd(a.path_used, np), vs = atomicAdd(a.vert_used, nv);
            a.poly_path_start[k] = ps;
            if (ps + np <= a.path_cap && vs + nv <= a.vert_cap) {
                for (int q = 0; q < np; ++q) { a.path_vstart[ps + q] = vs + w.ra_off[q]; a.path_len[ps + q] = w.ra_off[q + 1] - w.ra_off[q]; }
                for (int t = 0; t < nv; ++t) { a.pvx[vs + t] = w.rax[t]; a.pvy[vs + t] = w.ray[t]; }
            }
        }
    }
}

template <class C>
__device__ __forceinline__ void resolve_pair(const NarrowArgs& a, int k, bool valid, szpf::Workspace<C>& w) { resolve_pair_impl<C, false>(a, k, valid, w); }

// class C (convex fast path): strictly convex floe-floe pairs, clip #1 by the four-edge sweep of sz_convex.cuh, sign
// test by its margin certificate; no arena.  A pair the fast path declines is appended to class S's list.
#ifndef SZ_C_TPB
#define SZ_C_TPB 512
#endif
#ifndef SZ_C_MINB
#define SZ_C_MINB 2
#endif
template <class C>
__global__ void __launch_bounds__(SZ_C_TPB, SZ_C_MINB) narrow_convex_kernel(const NarrowArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *a.list_count;
    if (blockIdx.x * blockDim.x >= n) return;           // whole CTA beyond the list: uniform exit
    szpf::WorkspaceLite<C> w;
    resolve_pair_impl<C, true>(a, t < n ? a.list[t] : 0, t < n, w);
}

// Experiment (sz_set_option "convex_split", off by default, not measured yet): class C as two kernels over the same work list.
// The fused kernel keeps ~1.5 KB of local memory in flight per thread -- 227 MB over the 151,552 resident threads, against
// 126 MB of L2 -- and moves 17x its algorithmic bytes through DRAM.  Split, the sweep needs the int64 outlines, the edge
// records and the deque, the force law the double outlines, the region and the Sutherland-Hodgman buffers: about half each.
// The intersection polygon (6 vertices on average) crosses in a strided global buffer.
template <class C>
__global__ void __launch_bounds__(SZ_C_TPB, SZ_C_MINB) narrow_convex_sweep_kernel(const NarrowArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *a.list_count;
    if (blockIdx.x * blockDim.x >= n) return;           // whole CTA beyond the list: uniform exit
    bool go = false;
    szpf::ClipInput subj, clip;
    subj.x = subj.y = nullptr; subj.dx = subj.dy = 0; subj.ix = subj.iy = nullptr; subj.n = 0; subj.ring = 0; subj.rot = 0;
    clip = subj;
    if (t < n) {
        const int k = a.list[t];
        const int i = a.pi[k], j = a.pj[k];
        const int si = a.esrc[i], sj = a.esrc[j];
        const int o1 = a.voff[si], n1 = a.voff[si + 1] - o1, o2 = a.voff[sj], n2 = a.voff[sj + 1] - o2;
        const int no1 = a.eno[i], no2 = a.eno[j];
        // the same pairs the fused kernel sweeps: both outlines strictly convex and small enough for the class
        if (a.econvex[i] && a.econvex[j] && no1 >= 3 && no2 >= 3 && n1 >= 1 && n2 >= 1 && n1 + 1 <= C::NV && n2 + 1 <= C::NV) {
            go = true;
            subj.x = a.vx + o1; subj.y = a.vy + o1; subj.dx = a.ex[i]; subj.dy = a.ey[i]; subj.n = no1; subj.rot = a.erot[i];     // floe_interactions.m:25
            clip.x = a.vx + o2; clip.y = a.vy + o2; clip.dx = a.ex[j]; clip.dy = a.ey[j]; clip.n = no2; clip.rot = a.erot[j];     // floe_interactions_all.m:105
        } else a.ho_st[t] = -1;
    }
    i64 svx[2 * C::NV], svy[2 * C::NV], dqx[C::RV], dqy[C::RV], ox[C::RV], oy[C::RV];
    const szpf::ConvexHandoff ho{a.ho_st, a.ho_x, a.ho_y, a.ho_stride, a.ho_cap, t};
    szpf::convex_sweep_only<C::NV>(go, subj, clip, svx, svy, dqx, dqy, ox, oy, C::RV, ho);
}
template <class C>
__global__ void __launch_bounds__(SZ_C_TPB, SZ_C_MINB) narrow_convex_force_kernel(const NarrowArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *a.list_count;
    if (blockIdx.x * blockDim.x >= n) return;
    szpf::WorkspaceLite<C> w;
    resolve_pair_impl<C, true, szpf::WorkspaceLite<C>, 2>(a, t < n ? a.list[t] : 0, t < n, w, t);
}

// class S launch shape: two 512-thread CTAs per SM (64 registers per thread).  With SZ_BLOCK_SYNC (default) all
// warps of a CTA walk the sweep phases together and share the instruction lines of each phase; measured on B200 at 1M
// floes (first version of the sweep): warp-synchronous 128-thread CTAs 181 ms, block-synchronous 512 threads 154 ms,
// 1024 threads 147 ms; after the work list was bucketed, 2 x 512 with one CTA vote per scanbeam is 5 % ahead of 1 x 1024.
#ifndef SZ_S_MINB
#define SZ_S_MINB 2
#endif
#ifndef SZ_S_TPB
#define SZ_S_TPB 512
#endif
// class S: arena in local memory, one thread per work item (pairs come through the bucketed work list)
template <class C>
__global__ void __launch_bounds__(SZ_S_TPB, SZ_S_MINB) narrow_local_kernel(const NarrowArgs a)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = a.list ? *a.list_count : a.n_work;
    if (blockIdx.x * blockDim.x >= n) return;           // whole CTA beyond the list: uniform exit
    szpf::Workspace<C> w;
    resolve_pair<C>(a, t < n ? (a.list ? a.list[t] : t) : 0, t < n, w);
}

// classes M/L: arena in HBM scratch, persistent threads striding over the work list (n_threads is a
// multiple of the block size, so whole warps stay together)
template <class C>
__global__ void __launch_bounds__(64) narrow_scratch_kernel(const NarrowArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    szpf::Workspace<C>& w = reinterpret_cast<szpf::Workspace<C>*>(a.scratch)[tid];
    const int n = *a.list_count;
    for (int base = 0; base < n; base += a.n_threads) {
        const int t = base + tid;
        resolve_pair<C>(a, t < n ? a.list[t] : 0, t < n, w);
    }
}

// One thread per selected floe (persistent threads over the list, class-L arena in HBM scratch; all threads of the CTA
// walk the two sweeps together like the narrow phase).  fracture_floe.m:17-50.
template <class C>
__global__ void __launch_bounds__(64) fracture_deform_kernel(const FractureArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    szpf::Workspace<C>& w = reinterpret_cast<szpf::Workspace<C>*>(a.scratch)[tid];
    for (int base = 0; base < a.count; base += a.n_threads) {
        const int q = base + tid;
        bool live = q < a.count;
        int i = 0, j = 0, st = 0;
        double fx = 0, fy = 0;
        if (live) {
            i = a.idx[q] - 1;
            a.changed[q] = 0; a.oxi[q] = a.x[i]; a.oyi[q] = a.y[i]; a.oarea[q] = a.area[i]; a.vstart[q] = 0; a.vcount[q] = 0; a.status[q] = 0;
            // a = interactions without the wall rows; [~,k] = max(a(:,7)): first maximum (:17-22)
            int k = -1; double best = 0;
            for (int r = a.row_off[i]; r < a.row_off[i + 1]; ++r) {
                const double* row = a.rows + (size_t)r * 7;
                if (row[0] == SZ_INF || row[0] == -SZ_INF) continue;
                if (k < 0 || row[6] > best) { best = row[6]; k = r; }
            }
            if (k < 0) live = false;
            else {
                const double* row = a.rows + (size_t)k * 7;
                if (!(row[0] < a.n0 + 1)) live = false;                 // :26: the partner must be an original floe
                else { j = (int)row[0] - 1; fx = row[1]; fy = row[2]; }
            }
        }
        if (live) {
            const int o1 = a.voff[i], n1 = a.voff[i + 1] - o1, o2 = a.voff[j], n2 = a.voff[j + 1] - o2;
            if (n1 + 1 > C::NV || n2 + 1 > C::NV || n1 < 1 || n2 < 1) { st = szpf::PS_CAPACITY; live = false; }
            else {
                w.n1 = n1; w.n2 = n2;
                for (int t = 0; t < n1; ++t) { w.c1x[t] = a.vx[o1 + t] + a.x[i]; w.c1y[t] = a.vy[o1 + t] + a.y[i]; }    // :23-24
                for (int t = 0; t < n2; ++t) { w.c2x[t] = a.vx[o2 + t] + a.x[j]; w.c2y[t] = a.vy[o2 + t] + a.y[j]; }    // :27-28
            }
        }
        szpf::ClipInput subj, clip;
        subj.x = w.c1x; subj.y = w.c1y; subj.dx = 0; subj.dy = 0; subj.ix = subj.iy = 0; subj.n = w.n1; subj.ring = 0; subj.rot = 0;
        clip.x = w.c2x; clip.y = w.c2y; clip.dx = 0; clip.dy = 0; clip.ix = clip.iy = 0; clip.n = w.n2; clip.ring = 0; clip.rot = 0;
        int cs = szpf::run_sweep(w.eng, live, 1, subj, clip);                                     // :29 polyclip(..., 'int')
        if (live && cs != szpf::PS_OK) { st = cs; live = false; }
        int nr = 0;
        if (live) {
            szpf::RegionSink<C> sink(w.rax, w.ray, w.ra_off);
            w.eng.emit(sink);
            if (sink.overflow) { st = szpf::PS_CAPACITY; live = false; }
            else if (sink.n_paths == 0) live = false;                                             // :32
            else nr = w.ra_off[1];
        }
        if (live) {
            double ar, xm, ym;
            szpf::ring_area_centroid(w.rax, w.ray, nr, ar, xm, ym);                               // :34 centroid(polyshape(Xt,Yt))
            if (nr + 1 > C::NP) { st = szpf::PS_CAPACITY; live = false; }
            else {
                for (int t = 0; t < nr; ++t) { w.px[t] = (double)w.rax[t] / SZ_SCALE; w.py[t] = (double)w.ray[t] / SZ_SCALE; }
                int nc = nr;
                if (w.px[0] != w.px[nr - 1] || w.py[0] != w.py[nr - 1]) { w.px[nr] = w.px[0]; w.py[nr] = w.py[0]; nc = nr + 1; }   // p_poly_dist.m:135-141
                if (!szpf::outline_ok_for_poly_dist_xy(w.px, w.py, nc)) { st = szpf::PS_BAD_POLY; live = false; }
                else {
                    const double d = szpf::abs_poly_dist_xy(w.px, w.py, nc, xm, ym);              // :35 |p_poly_dist|
                    const double F = sqrt(fx * fx + fy * fy);                                     // :36
                    clip.dx = fx * d / 2 / F; clip.dy = fy * d / 2 / F;                           // :37-39
                }
            }
        }
        cs = szpf::run_sweep(w.eng, live, 0, subj, clip);                                         // :40 polyclip(..., 'dif')
        if (live && cs != szpf::PS_OK) { st = cs; live = false; }
        if (live) {
            szpf::RegionSink<C> sink(w.rax, w.ray, w.ra_off);
            w.eng.emit(sink);
            if (sink.overflow) { st = szpf::PS_CAPACITY; live = false; }
            else if (sink.n_paths == 0) live = false;
            else nr = w.ra_off[1];
        }
        if (live) {
            const double anew = szpf::ring_polyarea(w.rax, w.ray, nr);                            // :43
            if (anew / a.area[i] > 0.9) {                                                         // :44
                double ar, xm, ym;
                szpf::ring_area_centroid(w.rax, w.ray, nr, ar, xm, ym);                           // :45
                const int vs = atomicAdd(a.vert_used, nr);
                a.changed[q] = 1; a.oxi[q] = xm; a.oyi[q] = ym; a.oarea[q] = anew; a.vstart[q] = vs; a.vcount[q] = nr;    // :46-48
                atomicAdd(a.n_changed, 1);
                if (vs + nr <= a.vert_cap) for (int t = 0; t < nr; ++t) { a.pvx[vs + t] = (double)w.rax[t] / SZ_SCALE - xm; a.pvy[vs + t] = (double)w.ray[t] / SZ_SCALE - ym; }
            }
        }
        if (q < a.count && st != 0) a.status[q] = st;
    }
}

template <class CC>
__device__ __forceinline__ void resolve_clip(const ClipArgs& a, int k, szclip::ClipEngine<CC>& eng)
{
    const i64 s0 = a.soff[k], c0 = a.coff[k];
    const int ns = (int)(a.soff[k + 1] - s0), nc = (int)(a.coff[k + 1] - c0);
    eng.begin(a.method[k]);
    PtrGetter gs{a.sx + s0, a.sy + s0}, gc{a.cx + c0, a.cy + c0};
    eng.add_path(gs, ns, 0);
    eng.add_path(gc, nc, 1);
    int st = eng.execute();
    if (st == szclip::ST_OVERFLOW) {
        if (a.next_list) { int t = atomicAdd(a.next_count, 1); a.next_list[t] = k; }
        else { a.status[k] = szpf::PS_CAPACITY; a.item_npaths[k] = 0; }
        return;
    }
    if (st != szclip::ST_OK) { a.status[k] = szpf::PS_CLIPPER_FAIL; a.item_npaths[k] = 0; return; }
    SizeSink sz; sz.paths = 0; sz.verts = 0;
    eng.emit(sz);
    a.status[k] = 0; a.item_npaths[k] = sz.paths;
    if (sz.paths > 0) {
        int ps = atomicAdd(a.path_used, sz.paths), vs = atomicAdd(a.vert_used, sz.verts);
        a.item_path_start[k] = ps;
        if (ps + sz.paths <= a.path_cap && vs + sz.verts <= a.vert_cap) {
            PoolSink sink{a.pvx, a.pvy, a.path_vstart, a.path_len, ps, vs};
            eng.emit(sink);
        }
    }
}
template <class CC>
__global__ void __launch_bounds__(128) clip_local_kernel(const ClipArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.count) return;
    szclip::ClipEngine<CC> eng;
    resolve_clip<CC>(a, k, eng);
}
template <class C>
__global__ void __launch_bounds__(64) clip_scratch_kernel(const ClipArgs a)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= a.n_threads) return;
    szpf::Workspace<C>& w = reinterpret_cast<szpf::Workspace<C>*>(a.scratch)[tid];
    const int n = *a.list_count;
    for (int t = tid; t < n; t += a.n_threads) resolve_clip<typename C::Clip>(a, a.list[t], w.eng);
}
// calc_eulerian_data.m:140-147, one (cell, floe) item per thread: the shared area by one Clipper sweep (sz_euler.cuh).
// Class S keeps the arena in local memory; an item whose arena overflows is appended to the class L list.
template <class CC>
__global__ void __launch_bounds__(128) euler_item_local_kernel(const szeul::EulerArgs a, int* next_list, int* next_count)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= a.n_items) return;
    szclip::ClipEngine<CC> eng;
    double ar = 0;
    const int st = szeul::item_area(eng, a, k, ar);
    if (st == szpf::PS_CAPACITY && next_list) { const int t = atomicAdd(next_count, 1); next_list[t] = k; return; }
    a.item_status[k] = st; a.item_area[k] = ar;
}
template <class C>
__global__ void __launch_bounds__(64) euler_item_scratch_kernel(const szeul::EulerArgs a, const int* list, const int* list_count, void* scratch, int n_threads)
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n_threads) return;
    szpf::Workspace<C>& w = reinterpret_cast<szpf::Workspace<C>*>(scratch)[tid];
    const int n = *list_count;
    for (int t = tid; t < n; t += n_threads) {
        const int k = list[t];
        double ar = 0;
        const int st = szeul::item_area(w.eng, a, k, ar);
        a.item_status[k] = st; a.item_area[k] = ar;
    }
}
#endif  // __CUDACC__

}  // namespace sznarrow

// launchers (one translation unit per class); all asynchronous on `stream`
extern "C" {
void sz_launch_narrow_C(const sznarrow::NarrowArgs* a, cudaStream_t stream);
void sz_launch_narrow_C_split(const sznarrow::NarrowArgs* a, cudaStream_t stream);
void sz_launch_narrow_S(const sznarrow::NarrowArgs* a, cudaStream_t stream);
void sz_launch_narrow_T(const sznarrow::NarrowArgs* a, cudaStream_t stream);
void sz_launch_narrow_M(const sznarrow::NarrowArgs* a, cudaStream_t stream);
void sz_launch_narrow_L(const sznarrow::NarrowArgs* a, cudaStream_t stream);
void sz_launch_clip_S(const sznarrow::ClipArgs* a, cudaStream_t stream);
void sz_launch_clip_M(const sznarrow::ClipArgs* a, cudaStream_t stream);
void sz_launch_clip_L(const sznarrow::ClipArgs* a, cudaStream_t stream);
void sz_launch_fracture_L(const sznarrow::FractureArgs* a, cudaStream_t stream);
void sz_launch_euler_S(const szeul::EulerArgs* a, int* next_list, int* next_count, cudaStream_t stream);
void sz_launch_euler_L(const szeul::EulerArgs* a, const int* list, const int* list_count, void* scratch, int n_threads, cudaStream_t stream);
size_t sz_workspace_bytes_M(void);
size_t sz_workspace_bytes_L(void);
}
