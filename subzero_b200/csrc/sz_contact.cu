// sz_contact.cu -- the contact step on one B200: ghost floes, spatial-hash broad phase, narrow-phase
// orchestration, contact-row assembly (mirror, torque, per-floe sums, stress, collision count) and
// the C ABI of include/subzero_b200.h.  Replaces floe_interactions_all.m:9-285 of the reference.
//
// Device data layout (all structure-of-arrays, FP64 unless noted):
//   originals [n0]       x y rmax h area u v ksi, alive u8, voff i32[n0+1]; vertex pool vx vy [V]
//                        (c_alpha about the centroid, closed) -- uploaded once per step / per topology
//   extended list [n]    ex ey (centroid, ghost-shifted), esrc (original sharing outline + body),
//                        efn (FloeNums, negative for ghosts), eparent, ealive  -- built by K0
//   cell grid            cell_start i32[ncell+1]; floes bucketed by cell: s_idx s_x s_y s_r
//   pairs [P]            pi pj ascending (i,j) = order of Floe(i).potentialInteractions; pair_off[n+1]
//   per pair             status nrows row_start ovl_state; contact-row pool 5 doubles per row
//   transpose            toff[n+1], tlist: for each floe the pairs in which it is the partner (mirror)
//   rows [K][7]          canonical order of Floe(m).interactions; row_off[n+1]
//
// Kernels: K0 ghost_flag/ghost_emit (+ scan), bbox; K1 cell_count/cell_fill, broad<count>/<fill>
// (warp per floe, ballot compaction, in-warp rank sort); K2+K3 narrow phase (sz_narrow*.cu);
// K4 tcount/tfill, rowcount, assemble, fold, kill fix-up.  Every FP64 expression keeps the
// reference's operation order; the file is compiled with -fmad=false.
#include "../../include/subzero_b200.h"
#include "sz_narrow.cuh"
#include "sz_corners.cuh"
#include "sz_apart.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <cmath>
#include <vector>
#include <algorithm>

using sznarrow::NarrowArgs;
using sznarrow::ClipArgs;
using szpf::Params;
using szpf::Body;
typedef long long i64;
typedef unsigned long long u64;

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static long long g_launches = 0;   // kernels launched by this library (sz_launch_count)
void sz_set_error(const char* fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
extern "C" const char* sz_last_error(void) { return g_err; }
extern "C" int sz_abi_version(void) { return 1; }
extern "C" long long sz_launch_count(void) { return g_launches; }

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    sz_set_error("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); return SZ_ERR_CUDA; } } while (0)

#define CKS(call) do { int r_ = (call); if (r_ != SZ_OK) return r_; } while (0)

template <class T>
struct DBuf {
    T* p = nullptr; size_t cap = 0;
    cudaError_t ensure(size_t n, bool keep = false)
    {
        if (n <= cap) return cudaSuccess;
        size_t ncap = n + n / 4 + 64;
        T* q = nullptr;
        cudaError_t e = cudaMalloc(&q, ncap * sizeof(T));
        if (e != cudaSuccess) return e;
        static const bool poison = getenv("SZ_DEBUG_POISON") != nullptr;      // debugging aid: fresh buffers start as 0xFF.. instead of whatever was there
        // (the memset runs on the NULL stream, which the contexts' non-blocking streams do not wait for: without the synchronisation it
        // can land AFTER the first real writes to the new buffer -- it did, and looked like a flaky physics bug for an hour)
        if (poison) { cudaMemset(q, 0x7F, ncap * sizeof(T)); cudaDeviceSynchronize(); }
        if (keep && p && cap) cudaMemcpy(q, p, cap * sizeof(T), cudaMemcpyDeviceToDevice);
        if (p) cudaFree(p);
        p = q; cap = ncap;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

#define SZ_BIN_N 16                       // vertex counts 0..15 per outline get their own bucket (larger ones share the last)
#define SZ_NSECT 16
#define SZ_NBINS (SZ_BIN_N * SZ_BIN_N * SZ_NSECT) // x 16 sectors of the partner's direction
// counters living in device memory, mirrored into pinned host memory with one copy
struct Counters {
    int n1, n;                     // extended-list sizes after the x pass / after the y pass
    int n_pairs;
    int row_used, path_used, vert_used;
    int listC, listS, listT, listM, listL, wlistT, wlistM, wlistL;
    int listC0, listS0;            // list lengths before class C ran (listS grows by the pairs class C declines)
    int total_rows;
    int n_pairs_force, n_fail, n_cap_fail, n_pairs_owned, n_bbox_reject, n_kill_events;
    int overflow;                  // a speculated capacity (list, pairs, rows) was too small: the step is repeated with measured sizes
    u64 bbox[4];                   // order-preserving encodings of min x, max x, min y, max y
    u64 rmax_bits;
    u64 n_fin_rows, n_inf_rows;
    int clip_listM, clip_listL, clip_path_used, clip_vert_used;
    int n_forced, n_no_points;     // ocean forcing: floes evaluated / floes without a Monte-Carlo point inside
    int fr_vert_used, fr_changed;  // fracture deformation: vertices of the new outlines, floes changed
    int cr_n1, cr_n;               // corners.m's own periodic list: sizes after the x pass / after the y pass
    int eu_n1, eu_n2, eu_n, eu_yflag, eu_listL, eu_fail, eu_cap;   // calc_eulerian_data: list sizes (alive, + x images, + y images), the stale-polygon flag, class L items, failures
    int any_concave;               // some AddPath-valid outline is not strictly convex (ext_prep_kernel): the classifier's edge-by-edge rule has work
};

struct SzContext {
    int device = 0;
    cudaStream_t stream = nullptr;          // the stream every launch and copy of the library goes to
    cudaStream_t own_stream = nullptr;      // created by sz_create; `stream` points elsewhere after sz_set_stream
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evp[4] = {nullptr, nullptr, nullptr, nullptr};
    float phase_ms[5] = {0, 0, 0, 0, 0};
    cudaEvent_t evk[10] = {};                  // start/stop of the narrow-phase launch of each size class (C, S, T, M, L)
    bool evk_used[5] = {false, false, false, false, false}; int class_pairs[5] = {0, 0, 0, 0, 0};
    int opt_convex_fast = 1;
    int opt_graph_safe = 0;          // the step may be captured into a CUDA graph: no host-carried scan state, no event timing in sz_step_enqueue
    struct { bool active = false; int n = 0, np = 0; long long rows_bound = 0; } pend; bool enq_mode = false;      // a step enqueued by sz_step_enqueue and not finished yet
    int opt_speculate = 1;           // sizes of the step (list, grid, pairs, rows) carried over from the previous step: one counter read per step instead of four
    bool plan_valid = false; int plan_n0 = -1, plan_nl0 = -1, plan_ncap = 0, plan_npcap = 0; long long plan_rowscap = 0; struct GridDescHost { double x0, y0, cell; int nx, ny; } plan_g = {0, 0, 1, 1, 1};
    int n_fast_steps = 0, n_slow_steps = 0;
    int opt_convex_split = 0;        // experiment: class C as two kernels (sweep, then force law)
    int opt_apart = 1;               // classifier: outlines of any shape certified apart edge by edge (sz_apart.cuh); 0 = bounding boxes only, as before
    int opt_euler_cell_warp = 1;     // calc_eulerian_data: a warp per cell (0: one thread per cell)
    Counters* d_cnt = nullptr; Counters* h_cnt = nullptr; Counters* h_init = nullptr;      // h_init: the step's initial counters (pinned; never the target of a read-back, so a replayed copy node finds them unchanged)
    // inputs
    bool have_input = false, have_step = false;
    bool have_rows = false;                 // the contact rows of the last step are still on the device (also after the integrator moved the floes)
    SzParams prm; Params dprm; bool have_bnd = false; Body bbody;
    int n0 = 0; i64 nverts = 0;
    DBuf<double> x, y, rmax, h, area, u, v, ksi, vx, vy, bx, by, boxx, boxy;
    DBuf<uint8_t> alive;
    DBuf<int> voff;
    int bn = 0, boxn = 0;
    // extended list
    int n = 0, n1 = 0;
    DBuf<double> ex, ey, erootx, erooty; DBuf<int> esrc, efn, eparent, gx_of, gy_of, egid; DBuf<uint8_t> ealive, eowned;
    bool ext_mode = false;          // extended list supplied by the caller (multi-GPU slabs), K0 skipped
    // slab mode (sz_slab_*): the rank's owned floes are the resident originals [0, n0); halo records follow them in the body arrays and the
    // extended list is rebuilt on the device every step
    bool slab = false, sl_configured = false, sl_built = false; int sl_rank = 0, sl_world = 1, sl_nglobal = 0, sl_cap_img = 0, sl_cap_rec = 0, sl_cap_vert = 0, sl_nl_cap = 0;
    struct SlabScratch* sl_scratch = nullptr; double sl_xlo = -SZ_INF, sl_xhi = SZ_INF;
    DBuf<int> sl_ogid, sl_flag, sl_pos, sl_opos, sl_g, sl_sendcnt, sl_keys, sl_slots, sl_hnv, sl_hvoff; DBuf<uint8_t> sl_cub; DBuf<double> sl_out, sl_rows; DBuf<int> sl_rcnt, sl_roff;
    int nout = 0;                   // entries with per-floe outputs: n0 (single GPU) or n (extended mode)
    DBuf<int> flag, pos, scan_tmp;
    DBuf<u64> scan_state; u64 scan_ticket_base = 0; unsigned scan_epoch = 0;      // single-pass scan: [0] ticket counter, [1..] tile status words
    // grid
    DBuf<int> cid, cell_cnt, cell_start, s_idx; DBuf<double> s_x, s_y, s_r;
    // pairs
    int n_pairs = 0;
    DBuf<int> pcnt, pair_off, pi, pj, pstatus, pnrows, prow_start; DBuf<double> povl;
    DBuf<int> stage, listC, listS, listT, wlistT, listM, listL, env, bins, bin_fill; /* bins: class-S work list buckets */ DBuf<i64> ebb, erec /* EntryRec [n], 8 words each */; DBuf<short> pkey; DBuf<uint8_t> evalid, econvex, erot, eno, papart /* per pair: certified apart */;
    DBuf<int> wstatus, wnrows, wrow_start, wlistM, wlistL; DBuf<double> wovl;
    DBuf<double> row_pool;
    DBuf<int> poly_path_start, poly_npaths, path_vstart, path_len; DBuf<i64> pvx, pvy;
    DBuf<unsigned char> scratchM, scratchL;
    DBuf<int> ho_st; DBuf<i64> ho_x, ho_y;      // experiment convex_split: clip #1 polygons between the two class C kernels
    // assembly
    DBuf<int> tcnt, toff, tlist, rcnt, row_off;
    DBuf<double> rows; i64 n_rows = 0;
    DBuf<double> osum;             // [n*3] own column sums Fx Fy tau
    DBuf<double> e_ov;             // [n] OverlapArea of every entry (the periodic images included: sz_get_ghost_outputs)
    DBuf<uint8_t> has_rows;
    DBuf<int> kill_i, transfer_i, tmax;
    // per-original outputs
    DBuf<double> o_fx, o_fy, o_tq, o_ov, o_stress, o_xi, o_yi; DBuf<uint8_t> o_alive; DBuf<int> o_kill, o_transfer;
    SzSummary summary;
    // integrator state (sz_trajectory_*): calc_trajectory.m fields kept on the device between steps
    bool have_traj = false; int traj_nz = 0;
    DBuf<double> t_mass, t_inertia, t_alpha, t_dXi_p, t_dYi_p, t_dUi_p, t_dVi_p, t_dalpha_p, t_dksi_p, t_FxOA, t_FyOA, t_torqueOA, c0x, c0y, t_stressH, t_stress;
    // fracture deformation (fracture_floe.m:12-52)
    DBuf<int> fr_idx, fr_vstart, fr_vcount, fr_status; DBuf<uint8_t> fr_changed; DBuf<double> fr_xi, fr_yi, fr_area, fr_vx, fr_vy; int fr_count = 0; i64 fr_verts = 0; bool have_fr = false;
    // corner mask (corners.m:10-88)
    DBuf<int> cr_idx, cr_nv, cr_off, cr_esrc; DBuf<uint8_t> cr_da, cr_ealive; DBuf<double> cr_ex, cr_ey; int cr_count = 0; i64 cr_verts = 0; bool have_cr = false;
    // coarse-grid averages (calc_eulerian_data.m)
    DBuf<int> eu_lsrc, eu_icnt, eu_ioff, eu_cell, eu_q, eu_status, eu_iota, eu_sorted, eu_keys, eu_ccnt, eu_coff, eu_listL; DBuf<double> eu_lx, eu_ly, eu_in, eu_area, eu_out; DBuf<uint8_t> eu_tmp;
    // ocean / atmosphere forcing (calc_trajectory.m:94-166)
    DBuf<double> oc_Xo, oc_Yo, oc_U, oc_V, oc_Wu, oc_Wv, pt_x, pt_y, t_strain; DBuf<uint8_t> pt_a, t_forced;
    int oc_nx = 0, oc_ny = 0, npts = 0; bool have_ocean = false, have_points = false, traj_do_int = false;
    double oc_fc = 0, oc_turn = 0, oc_rho0 = 1027, oc_Cd = 3e-3, oc_rho_air = 1.2, oc_Cd_atm = 1e-3;
    DBuf<int> t_scount, t_flags;
    // weld / FloeSimplify pair searches
    DBuf<int> se_bin, se_cnt, se_off, se_q, se_cid, se_ccnt, se_cstart, se_sidx, se_tmp, se_out; DBuf<double> se_sx, se_sy, se_sr;
    int se_nq = 0, se_np = 0, se_first = 0; bool se_weld = false, have_search = false;
    // clip batch
    int clip_count = 0; i64 clip_paths = 0, clip_verts = 0;
    DBuf<int> c_method, c_status, c_path_start, c_npaths, c_path_vstart, c_path_len, c_listM, c_listL;
    DBuf<i64> c_soff, c_coff, c_sx, c_sy, c_cx, c_cy, c_pvx, c_pvy;
};

// ------------------------------------------------------------------------------------------------ scan
// Exclusive prefix sum of int32: out[k] = sum in[0..k), k in [0, n_out).  Reads of in[k] for k >= n_in yield 0, so calling with
// n_out = n_in + 1 leaves the grand total in out[n_in].
// ONE launch (the step runs about ten scans over lists of 1e5..1e6 entries, where three launches each were a visible part
// of the step's fixed cost): single-pass scan with decoupled look-back.  Tiles take their number from a ticket counter, so a
// tile only ever waits for tiles that are already running; a tile publishes its aggregate, looks back over its predecessors'
// words until it meets an inclusive prefix, and publishes its own.  A status word carries the launch's epoch, so the array
// needs no reset between launches: [epoch 24 | flag 2 (1 aggregate, 2 inclusive prefix) | pad 6 | value 32].
#define SCAN_TPB 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_TPB * SCAN_ITEMS)
__global__ void __launch_bounds__(SCAN_TPB) scan_lookback_kernel(const int* __restrict__ in, int n_in, int* __restrict__ out, int n_out,
                                                                 volatile u64* status, u64* ticket, u64 ticket_base, unsigned epoch)
{
    __shared__ int warp_sums[SCAN_TPB / 32];
    __shared__ int s_tile, s_prefix;
    if (threadIdx.x == 0) s_tile = (int)(atomicAdd(ticket, 1ULL) - ticket_base);
    __syncthreads();
    const int tile = s_tile;
    const int base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS]; int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { int idx = base + k; v[k] = (idx < n_in) ? in[idx] : 0; s += v[k]; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int ws = (lane < SCAN_TPB / 32) ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_TPB / 32; d <<= 1) { int t = __shfl_up_sync(0xffffffffu, ws, d); if (lane >= d) ws += t; }
        if (lane < SCAN_TPB / 32) warp_sums[lane] = ws;   // inclusive
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int agg = warp_sums[SCAN_TPB / 32 - 1];
        const u64 tag = (u64)epoch << 40;
        int excl = 0;
        if (tile == 0) status[0] = tag | (2ULL << 38) | (unsigned)agg;
        else {
            status[tile] = tag | (1ULL << 38) | (unsigned)agg;
            for (int p = tile - 1; ; --p) {
                u64 w;
                do { w = status[p]; } while ((w >> 40) != epoch || ((w >> 38) & 3ULL) == 0);
                excl += (int)(unsigned)(w & 0xffffffffULL);
                if (((w >> 38) & 3ULL) == 2) break;
            }
            status[tile] = tag | (2ULL << 38) | (unsigned)(excl + agg);
        }
        s_prefix = excl;
    }
    __syncthreads();
    int run = s_prefix + incl - s + (wid > 0 ? warp_sums[wid - 1] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { int idx = base + k; if (idx < n_out) out[idx] = run; run += v[k]; }
}
static int exclusive_scan(SzContext* c, const int* in, int n_in, int* out, int n_out)
{
    if (n_out <= 0) return SZ_OK;
    const int tiles = (n_out + SCAN_TILE - 1) / SCAN_TILE;
    if ((size_t)tiles + 2 > c->scan_state.cap) {
        CK(cudaStreamSynchronize(c->stream));
        CK(c->scan_state.ensure((size_t)tiles + 2));
        CK(cudaMemset(c->scan_state.p, 0, c->scan_state.cap * 8));
        c->scan_ticket_base = 0; c->scan_epoch = 0;
    }
    if (c->opt_graph_safe) {
        // a launch that may be replayed from a CUDA graph cannot carry the host's ticket base and epoch: it resets the ticket
        // and the status words it is going to use instead (one small memset node)
        CK(cudaMemsetAsync(c->scan_state.p, 0, ((size_t)tiles + 1) * 8, c->stream));
        ++g_launches;
        scan_lookback_kernel<<<tiles, SCAN_TPB, 0, c->stream>>>(in, n_in, out, n_out, c->scan_state.p + 1, c->scan_state.p, 0ULL, 0xFFFFFFu);
        c->scan_ticket_base = 0; c->scan_epoch = 0;      // (the next ordinary launch starts from a clean array: epoch 0xFFFFFF words are stale for it)
        CK(cudaMemsetAsync(c->scan_state.p, 0, 8, c->stream));      // (the ticket)
        return SZ_OK;
    }
    if (++c->scan_epoch >= (1u << 24) - 1) { CK(cudaMemsetAsync(c->scan_state.p + 1, 0, (c->scan_state.cap - 1) * 8, c->stream)); c->scan_epoch = 1; }
    ++g_launches;
    scan_lookback_kernel<<<tiles, SCAN_TPB, 0, c->stream>>>(in, n_in, out, n_out, c->scan_state.p + 1, c->scan_state.p, c->scan_ticket_base, c->scan_epoch);
    c->scan_ticket_base += (u64)tiles;
    return SZ_OK;
}
static size_t scan_tmp_ints(size_t n) { (void)n; return 16; }      // (the single-pass scan keeps its state in the context)

// ------------------------------------------------------------------------------------------------ K0 ghosts
__device__ __forceinline__ double sgn_d(double v) { return (double)((v > 0) - (v < 0)); }

// flag[i] = alive(i) && max_v |c_alpha(axis,v) + centroid(axis)| > L      (floe_interactions_all.m:30-31, 51-52)
__global__ void ghost_flag_kernel(int axis, int n_bound, const int* __restrict__ n_dev, const double* __restrict__ ec,
                                  const int* __restrict__ esrc, const uint8_t* __restrict__ ealive,
                                  const int* __restrict__ voff, const double* __restrict__ vc, double L, int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bound) return;
    const int n = n_dev ? *n_dev : n_bound;
    int f = 0;
    if (i < n && ealive[i]) {
        const int s = esrc[i]; const double c = ec[i];
        double m = -SZ_INF;
        for (int t = voff[s]; t < voff[s + 1]; ++t) { double a = fabs(vc[t] + c); if (a > m) m = a; }
        f = (m > L);
    }
    flag[i] = f;
}
// appended copy with the centroid shifted by -2L*sign(centroid)             (:33-36, 54-57)
__global__ void ghost_emit_kernel(int axis, int n_bound, const int* __restrict__ n_dev, int n0, const int* __restrict__ flag, const int* __restrict__ pos,
                                  double* __restrict__ ex, double* __restrict__ ey, int* __restrict__ esrc, int* __restrict__ efn,
                                  int* __restrict__ eparent, uint8_t* __restrict__ ealive, int* __restrict__ gx_of, int* __restrict__ gy_of,
                                  double L, int* __restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_dev ? *n_dev : n_bound;
    if (i == 0) *n_out = n + pos[n_bound];
    if (i >= n || !flag[i]) return;
    const int g = n + pos[i];
    if (axis == 0) { ex[g] = ex[i] - 2 * L * sgn_d(ex[i]); ey[g] = ey[i]; if (i < n0) gx_of[i] = g; }
    else { ex[g] = ex[i]; ey[g] = ey[i] - 2 * L * sgn_d(ey[i]); if (i < n0) gy_of[i] = g; }
    esrc[g] = esrc[i]; efn[g] = -abs(efn[i]); eparent[g] = i + 1; ealive[g] = ealive[i];
}
__global__ void init_extended_kernel(int n0, const double* __restrict__ x, const double* __restrict__ y, const uint8_t* __restrict__ alive,
                                     double* ex, double* ey, int* esrc, int* efn, int* eparent, uint8_t* ealive, int* gx_of, int* gy_of)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0) return;
    ex[i] = x[i]; ey[i] = y[i]; esrc[i] = i; efn[i] = i + 1; eparent[i] = 0; ealive[i] = alive[i]; gx_of[i] = -1; gy_of[i] = -1;
}

// single-GPU mode: global position, ownership and root centroid of every entry of the extended list
__global__ void finish_extended_kernel(int n_bound, const int* __restrict__ n_dev, const double* __restrict__ x, const double* __restrict__ y,
                                       const int* __restrict__ esrc, int* __restrict__ egid, uint8_t* __restrict__ eowned,
                                       double* __restrict__ erootx, double* __restrict__ erooty)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_bound || e >= *n_dev) return;
    egid[e] = e + 1; eowned[e] = 1; erootx[e] = x[esrc[e]]; erooty[e] = y[esrc[e]];
}
// extended mode: the (at most two) ghost children of every entry, ascending = creation order (:242-245)
__global__ void child_init_kernel(int n, int* c0, int* c1) { const int e = blockIdx.x * blockDim.x + threadIdx.x; if (e < n) { c0[e] = 0x7fffffff; c1[e] = -1; } }
__global__ void child_mark_kernel(int n, const int* __restrict__ parent, int* c0, int* c1)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int p = parent[e] - 1;          // eparent is 1-based, 0 = none
    if (p >= 0 && p < n) { atomicMin(&c0[p], e); atomicMax(&c1[p], e); }
}
__global__ void child_final_kernel(int n, int* c0, int* c1)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    if (c0[e] == 0x7fffffff) c0[e] = -1;
    if (c1[e] == c0[e]) c1[e] = -1;
}

// order-preserving map double -> u64 so that atomicMin/atomicMax work
__device__ __host__ __forceinline__ u64 enc_d(double v) { u64 b; memcpy(&b, &v, 8); return (b >> 63) ? ~b : (b | 0x8000000000000000ULL); }
__device__ __host__ __forceinline__ double dec_d(u64 b) { b = (b >> 63) ? (b & 0x7FFFFFFFFFFFFFFFULL) : ~b; double v; memcpy(&v, &b, 8); return v; }

__global__ void bbox_kernel(int n_bound, const int* __restrict__ n_dev, const double* __restrict__ ex, const double* __restrict__ ey,
                            const int* __restrict__ esrc, const double* __restrict__ rmax, Counters* c)
{
    const int n = *n_dev;
    double xmn = SZ_INF, xmx = -SZ_INF, ymn = SZ_INF, ymx = -SZ_INF, rm = 0;
    // grid-stride: a few thousand warps, one set of atomics each (one warp per 32 floes serialised on five addresses)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_bound && i < n; i += gridDim.x * blockDim.x) {
        const double x = ex[i], y = ey[i];
        if (x == x && y == y) { xmn = fmin(xmn, x); xmx = fmax(xmx, x); ymn = fmin(ymn, y); ymx = fmax(ymx, y); }
        const double r = rmax[esrc[i]]; if (r > rm) rm = r;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        xmn = fmin(xmn, __shfl_xor_sync(0xffffffffu, xmn, d)); xmx = fmax(xmx, __shfl_xor_sync(0xffffffffu, xmx, d));
        ymn = fmin(ymn, __shfl_xor_sync(0xffffffffu, ymn, d)); ymx = fmax(ymx, __shfl_xor_sync(0xffffffffu, ymx, d));
        rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (xmn <= xmx) { atomicMin(&c->bbox[0], enc_d(xmn)); atomicMax(&c->bbox[1], enc_d(xmx)); atomicMin(&c->bbox[2], enc_d(ymn)); atomicMax(&c->bbox[3], enc_d(ymx)); }
        atomicMax(&c->rmax_bits, enc_d(rm));
    }
}

// ------------------------------------------------------------------------------------------------ K1 broad phase
struct GridDesc { double x0, y0, cell; int nx, ny; };
__device__ __forceinline__ int cell_coord(double v, double v0, double cell, int n) { int c = (int)((v - v0) / cell); return c < 0 ? 0 : (c >= n ? n - 1 : c); }

__global__ void cell_count_kernel(int n, GridDesc g, const double* __restrict__ ex, const double* __restrict__ ey, const uint8_t* __restrict__ ealive,
                                  int* __restrict__ cid, int* __restrict__ cell_cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = ex[i], y = ey[i];
    int c = -1;
    if (ealive[i] && x == x && y == y) { c = cell_coord(y, g.y0, g.cell, g.ny) * g.nx + cell_coord(x, g.x0, g.cell, g.nx); atomicAdd(&cell_cnt[c], 1); }
    cid[i] = c;
}
__global__ void cell_fill_kernel(int n, const int* __restrict__ cid, const int* __restrict__ cell_start, int* __restrict__ cell_pos,
                                 const double* __restrict__ ex, const double* __restrict__ ey, const int* __restrict__ esrc, const double* __restrict__ rmax,
                                 int* __restrict__ s_idx, double* __restrict__ s_x, double* __restrict__ s_y, double* __restrict__ s_r)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cid[i];
    if (c < 0) return;
    const int t = cell_start[c] + atomicAdd(&cell_pos[c], 1);
    s_idx[t] = i; s_x[t] = ex[i]; s_y[t] = ey[i]; s_r[t] = rmax[esrc[i]];
}

struct BroadArgs {
    int n, n0, Nb, collision; GridDesc g; double minL2; const u64* rmax_bits;     // largest rmax of the list, as bbox_kernel left it on the device
    int np_cap; int* overflow; int pair_boundary;      // pair_boundary: SzParams.pair_with_boundary_floes
    const double* ex; const double* ey; const int* esrc; const int* efn; const uint8_t* ealive; const double* rmax;
    const int* egid; const uint8_t* eowned; const double* erootx; const double* erooty;
    const int* cell_start; const int* s_idx; const double* s_x; const double* s_y; const double* s_r;
    int* pcnt; const int* pair_off; int* pi; int* pj;
    int* stage; int stage_cap;     // count pass: the first stage_cap partners of every floe, so that the fill pass need not search again
};
// The predicate of floe_interactions_all.m:103 for partner j of floe i (j > i checked by the caller):
//   alive(j) && sqrt((xi-xj)^2+(yi-yj)^2) < rmax_i+rmax_j && (~ismember(|FloeNums(j)|,mems) || 2(rmax_i+rmax_j) > min(2Lx,2Ly))
// `mems` (:93-99,113) in closed form: a ghost j is a member iff its original o_j is a forward partner of
// a = i (i an original) or of a = i's original (i a ghost), i.e. o_j > a and the pair (a, o_j) itself passes
// the distance test with a eligible for the pair loop.
__device__ __forceinline__ bool ghost_is_member(const BroadArgs& b, int i, int j)
{
    // a and o_j are identified by their FloeNums; their centroids travel with every image (erootx/erooty), their
    // rmax and alive flag are the image's own, so the test needs no other entry of the list (multi-GPU halos)
    const int A = abs(b.efn[i]), O = abs(b.efn[j]);
    if (!(O > A) || A <= b.Nb) return false;
    if (!b.ealive[i] || !b.ealive[j]) return false;
    const double xa = b.erootx[i], ya = b.erooty[i];
    if (xa != xa) return false;
    const double dx = xa - b.erootx[j], dy = ya - b.erooty[j];
    return sqrt(dx * dx + dy * dy) < (b.rmax[b.esrc[i]] + b.rmax[b.esrc[j]]);
}
// one warp per floe i: lanes stride over the three cell rows (each row's three cells are contiguous in
// the bucketed arrays), ballot + popc compacts accepted partners, then an in-warp rank sort restores
// ascending j (the order of Floe(i).potentialInteractions)
template <bool FILL, bool PB>      // PB: the opt-in pairing with topography floes (a separate instantiation keeps the default's registers: 40, not 54)
__global__ void __launch_bounds__(256) broad_kernel(const BroadArgs b)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int i = warp;
    if (i >= b.n) return;
    const double xi = b.ex[i], yi = b.ey[i];
    const bool active = b.egid[i] > b.Nb && b.ealive[i] && (xi == xi) && b.collision && (yi == yi);
    const bool own_i = b.eowned[i] != 0;
    int count = 0;
    const int off = FILL ? b.pair_off[i] : 0;
    if (FILL && b.pair_off[i + 1] > b.np_cap) { if (lane == 0) *b.overflow = 1; return; }     // more pairs than the speculated capacity: flagged, the step is repeated
    bool staged = false;
    if (FILL && b.stage) {
        count = b.pair_off[i + 1] - off;
        staged = count <= b.stage_cap;
        if (staged && lane < count) b.pj[off + lane] = b.stage[(size_t)i * b.stage_cap + lane];      // stage_cap <= 32
        if (!staged) count = 0;
    }
    if (active && !staged) {
        const double ri = b.rmax[b.esrc[i]];
#if defined(SZ_BROAD_SQUARE)
        const int cxi = cell_coord(xi, b.g.x0, b.g.cell, b.g.nx), cyi = cell_coord(yi, b.g.y0, b.g.cell, b.g.ny);
        // a partner has |dx|, |dy| < ri + rmax_j <= ri + max(rmax): that many cells each way (floor differences <= ceil)
        const int R = (int)((ri + dec_d(*b.rmax_bits)) / b.g.cell) + 1;
        const int cx0 = cxi - R > 0 ? cxi - R : 0, cx1 = cxi + R < b.g.nx ? cxi + R : b.g.nx - 1;
        const int cy0 = cyi - R > 0 ? cyi - R : 0, cy1 = cyi + R < b.g.ny ? cyi + R : b.g.ny - 1;
#else
        // a partner has |dx|, |dy| <= sqrt(dx^2 + dy^2) < ri + rmax_j <= ri + max(rmax) =: reach, so its centroid lies in
        // [xi - reach, xi + reach] x [yi - reach, yi + reach]; cell_coord (clamps included) is monotone, so its cell lies
        // between the cells of the corners -- on average 3.5 x 3.5 cells of the benchmark field instead of the
        // (2 floor(reach / cell) + 3)^2 = 5 x 5 that counting whole cells each way from the floe's own cell gives.  The
        // margin (1e-9 of the reach, 1e-12 of the coordinates, 1 nm) covers the last-bit rounding of the subtractions and the square root.
        double reach = ri + dec_d(*b.rmax_bits);
        reach = reach * (1.0 + 1e-9) + 1e-9 + (fabs(xi) + fabs(yi)) * 1e-12;
        const int cx0 = cell_coord(xi - reach, b.g.x0, b.g.cell, b.g.nx), cx1 = cell_coord(xi + reach, b.g.x0, b.g.cell, b.g.nx);
        const int cy0 = cell_coord(yi - reach, b.g.y0, b.g.cell, b.g.ny), cy1 = cell_coord(yi + reach, b.g.y0, b.g.cell, b.g.ny);
#endif
        for (int cy = cy0; cy <= cy1; ++cy) {
            const int t0 = b.cell_start[cy * b.g.nx + cx0], t1 = b.cell_start[cy * b.g.nx + cx1 + 1];
            for (int tb = t0; tb < t1; tb += 32) {
                const int t = tb + lane;
                bool ok = false; int j = -1;
                if (t < t1) {
                    j = b.s_idx[t];
                    // j > i (:103); opt-in: also the topography floes j <= Nb below it (never i themselves: `active` needs egid > Nb)
                    if ((j > i || (PB && b.egid[j] <= b.Nb)) && (own_i || b.eowned[j])) {     // a pair is resolved where either floe is owned
                        const double dx = xi - b.s_x[t], dy = yi - b.s_y[t], rs = ri + b.s_r[t];
                        if (sqrt(dx * dx + dy * dy) < rs) {
                            ok = true;
                            if (b.efn[j] < 0 && ghost_is_member(b, i, j) && !(2 * rs > b.minL2)) ok = false;
                        }
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (FILL && ok) b.pj[off + count + __popc(m & ((1u << lane) - 1))] = j;
                if (!FILL && ok && b.stage) { const int k = count + __popc(m & ((1u << lane) - 1)); if (k < b.stage_cap) b.stage[(size_t)i * b.stage_cap + k] = j; }
                count += __popc(m);
            }
        }
    }
    if (!FILL) { if (lane == 0) b.pcnt[i] = count; return; }
    if (count == 0) return;
    __syncwarp();
    if (count <= 32) {
        const int v = (lane < count) ? b.pj[off + lane] : 0x7fffffff;
        int rank = 0;
        for (int m = 0; m < count; ++m) rank += (__shfl_sync(0xffffffffu, v, m) < v);
        __syncwarp();
        if (lane < count) { b.pj[off + rank] = v; b.pi[off + lane] = i; }
    } else if (count <= 256) {
        // up to eight partners per lane: every lane ranks its own against the whole segment, then all write in place
        int v[8], r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = lane + 32 * q;
            v[q] = k < count ? b.pj[off + k] : 0x7fffffff; r[q] = 0;
        }
        for (int m = 0; m < count; ++m) {
            const int w = b.pj[off + m];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] += (w < v[q]);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) if (lane + 32 * q < count) { b.pj[off + r[q]] = v[q]; b.pi[off + lane + 32 * q] = i; }
    } else {
        if (lane == 0) {
            for (int a = 1; a < count; ++a) { int v = b.pj[off + a], k = a - 1; while (k >= 0 && b.pj[off + k] > v) { b.pj[off + k + 1] = b.pj[off + k]; --k; } b.pj[off + k + 1] = v; }
        }
        for (int k = lane; k < count; k += 32) b.pi[off + k] = i;
    }
}

// ------------------------------------------------------------------------------------------------ narrow-phase work list
// Everything the pair classifier needs about one entry of the extended list in ONE 64-byte record (two sectors): the classifier
// is a gather over randomly numbered partners, and it used to touch eight separate arrays per entry.
struct __align__(16) EntryRec { i64 bb[4]; double x, y; int vo; short nv; unsigned char no, valid; int pad; };
// Per entry of the extended list: the bounding box of its outline in Clipper's coordinates (polyclip.m:66) and
// whether the outline survives Clipper's AddPath (>= 3 vertices, not all collinear: clipper.cpp:1058,1119-1123).
__global__ void ext_prep_kernel(int n, const double* __restrict__ ex, const double* __restrict__ ey, const int* __restrict__ esrc, const int* __restrict__ voff,
                                const double* __restrict__ vx, const double* __restrict__ vy, i64* __restrict__ ebb, uint8_t* __restrict__ evalid, int* __restrict__ env,
                                uint8_t* __restrict__ econvex, uint8_t* __restrict__ erot, uint8_t* __restrict__ eno, EntryRec* __restrict__ erec, Counters* __restrict__ cnt)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const int s = esrc[e], o = voff[s], nv = voff[s + 1] - o;
    const double X = ex[e], Y = ey[e];
    i64 xmn = 0x7fffffffffffffffLL, xmx = -0x7fffffffffffffffLL - 1, ymn = xmn, ymx = xmx;
    szclip::P64 p0{0, 0}, p1{0, 0}; int have = 0; bool valid = false;
    for (int t = 0; t < nv; ++t) {
        szclip::P64 p; p.x = szpf::matlab_int64((vx[o + t] + X) * SZ_SCALE); p.y = szpf::matlab_int64((vy[o + t] + Y) * SZ_SCALE);
        xmn = p.x < xmn ? p.x : xmn; xmx = p.x > xmx ? p.x : xmx; ymn = p.y < ymn ? p.y : ymn; ymx = p.y > ymx ? p.y : ymx;
        if (have == 0) { p0 = p; have = 1; }
        else if (have == 1) { if (p != p0) { p1 = p; have = 2; } }
        else if (!valid && !szclip::slopes_eq3(p0, p1, p)) valid = true;
    }
    // coordinates must also stay far from Clipper's hiRange for the shortcut to be trusted
    if (xmx > (1LL << 61) || ymx > (1LL << 61) || xmn < -(1LL << 61) || ymn < -(1LL << 61)) valid = false;
    ebb[(size_t)e * 4] = xmn; ebb[(size_t)e * 4 + 1] = xmx; ebb[(size_t)e * 4 + 2] = ymn; ebb[(size_t)e * 4 + 3] = ymx;
    evalid[e] = valid; env[e] = nv;
    // strict convexity of the outline in Clipper's coordinates (closing vertex dropped): enables the margin-certified
    // sign test of szpf::convex_sign_test
    struct Get { const double* x; const double* y; double X, Y; int o;
                 __device__ szclip::P64 operator()(int i) const { szclip::P64 p; p.x = szpf::matlab_int64((x[o + i] + X) * SZ_SCALE); p.y = szpf::matlab_int64((y[o + i] + Y) * SZ_SCALE); return p; } };
    int no = nv;
    while (no > 1 && vx[o + no - 1] == vx[o] && vy[o + no - 1] == vy[o]) --no;
    const bool cvx = valid && no <= 255 && szpf::ring_is_strictly_convex(Get{vx, vy, X, Y, o}, no);
    econvex[e] = cvx;
    if (valid && !cvx) cnt->any_concave = 1;        // (every writer stores the same value)
    eno[e] = cvx ? (uint8_t)no : 0;
    erot[e] = cvx ? (uint8_t)szpf::ring_bottom_vertex(Get{vx, vy, X, Y, o}, no) : 0;     // start of the sweep input (PairHints)
    EntryRec r; r.bb[0] = xmn; r.bb[1] = xmx; r.bb[2] = ymn; r.bb[3] = ymx; r.x = X; r.y = Y; r.vo = o; r.nv = (short)(nv < 32767 ? nv : 32767); r.no = cvx ? (unsigned char)no : 0;
    r.valid = valid ? 1 : 0; r.pad = 0;
    erec[e] = r;
}
// Separating-axis test for two strictly convex outlines (sat_side_group below): true when an edge line of one outline has every
// vertex of the other at least 1 mm (4e6 Clipper units) on its outer side.  Then the outlines are disjoint with a margin a
// million times Clipper's rounding, the intersection is empty and the sweep cannot fail: zero force, overlap 0.
// Work list of the narrow phase.  classify: a pair whose outlines both survive AddPath and whose integer bounding boxes are
// strictly disjoint -- or, for two strictly convex outlines, that an edge line of one separates with 1 mm to spare -- has an
// empty Clipper intersection (and cannot fail), so floe_interactions returns zero force and overlap 0 (:43-51,71-74): it is
// answered here.  Every other pair gets a bucket key (n1, n2, class C: vertices inside the shared Y range = scanbeams of the
// convex sweep; class S: direction sector of the partner) so that the pairs a CTA sweeps together have similar event orders.
// The separating-axis test used to be a doubly nested loop over (edge of one outline) x (vertex of the other) with an early exit:
// 7.5 of 32 lanes active, 29 % of the FP64 pipe, 16 % of the whole step.  An edge line can only have EVERY vertex of the other
// outline beyond it if that outline's centroid (inside its convex hull) is beyond it too, so the vertex loop now runs only for
// the one to three edges that pass this one-product test: the same decisions for a quarter of the arithmetic.
// CLS_G lanes per pair (compile-time; 1 = a thread per pair): with CLS_G > 1 the lanes of a group take the edges in turn and
// vote -- measured slower at 8 (2.84 vs 1.81 ms at 1M floes: the per-pair gathers of boxes, flags and offsets are repeated by
// every lane of the group and dominate).
#ifndef CLS_G
#define CLS_G 1
#endif
__device__ __forceinline__ bool sat_side_group(const double* __restrict__ vx, const double* __restrict__ vy, int oa, int na, double XA, double YA,
                                               int ob, int nb, double XB, double YB, int gl)
{
    // (XB, YB) is the other outline's centroid: c_alpha is the outline about the centroid (initialize_floe_values.m:17)
    // orientation of A from its first turn (strictly convex: every turn has this sign)
    const double t = ((vx[oa + 1] - vx[oa]) * (vy[oa + 2] - vy[oa + 1]) - (vy[oa + 1] - vy[oa]) * (vx[oa + 2] - vx[oa + 1]));
    const double sg = t > 0 ? 1.0 : -1.0;       // inner side of an edge is where sg * cross > 0
    // a point inside B's convex hull: the mean of its vertices (the field's Xi, Yi need not be trusted for this)
    double cbx = 0, cby = 0;
    for (int q = 0; q < nb; ++q) { cbx += vx[ob + q]; cby += vy[ob + q]; }
    cbx = cbx / nb + XB; cby = cby / nb + YB;
    bool found = false;
    for (int e = gl; e < na; e += CLS_G) {
        const int e1 = (e + 1 == na) ? 0 : e + 1;
        const double ax = vx[oa + e] + XA, ay = vy[oa + e] + YA, dx = (vx[oa + e1] + XA) - ax, dy = (vy[oa + e1] + YA) - ay;
        const double lim = -1e-3 * sqrt(dx * dx + dy * dy);
        // necessary condition: a point of B's convex hull -- the mean of its vertices -- is beyond the line by the same margin
        if (!(sg * (dx * (cby - ay) - dy * (cbx - ax)) < lim)) continue;
        bool all_out = true;
        for (int q = 0; q < nb && all_out; ++q) {
            const double cr = sg * (dx * ((vy[ob + q] + YB) - ay) - dy * ((vx[ob + q] + XB) - ax));
            all_out = cr < lim;
        }
        found = found || all_out;
    }
    return found;
}
// Pre-pass of the classifier for fields with concave outlines (skipped at once on a field of strictly convex floes): a pair of
// AddPath-valid outlines that are not both strictly convex, with overlapping boxes, is certified apart edge by edge
// (sz_apart.cuh) -- the sweep would return nothing and the pair take the zero-force branch (floe_interactions.m:43-44,71-74).
// 40 % of the candidate pairs of a field of the reference's own floe shapes.  A kernel of its own: inlined into the classifier it
// raised that kernel from 73 to 122 registers for the convex field, which never uses it.
__global__ void __launch_bounds__(128) pair_apart_kernel(int np_cap, const int* __restrict__ np_dev, const int* __restrict__ pi, const int* __restrict__ pj, const EntryRec* __restrict__ erec,
                                                         const double* __restrict__ vx, const double* __restrict__ vy, uint8_t* __restrict__ apart, const Counters* __restrict__ c)
{
    if (!c->any_concave || c->overflow) return;                 // a convex field leaves after one load per thread (fixed, small grid)
    const int np = *np_dev < np_cap ? *np_dev : np_cap;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
    const EntryRec ri = erec[pi[p]], rj = erec[pj[p]];
    bool ap = false;
    const bool cvx = ri.no >= 3 && rj.no >= 3;                  // both strictly convex: the separating-axis rule of the classifier is complete for them
    if (ri.valid && rj.valid && !cvx && ri.nv < 32767 && rj.nv < 32767 && !(ri.bb[1] < rj.bb[0] || rj.bb[1] < ri.bb[0] || ri.bb[3] < rj.bb[2] || rj.bb[3] < ri.bb[2])) {
        const double s = 1.0 / SZ_SCALE;                        // the int64 boxes of ext_prep_kernel in metres (only a filter)
        ap = szapart::rings_apart(vx + ri.vo, vy + ri.vo, ri.nv, ri.x, ri.y, (double)ri.bb[0] * s, (double)ri.bb[1] * s, (double)ri.bb[2] * s, (double)ri.bb[3] * s,
                                  vx + rj.vo, vy + rj.vo, rj.nv, rj.x, rj.y, (double)rj.bb[0] * s, (double)rj.bb[1] * s, (double)rj.bb[2] * s, (double)rj.bb[3] * s);
    }
    apart[p] = ap;
    }
}
__global__ void __launch_bounds__(256) pair_classify_kernel(int np_cap, const int* __restrict__ np_dev, const int* __restrict__ pi, const int* __restrict__ pj, const EntryRec* __restrict__ erec,
                                     const double* __restrict__ vx, const double* __restrict__ vy, int want_polys,
                                     int* __restrict__ status, int* __restrict__ nrows, double* __restrict__ ovl, int* __restrict__ poly_npaths,
                                     int* __restrict__ bins, short* __restrict__ pkey, Counters* c, int cvx_key_mode, const uint8_t* __restrict__ apart)
{
    // per bucket: pairs of strictly convex outlines (class C) in the high half-word, the others (class S) in the low one
    __shared__ int sh[SZ_NBINS];
    for (int t = threadIdx.x; t < SZ_NBINS; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    const int np = c->overflow ? 0 : (*np_dev < np_cap ? *np_dev : np_cap);     // a flagged step is going to be repeated: its pair list may be incomplete
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) / CLS_G, gl = threadIdx.x & (CLS_G - 1);
    const unsigned gmask = ((1u << CLS_G) - 1u) << ((threadIdx.x & 31) & ~(CLS_G - 1));
    if (p < np) {                                   // group-uniform
        const int i = pi[p], j = pj[p];
        const EntryRec ri = erec[i], rj = erec[j];
        const i64* a = ri.bb; const i64* b = rj.bb;
        bool disjoint = ri.valid && rj.valid && (a[1] < b[0] || b[1] < a[0] || a[3] < b[2] || b[3] < a[2]);
        const int ni = ri.no, nj = rj.no;
        const bool cvx = ni >= 3 && nj >= 3;                    // both strictly convex (no is 0 otherwise)
        const int oi = ri.vo, oj = rj.vo;
        const double Xi = ri.x, Yi = ri.y, Xj = rj.x, Yj = rj.y;
        if (!disjoint && cvx) {
            bool f = sat_side_group(vx, vy, oi, ni, Xi, Yi, oj, nj, Xj, Yj, gl);
            f = f || sat_side_group(vx, vy, oj, nj, Xj, Yj, oi, ni, Xi, Yi, gl);
            disjoint = __any_sync(gmask, f);
        }
        if (!disjoint && !cvx && apart != nullptr && c->any_concave) disjoint = apart[p] != 0;      // pair_apart_kernel's certificate
        int cnt = 0;
        if (!disjoint && cvx && cvx_key_mode == 1) {
            // class C sweeps one scanbeam per CTA-synchronous iteration, and only over the Y range the two outlines share:
            // bucket by the number of vertices in that range (= iterations) instead of the direction
            const double lo = (double)(a[2] > b[2] ? a[2] : b[2]), hi = (double)(a[3] < b[3] ? a[3] : b[3]);
            for (int t = gl; t < ni; t += CLS_G) { const double y = (vy[oi + t] + Yi) * SZ_SCALE; cnt += (y >= lo && y <= hi); }
            for (int t = gl; t < nj; t += CLS_G) { const double y = (vy[oj + t] + Yj) * SZ_SCALE; cnt += (y >= lo && y <= hi); }
            cnt = __reduce_add_sync(gmask, cnt);
        }
        if (gl == 0) {
            int key = -1;
            if (disjoint) {
                status[p] = 0; nrows[p] = 0; ovl[p] = 0; if (want_polys) poly_npaths[p] = 0;
            } else {
                const int bi = ri.nv < SZ_BIN_N ? ri.nv : SZ_BIN_N - 1, bj = rj.nv < SZ_BIN_N ? rj.nv : SZ_BIN_N - 1;
                // pairs of one bucket have the same vertex counts and the partner in the same octant: similar event orders
                const double dx = Xj - Xi, dy = Yj - Yi;
                const double adx = fabs(dx), ady = fabs(dy), mn = adx < ady ? adx : ady, mx = adx < ady ? ady : adx;
                int oct = (dx > 0) | ((dy > 0) << 1) | ((adx > ady) << 2) | ((mn > 0.41421356237309503 * mx) << 3);
                if (cvx && cvx_key_mode == 1) oct = cnt < SZ_NSECT ? cnt : SZ_NSECT - 1;
                key = (bi * SZ_BIN_N + bj) * SZ_NSECT + oct;
                atomicAdd(&sh[key], cvx ? 65536 : 1);
            }
            pkey[p] = (short)(key < 0 ? key : (key | (cvx ? SZ_NBINS : 0)));
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < SZ_NBINS; t += blockDim.x) {
        const int v = sh[t];
        if (v >> 16) atomicAdd(&bins[t], v >> 16);
        if (v & 0xffff) atomicAdd(&bins[SZ_NBINS + t], v & 0xffff);
    }
}
// scatter: bins[] holds exclusive offsets (class C buckets, then class S buckets); every CTA reserves its share of every bucket,
// then places its pairs -- class C first, then class S through the same shared array
__global__ void __launch_bounds__(256) pair_scatter_kernel(int np_cap, const int* __restrict__ np_dev, const short* __restrict__ pkey, const int* __restrict__ bins, int* __restrict__ bin_fill,
                                                           int* __restrict__ listC, int* __restrict__ listS, const Counters* __restrict__ c)
{
    __shared__ int sh[SZ_NBINS];
    __shared__ int base[SZ_NBINS];
    for (int t = threadIdx.x; t < SZ_NBINS; t += blockDim.x) sh[t] = 0;
    __syncthreads();
    const int np = c->overflow ? 0 : (*np_dev < np_cap ? *np_dev : np_cap);
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    int key = -1, slot = 0; bool cvx = false;
    if (p < np) {
        key = pkey[p];
        if (key >= 0) { cvx = (key & SZ_NBINS) != 0; key &= SZ_NBINS - 1; const int old = atomicAdd(&sh[key], cvx ? 65536 : 1); slot = cvx ? (old >> 16) : (old & 0xffff); }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < SZ_NBINS; t += blockDim.x) base[t] = (sh[t] >> 16) ? atomicAdd(&bin_fill[t], sh[t] >> 16) : 0;
    __syncthreads();
    if (key >= 0 && cvx) listC[bins[key] + base[key] + slot] = p;
    __syncthreads();
    for (int t = threadIdx.x; t < SZ_NBINS; t += blockDim.x) base[t] = (sh[t] & 0xffff) ? atomicAdd(&bin_fill[SZ_NBINS + t], sh[t] & 0xffff) : 0;
    __syncthreads();
    if (key >= 0 && !cvx) listS[bins[SZ_NBINS + key] + base[key] + slot] = p;
}
// Exclusive offsets of the buckets in launch order: longest outlines first (the CTAs with the longest sweeps start
// first, which shortens the tail of the launch).  One thread per (n1, n2) combination; its rank in the order
// "n1 + n2 descending, n1 descending" is closed-form, a 256-wide shared-memory scan does the rest.
__global__ void __launch_bounds__(SZ_BIN_N * SZ_BIN_N) bins_scan_kernel(Counters* c, int np_cap, int* __restrict__ bins, int* __restrict__ bin_fill)
{
    const int np = c->n_pairs < np_cap ? c->n_pairs : np_cap;
    bins += blockIdx.x * SZ_NBINS; bin_fill += blockIdx.x * SZ_NBINS;      // block 0: class C buckets, block 1: class S buckets
    const int B = SZ_BIN_N, t = threadIdx.x, ni = t / B, nj = t % B, sum = ni + nj;
    __shared__ int tot[SZ_BIN_N * SZ_BIN_N], scan[SZ_BIN_N * SZ_BIN_N];
    int before = (sum < B - 1 ? sum : B - 1) - ni;                       // same sum, larger n1
    for (int u = sum + 1; u <= 2 * B - 2; ++u) before += (u < B) ? u + 1 : 2 * B - 1 - u;
    int mine = 0;
    for (int oct = 0; oct < SZ_NSECT; ++oct) mine += bins[t * SZ_NSECT + oct];
    tot[before] = mine;
    __syncthreads();
    scan[t] = tot[t];
    __syncthreads();
    for (int d = 1; d < B * B; d <<= 1) { const int v = (t >= d) ? scan[t - d] : 0; __syncthreads(); scan[t] += v; __syncthreads(); }
    int base = scan[before] - mine;                                       // exclusive
    for (int oct = 0; oct < SZ_NSECT; ++oct) { const int k = t * SZ_NSECT + oct; const int v = bins[k]; bins[k] = base; base += v; bin_fill[k] = 0; }
    if (t == 0) {
        if (blockIdx.x == 0) { c->listC = scan[B * B - 1]; atomicAdd(&c->n_bbox_reject, np - scan[B * B - 1]); }
        else { c->listS = scan[B * B - 1]; atomicAdd(&c->n_bbox_reject, -scan[B * B - 1]); }
    }
}

// ------------------------------------------------------------------------------------------------ K4 assembly
__global__ void tcount_kernel(int np_cap, const int* __restrict__ np_dev, const int* __restrict__ pi, const int* __restrict__ pj, const int* __restrict__ nrows, int* __restrict__ tcnt)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int np = *np_dev < np_cap ? *np_dev : np_cap;
    if (p < np && nrows[p] > 0 && pj[p] > pi[p]) atomicAdd(&tcnt[pj[p]], 1);      // mirrored only to a partner behind i (:196)
}
__global__ void tfill_kernel(int np_cap, const int* __restrict__ np_dev, const int* __restrict__ pi, const int* __restrict__ pj, const int* __restrict__ nrows, const int* __restrict__ toff, int* __restrict__ tpos, int* __restrict__ tlist)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int np = *np_dev < np_cap ? *np_dev : np_cap;
    if (p < np && nrows[p] > 0 && pj[p] > pi[p]) { const int j = pj[p]; tlist[toff[j] + atomicAdd(&tpos[j], 1)] = p; }
}
// rows of floe m = own pairs + wall + mirrored (floe_interactions_all.m:136,167,196)
__global__ void rowcount_kernel(int n, const uint8_t* __restrict__ eowned, const int* __restrict__ pair_off, const int* __restrict__ nrows, const int* __restrict__ wnrows,
                                const int* __restrict__ toff, int* __restrict__ tlist, int* __restrict__ rcnt)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    if (!eowned[m]) { rcnt[m] = 0; return; }
    int c = 0;
    for (int p = pair_off[m]; p < pair_off[m + 1]; ++p) c += nrows[p];
    if (wnrows) c += wnrows[m];
    const int t0 = toff[m], t1 = toff[m + 1];
    for (int a = t0 + 1; a < t1; ++a) { int v = tlist[a], k = a - 1; while (k >= t0 && tlist[k] > v) { tlist[k + 1] = tlist[k]; --k; } tlist[k + 1] = v; }   // ascending i
    for (int t = t0; t < t1; ++t) c += nrows[tlist[t]];
    rcnt[m] = c;
}

// polygon_operations/inpolygon.m:64-224 for one point against a double polygon (used for the
// centroid-in-domain test, floe_interactions_all.m:152)
__device__ bool in_polygon_d(double x, double y, const double* xv, const double* yv, int nv)
{
    if (nv < 1) return false;
    double xmin = xv[0], xmax = xv[0], ymin = yv[0], ymax = yv[0];
    for (int k = 1; k < nv; ++k) { xmin = fmin(xmin, xv[k]); xmax = fmax(xmax, xv[k]); ymin = fmin(ymin, yv[k]); ymax = fmax(ymax, yv[k]); }
    if (!(x >= xmin && x <= xmax && y >= ymin && y <= ymax)) return false;
    const bool closed = !(nv >= 3 && (xv[0] != xv[nv - 1] || yv[0] != yv[nv - 1]));
    const int ne = closed ? nv - 1 : nv;      // edges
    if (ne < 1) return false;
    double sumdq = 0; bool on = false;
    double ax = xv[0], ay = yv[0];
    double vx0 = ax - x, vy0 = ay - y;
    bool px0 = vx0 > 0, py0 = vy0 > 0;
    double q0 = (double)((!px0 && py0) + 2 * (!px0 && !py0) + 3 * (px0 && !py0));
    for (int m = 0; m < ne; ++m) {
        const int j = (m + 1 < nv) ? m + 1 : 0;
        const double bx = xv[j], by = yv[j];
        const double avx = fabs(0.5 * (ax + bx)), avy = fabs(0.5 * (ay + by));
        double sf = avx > avy ? avx : avy; const double pr = avx * avy; if (pr > sf) sf = pr;
        const double seps = sf * SZ_EPS * 3;
        const double vx1 = bx - x, vy1 = by - y;
        const bool px1 = vx1 > 0, py1 = vy1 > 0;
        const double q1 = (double)((!px1 && py1) + 2 * (!px1 && !py1) + 3 * (px1 && !py1));
        const double cross = vx0 * vy1 - vx1 * vy0;
        double sg = (double)((cross > 0) - (cross < 0));
        if (fabs(cross) < seps) sg = 0;
        const double dot = vx0 * vx1 + vy0 * vy1;
        double dq = q1 - q0;
        if (fabs(dq) == 3) dq = -dq / 3; else if (fabs(dq) == 2) dq = 2 * sg;
        sumdq += dq;
        if (sg == 0 && dot <= 0) on = true;
        ax = bx; ay = by; vx0 = vx1; vy0 = vy1; q0 = q1;
    }
    return (sumdq != 0) || on;
}

struct AssembleArgs {
    int n, nout, Nb, periodic, wall; double Lx, Ly;
    const int* egid; const int* efn; const uint8_t* eowned;
    const double* ex; const double* ey; const int* esrc; const uint8_t* ealive; const double* area; const double* h;
    const int* pair_off; const int* pi; const int* pj; const int* nrows; const int* row_start; const double* ovl; const int* pstatus;
    const int* wnrows; const int* wrow_start; const int* wstatus;
    const int* toff; const int* tlist; const int* row_off; const double* pool;
    const double* boxx; const double* boxy; int boxn;
    double* rows; long long rows_cap; double* osum; double* e_ov; uint8_t* has_rows; int* kill_i; int* transfer_i;
    double* o_ov; double* o_stress; double* o_xi; double* o_yi; uint8_t* o_alive;
    Counters* cnt;
};
// one thread per floe of the extended list writes its rows in the reference's canonical order and, walking
// them top to bottom like MATLAB's column sum, accumulates torque, force sums, overlap area and stress
__global__ void __launch_bounds__(128) assemble_kernel(const AssembleArgs a)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= a.n) return;
    // an earlier phase outgrew its (speculated) capacity, or there are more rows than the speculated row capacity: the step is
    // flagged and will be repeated; nothing may be read through its incomplete lists, and the kill fix-up must find no event
    if (a.cnt->overflow || a.row_off[m + 1] > a.rows_cap) { a.cnt->overflow = 1; a.kill_i[m] = 0; a.transfer_i[m] = 0; a.has_rows[m] = 0; return; }
    const bool owned = a.eowned[m] != 0, orig = a.efn[m] > 0, pairing = a.egid[m] > a.Nb;   // pairing: i >= 1+Nb (:125)
    if (!owned) {
        a.osum[(size_t)m * 3] = a.osum[(size_t)m * 3 + 1] = a.osum[(size_t)m * 3 + 2] = 0; a.e_ov[m] = 0; a.has_rows[m] = 0; a.kill_i[m] = 0; a.transfer_i[m] = 0;
        if (m < a.nout) { a.o_ov[m] = 0; a.o_xi[m] = a.ex[m]; a.o_yi[m] = a.ey[m]; a.o_alive[m] = a.ealive[m]; double* S = a.o_stress + (size_t)m * 4; S[0] = S[1] = S[2] = S[3] = 0; }
        return;
    }
    const double xm = a.ex[m], ym = a.ey[m];
    // periodic wrap of the centroid (:267-277); the stress uses the wrapped centroid
    double xw = xm, yw = ym;
    if (a.periodic && pairing && orig) {
        if (fabs(xw) > a.Lx) xw = xw - 2 * a.Lx * sgn_d(xw);
        if (fabs(yw) > a.Ly) yw = yw - 2 * a.Ly * sgn_d(yw);
    }
    double* R = a.rows + (size_t)a.row_off[m] * 7;
    const int nr_total = a.row_off[m + 1] - a.row_off[m];
    double sfx = 0, sfy = 0, st = 0, ova = 0;
    double s11 = 0, s12 = 0, s21 = 0, s22 = 0, t11 = 0, t12 = 0, t21 = 0, t22 = 0;
    int nfin = 0, ninf = 0, kill = 0, transfer = 0;
    auto put = [&](double partner, double fx, double fy, double px, double py, double ov) {
        const double tau = (px - xm) * fy - (py - ym) * fx;                 // cross([P-r 0],[F 0]) (:255-259)
        R[0] = partner; R[1] = fx; R[2] = fy; R[3] = px; R[4] = py; R[5] = tau; R[6] = ov; R += 7;
        sfx += fx; sfy += fy; st += tau;
        s11 += (px - xw) * fx; s12 += (py - yw) * fx; s21 += (px - xw) * fy; s22 += (py - yw) * fy;   // calc_trajectory.m:12
        t11 += fx * (px - xw); t12 += fy * (px - xw); t21 += fx * (py - yw); t22 += fy * (py - yw);
    };
    // own pairs, ascending partner (:125-147)
    for (int p = a.pair_off[m]; p < a.pair_off[m + 1]; ++p) {
        const int nr = a.nrows[p];
        if (nr > 0) {
            const double* src = a.pool + (size_t)a.row_start[p] * 5;
            double so = 0;
            for (int q = 0; q < nr; ++q) { put((double)a.egid[a.pj[p]], src[q * 5], src[q * 5 + 1], src[q * 5 + 2], src[q * 5 + 3], src[q * 5 + 4]); so += src[q * 5 + 4]; ++nfin; }
            ova = so + ova;
        } else if (a.pstatus[p] == 0) {
            const double ov = a.ovl[p];
            if ((ov == SZ_INF || ov == -SZ_INF) && pairing && a.egid[a.pj[p]] > a.Nb) {          // :138-145 (an opt-in topography partner raises no kill / transfer)
                if (orig && ov > 0) { kill = a.egid[m]; transfer = a.egid[a.pj[p]]; }
                else if (a.efn[a.pj[p]] > 0) kill = a.egid[a.pj[p]];
            }
        }
    }
    // wall rows (:150-172)
    uint8_t alive_out = a.ealive[m];
    if (a.wall && pairing && a.wstatus[m] == 0) {
        if (!in_polygon_d(xm, ym, a.boxx, a.boxy, a.boxn)) alive_out = 0;  // :152-155
        const int nr = a.wnrows[m];
        if (nr > 0) {
            const double* src = a.pool + (size_t)a.wrow_start[m] * 5;
            double so = 0;
            for (int q = 0; q < nr; ++q) {
                double fx = src[q * 5], fy = src[q * 5 + 1];
                if (fabs(src[q * 5 + 3]) == a.Ly) fx = 0;                    // :160-162
                if (fabs(src[q * 5 + 2]) == a.Lx) fy = 0;                    // :163-165
                put(SZ_INF, fx, fy, src[q * 5 + 2], src[q * 5 + 3], src[q * 5 + 4]); so += src[q * 5 + 4]; ++ninf;
            }
            ova = so + ova;
        }
    }
    // mirrored rows from lower-numbered floes, ascending i (:187-214)
    for (int t = a.toff[m]; t < a.toff[m + 1]; ++t) {
        const int p = a.tlist[t]; const int nr = a.nrows[p];
        const double* src = a.pool + (size_t)a.row_start[p] * 5;
        for (int q = 0; q < nr; ++q) { put((double)a.egid[a.pi[p]], -src[q * 5], -src[q * 5 + 1], src[q * 5 + 2], src[q * 5 + 3], src[q * 5 + 4]); ova = ova + src[q * 5 + 4]; ++nfin; }
    }
    a.osum[(size_t)m * 3] = sfx; a.osum[(size_t)m * 3 + 1] = sfy; a.osum[(size_t)m * 3 + 2] = st; a.e_ov[m] = ova;
    a.has_rows[m] = nr_total > 0;
    a.kill_i[m] = kill; a.transfer_i[m] = transfer;
    if (kill) atomicAdd(&a.cnt->n_kill_events, 1);      // rare (merges)
    if (m < a.nout) {
        a.o_ov[m] = ova; a.o_xi[m] = xw; a.o_yi[m] = yw; a.o_alive[m] = alive_out;
        double* S = a.o_stress + (size_t)m * 4;
        if (pairing && orig && alive_out && nr_total > 0) {
            const double k = 1 / (2 * a.area[a.esrc[m]] * a.h[a.esrc[m]]);
            S[0] = k * (s11 + t11); S[1] = k * (s12 + t12); S[2] = k * (s21 + t21); S[3] = k * (s22 + t22);
        } else { S[0] = S[1] = S[2] = S[3] = 0; }
    }
    // calc_collisionNum runs over the original floes.  One atomic per group of lanes that arrive here together instead of
    // one per floe (a million atomics on one address otherwise).
    {
        const int fin = (m < a.nout && orig) ? nfin : 0, inf = (m < a.nout && orig) ? ninf : 0;
        const unsigned mk = __activemask();
        const int sf = __reduce_add_sync(mk, fin), si = __reduce_add_sync(mk, inf);
        if ((int)(threadIdx.x & 31) == __ffs(mk) - 1) {
            if (sf) atomicAdd(&a.cnt->n_fin_rows, (u64)sf);
            if (si) atomicAdd(&a.cnt->n_inf_rows, (u64)si);
        }
    }
}
// ghost sums folded into their parents in creation order (:242-245), then the floe's own column sums (:262-263)
__global__ void fold_kernel(int nout, int n, int Nb, const int* __restrict__ egid, const int* __restrict__ gx_of, const int* __restrict__ gy_of, const double* __restrict__ osum,
                            const uint8_t* __restrict__ has_rows, double* __restrict__ fx, double* __restrict__ fy, double* __restrict__ tq)
{
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nout) return;
    double f0 = 0, f1 = 0, t = 0;
    int gx = gx_of[m], gy = gy_of[m];
    if (gx >= n) gx = -1;            // an image beyond the (speculated) list length: the step is flagged and repeated
    if (gy >= n) gy = -1;
    if (gx >= 0) { f0 = f0 + (has_rows[gx] ? osum[(size_t)gx * 3] : 0.0); f1 = f1 + (has_rows[gx] ? osum[(size_t)gx * 3 + 1] : 0.0); t = t + (has_rows[gx] ? osum[(size_t)gx * 3 + 2] : 0.0); }
    if (gy >= 0) { f0 = f0 + (has_rows[gy] ? osum[(size_t)gy * 3] : 0.0); f1 = f1 + (has_rows[gy] ? osum[(size_t)gy * 3 + 1] : 0.0); t = t + (has_rows[gy] ? osum[(size_t)gy * 3 + 2] : 0.0); }
    if (egid[m] > Nb && has_rows[m]) { f0 = osum[(size_t)m * 3] + f0; f1 = osum[(size_t)m * 3 + 1] + f1; t = osum[(size_t)m * 3 + 2] + t; }
    fx[m] = f0; fy[m] = f1; tq[m] = t;
}
// :175-179  for i=1:length(kill): if kill(i) ~= i && kill(i) > 0, transfer(kill(i)) = i   (serial: the largest i wins)
__global__ void kill_mark_kernel(int n, const int* __restrict__ kill_i, int* __restrict__ tmax)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int k = kill_i[i];
    if (k > 0 && k != i + 1) atomicMax(&tmax[k - 1], i + 1);
}
__global__ void kill_final_kernel(int n0, const int* __restrict__ kill_i, const int* __restrict__ transfer_i, const int* __restrict__ tmax, int* kill, int* transfer)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0) return;
    kill[i] = kill_i[i]; transfer[i] = tmax[i] ? tmax[i] : transfer_i[i];
}
__global__ void pair_stats_kernel(int np_cap, const int* __restrict__ np_dev, const int* __restrict__ status, const int* __restrict__ nrows, const int* __restrict__ pi, const uint8_t* __restrict__ eowned,
                                  int count_force, Counters* c)
{
    int f = 0, e = 0, k = 0, u = 0;
    const int np = c->overflow ? 0 : ((np_dev && *np_dev < np_cap) ? *np_dev : np_cap);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < np; p += gridDim.x * blockDim.x) {
        if (!eowned[pi ? pi[p] : p]) continue;
        const int s = status[p];
        ++u; f += (s == 0 && nrows[p] > 0); e += (s == szpf::PS_CLIPPER_FAIL || s == szpf::PS_BAD_POLY); k += (s == szpf::PS_CAPACITY);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        f += __shfl_xor_sync(0xffffffffu, f, d); e += __shfl_xor_sync(0xffffffffu, e, d);
        k += __shfl_xor_sync(0xffffffffu, k, d); u += __shfl_xor_sync(0xffffffffu, u, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (u && count_force) atomicAdd(&c->n_pairs_owned, u);
        if (f && count_force) atomicAdd(&c->n_pairs_force, f);
        if (e) atomicAdd(&c->n_fail, e);
        if (k) atomicAdd(&c->n_cap_fail, k);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static inline int nblk(i64 n, int tpb) { return (int)((n + tpb - 1) / tpb); }

extern "C" void sz_default_params(SzParams* p)
{
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->nu = 0.3; p->mu = 0.2; p->merge_frac = 0.55; p->wall_frac = 0.75; p->amin_per_vertex = 100.0 / 1.75;
    p->vertex_match_tol = 1; p->on_edge_tol = 1e-8; p->dl_min = 0.1; p->close_gap = 1; p->big_floe_r = 1e5; p->domain_area_frac = 0.95;
    p->dt = 10; p->collision = 1; p->periodic = 0; p->Nb = 0; p->want_clip_polys = 0; p->pair_with_boundary_floes = 0;
}

extern "C" int sz_create(SzContext** out, int device)
{
    if (!out) { sz_set_error("sz_create: out is NULL"); return SZ_ERR_ARG; }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { sz_set_error("sz_create: no CUDA device is usable (there is no CPU fallback)"); return SZ_ERR_CUDA; }
    if (device < 0 || device >= ndev) { sz_set_error("sz_create: device %d out of range (%d devices)", device, ndev); return SZ_ERR_ARG; }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) { sz_set_error("sz_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return SZ_ERR_CUDA; }
    CK(cudaSetDevice(device));
    SzContext* c = new SzContext;
    c->device = device;
    if (getenv("SZ_NO_SPECULATE")) c->opt_speculate = 0;      // debugging aid: measure every size as it is needed
    CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)); c->stream = c->own_stream;
    CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
    for (auto& e : c->evp) CK(cudaEventCreate(&e));
    for (auto& e : c->evk) CK(cudaEventCreate(&e));
    CK(cudaMalloc(&c->d_cnt, sizeof(Counters)));
    CK(cudaMallocHost(&c->h_cnt, sizeof(Counters))); CK(cudaMallocHost(&c->h_init, sizeof(Counters)));
    memset(&c->summary, 0, sizeof(c->summary));
    *out = c;
    return SZ_OK;
}

extern "C" void sz_destroy(SzContext* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    DBuf<double>* db[] = {&c->x, &c->y, &c->rmax, &c->h, &c->area, &c->u, &c->v, &c->ksi, &c->vx, &c->vy, &c->bx, &c->by, &c->boxx, &c->boxy, &c->ex, &c->ey,
                          &c->erootx, &c->erooty, &c->s_x, &c->s_y, &c->s_r, &c->povl, &c->wovl, &c->row_pool, &c->rows, &c->osum, &c->e_ov, &c->o_fx, &c->o_fy, &c->o_tq, &c->o_ov, &c->o_stress, &c->o_xi, &c->o_yi,
                          &c->oc_Xo, &c->oc_Yo, &c->oc_U, &c->oc_V, &c->oc_Wu, &c->oc_Wv, &c->pt_x, &c->pt_y, &c->t_strain,
                          &c->fr_xi, &c->fr_yi, &c->fr_area, &c->fr_vx, &c->fr_vy, &c->cr_ex, &c->cr_ey, &c->eu_lx, &c->eu_ly, &c->eu_in, &c->eu_area, &c->eu_out, &c->se_sx, &c->se_sy, &c->se_sr};
    for (auto* b : db) b->release();
    DBuf<int>* ib[] = {&c->egid, &c->voff, &c->esrc, &c->efn, &c->eparent, &c->gx_of, &c->gy_of, &c->flag, &c->pos, &c->scan_tmp, &c->cid, &c->cell_cnt, &c->cell_start, &c->s_idx,
                       &c->pcnt, &c->pair_off, &c->pi, &c->pj, &c->pstatus, &c->pnrows, &c->prow_start, &c->bins, &c->bin_fill, &c->stage, &c->listC, &c->listS, &c->listT, &c->wlistT, &c->env, &c->listM, &c->listL, &c->wstatus, &c->wnrows, &c->wrow_start,
                       &c->wlistM, &c->wlistL, &c->poly_path_start, &c->poly_npaths, &c->path_vstart, &c->path_len, &c->tcnt, &c->toff, &c->tlist, &c->rcnt, &c->row_off,
                       &c->kill_i, &c->transfer_i, &c->tmax, &c->o_kill, &c->o_transfer, &c->c_method, &c->c_status, &c->c_path_start, &c->c_npaths, &c->c_path_vstart,
                       &c->c_path_len, &c->c_listM, &c->c_listL, &c->fr_idx, &c->fr_vstart, &c->fr_vcount, &c->fr_status, &c->cr_idx, &c->cr_nv, &c->cr_off, &c->cr_esrc, &c->eu_lsrc, &c->eu_icnt, &c->eu_ioff, &c->eu_cell, &c->eu_q, &c->eu_status, &c->eu_iota, &c->eu_sorted, &c->eu_keys, &c->eu_ccnt, &c->eu_coff, &c->eu_listL, &c->ho_st, &c->se_bin, &c->se_cnt, &c->se_off, &c->se_q, &c->se_cid, &c->se_ccnt, &c->se_cstart, &c->se_sidx, &c->se_tmp, &c->se_out};
    for (auto* b : ib) b->release();
    DBuf<uint8_t>* ub[] = {&c->evalid, &c->econvex, &c->erot, &c->eno, &c->eowned, &c->alive, &c->ealive, &c->has_rows, &c->o_alive, &c->scratchM, &c->scratchL, &c->pt_a, &c->t_forced, &c->fr_changed, &c->cr_da, &c->cr_ealive, &c->eu_tmp};
    for (auto* b : ub) b->release();
    DBuf<i64>* lb[] = {&c->erec, &c->ebb, &c->pvx, &c->pvy, &c->c_soff, &c->c_coff, &c->c_sx, &c->c_sy, &c->c_cx, &c->c_cy, &c->c_pvx, &c->c_pvy, &c->ho_x, &c->ho_y};
    for (auto* b : lb) b->release();
    c->scan_state.release();
    c->pkey.release(); c->papart.release();
    { DBuf<double>* tb[] = {&c->t_mass, &c->t_inertia, &c->t_alpha, &c->t_dXi_p, &c->t_dYi_p, &c->t_dUi_p, &c->t_dVi_p, &c->t_dalpha_p, &c->t_dksi_p, &c->t_FxOA, &c->t_FyOA, &c->t_torqueOA,
                          &c->c0x, &c->c0y, &c->t_stressH, &c->t_stress};
      for (auto* b : tb) b->release(); c->t_scount.release(); c->t_flags.release(); }
    { DBuf<int>* sb[] = {&c->sl_ogid, &c->sl_flag, &c->sl_pos, &c->sl_opos, &c->sl_g, &c->sl_sendcnt, &c->sl_keys, &c->sl_slots, &c->sl_hnv, &c->sl_hvoff};
      for (auto* b : sb) b->release(); c->sl_cub.release(); c->sl_out.release(); c->sl_rows.release(); c->sl_rcnt.release(); c->sl_roff.release(); if (c->sl_scratch) cudaFree(c->sl_scratch); }
    if (c->d_cnt) cudaFree(c->d_cnt);
    if (c->h_cnt) cudaFreeHost(c->h_cnt);
    if (c->h_init) cudaFreeHost(c->h_init);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (auto& e : c->evp) if (e) cudaEventDestroy(e);
    for (auto& e : c->evk) if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

static void fill_device_params(const SzParams* p, const SzBoundary* bnd, Params& P)
{
    P.Lx = p->Lx; P.Ly = p->Ly; P.modulus = p->modulus; P.dt = p->dt; P.nu = p->nu; P.mu = p->mu; P.merge_frac = p->merge_frac;
    P.wall_frac = p->wall_frac; P.amin_per_vertex = p->amin_per_vertex; P.vertex_match_tol = p->vertex_match_tol;
    P.on_edge_tol = p->on_edge_tol; P.dl_min = p->dl_min; P.close_gap = p->close_gap; P.big_floe_r = p->big_floe_r;
    P.domain_area_frac = p->domain_area_frac; P.Nb = p->Nb; P.periodic = p->periodic; P.collision = p->collision;
    P.has_box = (bnd && bnd->box_n > 0) ? 1 : 0;
    P.bxmin = P.bxmax = P.bymin = P.bymax = P.barea = 0;
    if (P.has_box) {
        const double* bx = bnd->box_x; const double* by = bnd->box_y; const int nb = bnd->box_n;
        P.bxmin = P.bxmax = bx[0]; P.bymin = P.bymax = by[0];
        for (int i = 1; i < nb; ++i) { P.bxmin = std::min(P.bxmin, bx[i]); P.bxmax = std::max(P.bxmax, bx[i]); P.bymin = std::min(P.bymin, by[i]); P.bymax = std::max(P.bymax, by[i]); }
        // area(polyshape(c2_boundary')) (floe_interactions.m:54): vertex-0-relative shoelace, not contracted
        volatile double a2 = 0;
        for (int i = 0; i < nb; ++i) {
            const int j = (i + 1) % nb;
            volatile double xi = bx[i] - bx[0], yi = by[i] - by[0], xj = bx[j] - bx[0], yj = by[j] - by[0];
            volatile double t1 = xi * yj, t2 = xj * yi; volatile double c = t1 - t2;
            a2 = a2 + c;
        }
        P.barea = fabs(a2) / 2;
    }
}

extern "C" int sz_upload(SzContext* c, const SzParams* prm, const SzFloesSoA* f, const SzBoundary* bnd)
{
    if (!c || !prm || !f) { sz_set_error("sz_upload: NULL argument"); return SZ_ERR_ARG; }
    if (f->n < 0 || f->nverts < 0 || (f->n > 0 && (!f->x || !f->y || !f->rmax || !f->h || !f->area || !f->u || !f->v || !f->ksi || !f->alive || !f->voff)) ||
        (f->nverts > 0 && (!f->vx || !f->vy))) { sz_set_error("sz_upload: floe arrays missing"); return SZ_ERR_ARG; }
    {   // voff sanity (only when the offsets are host memory; device pointers are taken on trust)
        cudaPointerAttributes pa; bool host = true;
        if (f->n > 0 && cudaPointerGetAttributes(&pa, f->voff) == cudaSuccess) host = (pa.type == cudaMemoryTypeUnregistered || pa.type == cudaMemoryTypeHost);
        else cudaGetLastError();
        if (f->n > 0 && host && (f->voff[0] != 0 || (i64)f->voff[f->n] != f->nverts)) { sz_set_error("sz_upload: voff[0] must be 0 and voff[n] == nverts"); return SZ_ERR_ARG; }
    }
    if (f->n > 500000000 / 4) { sz_set_error("sz_upload: too many floes"); return SZ_ERR_ARG; }
    if (prm->Nb < 0 || !(prm->Lx > 0) || !(prm->Ly > 0)) { sz_set_error("sz_upload: bad Nb/Lx/Ly"); return SZ_ERR_ARG; }
    if (!prm->periodic && bnd && (bnd->n < 3 || !bnd->x || !bnd->y)) { sz_set_error("sz_upload: boundary polygon needs >= 3 vertices"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    const int n = f->n; const size_t nv = (size_t)f->nverts;
    c->have_step = false; c->have_rows = false;
    CK(c->x.ensure(n)); CK(c->y.ensure(n)); CK(c->rmax.ensure(n)); CK(c->h.ensure(n)); CK(c->area.ensure(n));
    CK(c->u.ensure(n)); CK(c->v.ensure(n)); CK(c->ksi.ensure(n)); CK(c->alive.ensure(n)); CK(c->voff.ensure(n + 1));
    CK(c->vx.ensure(nv)); CK(c->vy.ensure(nv));
    cudaStream_t st = c->stream;
    if (n > 0) {
        const size_t b = (size_t)n * 8;
        CK(cudaMemcpyAsync(c->x.p, f->x, b, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->y.p, f->y, b, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->rmax.p, f->rmax, b, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->h.p, f->h, b, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->area.p, f->area, b, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->u.p, f->u, b, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->v.p, f->v, b, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->ksi.p, f->ksi, b, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->alive.p, f->alive, n, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->voff.p, f->voff, (size_t)(n + 1) * 4, cudaMemcpyDefault, st));
        if (nv) { CK(cudaMemcpyAsync(c->vx.p, f->vx, nv * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->vy.p, f->vy, nv * 8, cudaMemcpyDefault, st)); }
    }
    c->have_bnd = false; c->bn = 0; c->boxn = 0;
    if (bnd && !prm->periodic) {
        c->have_bnd = true; c->bn = bnd->n; c->boxn = bnd->box_n;
        CK(c->bx.ensure(bnd->n)); CK(c->by.ensure(bnd->n)); CK(c->boxx.ensure(std::max(1, bnd->box_n))); CK(c->boxy.ensure(std::max(1, bnd->box_n)));
        CK(cudaMemcpyAsync(c->bx.p, bnd->x, (size_t)bnd->n * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->by.p, bnd->y, (size_t)bnd->n * 8, cudaMemcpyDefault, st));
        if (bnd->box_n > 0) { CK(cudaMemcpyAsync(c->boxx.p, bnd->box_x, (size_t)bnd->box_n * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->boxy.p, bnd->box_y, (size_t)bnd->box_n * 8, cudaMemcpyDefault, st)); }
        c->bbody.h = bnd->h; c->bbody.area = bnd->area; c->bbody.Xi = bnd->xi; c->bbody.Yi = bnd->yi; c->bbody.Ui = bnd->u; c->bbody.Vi = bnd->v; c->bbody.ksi = bnd->ksi;
    }
    c->prm = *prm; c->ext_mode = false; c->slab = false; c->have_traj = false;
    if (c->plan_n0 != f->n) c->plan_valid = false;
    fill_device_params(prm, (bnd && !prm->periodic) ? bnd : nullptr, c->dprm);
    c->n0 = n; c->nverts = f->nverts;
    CK(cudaStreamSynchronize(st));   // the caller may reuse its buffers
    c->have_input = true;
    return SZ_OK;
}

extern "C" int sz_upload_extended(SzContext* c, const SzParams* prm, const SzFloesSoA* f, const SzBoundary* bnd, const SzExtendedList* e)
{
    if (!e || (f && f->n > 0 && (!e->gid || !e->floe_num || !e->root_x || !e->root_y || !e->owned || !e->parent))) { sz_set_error("sz_upload_extended: extended-list arrays missing"); return SZ_ERR_ARG; }
    int r = sz_upload(c, prm, f, bnd);
    if (r != SZ_OK) return r;
    const int n = f->n;
    CK(c->egid.ensure(n + 1)); CK(c->efn.ensure(n + 1)); CK(c->erootx.ensure(n + 1)); CK(c->erooty.ensure(n + 1)); CK(c->eowned.ensure(n + 1)); CK(c->eparent.ensure(n + 1));
    CK(c->esrc.ensure(n + 1)); CK(c->ex.ensure(n + 1)); CK(c->ey.ensure(n + 1)); CK(c->ealive.ensure(n + 1));
    if (n > 0) {
        cudaStream_t st = c->stream;
        CK(cudaMemcpyAsync(c->egid.p, e->gid, (size_t)n * 4, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->efn.p, e->floe_num, (size_t)n * 4, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->erootx.p, e->root_x, (size_t)n * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->erooty.p, e->root_y, (size_t)n * 8, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->eowned.p, e->owned, (size_t)n, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->eparent.p, e->parent, (size_t)n * 4, cudaMemcpyDefault, st));
        std::vector<int> ident((size_t)n); for (int k = 0; k < n; ++k) ident[k] = k;
        CK(cudaMemcpyAsync(c->esrc.p, ident.data(), (size_t)n * 4, cudaMemcpyDefault, st));
        CK(cudaStreamSynchronize(st));
    }
    c->ext_mode = true;
    return SZ_OK;
}

// ------------------------------------------------------------------------------------------------ multi-GPU slabs
// One process per GPU; every rank owns a set of floes (ascending global numbers, the numbering of the single-GPU run) and keeps
// their state -- and the integrator's -- resident here.  Every step the part of the GLOBAL extended floe list this rank needs is
// rebuilt on the device from the current state, so the step is exact for a field that translates, rotates, thins and loses
// floes (calc_trajectory.m:75-79,170-222): there is no plan that could go stale.
//   sz_slab_prepare   periodic-image flags of the owned floes (floe_interactions_all.m:28-31,49-52) from their CURRENT outlines,
//                     x-extents, largest rmax -> this rank's meta record (the caller all-gathers the records)
//   sz_slab_pack      global numbers of the rank's images (their rank among ALL ranks' images, :33-36,54-57), halo selection:
//                     every own entry within reach (2 max rmax) of another rank's x-extent is packed for that rank -- state,
//                     FloeNums, root centroid AND its current outline (the caller runs one all-to-all of fixed-size blocks)
//   sz_slab_build     received entries sorted by global number and merged with the own ones into the resident extended list
//                     (ascending global position, so `j > i`, partner ids and row order are the single-GPU ones); then
//                     sz_step_resident resolves every pair with an owned floe, and sz_trajectory_step integrates the owned floes
// Meta record of a rank, SZ_SLAB_META(cap_img) doubles: [0] x-images [1] y-images of originals [2] y-images of x-images
// [3,4] x-extent of the originals [5,6] x-extent of the x-images [7] largest rmax, then the global numbers of the originals that
// have an x-image / a y-image / both (cap_img each, ascending).
// Block for one peer, SZ_SLAB_BLOCK(cap_rec, cap_vert) doubles: [0] records [1] vertices, cap_rec records of SL_REC doubles
// (gid FloeNum X Y rootX rootY rmax h area u v ksi alive nverts vstart -), cap_vert (x, y) pairs.
#define SL_REC 16
#define SL_META_HDR 8
struct SlabScratch { u64 ext[4]; u64 rmax_bits; int overflow; int n_list; int n_outside; int pad; };     // a_lo a_hi b_lo b_hi
__device__ __forceinline__ int lower_bound_d(const double* a, int n, double v) { int lo = 0, hi = n; while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; } return lo; }
__device__ __forceinline__ int lower_bound_i(const int* a, int n, int v) { int lo = 0, hi = n; while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < v) lo = mid + 1; else hi = mid; } return lo; }

__global__ void slab_flag_kernel(int n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ rmax, const uint8_t* __restrict__ alive,
                                 const int* __restrict__ voff, const double* __restrict__ vx, const double* __restrict__ vy, double Lx, double Ly, int periodic,
                                 int* __restrict__ fx, int* __restrict__ fy, int* __restrict__ fxy, SlabScratch* s, double xlo, double xhi)
{
    int outside = 0;
    double alo = SZ_INF, ahi = -SZ_INF, blo = SZ_INF, bhi = -SZ_INF, rm = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double X = x[i], Y = y[i];
        int a = 0, b = 0;
        if (periodic && alive[i]) {
            double mx = -SZ_INF, my = -SZ_INF;
            for (int t = voff[i]; t < voff[i + 1]; ++t) { const double ax = fabs(vx[t] + X), ay = fabs(vy[t] + Y); if (ax > mx) mx = ax; if (ay > my) my = ay; }
            a = mx > Lx; b = my > Ly;            // :31,52 -- an x-image keeps its parent's Yi and outline, so its y flag is the parent's
        }
        fx[i] = a; fy[i] = b; fxy[i] = a & b;
        if (X == X) { outside += (alive[i] && (X < xlo || X >= xhi)); alo = fmin(alo, X); ahi = fmax(ahi, X); if (a) { const double Xg = X - 2 * Lx * sgn_d(X); blo = fmin(blo, Xg); bhi = fmax(bhi, Xg); } }
        rm = fmax(rm, rmax[i]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        alo = fmin(alo, __shfl_xor_sync(0xffffffffu, alo, d)); ahi = fmax(ahi, __shfl_xor_sync(0xffffffffu, ahi, d));
        blo = fmin(blo, __shfl_xor_sync(0xffffffffu, blo, d)); bhi = fmax(bhi, __shfl_xor_sync(0xffffffffu, bhi, d));
        rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, d)); outside += __shfl_xor_sync(0xffffffffu, outside, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (outside) atomicAdd(&s->n_outside, outside);
        if (alo <= ahi) { atomicMin(&s->ext[0], enc_d(alo)); atomicMax(&s->ext[1], enc_d(ahi)); }
        if (blo <= bhi) { atomicMin(&s->ext[2], enc_d(blo)); atomicMax(&s->ext[3], enc_d(bhi)); }
        atomicMax(&s->rmax_bits, enc_d(rm));
    }
}
__global__ void slab_scratch_init_kernel(SlabScratch* s)
{
    s->ext[0] = enc_d(SZ_INF); s->ext[1] = enc_d(-SZ_INF); s->ext[2] = enc_d(SZ_INF); s->ext[3] = enc_d(-SZ_INF); s->rmax_bits = enc_d(0.0); s->n_list = 0; s->n_outside = 0;
    // overflow is sticky until the host has seen it (cleared by sz_slab_prepare)
}
__global__ void slab_meta_kernel(int n, int cap_img, const int* __restrict__ ogid, const int* __restrict__ fx, const int* __restrict__ fy, const int* __restrict__ fxy,
                                 const int* __restrict__ px, const int* __restrict__ py, const int* __restrict__ pxy, SlabScratch* s, double* __restrict__ meta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        meta[0] = (double)px[n]; meta[1] = (double)py[n]; meta[2] = (double)pxy[n];
        meta[3] = dec_d(s->ext[0]); meta[4] = dec_d(s->ext[1]); meta[5] = dec_d(s->ext[2]); meta[6] = dec_d(s->ext[3]); meta[7] = dec_d(s->rmax_bits);
        if (px[n] > cap_img || py[n] > cap_img) s->overflow = 1;
    }
    if (i >= n) return;
    const double g = (double)ogid[i];
    if (fx[i] && px[i] < cap_img) meta[SL_META_HDR + px[i]] = g;
    if (fy[i] && py[i] < cap_img) meta[SL_META_HDR + cap_img + py[i]] = g;
    if (fxy[i] && pxy[i] < cap_img) meta[SL_META_HDR + 2 * cap_img + pxy[i]] = g;
}
struct SlabPackArgs {
    int n, rank, world, cap_img, cap_rec, cap_vert, n_global, periodic; double Lx, Ly;
    const double* all_meta; int meta_stride;
    const int* ogid; const double* x; const double* y; const double* rmax; const double* h; const double* area; const double* u; const double* v; const double* ksi; const uint8_t* alive;
    const int* voff; const double* vx; const double* vy;
    const int* fx; const int* fy; const int* fxy; const int* px; const int* py; const int* pxy;
    int* gx; int* gyo; int* gyx;                 // global positions of the own images (ascending)
    double* send; long long block; int* send_cnt; SlabScratch* s;
};
// rank (0-based) of global number G among ALL ranks' originals of one image class = sum of lower bounds in the ranks' lists
__device__ __forceinline__ int slab_global_rank(const SlabPackArgs& a, int cls, double G)
{
    int r = 0;
    for (int p = 0; p < a.world; ++p) {
        const double* m = a.all_meta + (size_t)p * a.meta_stride;
        int cnt = (int)m[cls]; if (cnt > a.cap_img) cnt = a.cap_img;
        r += lower_bound_d(m + SL_META_HDR + cls * a.cap_img, cnt, G);
    }
    return r;
}
__device__ void slab_send(const SlabPackArgs& a, double reach, int i, int gid, int fnum, double X, double Y)
{
    for (int p = 0; p < a.world; ++p) {
        if (p == a.rank) continue;
        const double* m = a.all_meta + (size_t)p * a.meta_stride;
        const bool near_a = X >= m[3] - reach && X <= m[4] + reach, near_b = X >= m[5] - reach && X <= m[6] + reach;
        if (!(near_a || near_b)) continue;
        const int nv = a.voff[i + 1] - a.voff[i];
        const int slot = atomicAdd(&a.send_cnt[2 * p], 1), vs = atomicAdd(&a.send_cnt[2 * p + 1], nv);
        if (slot >= a.cap_rec || vs + nv > a.cap_vert) { a.s->overflow = 1; continue; }
        double* blk = a.send + (size_t)p * a.block;
        double* r = blk + 2 + (size_t)slot * SL_REC;
        r[0] = gid; r[1] = fnum; r[2] = X; r[3] = Y; r[4] = a.x[i]; r[5] = a.y[i]; r[6] = a.rmax[i]; r[7] = a.h[i]; r[8] = a.area[i];
        r[9] = a.u[i]; r[10] = a.v[i]; r[11] = a.ksi[i]; r[12] = a.alive[i]; r[13] = nv; r[14] = vs; r[15] = 0;
        double* vdst = blk + 2 + (size_t)a.cap_rec * SL_REC + 2 * (size_t)vs;
        for (int t = 0; t < nv; ++t) { vdst[2 * t] = a.vx[a.voff[i] + t]; vdst[2 * t + 1] = a.vy[a.voff[i] + t]; }
    }
}
// one thread per owned floe: numbers its images in the global list and packs every entry another rank needs
__global__ void slab_pack_kernel(const SlabPackArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    double rm = 0; int cx_tot = 0, cyo_tot = 0;
    for (int p = 0; p < a.world; ++p) { const double* m = a.all_meta + (size_t)p * a.meta_stride; rm = fmax(rm, m[7]); cx_tot += (int)m[0]; cyo_tot += (int)m[1]; }
    const double reach = 2 * rm;
    const int G = a.ogid[i];
    const double X = a.x[i], Y = a.y[i];
    const int n1g = a.n_global + cx_tot, n2g = n1g + cyo_tot;
    slab_send(a, reach, i, G, G, X, Y);
    if (!a.periodic) return;
    double Xg = X, Yg = Y;
    if (a.fx[i]) {
        Xg = X - 2 * a.Lx * sgn_d(X);                                               // :34
        const int g = a.n_global + slab_global_rank(a, 0, (double)G) + 1;
        if (a.px[i] < a.cap_img) a.gx[a.px[i]] = g;
        slab_send(a, reach, i, g, -G, Xg, Y);
    }
    if (a.fy[i]) {
        Yg = Y - 2 * a.Ly * sgn_d(Y);                                               // :55
        const int g = n1g + slab_global_rank(a, 1, (double)G) + 1;
        if (a.py[i] < a.cap_img) a.gyo[a.py[i]] = g;
        slab_send(a, reach, i, g, -G, X, Yg);
    }
    if (a.fxy[i]) {
        const int g = n2g + slab_global_rank(a, 2, (double)G) + 1;
        if (a.pxy[i] < a.cap_img) a.gyx[a.pxy[i]] = g;
        slab_send(a, reach, i, g, -G, Xg, Yg);
    }
}
__global__ void slab_header_kernel(int world, int cap_rec, int cap_vert, const int* __restrict__ send_cnt, double* send, long long block)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= world) return;
    const int nr = send_cnt[2 * p], nv = send_cnt[2 * p + 1];
    send[(size_t)p * block] = (double)(nr < cap_rec ? nr : cap_rec);
    send[(size_t)p * block + 1] = (double)(nv < cap_vert ? nv : cap_vert);
}
// received records: sort key (global position, INT_MAX for unused slots), vertex count per slot
__global__ void slab_keys_kernel(int world, int rank, int cap_rec, const double* __restrict__ recv, long long block, int* __restrict__ keys, int* __restrict__ slots, int* __restrict__ hnv)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= world * cap_rec) return;
    const int p = k / cap_rec, s = k - p * cap_rec;
    const double* blk = recv + (size_t)p * block;
    int key = 0x7fffffff, nv = 0;
    if (p != rank && s < (int)blk[0]) { const double* r = blk + 2 + (size_t)s * SL_REC; key = (int)r[0]; nv = (int)r[13]; }
    keys[k] = key; slots[k] = k; hnv[k] = nv;
}
struct SlabBuildArgs {
    int n, rank, world, cap_img, cap_rec, cap_vert, nl_cap, periodic; double Lx, Ly; long long vown;
    const double* recv; long long block;
    const int* keys_sorted; const int* slots_sorted; const int* hvoff;      // hvoff: exclusive scan of the slots' vertex counts
    const int* ogid; const int* fx; const int* fy; const int* fxy; const int* px; const int* py; const int* pxy; const int* gx; const int* gyo; const int* gyx;
    double* x; double* y; double* rmax; double* h; double* area; double* u; double* v; double* ksi; uint8_t* alive; int* voff; double* vx; double* vy;     // body records: [owned | halo slots]
    double* ex; double* ey; double* erootx; double* erooty; int* esrc; int* efn; int* eparent; int* egid; uint8_t* ealive; uint8_t* eowned;
    int* opos; int* child0; int* child1; SlabScratch* s;      // child0/1: the (at most two) images of an entry, in creation order (:242-245)
};
// halo records: body record n + slot, outline copied behind the owned outlines, list entry at (rank among received) + (own entries before)
__global__ void slab_build_halo_kernel(const SlabBuildArgs a)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;          // position among the sorted received keys
    const int nslots = a.world * a.cap_rec;
    if (k >= nslots) return;
    const int slot = a.slots_sorted[k], key = a.keys_sorted[k];
    const int src = a.n + slot;
    // every slot gets a (possibly empty) outline range, so that voff stays monotone over [owned | slots]
    a.voff[src + 1] = (int)a.vown + a.hvoff[slot + 1];
    if (key == 0x7fffffff) return;
    const int p = slot / a.cap_rec, s = slot - p * a.cap_rec;
    const double* blk = a.recv + (size_t)p * a.block;
    const double* r = blk + 2 + (size_t)s * SL_REC;
    const int cx = a.px[a.n], cyo = a.py[a.n], cyx = a.pxy[a.n];
    const int own_before = lower_bound_i(a.ogid, a.n, key) + lower_bound_i(a.gx, cx < a.cap_img ? cx : a.cap_img, key)
                         + lower_bound_i(a.gyo, cyo < a.cap_img ? cyo : a.cap_img, key) + lower_bound_i(a.gyx, cyx < a.cap_img ? cyx : a.cap_img, key);
    const int pos = k + own_before;
    if (pos >= a.nl_cap) { a.s->overflow = 1; return; }
    a.rmax[src] = r[6]; a.h[src] = r[7]; a.area[src] = r[8]; a.u[src] = r[9]; a.v[src] = r[10]; a.ksi[src] = r[11]; a.alive[src] = (uint8_t)r[12]; a.x[src] = r[4]; a.y[src] = r[5];
    const int nv = (int)r[13], vs = (int)r[14];
    const double* vsrc = blk + 2 + (size_t)a.cap_rec * SL_REC + 2 * (size_t)vs;
    const size_t o = (size_t)a.vown + a.hvoff[slot];
    for (int t = 0; t < nv; ++t) { a.vx[o + t] = vsrc[2 * t]; a.vy[o + t] = vsrc[2 * t + 1]; }
    a.ex[pos] = r[2]; a.ey[pos] = r[3]; a.erootx[pos] = r[4]; a.erooty[pos] = r[5]; a.esrc[pos] = src; a.efn[pos] = (int)r[1]; a.eparent[pos] = 0; a.egid[pos] = key;
    a.ealive[pos] = (uint8_t)r[12]; a.eowned[pos] = 0; a.child0[pos] = -1; a.child1[pos] = -1;
}
// owned floes and their images: list position = (own entries before) + (received entries with a smaller global position)
__global__ void slab_build_own_kernel(const SlabBuildArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int nslots = a.world * a.cap_rec;
    const int cx = a.px[a.n], cyo = a.py[a.n], cyx = a.pxy[a.n];
    if (i == 0) {
        int n_recv = lower_bound_i(a.keys_sorted, nslots, 0x7fffffff);
        a.s->n_list = a.n + cx + cyo + cyx + n_recv;
        if (a.s->n_list > a.nl_cap) a.s->overflow = 1;
        a.voff[a.n] = (int)a.vown;      // first halo slot starts behind the owned outlines (== ovoff[n])
    }
    if (i >= a.n) return;
    const int G = a.ogid[i];
    const double X = a.x[i], Y = a.y[i];
    auto put = [&](int own_rank, int g, int fnum, double ex, double ey, int parent_pos) {
        const int pos = own_rank + lower_bound_i(a.keys_sorted, nslots, g);
        if (pos >= a.nl_cap) { a.s->overflow = 1; return -1; }
        a.ex[pos] = ex; a.ey[pos] = ey; a.erootx[pos] = X; a.erooty[pos] = Y; a.esrc[pos] = i; a.efn[pos] = fnum; a.eparent[pos] = parent_pos + 1; a.egid[pos] = g;
        a.ealive[pos] = a.alive[i]; a.eowned[pos] = 1; a.child0[pos] = -1; a.child1[pos] = -1;
        return pos;
    };
    const int p0 = put(i, G, G, X, Y, -1);
    a.opos[i] = p0;
    if (!a.periodic) return;
    double Xg = X, Yg = Y; int pxg = -1;
    int pyg = -1, pxyg = -1;
    if (a.fx[i] && a.px[i] < a.cap_img) { Xg = X - 2 * a.Lx * sgn_d(X); pxg = put(a.n + a.px[i], a.gx[a.px[i]], -G, Xg, Y, p0); }
    if (a.fy[i] && a.py[i] < a.cap_img) { Yg = Y - 2 * a.Ly * sgn_d(Y); pyg = put(a.n + cx + a.py[i], a.gyo[a.py[i]], -G, X, Yg, p0); }
    if (a.fxy[i] && a.pxy[i] < a.cap_img) pxyg = put(a.n + cx + cyo + a.pxy[i], a.gyx[a.pxy[i]], -G, Xg, Yg, pxg);
    // images of an entry in creation order: the floe's x-image, then its y-image; an x-image's only image is the xy one
    if (p0 >= 0) { a.child0[p0] = pxg >= 0 ? pxg : pyg; a.child1[p0] = pxg >= 0 ? pyg : -1; }
    if (pxg >= 0) a.child0[pxg] = pxyg;
}
// entries behind the list's end are inert: dead, unowned, nowhere
__global__ void slab_build_tail_kernel(int nl_cap, const int* __restrict__ n_dev, double* ex, double* ey, double* erootx, double* erooty, int* esrc, int* efn, int* eparent, int* egid,
                                       uint8_t* ealive, uint8_t* eowned, int* child0, int* child1)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nl_cap || e < *n_dev) return;
    if (child0) { child0[e] = -1; child1[e] = -1; }
    const double nan = SZ_INF - SZ_INF;
    ex[e] = nan; ey[e] = nan; erootx[e] = nan; erooty[e] = nan; esrc[e] = 0; efn[e] = 0; eparent[e] = 0; egid[e] = 0x7fffffff; ealive[e] = 0; eowned[e] = 0;
}
__global__ void slab_count_halo_kernel(int n, int rank, int world, int periodic, double Lx, double Ly, double reach, const double* __restrict__ ext4,
                                       const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ voff,
                                       const int* __restrict__ fx, const int* __restrict__ fy, const int* __restrict__ fxy, unsigned long long* __restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double X = x[i], Y = y[i];
    const double Xg = X - 2 * Lx * sgn_d(X);
    const int nv = voff[i + 1] - voff[i];
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        const double* m = ext4 + 4 * p;
        const bool near0 = (X >= m[0] - reach && X <= m[1] + reach) || (X >= m[2] - reach && X <= m[3] + reach);
        const bool near1 = (Xg >= m[0] - reach && Xg <= m[1] + reach) || (Xg >= m[2] - reach && Xg <= m[3] + reach);
        int k = near0 ? 1 : 0;
        if (periodic) { if (fx[i] && near1) ++k; if (fy[i] && near0) ++k; if (fxy[i] && near1) ++k; }
        if (k) { atomicAdd(&cnt[2 * p], (unsigned long long)k); atomicAdd(&cnt[2 * p + 1], (unsigned long long)k * nv); }
    }
}
__global__ void slab_status_kernel(const SlabScratch* __restrict__ s, int* __restrict__ out) { out[0] = s->overflow; out[1] = s->n_outside; out[2] = s->n_list; out[3] = 0; }

extern "C" int64_t sz_slab_meta_doubles(int32_t cap_img) { return SL_META_HDR + 3 * (int64_t)cap_img; }
extern "C" int64_t sz_slab_block_doubles(int32_t cap_rec, int32_t cap_vert) { return 2 + (int64_t)cap_rec * SL_REC + 2 * (int64_t)cap_vert; }

extern "C" int sz_slab_upload(SzContext* c, const SzParams* prm, const SzFloesSoA* f, const SzBoundary* bnd, const int32_t* gid, int32_t n_global, int32_t rank, int32_t world)
{
    if (!c || !prm || !f || (f->n > 0 && !gid) || world < 1 || rank < 0 || rank >= world || n_global < f->n) { sz_set_error("sz_slab_upload: bad argument"); return SZ_ERR_ARG; }
    int r = sz_upload(c, prm, f, bnd);
    if (r != SZ_OK) return r;
    const int n = f->n;
    CK(c->sl_ogid.ensure(n + 1));
    if (n > 0) CK(cudaMemcpy(c->sl_ogid.p, gid, (size_t)n * 4, cudaMemcpyDefault));
    CK(c->sl_flag.ensure(3 * (size_t)(n + 1))); CK(c->sl_pos.ensure(3 * (size_t)(n + 2))); CK(c->sl_opos.ensure(n + 1));
    CK(c->scan_tmp.ensure(scan_tmp_ints((size_t)n + 2)));
    if (!c->sl_scratch) CK(cudaMalloc(&c->sl_scratch, sizeof(SlabScratch)));
    CK(cudaMemset(c->sl_scratch, 0, sizeof(SlabScratch)));
    c->slab = true; c->ext_mode = true; c->sl_xlo = -SZ_INF; c->sl_xhi = SZ_INF; c->sl_rank = rank; c->sl_world = world; c->sl_nglobal = n_global; c->sl_configured = false; c->sl_built = false;
    return SZ_OK;
}
// flags, scans and extents of the owned floes from their current state (shared by sz_slab_measure and sz_slab_prepare)
static int slab_flags(SzContext* c)
{
    cudaStream_t st = c->stream; const int n = c->n0; const SzParams& P = c->prm;
    int* fx = c->sl_flag.p; int* fy = fx + (n + 1); int* fxy = fy + (n + 1);
    int* px = c->sl_pos.p; int* py = px + (n + 2); int* pxy = py + (n + 2);
    ++g_launches; slab_scratch_init_kernel<<<1, 1, 0, st>>>(c->sl_scratch);
    if (n > 0) { ++g_launches; slab_flag_kernel<<<std::min(nblk(n, 256), 148 * 16), 256, 0, st>>>(n, c->x.p, c->y.p, c->rmax.p, c->alive.p, c->voff.p, c->vx.p, c->vy.p, P.Lx, P.Ly, P.periodic, fx, fy, fxy, c->sl_scratch, c->sl_xlo, c->sl_xhi); }
    CKS(exclusive_scan(c, fx, n, px, n + 1));
    CKS(exclusive_scan(c, fy, n, py, n + 1));
    CKS(exclusive_scan(c, fxy, n, pxy, n + 1));
    CK(cudaGetLastError());
    return SZ_OK;
}
// host-synchronous measurements for choosing capacities: local8 = this rank's meta header (counts, extents, largest rmax)
extern "C" int sz_slab_measure(SzContext* c, double* local8)
{
    if (!c || !local8 || !c->slab) { sz_set_error("sz_slab_measure: needs sz_slab_upload"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    CKS(slab_flags(c));
    const int n = c->n0; cudaStream_t st = c->stream;
    int cnt[3]; SlabScratch s;
    int* px = c->sl_pos.p;
    CK(cudaMemcpyAsync(&cnt[0], px + n, 4, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(&cnt[1], px + (n + 2) + n, 4, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(&cnt[2], px + 2 * (n + 2) + n, 4, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(&s, c->sl_scratch, sizeof(s), cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    local8[0] = cnt[0]; local8[1] = cnt[1]; local8[2] = cnt[2];
    for (int k = 0; k < 4; ++k) local8[3 + k] = dec_d(s.ext[k]);
    local8[7] = dec_d(s.rmax_bits);
    return SZ_OK;
}
// all8: every rank's header (host, [world*8]); counts of the records and vertices this rank would send to each peer
extern "C" int sz_slab_measure_halo(SzContext* c, const double* all8, int64_t* rec_counts, int64_t* vert_counts)
{
    if (!c || !all8 || !rec_counts || !vert_counts || !c->slab) { sz_set_error("sz_slab_measure_halo: needs sz_slab_upload"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const int n = c->n0, W = c->sl_world; cudaStream_t st = c->stream; const SzParams& P = c->prm;
    std::vector<double> ext(4 * (size_t)W); double rm = 0;
    for (int p = 0; p < W; ++p) { for (int k = 0; k < 4; ++k) ext[4 * p + k] = all8[8 * p + 3 + k]; rm = std::max(rm, all8[8 * p + 7]); }
    double* d_ext = nullptr; unsigned long long* d_cnt = nullptr;
    CK(cudaMalloc(&d_ext, ext.size() * 8)); CK(cudaMalloc(&d_cnt, 2 * (size_t)W * 8));
    CK(cudaMemcpyAsync(d_ext, ext.data(), ext.size() * 8, cudaMemcpyDefault, st)); CK(cudaMemsetAsync(d_cnt, 0, 2 * (size_t)W * 8, st));
    int* fx = c->sl_flag.p; int* fy = fx + (n + 1); int* fxy = fy + (n + 1);
    if (n > 0) { ++g_launches; slab_count_halo_kernel<<<nblk(n, 256), 256, 0, st>>>(n, c->sl_rank, W, P.periodic, P.Lx, P.Ly, 2 * rm, d_ext, c->x.p, c->y.p, c->voff.p, fx, fy, fxy, d_cnt); }
    std::vector<unsigned long long> h(2 * (size_t)W);
    CK(cudaMemcpyAsync(h.data(), d_cnt, h.size() * 8, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    cudaFree(d_ext); cudaFree(d_cnt);
    for (int p = 0; p < W; ++p) { rec_counts[p] = (int64_t)h[2 * p]; vert_counts[p] = (int64_t)h[2 * p + 1]; }
    return SZ_OK;
}
extern "C" int sz_slab_configure(SzContext* c, int32_t cap_img, int32_t cap_rec, int32_t cap_vert)
{
    if (!c || !c->slab || cap_img < 1 || cap_rec < 1 || cap_vert < 1) { sz_set_error("sz_slab_configure: needs sz_slab_upload and positive capacities"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    const int n = c->n0, W = c->sl_world;
    const size_t nsrc = (size_t)n + (size_t)W * cap_rec, V = (size_t)c->nverts + (size_t)W * cap_vert;
    if (nsrc + 4 * (size_t)cap_img > 0x7ffffff0u || V > 0x7ffffff0u) { sz_set_error("sz_slab_configure: capacities overflow 32-bit indices"); return SZ_ERR_ARG; }
    DBuf<double>* body[] = {&c->x, &c->y, &c->rmax, &c->h, &c->area, &c->u, &c->v, &c->ksi};
    for (auto* b : body) CK(b->ensure(nsrc + 1, true));
    CK(c->alive.ensure(nsrc + 1, true)); CK(c->voff.ensure(nsrc + 2, true)); CK(c->vx.ensure(V + 1, true)); CK(c->vy.ensure(V + 1, true));
    const size_t nl = (size_t)n + 3 * (size_t)cap_img + (size_t)W * cap_rec;
    CK(c->ex.ensure(nl + 1)); CK(c->ey.ensure(nl + 1)); CK(c->erootx.ensure(nl + 1)); CK(c->erooty.ensure(nl + 1)); CK(c->esrc.ensure(nl + 1)); CK(c->efn.ensure(nl + 1));
    CK(c->eparent.ensure(nl + 1)); CK(c->egid.ensure(nl + 1)); CK(c->ealive.ensure(nl + 1)); CK(c->eowned.ensure(nl + 1)); CK(c->gx_of.ensure(nl + 1)); CK(c->gy_of.ensure(nl + 1));
    CK(c->sl_g.ensure(3 * (size_t)cap_img + 3)); CK(c->sl_sendcnt.ensure(2 * (size_t)W + 2));
    const size_t ns = (size_t)W * cap_rec;
    CK(c->sl_keys.ensure(2 * ns + 2)); CK(c->sl_slots.ensure(2 * ns + 2)); CK(c->sl_hnv.ensure(ns + 2)); CK(c->sl_hvoff.ensure(ns + 2));
    CK(c->scan_tmp.ensure(scan_tmp_ints(std::max<size_t>(ns + 2, (size_t)n + 2))));
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, c->sl_keys.p, c->sl_keys.p + ns, c->sl_slots.p, c->sl_slots.p + ns, (int)ns, 0, 32, c->stream);
    CK(c->sl_cub.ensure(bytes + 16));
    c->sl_cap_img = cap_img; c->sl_cap_rec = cap_rec; c->sl_cap_vert = cap_vert; c->sl_nl_cap = (int)nl; c->sl_configured = true; c->sl_built = false;
    return SZ_OK;
}
extern "C" int sz_slab_set_extent(SzContext* c, double xlo, double xhi)
{
    if (!c || !c->slab) { sz_set_error("sz_slab_set_extent: needs sz_slab_upload"); return SZ_ERR_STATE; }
    c->sl_xlo = xlo; c->sl_xhi = xhi;
    return SZ_OK;
}
extern "C" int sz_slab_prepare(SzContext* c, double* meta_dev)
{
    if (!c || !meta_dev || !c->slab || !c->sl_configured) { sz_set_error("sz_slab_prepare: needs sz_slab_upload and sz_slab_configure"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const int n = c->n0;
    c->sl_built = false; c->have_step = false;
    CK(cudaMemsetAsync(&c->sl_scratch->overflow, 0, 4, st));
    CKS(slab_flags(c));
    int* fx = c->sl_flag.p; int* fy = fx + (n + 1); int* fxy = fy + (n + 1);
    int* px = c->sl_pos.p; int* py = px + (n + 2); int* pxy = py + (n + 2);
    ++g_launches; slab_meta_kernel<<<std::max(1, nblk(n, 256)), 256, 0, st>>>(n, c->sl_cap_img, c->sl_ogid.p, fx, fy, fxy, px, py, pxy, c->sl_scratch, meta_dev);
    CK(cudaGetLastError());
    return SZ_OK;
}
extern "C" int sz_slab_pack(SzContext* c, const double* all_meta_dev, double* send_dev)
{
    if (!c || !all_meta_dev || !send_dev || !c->slab || !c->sl_configured) { sz_set_error("sz_slab_pack: needs sz_slab_prepare"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const int n = c->n0, W = c->sl_world; const SzParams& P = c->prm;
    SlabPackArgs a; memset(&a, 0, sizeof(a));
    a.n = n; a.rank = c->sl_rank; a.world = W; a.cap_img = c->sl_cap_img; a.cap_rec = c->sl_cap_rec; a.cap_vert = c->sl_cap_vert; a.n_global = c->sl_nglobal; a.periodic = P.periodic; a.Lx = P.Lx; a.Ly = P.Ly;
    a.all_meta = all_meta_dev; a.meta_stride = (int)sz_slab_meta_doubles(c->sl_cap_img);
    a.ogid = c->sl_ogid.p; a.x = c->x.p; a.y = c->y.p; a.rmax = c->rmax.p; a.h = c->h.p; a.area = c->area.p; a.u = c->u.p; a.v = c->v.p; a.ksi = c->ksi.p; a.alive = c->alive.p;
    a.voff = c->voff.p; a.vx = c->vx.p; a.vy = c->vy.p;
    a.fx = c->sl_flag.p; a.fy = a.fx + (n + 1); a.fxy = a.fy + (n + 1); a.px = c->sl_pos.p; a.py = a.px + (n + 2); a.pxy = a.py + (n + 2);
    a.gx = c->sl_g.p; a.gyo = a.gx + (c->sl_cap_img + 1); a.gyx = a.gyo + (c->sl_cap_img + 1);
    a.send = send_dev; a.block = sz_slab_block_doubles(c->sl_cap_rec, c->sl_cap_vert); a.send_cnt = c->sl_sendcnt.p; a.s = c->sl_scratch;
    CK(cudaMemsetAsync(c->sl_sendcnt.p, 0, 2 * (size_t)W * 4, st));
    if (n > 0) { ++g_launches; slab_pack_kernel<<<nblk(n, 128), 128, 0, st>>>(a); }
    ++g_launches; slab_header_kernel<<<1, std::max(32, W), 0, st>>>(W, c->sl_cap_rec, c->sl_cap_vert, c->sl_sendcnt.p, send_dev, a.block);
    CK(cudaGetLastError());
    return SZ_OK;
}
extern "C" int sz_slab_build(SzContext* c, const double* recv_dev, int32_t* status_dev)
{
    if (!c || !recv_dev || !c->slab || !c->sl_configured) { sz_set_error("sz_slab_build: needs sz_slab_pack"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const int n = c->n0, W = c->sl_world; const SzParams& P = c->prm;
    const int ns = W * c->sl_cap_rec;
    const long long block = sz_slab_block_doubles(c->sl_cap_rec, c->sl_cap_vert);
    int* keys = c->sl_keys.p; int* keys_s = keys + ns; int* slots = c->sl_slots.p; int* slots_s = slots + ns;
    ++g_launches; slab_keys_kernel<<<nblk(ns, 256), 256, 0, st>>>(W, c->sl_rank, c->sl_cap_rec, recv_dev, block, keys, slots, c->sl_hnv.p);
    size_t bytes = c->sl_cub.cap;
    ++g_launches; CK(cub::DeviceRadixSort::SortPairs(c->sl_cub.p, bytes, keys, keys_s, slots, slots_s, ns, 0, 32, st));
    CKS(exclusive_scan(c, c->sl_hnv.p, ns, c->sl_hvoff.p, ns + 1));
    SlabBuildArgs a; memset(&a, 0, sizeof(a));
    a.n = n; a.rank = c->sl_rank; a.world = W; a.cap_img = c->sl_cap_img; a.cap_rec = c->sl_cap_rec; a.cap_vert = c->sl_cap_vert; a.nl_cap = c->sl_nl_cap; a.periodic = P.periodic; a.Lx = P.Lx; a.Ly = P.Ly; a.vown = c->nverts;
    a.recv = recv_dev; a.block = block; a.keys_sorted = keys_s; a.slots_sorted = slots_s; a.hvoff = c->sl_hvoff.p;
    a.ogid = c->sl_ogid.p; a.fx = c->sl_flag.p; a.fy = a.fx + (n + 1); a.fxy = a.fy + (n + 1); a.px = c->sl_pos.p; a.py = a.px + (n + 2); a.pxy = a.py + (n + 2);
    a.gx = c->sl_g.p; a.gyo = a.gx + (c->sl_cap_img + 1); a.gyx = a.gyo + (c->sl_cap_img + 1);
    a.x = c->x.p; a.y = c->y.p; a.rmax = c->rmax.p; a.h = c->h.p; a.area = c->area.p; a.u = c->u.p; a.v = c->v.p; a.ksi = c->ksi.p; a.alive = c->alive.p; a.voff = c->voff.p; a.vx = c->vx.p; a.vy = c->vy.p;
    a.ex = c->ex.p; a.ey = c->ey.p; a.erootx = c->erootx.p; a.erooty = c->erooty.p; a.esrc = c->esrc.p; a.efn = c->efn.p; a.eparent = c->eparent.p; a.egid = c->egid.p; a.ealive = c->ealive.p; a.eowned = c->eowned.p;
    a.opos = c->sl_opos.p; a.child0 = c->gx_of.p; a.child1 = c->gy_of.p; a.s = c->sl_scratch;
    ++g_launches; slab_build_own_kernel<<<std::max(1, nblk(n, 128)), 128, 0, st>>>(a);
    ++g_launches; slab_build_halo_kernel<<<nblk(ns, 128), 128, 0, st>>>(a);
    ++g_launches; slab_build_tail_kernel<<<nblk(c->sl_nl_cap, 256), 256, 0, st>>>(c->sl_nl_cap, &c->sl_scratch->n_list, c->ex.p, c->ey.p, c->erootx.p, c->erooty.p, c->esrc.p, c->efn.p, c->eparent.p, c->egid.p, c->ealive.p, c->eowned.p, c->gx_of.p, c->gy_of.p);
    if (status_dev) { ++g_launches; slab_status_kernel<<<1, 1, 0, st>>>(c->sl_scratch, status_dev); }
    CK(cudaGetLastError());
    c->sl_built = true;
    return SZ_OK;
}
// list positions of the owned floes (0-based) in the resident extended list of the last sz_slab_build, and that list's length
extern "C" int sz_slab_get_positions(SzContext* c, int32_t* opos, int32_t* n_list)
{
    if (!c || !c->slab || !c->sl_built) { sz_set_error("sz_slab_get_positions: needs sz_slab_build"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    if (opos && c->n0 > 0) CK(cudaMemcpyAsync(opos, c->sl_opos.p, (size_t)c->n0 * 4, cudaMemcpyDefault, c->stream));
    SlabScratch s; CK(cudaMemcpyAsync(&s, c->sl_scratch, sizeof(s), cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (n_list) *n_list = s.n_list;
    return SZ_OK;
}

// debugging aid (SZ_DEBUG_SYNC=1): synchronise at the phase boundaries of the step and name the phase a CUDA error belongs to
static int dbg_sync(SzContext* c, const char* what)
{
    static const bool on = getenv("SZ_DEBUG_SYNC") != nullptr;
    if (!on) return SZ_OK;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { sz_set_error("CUDA error %s after phase '%s'", cudaGetErrorString(e), what); fprintf(stderr, "[sz] CUDA error %s after phase '%s'\n", cudaGetErrorString(e), what); return SZ_ERR_CUDA; }
    return SZ_OK;
}
static int read_counters(SzContext* c)
{
    CK(cudaMemcpyAsync(c->h_cnt, c->d_cnt, sizeof(Counters), cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
#define D_CNT(field) ((int*)((char*)c->d_cnt + offsetof(Counters, field)))

// runs the narrow phase over the pairs (wall = 0) or over floe-vs-wall (wall = 1), escalating S -> M -> L
static int run_narrow(SzContext* c, int wall, int n_work, bool fast)
{
    cudaStream_t st = c->stream;
    NarrowArgs a; memset(&a, 0, sizeof(a));
    a.ex = c->ex.p; a.ey = c->ey.p; a.esrc = c->esrc.p; a.econvex = c->econvex.p; a.erot = c->erot.p; a.eno = c->eno.p;
    a.h = c->h.p; a.area = c->area.p; a.u = c->u.p; a.v = c->v.p; a.ksi = c->ksi.p; a.voff = c->voff.p; a.vx = c->vx.p; a.vy = c->vy.p;
    a.pi = c->pi.p; a.pj = c->pj.p; a.n_work = n_work;
    a.row_pool = c->row_pool.p; a.row_cap = (int)std::min<size_t>(c->row_pool.cap / 5, 0x7fffffff); a.row_used = D_CNT(row_used);
    a.P = c->dprm;
    int* cntT; int* lstT; int* cntM; int* cntL; int* lstM; int* lstL;
    if (wall) {
        a.wall = 1; a.first_floe = 0; a.egid = c->egid.p; a.eowned = c->eowned.p; a.bx = c->bx.p; a.by = c->by.p; a.bn = c->bn; a.bbody = c->bbody;
        a.status = c->wstatus.p; a.nrows = c->wnrows.p; a.row_start = c->wrow_start.p; a.ovl_state = c->wovl.p;
        cntT = D_CNT(wlistT); lstT = c->wlistT.p; cntM = D_CNT(wlistM); cntL = D_CNT(wlistL); lstM = c->wlistM.p; lstL = c->wlistL.p;
    } else {
        a.status = c->pstatus.p; a.nrows = c->pnrows.p; a.row_start = c->prow_start.p; a.ovl_state = c->povl.p;
        a.want_polys = c->prm.want_clip_polys;
        a.poly_path_start = c->poly_path_start.p; a.poly_npaths = c->poly_npaths.p; a.path_vstart = c->path_vstart.p; a.path_len = c->path_len.p;
        a.path_cap = (int)c->path_vstart.cap; a.path_used = D_CNT(path_used); a.pvx = c->pvx.p; a.pvy = c->pvy.p; a.vert_cap = (int)c->pvx.cap; a.vert_used = D_CNT(vert_used);
        cntT = D_CNT(listT); lstT = c->listT.p; cntM = D_CNT(listM); cntL = D_CNT(listL); lstM = c->listM.p; lstL = c->listL.p;
    }
    a.next_list = lstT; a.next_count = cntT;
    if (!wall) {
        // work list of class S: bounding-box-disjoint pairs answered, the rest bucketed by vertex counts and direction
        static const int key_mode = getenv("SZ_CVX_KEY") ? atoi(getenv("SZ_CVX_KEY")) : 1;   // 0: direction sectors for class C too (experiments)
        CK(c->bins.ensure(2 * SZ_NBINS)); CK(c->bin_fill.ensure(2 * SZ_NBINS));
        CK(cudaMemsetAsync(c->bins.p, 0, 2 * SZ_NBINS * sizeof(int), st)); CK(cudaMemsetAsync(c->bin_fill.p, 0, 2 * SZ_NBINS * sizeof(int), st));
        if (c->opt_apart) { ++g_launches; pair_apart_kernel<<<std::min(nblk(n_work, 128), 148 * 16), 128, 0, st>>>(n_work, D_CNT(n_pairs), c->pi.p, c->pj.p, (const EntryRec*)c->erec.p, c->vx.p, c->vy.p, c->papart.p, c->d_cnt); }
        pair_classify_kernel<<<nblk((i64)n_work * CLS_G, 256), 256, 0, st>>>(n_work, D_CNT(n_pairs), c->pi.p, c->pj.p, (const EntryRec*)c->erec.p, c->vx.p, c->vy.p, c->prm.want_clip_polys,
                                                            c->pstatus.p, c->pnrows.p, c->povl.p, c->poly_npaths.p, c->bins.p, c->pkey.p, c->d_cnt, key_mode, c->opt_apart ? c->papart.p : nullptr);
        bins_scan_kernel<<<2, SZ_BIN_N * SZ_BIN_N, 0, st>>>(c->d_cnt, n_work, c->bins.p, c->bin_fill.p);
        pair_scatter_kernel<<<nblk(n_work, 256), 256, 0, st>>>(n_work, D_CNT(n_pairs), c->pkey.p, c->bins.p, c->bin_fill.p, c->listC.p, c->listS.p, c->d_cnt);
        g_launches += 3;
        CKS(dbg_sync(c, "pair classify + scatter"));
        // class C: strictly convex pairs through the four-edge sweep; what it declines is appended to class S's list.
        // Both launches are sized for all pairs (the list lengths are only known on the device; surplus CTAs exit at once).
        static const bool env_no_fast = getenv("SZ_NO_CONVEX_FAST") != nullptr;      // experiment switch: everything through class S
        const bool no_fast = env_no_fast || !c->opt_convex_fast;
        a.list = c->listC.p; a.list_count = D_CNT(listC); a.next_list = c->listS.p; a.next_count = D_CNT(listS);
        CK(cudaMemcpyAsync(D_CNT(listC0), D_CNT(listC), 4, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(D_CNT(listS0), D_CNT(listS), 4, cudaMemcpyDeviceToDevice, st));
        if (!c->enq_mode) CK(cudaEventRecord(c->evk[0], st));
        static const bool env_split = getenv("SZ_CONVEX_SPLIT") != nullptr;
        if (!no_fast && (c->opt_convex_split || env_split)) {
            enum { HO_CAP = 16 };                       // a larger intersection polygon sends the pair to class S
            CK(c->ho_st.ensure((size_t)n_work + 1)); CK(c->ho_x.ensure((size_t)HO_CAP * n_work + 1)); CK(c->ho_y.ensure((size_t)HO_CAP * n_work + 1));
            a.ho_st = c->ho_st.p; a.ho_x = c->ho_x.p; a.ho_y = c->ho_y.p; a.ho_stride = n_work; a.ho_cap = HO_CAP;
            g_launches += 2; sz_launch_narrow_C_split(&a, st); CK(cudaGetLastError());
        } else
        if (!no_fast) { ++g_launches; sz_launch_narrow_C(&a, st); CK(cudaGetLastError()); }
        else { a.list = c->listC.p; a.next_list = lstT; a.next_count = cntT; ++g_launches; sz_launch_narrow_S(&a, st); CK(cudaGetLastError()); }
        if (!c->enq_mode) CK(cudaEventRecord(c->evk[1], st)); c->evk_used[0] = !c->enq_mode;
        a.list = c->listS.p; a.list_count = D_CNT(listS); a.next_list = lstT; a.next_count = cntT;
    }
    if (!wall) if (!c->enq_mode) CK(cudaEventRecord(c->evk[2], st));
    ++g_launches; sz_launch_narrow_S(&a, st);
    if (!wall) { if (!c->enq_mode) CK(cudaEventRecord(c->evk[3], st)); c->evk_used[1] = !c->enq_mode; }
    a.list = nullptr; a.list_count = nullptr;
    CK(cudaGetLastError());
    // speculative step: the lists of the larger size classes were empty last step; whether they still are is checked with the
    // step's single counter read at the end (a non-empty list repeats the step on the synchronous path)
    if (fast) return SZ_OK;
    CKS(read_counters(c));
    if (!wall) { c->class_pairs[0] = c->h_cnt->listC0; c->class_pairs[1] = c->h_cnt->listS; }
    const int nT = wall ? c->h_cnt->wlistT : c->h_cnt->listT;
    if (nT > 0) {
        // class T: pairs that did not fit class S, arena still in local memory
        a.list = lstT; a.list_count = cntT; a.n_work = nT; a.next_list = lstM; a.next_count = cntM;
        if (!wall) { if (!c->enq_mode) CK(cudaEventRecord(c->evk[4], st)); c->class_pairs[2] = nT; }
        ++g_launches; sz_launch_narrow_T(&a, st);
        if (!wall) { if (!c->enq_mode) CK(cudaEventRecord(c->evk[5], st)); c->evk_used[2] = !c->enq_mode; }
        a.list = nullptr; a.list_count = nullptr;
        CK(cudaGetLastError());
        CKS(read_counters(c));
    }
    static const bool dbg = getenv("SZ_DEBUG_TIMING") != nullptr;
    cudaEvent_t d0, d1; float dms = 0;
    if (dbg) { cudaEventCreate(&d0); cudaEventCreate(&d1); }
    int nM = wall ? c->h_cnt->wlistM : c->h_cnt->listM;
    if (dbg) fprintf(stderr, "[sz] class S -> T %d, T -> M %d\n", nT, nM);
    if (nM > 0) {
        static const int m_tpsm = getenv("SZ_M_TPSM") ? atoi(getenv("SZ_M_TPSM")) : 768;   // persistent threads per SM of class M (160 KB of HBM scratch each: 18 GB at most)
        const int threads = std::min((nM + 63) / 64 * 64, 148 * m_tpsm);
        CK(c->scratchM.ensure((size_t)threads * sz_workspace_bytes_M()));
        a.list = lstM; a.list_count = cntM; a.next_list = lstL; a.next_count = cntL; a.scratch = c->scratchM.p; a.n_threads = threads;
        if (dbg) cudaEventRecord(d0, st);
        if (!wall) { if (!c->enq_mode) CK(cudaEventRecord(c->evk[6], st)); c->class_pairs[3] = nM; }
        ++g_launches; sz_launch_narrow_M(&a, st);
        CK(cudaGetLastError());
        if (!wall) { if (!c->enq_mode) CK(cudaEventRecord(c->evk[7], st)); c->evk_used[3] = !c->enq_mode; }
        if (dbg) cudaEventRecord(d1, st);
        CKS(read_counters(c));
        if (dbg) { cudaEventElapsedTime(&dms, d0, d1); fprintf(stderr, "[sz] class M: %d pairs on %d threads, %.2f ms\n", nM, threads, dms); }
        int nL = wall ? c->h_cnt->wlistL : c->h_cnt->listL;
        if (nL > 0) {
            static const int l_tpsm = getenv("SZ_L_TPSM") ? atoi(getenv("SZ_L_TPSM")) : 160;   // class L (1.1 MB of HBM scratch each: 26 GB at most; a thread sweeps one pair in ~0.3 s, so the launch lasts as many rounds as
                                                                                                 // pairs / threads: 512 + 128 -> 768 + 160 took the 57,600-floe raw field from 745 to 605 ms per step, profiles/r02i_lm_threads.txt)
            const int threadsL = std::min((nL + 63) / 64 * 64, 148 * l_tpsm);
            CK(c->scratchL.ensure((size_t)threadsL * sz_workspace_bytes_L()));
            a.list = lstL; a.list_count = cntL; a.next_list = nullptr; a.next_count = nullptr; a.scratch = c->scratchL.p; a.n_threads = threadsL;
            if (dbg) cudaEventRecord(d0, st);
            if (!wall) { if (!c->enq_mode) CK(cudaEventRecord(c->evk[8], st)); c->class_pairs[4] = nL; }
            ++g_launches; sz_launch_narrow_L(&a, st);
            CK(cudaGetLastError());
            if (!wall) { if (!c->enq_mode) CK(cudaEventRecord(c->evk[9], st)); c->evk_used[4] = !c->enq_mode; }
            if (dbg) cudaEventRecord(d1, st);
            CKS(read_counters(c));
            if (dbg) { cudaEventElapsedTime(&dms, d0, d1); fprintf(stderr, "[sz] class L: %d pairs on %d threads, %.2f ms\n", nL, threadsL, dms); }
        }
    }
    return SZ_OK;
}

// the cell grid of the broad phase from the list's bounding box and largest rmax: about eight floes per cell, never more than 8
// cells per reach (2 max(rmax)); every floe widens its own search by its radius, and any grid is CORRECT for any list (entries
// outside the box are clamped into the border cells, which only moves them towards their partners), so a step may use the grid
// of the step before
static GridDesc make_grid(const Counters* h, int n)
{
    GridDesc g; g.x0 = g.y0 = 0; g.cell = 1; g.nx = g.ny = 1;
    const double rm = dec_d(h->rmax_bits);
    const double xmn = dec_d(h->bbox[0]), xmx = dec_d(h->bbox[1]), ymn = dec_d(h->bbox[2]), ymx = dec_d(h->bbox[3]);
    if (xmn <= xmx && std::isfinite(xmn) && std::isfinite(xmx) && std::isfinite(ymn) && std::isfinite(ymx)) {
        g.x0 = xmn; g.y0 = ymn;
        double cell = 2 * rm; if (!(cell > 0) || !std::isfinite(cell)) cell = 1;
        // floes per cell: the broad phase reads whole cell rows, 32 candidates per warp iteration, over the cells that overlap the
        // floe's reach; measured at 1M floes: 4 / 6 / 8 / 12 per cell -> broad phase 1.59 / 1.54 / 1.51 / 1.59 ms (switch SZ_GRID_FLOES_PER_CELL)
        static const double per_cell = getenv("SZ_GRID_FLOES_PER_CELL") ? std::max(0.25, atof(getenv("SZ_GRID_FLOES_PER_CELL"))) : 8.0;
        const double dens = std::sqrt(per_cell * (xmx - xmn) * (ymx - ymn) / std::max(1, n));
        if (dens > 0 && std::isfinite(dens)) cell = std::min(cell, std::max(cell / 8, dens));
        while ((xmx - xmn) / cell * ((ymx - ymn) / cell) > 1.6e7) cell *= 2;      // keep the grid below ~16M cells
        g.cell = cell;
        g.nx = (int)((xmx - xmn) / cell) + 1; g.ny = (int)((ymx - ymn) / cell) + 1;
    }
    return g;
}

static int step_finish(SzContext* c, SzSummary* out, int n, int np, i64 rows_bound, bool fast, bool enqueued, float ms);
// mode 0: the whole step; mode 1: enqueue only (speculated sizes required), the counters travel to the host asynchronously
// and sz_step_finish completes the step
static int step_impl(SzContext* c, SzSummary* out, int mode)
{
    if (!c) { sz_set_error("sz_step_resident: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_input) { sz_set_error("sz_step_resident: no floes uploaded"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const bool enq = mode == 1;
    c->pend.active = false; c->enq_mode = enq;
    cudaStream_t st = c->stream;
    const SzParams& P = c->prm;
    const int n0 = c->n0, Nb = P.Nb;
    const bool ext = c->ext_mode, slab = c->slab;
    if (slab && !c->sl_built) { sz_set_error("sz_step_resident: slab mode needs sz_slab_build before every step"); return SZ_ERR_STATE; }
    const int nl0 = slab ? c->sl_nl_cap : n0;                    // length of a caller-supplied / device-built extended list (its tail may be inert)
    const int ncap = (P.periodic && !ext) ? 4 * n0 : nl0;        // every floe has at most an x-, a y- and an xy-ghost
    c->have_step = false; c->have_rows = false;
    const bool fast = c->opt_speculate && c->plan_valid && c->plan_n0 == n0 && c->plan_nl0 == nl0 && !P.want_clip_polys && c->plan_ncap <= ncap;
    if (enq && !fast) { sz_set_error("sz_step_enqueue: no sizes to carry over yet (run sz_step_resident first; not with want_clip_polys)"); return SZ_ERR_STATE; }
    // NVTX ranges name the phases on a timeline (Nsight Systems; free when no tool is attached)
    struct Nvtx { Nvtx(const char* n) { nvtxRangePushA(n); } ~Nvtx() { nvtxRangePop(); } void next(const char* n) { nvtxRangePop(); nvtxRangePushA(n); } } nvtx("sz K0 extended list");
    if (!enq) CK(cudaEventRecord(c->ev0, st));
    CK(cudaMemsetAsync(c->d_cnt, 0, sizeof(Counters), st));

    // ---- K0: extended list
    CK(c->ex.ensure(ncap + 1)); CK(c->ey.ensure(ncap + 1)); CK(c->esrc.ensure(ncap + 1)); CK(c->efn.ensure(ncap + 1)); CK(c->eparent.ensure(ncap + 1));
    CK(c->ealive.ensure(ncap + 1)); CK(c->gx_of.ensure(n0 + 1)); CK(c->gy_of.ensure(n0 + 1));
    CK(c->egid.ensure(ncap + 1)); CK(c->eowned.ensure(ncap + 1)); CK(c->erootx.ensure(ncap + 1)); CK(c->erooty.ensure(ncap + 1));
    CK(c->flag.ensure(2 * (size_t)n0 + 2)); CK(c->pos.ensure(2 * (size_t)n0 + 2));
    if (n0 > 0 && !ext) { ++g_launches; init_extended_kernel<<<nblk(n0, 256), 256, 0, st>>>(n0, c->x.p, c->y.p, c->alive.p, c->ex.p, c->ey.p, c->esrc.p, c->efn.p, c->eparent.p, c->ealive.p, c->gx_of.p, c->gy_of.p); }
    {
        Counters init; memset(&init, 0, sizeof(init));
        init.n1 = nl0; init.n = nl0;
        init.bbox[0] = enc_d(SZ_INF); init.bbox[1] = enc_d(-SZ_INF); init.bbox[2] = enc_d(SZ_INF); init.bbox[3] = enc_d(-SZ_INF); init.rmax_bits = enc_d(0.0);
        *c->h_cnt = init; *c->h_init = init;
        CK(cudaMemcpyAsync(c->d_cnt, c->h_init, sizeof(Counters), cudaMemcpyDefault, st));
    }
    if (P.periodic && n0 > 0 && !ext) {
        CK(c->scan_tmp.ensure(scan_tmp_ints(2 * (size_t)n0 + 2)));
        // x pass over the originals
        ++g_launches; ghost_flag_kernel<<<nblk(n0, 128), 128, 0, st>>>(0, n0, nullptr, c->ex.p, c->esrc.p, c->ealive.p, c->voff.p, c->vx.p, P.Lx, c->flag.p);
        CKS(exclusive_scan(c, c->flag.p, n0, c->pos.p, n0 + 1));
        ++g_launches; ghost_emit_kernel<<<nblk(n0, 256), 256, 0, st>>>(0, n0, nullptr, n0, c->flag.p, c->pos.p, c->ex.p, c->ey.p, c->esrc.p, c->efn.p, c->eparent.p, c->ealive.p,
                                                         c->gx_of.p, c->gy_of.p, P.Lx, D_CNT(n1));
        // y pass over originals + x-ghosts (their number is only known on the device: bound 2*n0)
        ++g_launches; ghost_flag_kernel<<<nblk(2 * (i64)n0, 128), 128, 0, st>>>(1, 2 * n0, D_CNT(n1), c->ey.p, c->esrc.p, c->ealive.p, c->voff.p, c->vy.p, P.Ly, c->flag.p);
        CKS(exclusive_scan(c, c->flag.p, 2 * n0, c->pos.p, 2 * n0 + 1));
        ++g_launches; ghost_emit_kernel<<<nblk(2 * (i64)n0, 256), 256, 0, st>>>(1, 2 * n0, D_CNT(n1), n0, c->flag.p, c->pos.p, c->ex.p, c->ey.p, c->esrc.p, c->efn.p, c->eparent.p, c->ealive.p,
                                                                  c->gx_of.p, c->gy_of.p, P.Ly, D_CNT(n));
    }
    if (ext && nl0 > 0) {
        // the caller supplied the extended list (sz_upload_extended): every entry has its own outline and body record.
        // Slab mode (sz_slab_build) wrote the list itself, with esrc pointing into [owned floes | halo records].
        if (!slab) {
            CK(cudaMemcpyAsync(c->ex.p, c->x.p, (size_t)n0 * 8, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(c->ey.p, c->y.p, (size_t)n0 * 8, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(c->ealive.p, c->alive.p, (size_t)n0, cudaMemcpyDeviceToDevice, st));
        }
        if (!slab) {      // (sz_slab_build wrote the images of every entry itself)
            CK(c->gx_of.ensure(nl0 + 1)); CK(c->gy_of.ensure(nl0 + 1));
            ++g_launches; child_init_kernel<<<nblk(nl0, 256), 256, 0, st>>>(nl0, c->gx_of.p, c->gy_of.p);
            ++g_launches; child_mark_kernel<<<nblk(nl0, 256), 256, 0, st>>>(nl0, c->eparent.p, c->gx_of.p, c->gy_of.p);
            ++g_launches; child_final_kernel<<<nblk(nl0, 256), 256, 0, st>>>(nl0, c->gx_of.p, c->gy_of.p);
        }
    } else if (ncap > 0) {
        ++g_launches; finish_extended_kernel<<<nblk(ncap, 256), 256, 0, st>>>(ncap, D_CNT(n), c->x.p, c->y.p, c->esrc.p, c->egid.p, c->eowned.p, c->erootx.p, c->erooty.p);
    }
    if (ncap > 0) { ++g_launches; bbox_kernel<<<std::min(nblk(ncap, 256), 148 * 8), 256, 0, st>>>(ncap, D_CNT(n), c->ex.p, c->ey.p, c->esrc.p, c->rmax.p, c->d_cnt); }
    CK(cudaGetLastError());
    // Speculative step: list length, grid, pair and row capacities come from the previous step (with slack), every kernel
    // takes the true counts from device memory, and the host reads the counters ONCE, at the end; a step whose counts
    // outgrew a capacity (or that needs a larger size class) raises a flag there and is repeated on the synchronous path.
    int n;
    if (fast) {
        n = slab ? nl0 : c->plan_ncap;
        if (!ext && n > 0) { ++g_launches; slab_build_tail_kernel<<<nblk(n, 256), 256, 0, st>>>(n, D_CNT(n), c->ex.p, c->ey.p, c->erootx.p, c->erooty.p, c->esrc.p, c->efn.p, c->eparent.p, c->egid.p, c->ealive.p, c->eowned.p, nullptr, nullptr); }
    } else {
        CKS(read_counters(c));
        n = c->h_cnt->n; c->n1 = c->h_cnt->n1;
    }
    c->n = n;

    CKS(dbg_sync(c, fast ? "K0 (speculated)" : "K0"));
    if (!enq) CK(cudaEventRecord(c->evp[0], st));
    nvtx.next("sz K1 broad phase");
    // ---- K1: cell grid + candidate pairs
    GridDesc g;
    if (fast) { g.x0 = c->plan_g.x0; g.y0 = c->plan_g.y0; g.cell = c->plan_g.cell; g.nx = c->plan_g.nx; g.ny = c->plan_g.ny; }
    else g = make_grid(c->h_cnt, n);
    const int ncell = g.nx * g.ny;
    if (g.nx < 1 || g.ny < 1 || (double)g.nx * (double)g.ny > 3.2e7) {
        sz_set_error("sz_step_resident: bad cell grid %d x %d (cell %g, origin %g %g; %s path; bbox %g..%g x %g..%g, rmax %g, n %d)", g.nx, g.ny, g.cell, g.x0, g.y0, fast ? "speculated" : "measured",
                     dec_d(c->h_cnt->bbox[0]), dec_d(c->h_cnt->bbox[1]), dec_d(c->h_cnt->bbox[2]), dec_d(c->h_cnt->bbox[3]), dec_d(c->h_cnt->rmax_bits), n);
        return SZ_ERR_STATE;
    }
    CK(c->cid.ensure(n + 1)); CK(c->cell_cnt.ensure(ncell + 2)); CK(c->cell_start.ensure(ncell + 2));
    CK(c->s_idx.ensure(n + 1)); CK(c->s_x.ensure(n + 1)); CK(c->s_y.ensure(n + 1)); CK(c->s_r.ensure(n + 1));
    CK(c->pcnt.ensure(n + 2)); CK(c->pair_off.ensure(n + 2));
    CK(c->scan_tmp.ensure(scan_tmp_ints(std::max<size_t>({(size_t)ncell + 2, (size_t)n + 2, 2 * (size_t)n0 + 2}))));
    CK(cudaMemsetAsync(c->cell_cnt.p, 0, (size_t)(ncell + 1) * 4, st));
    CK(cudaMemsetAsync(c->pcnt.p, 0, (size_t)(n + 1) * 4, st));
    BroadArgs b; memset(&b, 0, sizeof(b));
    if (n > 0) {
        ++g_launches; cell_count_kernel<<<nblk(n, 256), 256, 0, st>>>(n, g, c->ex.p, c->ey.p, c->ealive.p, c->cid.p, c->cell_cnt.p);
        CKS(exclusive_scan(c, c->cell_cnt.p, ncell, c->cell_start.p, ncell + 1));
        CK(cudaMemsetAsync(c->cell_cnt.p, 0, (size_t)(ncell + 1) * 4, st));
        ++g_launches; cell_fill_kernel<<<nblk(n, 256), 256, 0, st>>>(n, c->cid.p, c->cell_start.p, c->cell_cnt.p, c->ex.p, c->ey.p, c->esrc.p, c->rmax.p, c->s_idx.p, c->s_x.p, c->s_y.p, c->s_r.p);
        b.n = n; b.n0 = n0; b.Nb = Nb; b.collision = P.collision; b.g = g; b.minL2 = std::min(2 * P.Lx, 2 * P.Ly); b.rmax_bits = (const u64*)((const char*)c->d_cnt + offsetof(Counters, rmax_bits)); b.overflow = D_CNT(overflow); b.pair_boundary = P.pair_with_boundary_floes;
        b.ex = c->ex.p; b.ey = c->ey.p; b.esrc = c->esrc.p; b.efn = c->efn.p; b.ealive = c->ealive.p; b.rmax = c->rmax.p;
        b.egid = c->egid.p; b.eowned = c->eowned.p; b.erootx = c->erootx.p; b.erooty = c->erooty.p;
        b.cell_start = c->cell_start.p; b.s_idx = c->s_idx.p; b.s_x = c->s_x.p; b.s_y = c->s_y.p; b.s_r = c->s_r.p;
        b.pcnt = c->pcnt.p; b.pair_off = c->pair_off.p;
        static const int stage_cap = getenv("SZ_BROAD_STAGE") ? atoi(getenv("SZ_BROAD_STAGE")) : 16;     // 0: the fill pass searches again
        if (stage_cap > 0) { CK(c->stage.ensure((size_t)n * stage_cap + 1)); b.stage = c->stage.p; b.stage_cap = stage_cap > 32 ? 32 : stage_cap; }
        { ++g_launches; if (b.pair_boundary) broad_kernel<false, true><<<nblk(32 * (i64)n, 256), 256, 0, st>>>(b); else broad_kernel<false, false><<<nblk(32 * (i64)n, 256), 256, 0, st>>>(b); }
    }
    CKS(dbg_sync(c, "cell grid + broad count"));
    CKS(exclusive_scan(c, c->pcnt.p, n, c->pair_off.p, n + 1));
    CK(cudaMemcpyAsync(D_CNT(n_pairs), c->pair_off.p + n, 4, cudaMemcpyDeviceToDevice, st));
    // per-entry preparation of the narrow phase (bounding boxes, convexity): independent of the pair list, so it is queued
    // before the host waits for the pair count
    CK(c->erec.ensure(8 * (size_t)n + 8)); CK(c->ebb.ensure(4 * (size_t)n + 4)); CK(c->evalid.ensure(n + 1)); CK(c->econvex.ensure(n + 1)); CK(c->erot.ensure(n + 1)); CK(c->eno.ensure(n + 1)); CK(c->env.ensure(n + 1));
    if (n > 0) { ++g_launches; ext_prep_kernel<<<nblk(n, 128), 128, 0, st>>>(n, c->ex.p, c->ey.p, c->esrc.p, c->voff.p, c->vx.p, c->vy.p, c->ebb.p, c->evalid.p, c->env.p, c->econvex.p, c->erot.p, c->eno.p, (EntryRec*)c->erec.p, c->d_cnt); }
    CK(cudaGetLastError());
    int np;
    if (fast) np = c->plan_npcap;
    else { CKS(read_counters(c)); np = c->h_cnt->n_pairs; }
    c->n_pairs = np; b.np_cap = np;
    CK(c->pi.ensure(np + 1)); CK(c->pj.ensure(np + 1)); CK(c->pstatus.ensure(np + 1)); CK(c->pnrows.ensure(np + 1)); CK(c->prow_start.ensure(np + 1)); CK(c->povl.ensure(np + 1));
    CK(c->listT.ensure(np + 1)); CK(c->listM.ensure(np + 1)); CK(c->listL.ensure(np + 1));
    if (np > 0) { b.pi = c->pi.p; b.pj = c->pj.p; ++g_launches; if (b.pair_boundary) broad_kernel<true, true><<<nblk(32 * (i64)n, 256), 256, 0, st>>>(b); else broad_kernel<true, false><<<nblk(32 * (i64)n, 256), 256, 0, st>>>(b); }

    CKS(dbg_sync(c, "ext_prep + broad fill"));
    if (!enq) CK(cudaEventRecord(c->evp[1], st));
    nvtx.next("sz K2+K3 narrow phase and force law");
    CK(c->listC.ensure(np + 1)); CK(c->listS.ensure(np + 1)); CK(c->pkey.ensure(np + 1)); CK(c->papart.ensure(np + 1));
    // ---- K2 + K3: narrow phase (pool capacities are guesses; exact needs come back in the counters)
    const bool wall = c->have_bnd && !P.periodic;
    if (wall) { CK(c->wstatus.ensure(n + 1)); CK(c->wnrows.ensure(n + 1)); CK(c->wrow_start.ensure(n + 1)); CK(c->wovl.ensure(n + 1)); CK(c->wlistT.ensure(n + 1)); CK(c->wlistM.ensure(n + 1)); CK(c->wlistL.ensure(n + 1)); }
    CK(c->row_pool.ensure(5 * ((size_t)np + (wall ? n : 0) + 256)));
    if (P.want_clip_polys) {
        CK(c->poly_path_start.ensure(np + 1)); CK(c->poly_npaths.ensure(np + 1));
        CK(c->path_vstart.ensure((size_t)np + 256)); CK(c->path_len.ensure(c->path_vstart.cap)); CK(c->pvx.ensure(8 * (size_t)np + 1024)); CK(c->pvy.ensure(c->pvx.cap));
    }
    for (int attempt = 0; attempt < 3; ++attempt) {
        if (!fast) {
            // (the host copy of the counters is current here: the pair count was just read)
            Counters z = *c->h_cnt;
            z.row_used = z.path_used = z.vert_used = z.n_bbox_reject = z.listC = z.listS = z.listC0 = z.listS0 = z.listT = z.listM = z.listL = z.wlistT = z.wlistM = z.wlistL = 0;
            *c->h_cnt = z;
            CK(cudaMemcpyAsync(c->d_cnt, c->h_cnt, sizeof(Counters), cudaMemcpyDefault, st));
        }
        CK(cudaMemsetAsync(c->pstatus.p, 0, (size_t)(np + 1) * 4, st)); CK(cudaMemsetAsync(c->pnrows.p, 0, (size_t)(np + 1) * 4, st));
        for (int k = 0; k < 5; ++k) { c->evk_used[k] = false; c->class_pairs[k] = 0; }
        if (np > 0) CKS(run_narrow(c, 0, np, fast));
        if (wall) {
            // floes below Nb take no part in the wall call (:125 loops i = 1+Nb:N)
            CK(cudaMemsetAsync(c->wstatus.p, 0, (size_t)(n + 1) * 4, st)); CK(cudaMemsetAsync(c->wnrows.p, 0, (size_t)(n + 1) * 4, st));
            if (n > 0) CKS(run_narrow(c, 1, n, fast));
        }
        if (fast) break;          // capacities are checked with the counters at the end of the step
        // run_narrow ends with a counter read-back and nothing was launched since: the host copy is current
        bool again = false;
        if ((size_t)c->h_cnt->row_used * 5 > c->row_pool.cap) { CK(c->row_pool.ensure((size_t)c->h_cnt->row_used * 5 + 1024)); again = true; }
        if (P.want_clip_polys) {
            if ((size_t)c->h_cnt->path_used > c->path_vstart.cap) { CK(c->path_vstart.ensure(c->h_cnt->path_used + 64)); CK(c->path_len.ensure(c->path_vstart.cap)); again = true; }
            if ((size_t)c->h_cnt->vert_used > c->pvx.cap) { CK(c->pvx.ensure(c->h_cnt->vert_used + 64)); CK(c->pvy.ensure(c->pvx.cap)); again = true; }
        }
        if (!again) break;
        if (attempt == 2) { sz_set_error("sz_step_resident: result pools kept overflowing"); return SZ_ERR_CAPACITY; }
    }

    CKS(dbg_sync(c, "narrow phase"));
    if (!enq) CK(cudaEventRecord(c->evp[2], st));
    nvtx.next("sz K4 mirror, rows, sums");
    // ---- K4: mirror, rows, sums
    CK(c->tcnt.ensure(n + 2)); CK(c->toff.ensure(n + 2)); CK(c->tlist.ensure(np + 1)); CK(c->rcnt.ensure(n + 2)); CK(c->row_off.ensure(n + 2));
    CK(cudaMemsetAsync(c->tcnt.p, 0, (size_t)(n + 1) * 4, st));
    if (np > 0) { ++g_launches; tcount_kernel<<<nblk(np, 256), 256, 0, st>>>(np, D_CNT(n_pairs), c->pi.p, c->pj.p, c->pnrows.p, c->tcnt.p); }
    CKS(exclusive_scan(c, c->tcnt.p, n, c->toff.p, n + 1));
    CK(cudaMemsetAsync(c->tcnt.p, 0, (size_t)(n + 1) * 4, st));
    if (np > 0) { ++g_launches; tfill_kernel<<<nblk(np, 256), 256, 0, st>>>(np, D_CNT(n_pairs), c->pi.p, c->pj.p, c->pnrows.p, c->toff.p, c->tcnt.p, c->tlist.p); }
    if (n > 0) { ++g_launches; rowcount_kernel<<<nblk(n, 128), 128, 0, st>>>(n, c->eowned.p, c->pair_off.p, c->pnrows.p, wall ? c->wnrows.p : nullptr, c->toff.p, c->tlist.p, c->rcnt.p); }
    CKS(exclusive_scan(c, c->rcnt.p, n, c->row_off.p, n + 1));
    CKS(dbg_sync(c, "tcount/tfill/rowcount"));
    CK(cudaMemcpyAsync(D_CNT(total_rows), c->row_off.p + n, 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
    // every row of the pool appears at most twice (its floe's own row and the partner's mirrored row): the row count is
    // bounded without waiting for it; the exact value comes back with the step's last counter read
    const i64 rows_bound = fast ? (i64)c->plan_rowscap : 2 * (i64)c->h_cnt->row_used;
    CK(c->rows.ensure((size_t)rows_bound * 7 + 7)); CK(c->osum.ensure((size_t)n * 3 + 3)); CK(c->e_ov.ensure(n + 1)); CK(c->has_rows.ensure(n + 1));
    const int nout = ext ? n : n0; c->nout = nout;
    CK(c->kill_i.ensure(n + 1)); CK(c->transfer_i.ensure(n + 1)); CK(c->tmax.ensure(nout + 1));
    CK(c->o_fx.ensure(nout + 1)); CK(c->o_fy.ensure(nout + 1)); CK(c->o_tq.ensure(nout + 1)); CK(c->o_ov.ensure(nout + 1)); CK(c->o_stress.ensure(4 * (size_t)nout + 4));
    CK(c->o_xi.ensure(nout + 1)); CK(c->o_yi.ensure(nout + 1)); CK(c->o_alive.ensure(nout + 1)); CK(c->o_kill.ensure(nout + 1)); CK(c->o_transfer.ensure(nout + 1));
    if (n > 0) {
        AssembleArgs a; memset(&a, 0, sizeof(a));
        a.n = n; a.nout = nout; a.Nb = Nb; a.periodic = P.periodic; a.wall = wall; a.Lx = P.Lx; a.Ly = P.Ly;
        a.egid = c->egid.p; a.efn = c->efn.p; a.eowned = c->eowned.p;
        a.ex = c->ex.p; a.ey = c->ey.p; a.esrc = c->esrc.p; a.ealive = c->ealive.p; a.area = c->area.p; a.h = c->h.p;
        a.pair_off = c->pair_off.p; a.pi = c->pi.p; a.pj = c->pj.p; a.nrows = c->pnrows.p; a.row_start = c->prow_start.p; a.ovl = c->povl.p; a.pstatus = c->pstatus.p;
        a.wnrows = c->wnrows.p; a.wrow_start = c->wrow_start.p; a.wstatus = c->wstatus.p;
        a.toff = c->toff.p; a.tlist = c->tlist.p; a.row_off = c->row_off.p; a.pool = c->row_pool.p;
        a.boxx = c->boxx.p; a.boxy = c->boxy.p; a.boxn = c->boxn;
        a.rows = c->rows.p; a.rows_cap = rows_bound; a.osum = c->osum.p; a.e_ov = c->e_ov.p; a.has_rows = c->has_rows.p; a.kill_i = c->kill_i.p; a.transfer_i = c->transfer_i.p;
        a.o_ov = c->o_ov.p; a.o_stress = c->o_stress.p; a.o_xi = c->o_xi.p; a.o_yi = c->o_yi.p; a.o_alive = c->o_alive.p; a.cnt = c->d_cnt;
        ++g_launches; assemble_kernel<<<nblk(n, 128), 128, 0, st>>>(a);
        CKS(dbg_sync(c, "assemble_kernel"));
        if (!ext) {
            CK(cudaMemsetAsync(c->tmax.p, 0, (size_t)(nout + 1) * 4, st));
            ++g_launches; kill_mark_kernel<<<nblk(n, 256), 256, 0, st>>>(n, c->kill_i.p, c->tmax.p);
            if (nout > 0) { ++g_launches; kill_final_kernel<<<nblk(nout, 256), 256, 0, st>>>(nout, c->kill_i.p, c->transfer_i.p, c->tmax.p, c->o_kill.p, c->o_transfer.p); }
        } else {
            // extended mode: raw per-entry kill/transfer (global ids); the cross-rank fix-up of :175-179 is the caller's
            CK(cudaMemcpyAsync(c->o_kill.p, c->kill_i.p, (size_t)nout * 4, cudaMemcpyDeviceToDevice, st)); CK(cudaMemcpyAsync(c->o_transfer.p, c->transfer_i.p, (size_t)nout * 4, cudaMemcpyDeviceToDevice, st));
        }
        CKS(dbg_sync(c, "kill fix-up"));
        if (nout > 0) { ++g_launches; fold_kernel<<<nblk(nout, 256), 256, 0, st>>>(nout, n, Nb, c->egid.p, c->gx_of.p, c->gy_of.p, c->osum.p, c->has_rows.p, c->o_fx.p, c->o_fy.p, c->o_tq.p); }
        CKS(dbg_sync(c, "fold"));
        if (np > 0) { ++g_launches; pair_stats_kernel<<<std::min(nblk(np, 256), 148 * 8), 256, 0, st>>>(np, D_CNT(n_pairs), c->pstatus.p, c->pnrows.p, c->pi.p, c->eowned.p, 1, c->d_cnt); }
        if (wall) { ++g_launches; pair_stats_kernel<<<std::min(nblk(n, 256), 148 * 8), 256, 0, st>>>(n, nullptr, c->wstatus.p, c->wnrows.p, nullptr, c->eowned.p, 0, c->d_cnt); }
    }
    CKS(dbg_sync(c, "assembly"));
    CK(cudaGetLastError());
    if (enq) {
        CK(cudaMemcpyAsync(c->h_cnt, c->d_cnt, sizeof(Counters), cudaMemcpyDefault, st));
        c->pend.active = true; c->pend.n = n; c->pend.np = np; c->pend.rows_bound = rows_bound;
        return SZ_OK;
    }
    CK(cudaEventRecord(c->ev1, st));
    CKS(read_counters(c));
    float ms = 0; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    CK(cudaEventElapsedTime(&c->phase_ms[0], c->ev0, c->evp[0])); CK(cudaEventElapsedTime(&c->phase_ms[1], c->evp[0], c->evp[1]));
    CK(cudaEventElapsedTime(&c->phase_ms[2], c->evp[1], c->evp[2])); CK(cudaEventElapsedTime(&c->phase_ms[3], c->evp[2], c->ev1)); c->phase_ms[4] = ms;
    return step_finish(c, out, n, np, rows_bound, fast, false, ms);
}
// the host half of a step: the counters are in h_cnt; validates a speculated step, records what the next one may assume, fills the summary.
// Returns 1 when an ENQUEUED step has to be repeated (the caller runs sz_step_resident on the unchanged state).
static int step_finish(SzContext* c, SzSummary* out, int n, int np, i64 rows_bound, bool fast, bool enqueued, float ms)
{
    const SzParams& P = c->prm;
    const int n0 = c->n0; const bool ext = c->ext_mode, slab = c->slab;
    const int nl0 = slab ? c->sl_nl_cap : n0;
    const int ncap = (P.periodic && !ext) ? 4 * n0 : nl0;
    {
        const Counters& H = *c->h_cnt;
        const bool big_lists = H.listT || H.listM || H.listL || H.wlistT || H.wlistM || H.wlistL;
        if (fast) {
            const bool bad = H.overflow || H.n > n || H.n_pairs > np || (size_t)H.row_used * 5 > c->row_pool.cap || (i64)H.total_rows > rows_bound || big_lists;
            if (bad) { c->plan_valid = false; ++c->n_slow_steps; if (enqueued) return 1; return step_impl(c, out, 0); }      // nothing was consumed: the same step again, with measured sizes
            ++c->n_fast_steps;
            c->class_pairs[0] = H.listC0; c->class_pairs[1] = H.listS;
            if (!slab) c->n = H.n;
            c->n1 = H.n1; c->n_pairs = H.n_pairs;
        }
        // what the next step may assume
        const int n_act = slab ? nl0 : H.n;
        c->plan_n0 = n0; c->plan_nl0 = nl0;
        c->plan_ncap = slab ? nl0 : std::min(ncap, n_act + std::max(4096, n_act / 32));
        const GridDesc pg = make_grid(c->h_cnt, std::max(1, slab ? H.n : n_act));
        c->plan_g.x0 = pg.x0; c->plan_g.y0 = pg.y0; c->plan_g.cell = pg.cell; c->plan_g.nx = pg.nx; c->plan_g.ny = pg.ny;
        c->plan_npcap = H.n_pairs + H.n_pairs / 16 + 8192;
        c->plan_rowscap = (i64)H.total_rows + H.total_rows / 8 + 8192;
        c->plan_valid = !big_lists && H.n_fail == 0 && H.n_cap_fail == 0;
    }
    const int n_list = c->n, np_act = c->n_pairs;
    const i64 nrows = c->h_cnt->total_rows; c->n_rows = nrows;
    if (nrows > rows_bound) { sz_set_error("sz_step_resident: %lld contact rows exceed the bound %lld", (long long)nrows, (long long)rows_bound); return SZ_ERR_CAPACITY; }
    SzSummary& s = c->summary; memset(&s, 0, sizeof(s));
    s.n0 = n0; s.n = n_list; s.n_pairs = np_act; s.n_pairs_force = c->h_cnt->n_pairs_force; s.n_rows = nrows; s.n_pairs_owned = c->h_cnt->n_pairs_owned;
    s.n_clip_paths = P.want_clip_polys ? c->h_cnt->path_used : 0; s.n_clip_verts = P.want_clip_polys ? c->h_cnt->vert_used : 0;
    s.collision_count = (double)c->h_cnt->n_fin_rows / 2 + (double)c->h_cnt->n_inf_rows;   // calc_collisionNum.m:6
    s.n_clipper_fail = c->h_cnt->n_fail; s.n_capacity_fail = c->h_cnt->n_cap_fail; s.ms_device = ms; s.n_kill_events = c->h_cnt->n_kill_events;
    c->have_step = true; c->have_rows = true;
    if (out) *out = s;
    if (s.n_capacity_fail > 0) { sz_set_error("%d pair(s) exceed the largest narrow-phase size class (1299 vertices per outline)", s.n_capacity_fail); return SZ_ERR_CAPACITY; }
    if (s.n_clipper_fail > 0) { sz_set_error("Clipper Error. (%d pair(s); per-pair status via sz_get_pairs)", s.n_clipper_fail); return SZ_ERR_CLIPPER; }
    return SZ_OK;
}

extern "C" int sz_step_resident(SzContext* c, SzSummary* out) { return step_impl(c, out, 0); }
extern "C" int sz_step_enqueue(SzContext* c) { return step_impl(c, nullptr, 1); }
extern "C" int sz_step_finish(SzContext* c, SzSummary* out)
{
    if (!c) { sz_set_error("sz_step_finish: NULL context"); return SZ_ERR_ARG; }
    // (the enqueued launches may have been captured into a CUDA graph and replayed: the sizes they carry stay valid until an
    // ordinary step or an upload replaces them)
    if (!c->pend.active) { sz_set_error("sz_step_finish: no enqueued step"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 5; ++k) c->phase_ms[k] = 0;
    if (c->slab) c->sl_built = true;          // the step that just ran (launched or replayed) followed a list build on the same stream
    return step_finish(c, out, c->pend.n, c->pend.np, c->pend.rows_bound, true, true, 0.0f);
}
extern "C" void sz_add_launches(long long n) { g_launches += n; }
extern "C" int sz_contact_step(SzContext* c, const SzParams* prm, const SzFloesSoA* f, const SzBoundary* bnd, SzSummary* out)
{
    int r = sz_upload(c, prm, f, bnd);
    if (r != SZ_OK) return r;
    return sz_step_resident(c, out);
}

#define NEED_STEP(name) do { if (!c) { sz_set_error(name ": NULL context"); return SZ_ERR_ARG; } \
    if (!c->have_step) { sz_set_error(name ": no step has been run"); return SZ_ERR_STATE; } CK(cudaSetDevice(c->device)); } while (0)
#define D2H(dst, src, bytes) do { if ((dst) && (bytes) > 0) CK(cudaMemcpyAsync((dst), (src), (bytes), cudaMemcpyDefault, c->stream)); } while (0)
// ------------------------------------------------------------------------------------------------ trajectory
struct TrajArgs {
    int n0, nz, Nb; double dt, HFo, xo_min, xo_max, yo_min, yo_max;
    const double* cfx; const double* cfy; const double* ctq; const double* stress_now; const uint8_t* has_rows; const uint8_t* alive_step; const double* xw; const double* yw;
    const double* area; double* x; double* y; double* u; double* v; double* ksi; double* h; uint8_t* alive;
    double* mass; double* inertia; double* alpha; double* dXi_p; double* dYi_p; double* dUi_p; double* dVi_p; double* dalpha_p; double* dksi_p;
    const double* FxOA; const double* FyOA; const double* torqueOA;
    const int* voff; const double* c0x; const double* c0y; double* cax; double* cay;
    double* stress_h; int* scount; int* flags; Counters* cnt;
    const int* omap;              // slab mode: list position of owned floe i (the contact step's per-entry outputs are indexed by it); NULL: i
    const uint8_t* forced; int do_int; double* strain;     // forcing evaluated this step (h < 0.1 is then fine); doInt.flag: floe.strain (:224-234)
};
// calc_trajectory.m for one floe per thread: the branch with doInt.flag = false and the ocean/atmosphere tendencies
// FxOA, FyOA, torqueOA carried over (:3-46 stress history slot + force clamp, :67-80 thermodynamic thinning, :89,116-117
// sacking, :170-222 second-order Adams-Bashforth update of position, heading, velocities, outline rotation).
// flags: bit 0 sacked (the reference returns [] and the caller keeps the old struct), bit 1 would need the ocean (h < 0.1).
__global__ void trajectory_kernel(const TrajArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n0) return;
    a.flags[i] = 0;
    if (i < a.Nb) return;            // the timestepping loop is `parfor i=1+Nb:N0` (floe_interactions_all.m:249): topography floes are never wrapped, thinned or moved
    // the contact step's per-floe results replace the inputs of the integrator (floe_interactions_all.m:152-155,267-277)
    const int m = a.omap ? a.omap[i] : i;
    uint8_t alive = a.alive_step[m];
    double X = a.xw[m], Y = a.yw[m];
    a.alive[i] = alive; a.x[i] = X; a.y[i] = Y;
    if (!alive) return;                                                            // :280
    double ext_fx = a.cfx[m], ext_fy = a.cfy[m], ext_t = a.ctq[m];
    double mass = a.mass[i], inertia = a.inertia[i], h = a.h[i];
    int sc = a.scount[i];
    if (sc > a.nz) sc = 1;                                                        // :15-17
    if (h > 10) h = 10; else if (mass < 100) { mass = 1e3; alive = 0; }           // :36-41
    while (fmax(fabs(ext_fx), fabs(ext_fy)) > mass / (5 * a.dt)) { ext_fx = ext_fx / 10; ext_fy = ext_fy / 10; ext_t = ext_t / 10; }   // :42-46
    const double floe_area = a.area[i];
    const double dh = a.HFo * a.dt / h;                                           // :75-79
    const double floe_mass = (h - dh) / h * mass, floe_inertia = (h - dh) / h * inertia;
    const double h_new = h - dh;
    bool sack = (X != X);                                                         // :89
    if (!sack && h_new < 0.1 && !(a.forced && a.forced[i])) { a.flags[i] = 2; atomicAdd(&a.cnt->n_fail, 1); return; }
    if (!sack) {
        double cmaxx = -SZ_INF, cminx = SZ_INF, cmaxy = -SZ_INF, cminy = SZ_INF;
        for (int t = a.voff[i]; t < a.voff[i + 1]; ++t) { cmaxx = fmax(cmaxx, a.cax[t]); cminx = fmin(cminx, a.cax[t]); cmaxy = fmax(cmaxy, a.cay[t]); cminy = fmin(cminy, a.cay[t]); }
        sack = (cmaxx + X > a.xo_max || cminx + X < a.xo_min || cmaxy + Y > a.yo_max || cminy + Y < a.yo_min);   // :116-117
    }
    if (sack) { a.flags[i] = 1; atomicAdd(&a.cnt->n_cap_fail, 1); return; }       // state untouched, like the caller's `kill(i) = i`
    // commit: stress history slot (:18-19), clamps, thinning
    double* slot = a.stress_h + ((size_t)i * a.nz + (sc - 1)) * 4;
    for (int k = 0; k < 4; ++k) slot[k] = a.has_rows[m] ? a.stress_now[(size_t)m * 4 + k] : 0.0;
    a.scount[i] = sc + 1;
    a.mass[i] = floe_mass; a.inertia[i] = floe_inertia; a.h[i] = h_new; a.alive[i] = alive;
    if (alive != 1) return;                                                       // :118
    const double dt = a.dt, U = a.u[i], V = a.v[i], K = a.ksi[i];
    a.x[i] = X + (1.5 * dt * U - 0.5 * dt * a.dXi_p[i]); a.dXi_p[i] = U;          // :174-176
    a.y[i] = Y + (1.5 * dt * V - 0.5 * dt * a.dYi_p[i]); a.dYi_p[i] = V;
    const double alpha = a.alpha[i] + 1.5 * dt * K - 0.5 * dt * a.dalpha_p[i];     // :177
    a.alpha[i] = alpha; a.dalpha_p[i] = K;
    const double ax0 = a.FxOA[i] * floe_area + ext_fx, ay0 = a.FyOA[i] * floe_area + ext_fy;   // :181-182
    double dU = ax0 / floe_mass, dV = ay0 / floe_mass;
    bool have_frac = false; double frac = 0;
    const double lim = 0.5 * h_new;
    if (fabs(dt * dU) > lim && fabs(dt * dV) > lim) {                             // :184-191
        dU = sgn_d(dU) * 0.5 * h_new / dt; dV = sgn_d(dV) * 0.5 * h_new / dt;
        const double f1 = dU / ax0 * floe_mass, f2 = dV / ay0 * floe_mass;
        frac = f1 < f2 ? f1 : f2; have_frac = true;
        dU = ax0 / floe_mass; dV = ay0 / floe_mass; dU = frac * dU; dV = frac * dV;
    } else if (fabs(dt * dU) > lim && fabs(dt * dV) < lim) {                      // :192-197
        dU = sgn_d(dU) * 0.5 * h_new / dt;
        frac = dU / ax0 * floe_mass; have_frac = true;
        dU = ax0 / floe_mass; dV = ay0 / floe_mass; dU = frac * dU; dV = frac * dV;
    } else if (fabs(dt * dU) < lim && fabs(dt * dV) > lim) {                      // :198-203
        dV = sgn_d(dV) * 0.5 * h_new / dt;
        frac = dV / ay0 * floe_mass; have_frac = true;
        dU = ax0 / floe_mass; dV = ay0 / floe_mass; dU = frac * dU; dV = frac * dV;
    }
    a.u[i] = U + 1.5 * dt * dU - 0.5 * dt * a.dUi_p[i];                           // :204-207
    a.v[i] = V + 1.5 * dt * dV - 0.5 * dt * a.dVi_p[i];
    a.dUi_p[i] = dU; a.dVi_p[i] = dV;
    double dksi = (a.torqueOA[i] * floe_area + ext_t) / floe_inertia;             // :209-219
    if (have_frac) dksi = frac * dksi;
    double k2 = K + 1.5 * dt * dksi - 0.5 * dt * a.dksi_p[i];
    if (fabs(k2) > 1e-5) k2 = sgn_d(k2) * 1e-5;
    a.ksi[i] = k2; a.dksi_p[i] = dksi;
    const double ca = cos(alpha), sa = sin(alpha);                                // :221-222
    for (int t = a.voff[i]; t < a.voff[i + 1]; ++t) { const double px = a.c0x[t], py = a.c0y[t]; a.cax[t] = ca * px + (-sa) * py; a.cay[t] = sa * px + ca * py; }
    if (a.do_int && a.strain) {                                                   // :224-234
        const int o = a.voff[i], m = a.voff[i + 1] - o;
        const double Un = a.u[i], Vn = a.v[i];
        double sxu = 0, syu = 0, sxv = 0, syv = 0, U0 = 0, V0 = 0, Uf = 0, Vf = 0;
        for (int t = 0; t <= m && m > 0; ++t) {
            const int tt = (t == m) ? 0 : t;
            double U1, V1;
            if (t == m) { U1 = Uf; V1 = Vf; }
            else {
                const double theta = atan2(a.cay[o + tt], a.cax[o + tt]), rho = hypot(a.cax[o + tt], a.cay[o + tt]);
                U1 = Un - rho * k2 * sin(theta); V1 = Vn + rho * k2 * cos(theta);
                if (t == 0) { Uf = U1; Vf = V1; }
            }
            if (t > 0) {
                const int tp = t - 1;
                const double dxc = a.cax[o + tt] - a.cax[o + tp], dyc = a.cay[o + tt] - a.cay[o + tp];
                sxu += (U1 - U0) * dyc; syu += (U1 - U0) * dxc; sxv += (V1 - V0) * dyc; syv += (V1 - V0) * dxc;
            }
            U0 = U1; V0 = V1;
        }
        const double du_dx = 0.5 * sxu / floe_area, du_dy = 0.5 * syu / floe_area, dv_dx = 0.5 * sxv / floe_area, dv_dy = 0.5 * syv / floe_area;
        double* E = a.strain + (size_t)i * 4;
        E[0] = 0.5 * (du_dx + du_dx); E[1] = 0.5 * (du_dy + dv_dx); E[2] = 0.5 * (dv_dx + du_dy); E[3] = 0.5 * (dv_dy + dv_dy);
    }
}

// ------------------------------------------------------------------------------------------------ ocean / atmosphere forcing
struct OceanArgs {
    int n0, npts, nx, ny, do_int, Nb; double dt, HFo, xo_min, xo_max, yo_min, yo_max, fc, turn, rho0, Cd, rho_air, Cd_atm;
    const uint8_t* alive_step; const double* xw; const double* yw; const double* u; const double* v; const double* ksi; const double* h;
    const double* mass; const double* area; const double* alpha; const int* voff; const double* cax; const double* cay;
    const double* PX; const double* PY; const uint8_t* PA;
    const double* Xo; const double* Yo; const double* U; const double* V; const double* Wu; const double* Wv;
    double* FxOA; double* FyOA; double* torqueOA; uint8_t* forced; int* flags; Counters* cnt;
    const int* omap;
};
// interp2(X, Y, V, xq, yq), 'linear', NaN outside the grid; V column-major (iy + ix*ny) like MATLAB (calc_trajectory.m:135-138)
__device__ __forceinline__ void interp_cell(const double* __restrict__ X, int nx, double xq, int& ix, double& t)
{
    int lo = 0, hi = nx - 1;                       // largest ix with X[ix] <= xq, at most nx - 2
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (X[mid] <= xq) lo = mid; else hi = mid; }
    ix = lo; t = (xq - X[lo]) / (X[lo + 1] - X[lo]);
}
__device__ __forceinline__ double interp_val(const double* __restrict__ V, int ny, int ix, int iy, double t, double s)
{
    const double v00 = V[iy + (size_t)ix * ny], v10 = V[iy + (size_t)(ix + 1) * ny], v01 = V[iy + 1 + (size_t)ix * ny], v11 = V[iy + 1 + (size_t)(ix + 1) * ny];
    return (v00 * (1 - t) + v10 * t) * (1 - s) + (v01 * (1 - t) + v11 * t) * s;
}
__device__ __forceinline__ double warp_sum(double v) { for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d); return v; }
// calc_trajectory.m:94-166, one warp per floe: lanes stride over the floe's Monte-Carlo points, butterfly sums.
// The bounds / thinning of :36-41,67-80 and the sacking tests of :89,116-117 are repeated to select the floes and to get
// floe_mass exactly as the integrator will (trajectory_kernel runs after this kernel on the same state).
__global__ void __launch_bounds__(256) ocean_forcing_kernel(const OceanArgs a)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= a.n0 || i < a.Nb) return;                                             // floe_interactions_all.m:249 (i = 1+Nb:N0)
    const int mo = a.omap ? a.omap[i] : i;
    if (!a.alive_step[mo]) return;                                                 // floe_interactions_all.m:280
    double hh = a.h[i], m = a.mass[i]; int alive = a.alive_step[mo];
    if (hh > 10) hh = 10; else if (m < 100) { m = 1e3; alive = 0; }                // :36-41
    const double dh = a.HFo * a.dt / hh;
    const double floe_mass = (hh - dh) / hh * m, h_new = hh - dh, floe_area = a.area[i];
    const double Xi = a.xw[mo], Yi = a.yw[mo];
    if (Xi != Xi) return;                                                          // :89
    if (!(a.do_int || h_new < 0.1)) return;                                        // :94
    double cmaxx = -SZ_INF, cminx = SZ_INF, cmaxy = -SZ_INF, cminy = SZ_INF;
    for (int t = a.voff[i] + lane; t < a.voff[i + 1]; t += 32) { cmaxx = fmax(cmaxx, a.cax[t]); cminx = fmin(cminx, a.cax[t]); cmaxy = fmax(cmaxy, a.cay[t]); cminy = fmin(cminy, a.cay[t]); }
    for (int d = 16; d > 0; d >>= 1) {
        cmaxx = fmax(cmaxx, __shfl_xor_sync(0xffffffffu, cmaxx, d)); cminx = fmin(cminx, __shfl_xor_sync(0xffffffffu, cminx, d));
        cmaxy = fmax(cmaxy, __shfl_xor_sync(0xffffffffu, cmaxy, d)); cminy = fmin(cminy, __shfl_xor_sync(0xffffffffu, cminy, d));
    }
    if (cmaxx + Xi > a.xo_max || cminx + Xi < a.xo_min || cmaxy + Yi > a.yo_max || cminy + Yi < a.yo_min) return;   // :116-117
    if (alive != 1) return;                                                        // :118
    const double* px = a.PX + (size_t)i * a.npts; const double* py = a.PY + (size_t)i * a.npts; const uint8_t* pa = a.PA + (size_t)i * a.npts;
    double cnt = 0;
    for (int k = lane; k < a.npts; k += 32) cnt += pa[k] != 0;
    cnt = warp_sum(cnt);
    if (cnt == 0) { if (lane == 0) { a.flags[i] |= 4; atomicAdd(&a.cnt->n_no_points, 1); } return; }   // :100-111 draws new random points
    const double ca = cos(a.alpha[i]), sa = sin(a.alpha[i]);
    const double Ui = a.u[i], Vi = a.v[i], K = a.ksi[i];
    double su = 0, sv = 0;                                                         // winds over the floe (:140)
    for (int k = lane; k < a.npts; k += 32) if (pa[k]) {
        const double xr = ca * px[k] + (-sa) * py[k], yr = sa * px[k] + ca * py[k];
        const double xq = xr + Xi, yq = yr + Yi;
        if (!(xq >= a.Xo[0] && xq <= a.Xo[a.nx - 1] && yq >= a.Yo[0] && yq <= a.Yo[a.ny - 1])) { su = sv = SZ_INF - SZ_INF; continue; }
        int ix, iy; double t, s; interp_cell(a.Xo, a.nx, xq, ix, t); interp_cell(a.Yo, a.ny, yq, iy, s);
        su += interp_val(a.Wu, a.ny, ix, iy, t, s); sv += interp_val(a.Wv, a.ny, ix, iy, t, s);
    }
    const double U10 = warp_sum(su) / cnt, V10 = warp_sum(sv) / cnt;
    const double Fx_atm = a.rho_air * a.Cd_atm * sqrt(U10 * U10 + V10 * V10) * U10, Fy_atm = a.rho_air * a.Cd_atm * sqrt(U10 * U10 + V10 * V10) * V10;   // :141-142
    const double mfa = floe_mass / floe_area, cturn = cos(a.turn), sturn = sin(a.turn);
    double sfx = 0, sfy = 0, stq = 0;
    for (int k = lane; k < a.npts; k += 32) if (pa[k]) {
        const double xr = ca * px[k] + (-sa) * py[k], yr = sa * px[k] + ca * py[k];
        const double theta = atan2(yr, xr), rho = hypot(xr, yr);                   // cart2pol :124
        const double sth = sin(theta), cth = cos(theta);
        const double Uice = Ui - rho * K * sth, Vice = Vi + rho * K * cth;         // :127-128
        const double xq = xr + Xi, yq = yr + Yi;
        double uo, vo;
        if (!(xq >= a.Xo[0] && xq <= a.Xo[a.nx - 1] && yq >= a.Yo[0] && yq <= a.Yo[a.ny - 1])) uo = vo = SZ_INF - SZ_INF;
        else { int ix, iy; double t, s; interp_cell(a.Xo, a.nx, xq, ix, t); interp_cell(a.Yo, a.ny, yq, iy, s); uo = interp_val(a.U, a.ny, ix, iy, t, s); vo = interp_val(a.V, a.ny, ix, iy, t, s); }
        const double fxp = -mfa * a.fc * vo, fyp = +mfa * a.fc * uo;               // :144-145
        const double du = uo - Uice, dv = vo - Vice;                               // :147
        const double sp = sqrt(du * du + dv * dv);
        const double tx = a.rho0 * a.Cd * sp * (cturn * du - sturn * dv), ty = a.rho0 * a.Cd * sp * (sturn * du + cturn * dv);   // :149-150
        double Fx = tx + Fx_atm + fxp, Fy = ty + Fy_atm + fyp;                     // :152-153
        const double tq = (-Fx * sth + Fy * cth) * rho;                            // :157
        Fx = Fx + mfa * a.fc * Vi; Fy = Fy - mfa * a.fc * Ui;                      // :160-161
        sfx += Fx; sfy += Fy; stq += tq;
    }
    sfx = warp_sum(sfx); sfy = warp_sum(sfy); stq = warp_sum(stq);
    if (lane == 0) {
        a.FxOA[i] = sfx / cnt; a.FyOA[i] = sfy / cnt; a.torqueOA[i] = stq / cnt;   // :164-166
        a.forced[i] = 1; atomicAdd(&a.cnt->n_forced, 1);
    }
}
// floe.Stress = mean(StressH, 3) (calc_trajectory.m:20,28), evaluated when asked for
__global__ void stress_mean_kernel(int n0, int nz, const double* __restrict__ stress_h, double* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n0 * 4) return;
    const int i = t >> 2, k = t & 3;
    double sm = 0;
    for (int z = 0; z < nz; ++z) sm += stress_h[((size_t)i * nz + z) * 4 + k];
    out[t] = sm / nz;
}

extern "C" int sz_trajectory_init(SzContext* c, const SzTrajectoryInit* in)
{
    if (!c || !in) { sz_set_error("sz_trajectory_init: NULL argument"); return SZ_ERR_ARG; }
    if (!c->have_input || (c->ext_mode && !c->slab)) { sz_set_error("sz_trajectory_init: upload the floes first (single-GPU list or sz_slab_upload)"); return SZ_ERR_STATE; }
    if (in->nz < 1 || !in->mass || !in->inertia) { sz_set_error("sz_trajectory_init: mass, inertia and nz >= 1 are required"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->n0, nv = (size_t)c->nverts; cudaStream_t st = c->stream;
    struct { DBuf<double>* b; const double* src; } arr[] = {{&c->t_mass, in->mass}, {&c->t_inertia, in->inertia}, {&c->t_alpha, in->alpha}, {&c->t_dXi_p, in->dXi_p}, {&c->t_dYi_p, in->dYi_p},
        {&c->t_dUi_p, in->dUi_p}, {&c->t_dVi_p, in->dVi_p}, {&c->t_dalpha_p, in->dalpha_p}, {&c->t_dksi_p, in->dksi_p}, {&c->t_FxOA, in->FxOA}, {&c->t_FyOA, in->FyOA}, {&c->t_torqueOA, in->torqueOA}};
    for (auto& a : arr) {
        CK(a.b->ensure(n + 1));
        if (a.src) CK(cudaMemcpyAsync(a.b->p, a.src, n * 8, cudaMemcpyDefault, st)); else CK(cudaMemsetAsync(a.b->p, 0, n * 8, st));
    }
    CK(c->c0x.ensure(nv + 1)); CK(c->c0y.ensure(nv + 1));
    // c0 = the unrotated outline (initialize_floe_values.m:18); default: the current c_alpha, i.e. alpha_i = 0
    CK(cudaMemcpyAsync(c->c0x.p, in->c0x ? in->c0x : c->vx.p, nv * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->c0y.p, in->c0y ? in->c0y : c->vy.p, nv * 8, cudaMemcpyDefault, st));
    CK(c->t_stressH.ensure(n * (size_t)in->nz * 4 + 4)); CK(c->t_stress.ensure(n * 4 + 4)); CK(c->t_scount.ensure(n + 1)); CK(c->t_flags.ensure(n + 1));
    if (in->stress_h) CK(cudaMemcpyAsync(c->t_stressH.p, in->stress_h, n * (size_t)in->nz * 32, cudaMemcpyDefault, st));
    else CK(cudaMemsetAsync(c->t_stressH.p, 0, n * (size_t)in->nz * 32, st));       // StressH = zeros(2,2,1000), StressCount = 1 (:24-25)
    if (in->stress_count) { CK(cudaMemcpyAsync(c->t_scount.p, in->stress_count, n * 4, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st)); }
    else { std::vector<int> ones(n, 1); CK(cudaMemcpyAsync(c->t_scount.p, ones.data(), n * 4, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st)); }
    CK(cudaMemsetAsync(c->t_flags.p, 0, n * 4, st));
    CK(c->t_strain.ensure(n * 4 + 4)); CK(c->t_forced.ensure(n + 1));
    CK(cudaMemsetAsync(c->t_strain.p, 0, n * 32, st)); CK(cudaMemsetAsync(c->t_forced.p, 0, n, st));      // floe.strain starts at zero
    CK(cudaStreamSynchronize(st));
    c->traj_nz = in->nz; c->have_traj = true; c->traj_do_int = false;
    return SZ_OK;
}
extern "C" int sz_trajectory_step(SzContext* c, const SzTrajectoryParams* p, int32_t* n_sacked, int32_t* n_needs_ocean)
{
    if (!c || !p) { sz_set_error("sz_trajectory_step: NULL argument"); return SZ_ERR_ARG; }
    if (!c->have_traj || !c->have_step || (c->ext_mode && !c->slab)) { sz_set_error("sz_trajectory_step: needs sz_trajectory_init and a contact step"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const int n0 = c->n0;
    CK(cudaMemsetAsync(c->d_cnt, 0, sizeof(Counters), st));
    TrajArgs a; memset(&a, 0, sizeof(a));
    a.n0 = n0; a.nz = c->traj_nz; a.Nb = c->prm.Nb; a.dt = p->dt; a.HFo = p->HFo; a.xo_min = p->xo_min; a.xo_max = p->xo_max; a.yo_min = p->yo_min; a.yo_max = p->yo_max;
    a.cfx = c->o_fx.p; a.cfy = c->o_fy.p; a.ctq = c->o_tq.p; a.stress_now = c->o_stress.p; a.has_rows = c->has_rows.p; a.alive_step = c->o_alive.p; a.xw = c->o_xi.p; a.yw = c->o_yi.p;
    a.area = c->area.p; a.x = c->x.p; a.y = c->y.p; a.u = c->u.p; a.v = c->v.p; a.ksi = c->ksi.p; a.h = c->h.p; a.alive = c->alive.p;
    a.mass = c->t_mass.p; a.inertia = c->t_inertia.p; a.alpha = c->t_alpha.p; a.dXi_p = c->t_dXi_p.p; a.dYi_p = c->t_dYi_p.p; a.dUi_p = c->t_dUi_p.p; a.dVi_p = c->t_dVi_p.p;
    a.dalpha_p = c->t_dalpha_p.p; a.dksi_p = c->t_dksi_p.p; a.FxOA = c->t_FxOA.p; a.FyOA = c->t_FyOA.p; a.torqueOA = c->t_torqueOA.p;
    a.voff = c->voff.p; a.c0x = c->c0x.p; a.c0y = c->c0y.p; a.cax = c->vx.p; a.cay = c->vy.p;
    a.stress_h = c->t_stressH.p; a.scount = c->t_scount.p; a.flags = c->t_flags.p; a.cnt = c->d_cnt;
    a.forced = c->t_forced.p; a.do_int = c->traj_do_int ? 1 : 0; a.strain = c->t_strain.p; a.omap = c->slab ? c->sl_opos.p : nullptr;
    if (n0 > 0) { ++g_launches; trajectory_kernel<<<nblk(n0, 128), 128, 0, st>>>(a); }
    CK(cudaMemsetAsync(c->t_forced.p, 0, (size_t)n0, st)); c->traj_do_int = false;
    CK(cudaGetLastError());
    CKS(read_counters(c));
    c->have_step = false;            // the contact results belong to the previous positions now
    if (c->slab) c->sl_built = false; // and so does the extended list
    if (n_sacked) *n_sacked = c->h_cnt->n_cap_fail;
    if (n_needs_ocean) *n_needs_ocean = c->h_cnt->n_fail;
    if (c->h_cnt->n_fail > 0) { sz_set_error("%d floe(s) thinner than 0.1 m need the ocean forcing re-evaluated (calc_trajectory.m:94): not part of this path", c->h_cnt->n_fail); return SZ_ERR_STATE; }
    return SZ_OK;
}
extern "C" int sz_get_stress_history(SzContext* c, double* stress_h, int32_t* stress_count)
{
    if (!c) { sz_set_error("sz_get_stress_history: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_traj) { sz_set_error("sz_get_stress_history: no integrator state"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->n0;
    D2H(stress_h, c->t_stressH.p, n * (size_t)c->traj_nz * 32); D2H(stress_count, c->t_scount.p, n * 4);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
extern "C" int sz_trajectory_set_ocean(SzContext* c, const SzOcean* o)
{
    if (!c) { sz_set_error("sz_trajectory_set_ocean: NULL context"); return SZ_ERR_ARG; }
    if (!o) { c->have_ocean = false; return SZ_OK; }
    if (o->nx < 2 || o->ny < 2 || !o->Xo || !o->Yo || !o->Uocn || !o->Vocn || !o->Uwinds || !o->Vwinds) { sz_set_error("sz_trajectory_set_ocean: grid vectors (>= 2 points each) and the four fields are required"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const size_t nx = o->nx, ny = o->ny;
    CK(c->oc_Xo.ensure(nx)); CK(c->oc_Yo.ensure(ny)); CK(c->oc_U.ensure(nx * ny)); CK(c->oc_V.ensure(nx * ny)); CK(c->oc_Wu.ensure(nx * ny)); CK(c->oc_Wv.ensure(nx * ny));
    CK(cudaMemcpyAsync(c->oc_Xo.p, o->Xo, nx * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->oc_Yo.p, o->Yo, ny * 8, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(c->oc_U.p, o->Uocn, nx * ny * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->oc_V.p, o->Vocn, nx * ny * 8, cudaMemcpyDefault, st));
    CK(cudaMemcpyAsync(c->oc_Wu.p, o->Uwinds, nx * ny * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->oc_Wv.p, o->Vwinds, nx * ny * 8, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    c->oc_nx = o->nx; c->oc_ny = o->ny; c->oc_fc = o->fCoriolis; c->oc_turn = o->turn_angle;
    c->oc_rho0 = o->rho0 != 0 ? o->rho0 : 1027.0; c->oc_Cd = o->Cd != 0 ? o->Cd : 3e-3;          // calc_trajectory.m:57-64
    c->oc_rho_air = o->rho_air != 0 ? o->rho_air : 1.2; c->oc_Cd_atm = o->Cd_atm != 0 ? o->Cd_atm : 1e-3;
    c->have_ocean = true;
    return SZ_OK;
}
extern "C" int sz_trajectory_set_points(SzContext* c, int32_t npts, const double* X, const double* Y, const uint8_t* A)
{
    if (!c || npts < 1 || !X || !Y || !A) { sz_set_error("sz_trajectory_set_points: NULL argument or npts < 1"); return SZ_ERR_ARG; }
    if (!c->have_input || (c->ext_mode && !c->slab)) { sz_set_error("sz_trajectory_set_points: upload the floes first (single-GPU list or sz_slab_upload)"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const size_t tot = (size_t)c->n0 * npts;
    CK(c->pt_x.ensure(tot + 1)); CK(c->pt_y.ensure(tot + 1)); CK(c->pt_a.ensure(tot + 1));
    CK(cudaMemcpyAsync(c->pt_x.p, X, tot * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->pt_y.p, Y, tot * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->pt_a.p, A, tot, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    c->npts = npts; c->have_points = true;
    return SZ_OK;
}
extern "C" int sz_trajectory_ocean_forcing(SzContext* c, const SzTrajectoryParams* p, int32_t do_int, int32_t* n_evaluated, int32_t* n_no_points)
{
    if (!c || !p) { sz_set_error("sz_trajectory_ocean_forcing: NULL argument"); return SZ_ERR_ARG; }
    if (!c->have_traj || !c->have_step || (c->ext_mode && !c->slab)) { sz_set_error("sz_trajectory_ocean_forcing: needs sz_trajectory_init and a contact step"); return SZ_ERR_STATE; }
    if (!c->have_ocean || !c->have_points) { sz_set_error("sz_trajectory_ocean_forcing: needs sz_trajectory_set_ocean and sz_trajectory_set_points"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const int n0 = c->n0;
    CK(cudaMemsetAsync(c->d_cnt, 0, sizeof(Counters), st));
    CK(cudaMemsetAsync(c->t_forced.p, 0, (size_t)n0, st));
    OceanArgs a; memset(&a, 0, sizeof(a));
    a.n0 = n0; a.npts = c->npts; a.nx = c->oc_nx; a.ny = c->oc_ny; a.do_int = do_int != 0; a.Nb = c->prm.Nb;
    a.dt = p->dt; a.HFo = p->HFo; a.xo_min = p->xo_min; a.xo_max = p->xo_max; a.yo_min = p->yo_min; a.yo_max = p->yo_max;
    a.fc = c->oc_fc; a.turn = c->oc_turn; a.rho0 = c->oc_rho0; a.Cd = c->oc_Cd; a.rho_air = c->oc_rho_air; a.Cd_atm = c->oc_Cd_atm;
    a.alive_step = c->o_alive.p; a.xw = c->o_xi.p; a.yw = c->o_yi.p; a.u = c->u.p; a.v = c->v.p; a.ksi = c->ksi.p; a.h = c->h.p;
    a.mass = c->t_mass.p; a.area = c->area.p; a.alpha = c->t_alpha.p; a.voff = c->voff.p; a.cax = c->vx.p; a.cay = c->vy.p;
    a.PX = c->pt_x.p; a.PY = c->pt_y.p; a.PA = c->pt_a.p;
    a.Xo = c->oc_Xo.p; a.Yo = c->oc_Yo.p; a.U = c->oc_U.p; a.V = c->oc_V.p; a.Wu = c->oc_Wu.p; a.Wv = c->oc_Wv.p;
    a.FxOA = c->t_FxOA.p; a.FyOA = c->t_FyOA.p; a.torqueOA = c->t_torqueOA.p; a.forced = c->t_forced.p; a.flags = c->t_flags.p; a.cnt = c->d_cnt; a.omap = c->slab ? c->sl_opos.p : nullptr;
    if (n0 > 0) { ++g_launches; ocean_forcing_kernel<<<nblk(32 * (i64)n0, 256), 256, 0, st>>>(a); }
    CK(cudaGetLastError());
    CKS(read_counters(c));
    c->traj_do_int = do_int != 0;
    if (n_evaluated) *n_evaluated = c->h_cnt->n_forced;
    if (n_no_points) *n_no_points = c->h_cnt->n_no_points;
    if (c->h_cnt->n_no_points > 0) { sz_set_error("%d floe(s) have no Monte-Carlo point inside their outline: the reference draws new random points (calc_trajectory.m:100-111); supply new points with sz_trajectory_set_points", c->h_cnt->n_no_points); return SZ_ERR_STATE; }
    return SZ_OK;
}
extern "C" int sz_get_trajectory_forcing(SzContext* c, double* FxOA, double* FyOA, double* torqueOA, double* strain)
{
    if (!c) { sz_set_error("sz_get_trajectory_forcing: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_traj) { sz_set_error("sz_get_trajectory_forcing: no integrator state"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->n0;
    D2H(FxOA, c->t_FxOA.p, n * 8); D2H(FyOA, c->t_FyOA.p, n * 8); D2H(torqueOA, c->t_torqueOA.p, n * 8); D2H(strain, c->t_strain.p, n * 32);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
extern "C" int sz_get_trajectory(SzContext* c, double* x, double* y, double* u, double* v, double* ksi, double* h, uint8_t* alive, double* mass, double* inertia, double* alpha,
                                 double* dXi_p, double* dYi_p, double* dUi_p, double* dVi_p, double* dalpha_p, double* dksi_p, double* stress, int32_t* flags, double* cax, double* cay)
{
    if (!c) { sz_set_error("sz_get_trajectory: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_traj) { sz_set_error("sz_get_trajectory: no integrator state"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->n0, nv = (size_t)c->nverts;
    if (stress && n) { ++g_launches; stress_mean_kernel<<<nblk(4 * (i64)n, 256), 256, 0, c->stream>>>((int)n, c->traj_nz, c->t_stressH.p, c->t_stress.p); }
    D2H(x, c->x.p, n * 8); D2H(y, c->y.p, n * 8); D2H(u, c->u.p, n * 8); D2H(v, c->v.p, n * 8); D2H(ksi, c->ksi.p, n * 8); D2H(h, c->h.p, n * 8); D2H(alive, c->alive.p, n);
    D2H(mass, c->t_mass.p, n * 8); D2H(inertia, c->t_inertia.p, n * 8); D2H(alpha, c->t_alpha.p, n * 8); D2H(dXi_p, c->t_dXi_p.p, n * 8); D2H(dYi_p, c->t_dYi_p.p, n * 8);
    D2H(dUi_p, c->t_dUi_p.p, n * 8); D2H(dVi_p, c->t_dVi_p.p, n * 8); D2H(dalpha_p, c->t_dalpha_p.p, n * 8); D2H(dksi_p, c->t_dksi_p.p, n * 8);
    D2H(stress, c->t_stress.p, n * 32); D2H(flags, c->t_flags.p, n * 4); D2H(cax, c->vx.p, nv * 8); D2H(cay, c->vy.p, nv * 8);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}

// ------------------------------------------------------------------------------------------------ getters

extern "C" int sz_get_floe_outputs(SzContext* c, double* fx, double* fy, double* torque, double* overlap_area, double* stress,
                                   double* xi, double* yi, uint8_t* alive, int32_t* kill, int32_t* transfer)
{
    NEED_STEP("sz_get_floe_outputs");
    const size_t n = (size_t)c->nout;
    D2H(fx, c->o_fx.p, n * 8); D2H(fy, c->o_fy.p, n * 8); D2H(torque, c->o_tq.p, n * 8); D2H(overlap_area, c->o_ov.p, n * 8);
    D2H(stress, c->o_stress.p, n * 32); D2H(xi, c->o_xi.p, n * 8); D2H(yi, c->o_yi.p, n * 8); D2H(alive, c->o_alive.p, n);
    D2H(kill, c->o_kill.p, n * 4); D2H(transfer, c->o_transfer.p, n * 4);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
// slab mode: the per-floe outputs of the OWNED floes, in their order (per-entry results gathered through the list positions)
__global__ void slab_gather_outputs_kernel(int n, const int* __restrict__ opos, const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ tq, const double* __restrict__ ov,
                                           const double* __restrict__ stress, const double* __restrict__ xi, const double* __restrict__ yi, const uint8_t* __restrict__ alive,
                                           const int* __restrict__ kill, const int* __restrict__ transfer, double* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int m = opos[i];
    out[i] = fx[m]; out[(size_t)n + i] = fy[m]; out[2 * (size_t)n + i] = tq[m]; out[3 * (size_t)n + i] = ov[m];
    for (int k = 0; k < 4; ++k) out[4 * (size_t)n + 4 * (size_t)i + k] = stress[4 * (size_t)m + k];
    out[8 * (size_t)n + i] = xi[m]; out[9 * (size_t)n + i] = yi[m]; out[10 * (size_t)n + i] = alive[m]; out[11 * (size_t)n + i] = kill[m]; out[12 * (size_t)n + i] = transfer[m];
}
// the resident extended list of the last sz_slab_build, [n_list] each (any pointer may be NULL)
extern "C" int sz_slab_get_list(SzContext* c, int32_t* gid, int32_t* floe_num, uint8_t* owned, double* x, double* y)
{
    if (!c || !c->slab || !c->sl_built) { sz_set_error("sz_slab_get_list: needs sz_slab_build"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    SlabScratch s; CK(cudaMemcpyAsync(&s, c->sl_scratch, sizeof(s), cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    const size_t n = (size_t)std::min(s.n_list, c->sl_nl_cap);
    D2H(gid, c->egid.p, n * 4); D2H(floe_num, c->efn.p, n * 4); D2H(owned, c->eowned.p, n); D2H(x, c->ex.p, n * 8); D2H(y, c->ey.p, n * 8);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
extern "C" int sz_slab_get_outputs(SzContext* c, double* fx, double* fy, double* torque, double* overlap_area, double* stress, double* xi, double* yi,
                                   uint8_t* alive, int32_t* kill, int32_t* transfer)
{
    NEED_STEP("sz_slab_get_outputs");
    if (!c->slab) { sz_set_error("sz_slab_get_outputs: slab mode only"); return SZ_ERR_STATE; }
    const int n = c->n0; const size_t N = (size_t)n;
    if (n == 0) return SZ_OK;
    CK(c->sl_out.ensure(13 * N + 1));
    ++g_launches; slab_gather_outputs_kernel<<<nblk(n, 256), 256, 0, c->stream>>>(n, c->sl_opos.p, c->o_fx.p, c->o_fy.p, c->o_tq.p, c->o_ov.p, c->o_stress.p, c->o_xi.p, c->o_yi.p, c->o_alive.p,
                                                                                 c->o_kill.p, c->o_transfer.p, c->sl_out.p);
    CK(cudaGetLastError());
    const double* o = c->sl_out.p;
    D2H(fx, o, N * 8); D2H(fy, o + N, N * 8); D2H(torque, o + 2 * N, N * 8); D2H(overlap_area, o + 3 * N, N * 8); D2H(stress, o + 4 * N, N * 32); D2H(xi, o + 8 * N, N * 8); D2H(yi, o + 9 * N, N * 8);
    std::vector<double> t(3 * N);
    CK(cudaMemcpyAsync(t.data(), o + 10 * N, 3 * N * 8, cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < N; ++i) { if (alive) alive[i] = (uint8_t)t[i]; if (kill) kill[i] = (int32_t)t[N + i]; if (transfer) transfer[i] = (int32_t)t[2 * N + i]; }
    return SZ_OK;
}
// slab mode: Floe(i).interactions of the OWNED floes, in their order, gathered on the device through the list positions
__global__ void slab_rowcount_kernel(int n, const int* __restrict__ opos, const int* __restrict__ row_off, int* __restrict__ cnt)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int m = opos[i]; cnt[i] = row_off[m + 1] - row_off[m]; }
}
__global__ void __launch_bounds__(256) slab_rowcopy_kernel(int n, const int* __restrict__ opos, const int* __restrict__ row_off, const int* __restrict__ out_off,
                                                           const double* __restrict__ rows, double* __restrict__ out)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n) return;
    const int m = opos[i];
    const size_t src = (size_t)row_off[m] * 7, dst = (size_t)out_off[i] * 7;
    const int cnt = (row_off[m + 1] - row_off[m]) * 7;
    for (int t = lane; t < cnt; t += 32) out[dst + t] = rows[src + t];
}
extern "C" int sz_slab_get_rows(SzContext* c, int64_t* row_off, double* rows, int64_t rows_cap, int64_t* n_rows)
{
    NEED_STEP("sz_slab_get_rows");
    if (!c->slab) { sz_set_error("sz_slab_get_rows: slab mode only"); return SZ_ERR_STATE; }
    const int n = c->n0; cudaStream_t st = c->stream;
    if (n_rows) *n_rows = 0;
    if (n == 0) { if (row_off) row_off[0] = 0; return SZ_OK; }
    CK(c->sl_rcnt.ensure(n + 2)); CK(c->sl_roff.ensure(n + 2));
    ++g_launches; slab_rowcount_kernel<<<nblk(n, 256), 256, 0, st>>>(n, c->sl_opos.p, c->row_off.p, c->sl_rcnt.p);
    CKS(exclusive_scan(c, c->sl_rcnt.p, n, c->sl_roff.p, n + 1));
    std::vector<int> off(n + 1);
    CK(cudaMemcpyAsync(off.data(), c->sl_roff.p, (size_t)(n + 1) * 4, cudaMemcpyDefault, st));
    CK(cudaStreamSynchronize(st));
    const i64 total = off[n];
    if (n_rows) *n_rows = total;
    if (row_off) for (int k = 0; k <= n; ++k) row_off[k] = off[k];
    if (rows && total > 0) {
        if (total > rows_cap) { sz_set_error("sz_slab_get_rows: %lld rows, room for %lld", (long long)total, (long long)rows_cap); return SZ_ERR_CAPACITY; }
        CK(c->sl_rows.ensure((size_t)total * 7 + 7));
        ++g_launches; slab_rowcopy_kernel<<<nblk(32 * (i64)n, 256), 256, 0, st>>>(n, c->sl_opos.p, c->row_off.p, c->sl_roff.p, c->rows.p, c->sl_rows.p);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(rows, c->sl_rows.p, (size_t)total * 56, cudaMemcpyDefault, st));
        CK(cudaStreamSynchronize(st));
    }
    return SZ_OK;
}
extern "C" int sz_get_ghosts(SzContext* c, int32_t* parent, int32_t* floe_num, double* gx, double* gy)
{
    NEED_STEP("sz_get_ghosts");
    if (c->ext_mode) { sz_set_error("sz_get_ghosts: not available for a caller-supplied extended list"); return SZ_ERR_STATE; }
    const size_t g = (size_t)(c->n - c->n0); const int n0 = c->n0;
    D2H(parent, c->eparent.p + n0, g * 4); D2H(floe_num, c->efn.p + n0, g * 4); D2H(gx, c->ex.p + n0, g * 8); D2H(gy, c->ey.p + n0, g * 8);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
// what the reference leaves in the ghost structs Floe(N0+1:N) (floe_interactions_all.m:218-238): collision_force and
// collision_torque = the column sums of the image's own rows (zero when it has none), OverlapArea (:137,198)
extern "C" int sz_get_ghost_outputs(SzContext* c, double* fx, double* fy, double* torque, double* overlap_area)
{
    NEED_STEP("sz_get_ghost_outputs");
    if (c->ext_mode) { sz_set_error("sz_get_ghost_outputs: not available for a caller-supplied extended list"); return SZ_ERR_STATE; }
    const size_t g = (size_t)(c->n - c->n0); const int n0 = c->n0;
    if (g == 0) return SZ_OK;
    std::vector<double> sums(g * 3); std::vector<uint8_t> has(g);
    CK(cudaMemcpyAsync(sums.data(), c->osum.p + (size_t)n0 * 3, g * 24, cudaMemcpyDefault, c->stream));
    CK(cudaMemcpyAsync(has.data(), c->has_rows.p + n0, g, cudaMemcpyDefault, c->stream));
    D2H(overlap_area, c->e_ov.p + n0, g * 8);
    CK(cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < g; ++k) {
        if (fx) fx[k] = has[k] ? sums[k * 3] : 0.0;
        if (fy) fy[k] = has[k] ? sums[k * 3 + 1] : 0.0;
        if (torque) torque[k] = has[k] ? sums[k * 3 + 2] : 0.0;
    }
    return SZ_OK;
}
extern "C" int sz_get_pairs(SzContext* c, int32_t* pi, int32_t* pj, double* overlap_state, int32_t* n_regions, int32_t* status)
{
    NEED_STEP("sz_get_pairs");
    const size_t np = (size_t)c->n_pairs;
    D2H(pi, c->pi.p, np * 4); D2H(pj, c->pj.p, np * 4); D2H(overlap_state, c->povl.p, np * 8); D2H(n_regions, c->pnrows.p, np * 4); D2H(status, c->pstatus.p, np * 4);
    CK(cudaStreamSynchronize(c->stream));
    if (pi || pj) {                                             // 1-based positions in the (global) extended list
        std::vector<int> gid((size_t)c->n);
        CK(cudaMemcpy(gid.data(), c->egid.p, (size_t)c->n * 4, cudaMemcpyDefault));
        if (pi) for (size_t k = 0; k < np; ++k) pi[k] = gid[pi[k]];
        if (pj) for (size_t k = 0; k < np; ++k) pj[k] = gid[pj[k]];
    }
    if (overlap_state && status) for (size_t k = 0; k < np; ++k) if (status[k] != 0) overlap_state[k] = 0;
    return SZ_OK;
}
extern "C" int sz_get_rows(SzContext* c, int64_t* row_off, double* rows)
{
    NEED_STEP("sz_get_rows");
    const int n = c->n;
    if (row_off) {
        std::vector<int> tmp(n + 1);
        CK(cudaMemcpyAsync(tmp.data(), c->row_off.p, (size_t)(n + 1) * 4, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k <= n; ++k) row_off[k] = tmp[k];
    }
    D2H(rows, c->rows.p, (size_t)c->n_rows * 56);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
extern "C" int sz_get_phase_ms(SzContext* c, float* ms5)
{
    NEED_STEP("sz_get_phase_ms");
    if (ms5) for (int k = 0; k < 5; ++k) ms5[k] = c->phase_ms[k];
    return SZ_OK;
}
extern "C" int sz_get_narrow_class_ms(SzContext* c, float* ms5, int32_t* pairs5)
{
    NEED_STEP("sz_get_narrow_class_ms");
    for (int k = 0; k < 5; ++k) {
        float ms = 0;
        if (c->evk_used[k]) CK(cudaEventElapsedTime(&ms, c->evk[2 * k], c->evk[2 * k + 1]));
        if (ms5) ms5[k] = ms;
        if (pairs5) pairs5[k] = c->class_pairs[k];
    }
    return SZ_OK;
}
extern "C" int sz_set_stream(SzContext* c, void* cuda_stream)
{
    if (!c) { sz_set_error("sz_set_stream: NULL context"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;      // (cudaStreamLegacy / cudaStreamPerThread are valid handles)
    return SZ_OK;
}
extern "C" int sz_set_option(SzContext* c, const char* name, int32_t value)
{
    if (!c || !name) { sz_set_error("sz_set_option: NULL argument"); return SZ_ERR_ARG; }
    if (strcmp(name, "convex_fast") == 0) { c->opt_convex_fast = value != 0; return SZ_OK; }
    if (strcmp(name, "convex_split") == 0) { c->opt_convex_split = value != 0; return SZ_OK; }
    if (strcmp(name, "apart") == 0) { c->opt_apart = value != 0; return SZ_OK; }
    if (strcmp(name, "graph_safe") == 0) { c->opt_graph_safe = value != 0; return SZ_OK; }
    if (strcmp(name, "speculate") == 0) { c->opt_speculate = value != 0; c->plan_valid = false; return SZ_OK; }
    if (strcmp(name, "euler_cell_warp") == 0) { c->opt_euler_cell_warp = value != 0; return SZ_OK; }
    sz_set_error("sz_set_option: unknown option '%s'", name);
    return SZ_ERR_ARG;
}
extern "C" int sz_get_stat(SzContext* c, const char* name, int64_t* value)
{
    if (!c || !name || !value) { sz_set_error("sz_get_stat: NULL argument"); return SZ_ERR_ARG; }
    if (strcmp(name, "speculated_steps") == 0) { *value = c->n_fast_steps; return SZ_OK; }
    if (strcmp(name, "repeated_steps") == 0) { *value = c->n_slow_steps; return SZ_OK; }
    if (strcmp(name, "classifier_answered") == 0) { *value = c->h_cnt ? c->h_cnt->n_bbox_reject : 0; return SZ_OK; }     // pairs of the last step that needed no sweep (boxes, separating axis, edge-by-edge certificate)
    sz_set_error("sz_get_stat: unknown statistic '%s'", name);
    return SZ_ERR_ARG;
}
// ------------------------------------------------------------------------------------------------ fracture deformation
extern "C" int sz_fracture_deform(SzContext* c, int32_t count, const int32_t* floe_idx, int64_t* n_changed, int64_t* n_verts)
{
    if (!c) { sz_set_error("sz_fracture_deform: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_rows) { sz_set_error("sz_fracture_deform: no contact rows on the device (run a contact step first)"); return SZ_ERR_STATE; }
    if (c->ext_mode) { sz_set_error("sz_fracture_deform: single-GPU lists only"); return SZ_ERR_STATE; }
    if (count < 0 || (count > 0 && !floe_idx)) { sz_set_error("sz_fracture_deform: bad arguments"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->have_fr = false; c->fr_count = count; c->fr_verts = 0;
    if (n_changed) *n_changed = 0;
    if (n_verts) *n_verts = 0;
    if (count == 0) { c->have_fr = true; return SZ_OK; }
    {   // the numbers must name floes of the list
        std::vector<int32_t> h(count);
        CK(cudaMemcpy(h.data(), floe_idx, (size_t)count * 4, cudaMemcpyDefault));
        for (int32_t v : h) if (v < 1 || v > c->n0) { sz_set_error("sz_fracture_deform: floe number %d outside 1..%d", v, c->n0); return SZ_ERR_ARG; }
        CK(c->fr_idx.ensure(count + 1));
        CK(cudaMemcpyAsync(c->fr_idx.p, h.data(), (size_t)count * 4, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st));
    }
    CK(c->fr_changed.ensure(count + 1)); CK(c->fr_xi.ensure(count + 1)); CK(c->fr_yi.ensure(count + 1)); CK(c->fr_area.ensure(count + 1));
    CK(c->fr_vstart.ensure(count + 1)); CK(c->fr_vcount.ensure(count + 1)); CK(c->fr_status.ensure(count + 1));
    CK(c->fr_vx.ensure((size_t)c->nverts + 64 * (size_t)count + 64)); CK(c->fr_vy.ensure(c->fr_vx.cap));
    const int threads = std::min((count + 63) / 64 * 64, 148 * 16);
    CK(c->scratchL.ensure((size_t)threads * sz_workspace_bytes_L()));
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(cudaMemsetAsync(c->d_cnt, 0, sizeof(Counters), st));
        sznarrow::FractureArgs a; memset(&a, 0, sizeof(a));
        a.count = count; a.idx = c->fr_idx.p; a.n0 = c->n0;
        a.x = c->x.p; a.y = c->y.p; a.area = c->area.p; a.voff = c->voff.p; a.vx = c->vx.p; a.vy = c->vy.p;
        a.row_off = c->row_off.p; a.rows = c->rows.p;
        a.changed = c->fr_changed.p; a.oxi = c->fr_xi.p; a.oyi = c->fr_yi.p; a.oarea = c->fr_area.p; a.vstart = c->fr_vstart.p; a.vcount = c->fr_vcount.p; a.status = c->fr_status.p;
        a.pvx = c->fr_vx.p; a.pvy = c->fr_vy.p; a.vert_cap = (int)std::min<size_t>(c->fr_vx.cap, 0x7fffffff); a.vert_used = D_CNT(fr_vert_used); a.n_changed = D_CNT(fr_changed);
        a.scratch = c->scratchL.p; a.n_threads = threads;
        ++g_launches; sz_launch_fracture_L(&a, st);
        CK(cudaGetLastError());
        CKS(read_counters(c));
        if ((size_t)c->h_cnt->fr_vert_used <= c->fr_vx.cap) break;
        if (attempt == 1) { sz_set_error("sz_fracture_deform: vertex pool kept overflowing"); return SZ_ERR_CAPACITY; }
        CK(c->fr_vx.ensure((size_t)c->h_cnt->fr_vert_used + 64)); CK(c->fr_vy.ensure(c->fr_vx.cap));
    }
    c->fr_verts = c->h_cnt->fr_vert_used;
    std::vector<int> stt(count);
    CK(cudaMemcpy(stt.data(), c->fr_status.p, (size_t)count * 4, cudaMemcpyDefault));
    for (int k = 0; k < count; ++k) if (stt[k] != 0) {
        const int code = stt[k] == szpf::PS_CLIPPER_FAIL ? SZ_ERR_CLIPPER : (stt[k] == szpf::PS_CAPACITY ? SZ_ERR_CAPACITY : SZ_ERR_ARG);
        sz_set_error("sz_fracture_deform: item %d failed (%s)", k, code == SZ_ERR_CLIPPER ? "Clipper Error." : code == SZ_ERR_CAPACITY ? "outline beyond the largest size class" : "p_poly_dist would raise: degenerate region outline");
        return code;
    }
    c->have_fr = true;
    if (n_changed) *n_changed = c->h_cnt->fr_changed;
    if (n_verts) *n_verts = c->fr_verts;
    return SZ_OK;
}
extern "C" int sz_get_fracture_deform(SzContext* c, uint8_t* changed, double* xi, double* yi, double* area, int64_t* vert_off, double* cx, double* cy)
{
    if (!c) { sz_set_error("sz_get_fracture_deform: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_fr) { sz_set_error("sz_get_fracture_deform: no result"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->fr_count;
    D2H(changed, c->fr_changed.p, n); D2H(xi, c->fr_xi.p, n * 8); D2H(yi, c->fr_yi.p, n * 8); D2H(area, c->fr_area.p, n * 8);
    CK(cudaStreamSynchronize(c->stream));
    if (vert_off || cx || cy) {
        std::vector<int> vs(n), vc(n); std::vector<double> px((size_t)c->fr_verts), py((size_t)c->fr_verts);
        if (n) { CK(cudaMemcpy(vs.data(), c->fr_vstart.p, n * 4, cudaMemcpyDefault)); CK(cudaMemcpy(vc.data(), c->fr_vcount.p, n * 4, cudaMemcpyDefault)); }
        if (c->fr_verts) { CK(cudaMemcpy(px.data(), c->fr_vx.p, (size_t)c->fr_verts * 8, cudaMemcpyDefault)); CK(cudaMemcpy(py.data(), c->fr_vy.p, (size_t)c->fr_verts * 8, cudaMemcpyDefault)); }
        int64_t pos = 0;
        if (vert_off) vert_off[0] = 0;
        for (size_t k = 0; k < n; ++k) {
            for (int t = 0; t < vc[k]; ++t) { if (cx) cx[pos + t] = px[(size_t)vs[k] + t]; if (cy) cy[pos + t] = py[(size_t)vs[k] + t]; }
            pos += vc[k];
            if (vert_off) vert_off[k + 1] = pos;
        }
    }
    return SZ_OK;
}
// ------------------------------------------------------------------------------------------------ corner mask
// corners.m:13-51 rebuilds the periodic list from the current positions, whether or not the run is periodic; only the
// centroid, the source floe and the alive flag of an image are needed here.
__global__ void corner_ext_init_kernel(int n0, const double* __restrict__ x, const double* __restrict__ y, const uint8_t* __restrict__ alive,
                                       double* __restrict__ ex, double* __restrict__ ey, int* __restrict__ esrc, uint8_t* __restrict__ ealive)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0) return;
    ex[i] = x[i]; ey[i] = y[i]; esrc[i] = i; ealive[i] = alive[i];
}
__global__ void corner_ext_emit_kernel(int axis, int n_bound, const int* __restrict__ n_dev, const int* __restrict__ flag, const int* __restrict__ pos,
                                       double* __restrict__ ex, double* __restrict__ ey, int* __restrict__ esrc, uint8_t* __restrict__ ealive, double L, int* __restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_dev ? *n_dev : n_bound;
    if (i == 0) *n_out = n + pos[n_bound];
    if (i >= n || !flag[i]) return;
    const int g = n + pos[i];
    if (axis == 0) { ex[g] = ex[i] - 2 * L * sgn_d(ex[i]); ey[g] = ey[i]; }                      // corners.m:21-30
    else { ex[g] = ex[i]; ey[g] = ey[i] - 2 * L * sgn_d(ey[i]); }                                // :38-46
    esrc[g] = esrc[i]; ealive[g] = ealive[i];
}
__global__ void corner_count_kernel(const szcorn::CornerArgs a, int* __restrict__ nv)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < a.count) nv[q] = szcorn::open_count(a, a.idx[q] - 1);
}
template <int G>
__global__ void __launch_bounds__(256) corner_mask_kernel(const szcorn::CornerArgs a)
{
    const int lane = threadIdx.x % G;
    const unsigned mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((threadIdx.x & 31) / G * G));
    const int groups = gridDim.x * (blockDim.x / G);
    for (int q = blockIdx.x * (blockDim.x / G) + threadIdx.x / G; q < a.count; q += groups) szcorn::corner_mask_floe<G>(a, q, lane, mask);
}

extern "C" int sz_corner_mask(SzContext* c, int32_t count, const int32_t* floe_idx, int32_t nb_skip, int64_t* n_verts)
{
    if (!c) { sz_set_error("sz_corner_mask: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_rows) { sz_set_error("sz_corner_mask: no contact rows on the device (run a contact step first)"); return SZ_ERR_STATE; }
    if (c->ext_mode) { sz_set_error("sz_corner_mask: single-GPU lists only"); return SZ_ERR_STATE; }
    if (count < 0 || nb_skip < 0 || (count > 0 && !floe_idx)) { sz_set_error("sz_corner_mask: bad arguments"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->have_cr = false; c->cr_count = count; c->cr_verts = 0;
    if (n_verts) *n_verts = 0;
    if (count == 0) { c->have_cr = true; return SZ_OK; }
    const int n0 = c->n0;
    {   // the numbers must name floes of the list
        std::vector<int32_t> h(count);
        CK(cudaMemcpy(h.data(), floe_idx, (size_t)count * 4, cudaMemcpyDefault));
        for (int32_t v : h) if (v < 1 || v > n0) { sz_set_error("sz_corner_mask: floe number %d outside 1..%d", v, n0); return SZ_ERR_ARG; }
        CK(c->cr_idx.ensure(count + 1));
        CK(cudaMemcpyAsync(c->cr_idx.p, h.data(), (size_t)count * 4, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st));
    }
    // ---- corners.m:13-51: originals, x images, then y images over the extended list (at most 4 n0 entries)
    const size_t ncap = 4 * (size_t)n0 + 1;
    CK(c->cr_ex.ensure(ncap)); CK(c->cr_ey.ensure(ncap)); CK(c->cr_esrc.ensure(ncap)); CK(c->cr_ealive.ensure(ncap));
    CK(c->flag.ensure(2 * (size_t)n0 + 2)); CK(c->pos.ensure(2 * (size_t)n0 + 2)); CK(c->scan_tmp.ensure(scan_tmp_ints(std::max(2 * (size_t)n0 + 2, (size_t)count + 2))));
    const double Lx = c->prm.Lx, Ly = c->prm.Ly;
    ++g_launches; corner_ext_init_kernel<<<nblk(n0, 256), 256, 0, st>>>(n0, c->x.p, c->y.p, c->alive.p, c->cr_ex.p, c->cr_ey.p, c->cr_esrc.p, c->cr_ealive.p);
    ++g_launches; ghost_flag_kernel<<<nblk(n0, 128), 128, 0, st>>>(0, n0, nullptr, c->cr_ex.p, c->cr_esrc.p, c->cr_ealive.p, c->voff.p, c->vx.p, Lx, c->flag.p);
    CKS(exclusive_scan(c, c->flag.p, n0, c->pos.p, n0 + 1));
    ++g_launches; corner_ext_emit_kernel<<<nblk(n0, 256), 256, 0, st>>>(0, n0, nullptr, c->flag.p, c->pos.p, c->cr_ex.p, c->cr_ey.p, c->cr_esrc.p, c->cr_ealive.p, Lx, D_CNT(cr_n1));
    ++g_launches; ghost_flag_kernel<<<nblk(2 * (i64)n0, 128), 128, 0, st>>>(1, 2 * n0, D_CNT(cr_n1), c->cr_ey.p, c->cr_esrc.p, c->cr_ealive.p, c->voff.p, c->vy.p, Ly, c->flag.p);
    CKS(exclusive_scan(c, c->flag.p, 2 * n0, c->pos.p, 2 * n0 + 1));
    ++g_launches; corner_ext_emit_kernel<<<nblk(2 * (i64)n0, 256), 256, 0, st>>>(1, 2 * n0, D_CNT(cr_n1), c->flag.p, c->pos.p, c->cr_ex.p, c->cr_ey.p, c->cr_esrc.p, c->cr_ealive.p, Ly, D_CNT(cr_n));
    // ---- one slot of da per polyshape vertex of every selected floe
    szcorn::CornerArgs a; memset(&a, 0, sizeof(a));
    a.count = count; a.idx = c->cr_idx.p; a.nb_skip = nb_skip;
    a.x = c->x.p; a.y = c->y.p; a.voff = c->voff.p; a.vx = c->vx.p; a.vy = c->vy.p;
    a.row_off = c->row_off.p; a.rows = c->rows.p;
    a.cex = c->cr_ex.p; a.cey = c->cr_ey.p; a.cesrc = c->cr_esrc.p; a.n_ext = D_CNT(cr_n);
    a.boxx = c->boxx.p; a.boxy = c->boxy.p; a.nbox = c->have_bnd ? c->boxn : 0;
    CK(c->cr_nv.ensure(count + 1)); CK(c->cr_off.ensure(count + 2));
    ++g_launches; corner_count_kernel<<<nblk(count, 256), 256, 0, st>>>(a, c->cr_nv.p);
    CKS(exclusive_scan(c, c->cr_nv.p, count, c->cr_off.p, count + 1));
    CK(cudaGetLastError());
    int total = 0;
    CK(cudaMemcpyAsync(&total, c->cr_off.p + count, 4, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st));
    if (total < 0) { sz_set_error("sz_corner_mask: more than 2^31 vertices selected"); return SZ_ERR_CAPACITY; }
    CK(c->cr_da.ensure((size_t)total + 1));
    a.da_off = c->cr_off.p; a.da = c->cr_da.p;
    // 8 lanes per floe for Voronoi-sized outlines, a warp per floe for real shapes
    const bool small = c->nverts <= 12 * (i64)n0;
    const int per_block = small ? 256 / 8 : 256 / 32;
    const int blocks = std::max(1, std::min(nblk(count, per_block), 148 * 16));
    ++g_launches;
    if (small) corner_mask_kernel<8><<<blocks, 256, 0, st>>>(a); else corner_mask_kernel<32><<<blocks, 256, 0, st>>>(a);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    c->cr_verts = total; c->have_cr = true;
    if (n_verts) *n_verts = total;
    return SZ_OK;
}
extern "C" int sz_get_corner_mask(SzContext* c, int64_t* da_off, uint8_t* da)
{
    if (!c) { sz_set_error("sz_get_corner_mask: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_cr) { sz_set_error("sz_get_corner_mask: no result"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->cr_count;
    if (da_off) {
        da_off[0] = 0;
        if (n) {
            std::vector<int> off(n + 1);
            CK(cudaMemcpy(off.data(), c->cr_off.p, (n + 1) * 4, cudaMemcpyDefault));
            for (size_t k = 0; k <= n; ++k) da_off[k] = off[k];
        }
    }
    D2H(da, c->cr_da.p, (size_t)c->cr_verts);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
// ------------------------------------------------------------------------------------------------ coarse-grid averages
// calc_eulerian_data.m:7-65, the floe list: dead floes dropped (:7-8), an x image for every floe with a vertex beyond +-Lx
// (:39-48), and the y pass as written (:56-65): it tests the polygon left over from the LAST iteration of the x loop, so
// either every entry (x images included) gets a y image or none does.
__global__ void euler_alive_kernel(int n0, const uint8_t* __restrict__ alive, int* __restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n0) flag[i] = alive[i] ? 1 : 0;
}
__global__ void euler_list_init_kernel(int n0, const int* __restrict__ flag, const int* __restrict__ pos, const double* __restrict__ x, const double* __restrict__ y,
                                       int* __restrict__ lsrc, double* __restrict__ lx, double* __restrict__ ly, int* __restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *n_out = pos[n0];
    if (i >= n0 || !flag[i]) return;
    const int g = pos[i];
    lsrc[g] = i; lx[g] = x[i]; ly[g] = y[i];
}
__global__ void euler_xflag_kernel(int n_bound, const int* __restrict__ n_dev, const int* __restrict__ lsrc, const double* __restrict__ lx, const double* __restrict__ ly,
                                   const int* __restrict__ voff, const double* __restrict__ vx, const double* __restrict__ vy, double Lx, double Ly,
                                   int* __restrict__ flag, int* __restrict__ yflag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_bound) return;
    const int n = *n_dev;
    int f = 0;
    if (i < n) {
        const int s = lsrc[i];
        double mx = 0, my = 0;
        for (int t = voff[s]; t < voff[s + 1]; ++t) { const double a = fabs(vx[t] + lx[i]), b = fabs(vy[t] + ly[i]); if (a > mx) mx = a; if (b > my) my = b; }
        f = (mx > Lx);
        if (i == n - 1) *yflag = (my > Ly) ? 1 : 0;
    }
    flag[i] = f;
}
__global__ void euler_xemit_kernel(int n_bound, const int* __restrict__ n_dev, const int* __restrict__ flag, const int* __restrict__ pos,
                                   int* __restrict__ lsrc, double* __restrict__ lx, double* __restrict__ ly, double Lx, int* __restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *n_dev;
    if (i == 0) *n_out = n + pos[n_bound];
    if (i >= n || !flag[i]) return;
    const int g = n + pos[i];
    lsrc[g] = lsrc[i]; lx[g] = lx[i] - 2 * Lx * sgn_d(lx[i]); ly[g] = ly[i];
}
__global__ void euler_yemit_kernel(int n_bound, const int* __restrict__ n_dev, const int* __restrict__ yflag,
                                   int* __restrict__ lsrc, double* __restrict__ lx, double* __restrict__ ly, double Ly, int* __restrict__ n_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *n_dev, yf = *yflag;
    if (i == 0) *n_out = yf ? 2 * n : n;
    if (i >= n_bound || i >= n || !yf) return;
    const int g = n + i;
    lsrc[g] = lsrc[i]; lx[g] = lx[i]; ly[g] = ly[i] - 2 * Ly * sgn_d(ly[i]);
}
// items: the candidate cells of every list entry (:113-119), entry by entry
__global__ void euler_item_count_kernel(const szeul::EulerArgs a, int* __restrict__ icnt)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.n_list) return;
    int i0, i1, j0, j1, c = 0;
    if (szeul::cell_range(a, q, i0, i1, j0, j1))
        for (int jj = j0; jj <= j1; ++jj) for (int ii = i0; ii <= i1; ++ii) c += szeul::is_candidate(a, q, ii, jj) ? 1 : 0;
    icnt[q] = c;
}
__global__ void euler_item_fill_kernel(const szeul::EulerArgs a, const int* __restrict__ ioff, int* __restrict__ item_cell, int* __restrict__ item_q, int* __restrict__ cell_cnt)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.n_list) return;
    int i0, i1, j0, j1, k = ioff[q];
    if (szeul::cell_range(a, q, i0, i1, j0, j1))
        for (int jj = j0; jj <= j1; ++jj) for (int ii = i0; ii <= i1; ++ii)
            if (szeul::is_candidate(a, q, ii, jj)) { const int cell = jj * a.g.Nx + ii; item_cell[k] = cell; item_q[k] = q; ++k; atomicAdd(&cell_cnt[cell], 1); }
}
__global__ void euler_iota_kernel(int n, int* __restrict__ v) { const int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < n) v[k] = k; }
__global__ void euler_status_kernel(int n, const int* __restrict__ status, int* __restrict__ n_fail, int* __restrict__ n_cap)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || status[k] == 0) return;
    atomicAdd(status[k] == szpf::PS_CLIPPER_FAIL ? n_fail : n_cap, 1);
}
__global__ void euler_cell_kernel(const szeul::EulerArgs a)      // fallback (option euler_cell_warp = 0): one thread per cell
{
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell < a.g.Nx * a.g.Ny) szeul::cell_reduce(a, cell);
}
// A warp per cell.  The reference adds a cell's items up in list order; on a coarse grid over a large field a cell has
// thousands of them and one thread would chase their scattered state one item at a time.  Here the 32 lanes compute the terms of
// 32 consecutive items at once (item_terms: the gathers and the products run in parallel), then the terms are added in item
// order: lane j owns sum j and takes term j of item l, l = 0, 1, ..., by shuffle.  Same terms, same order of additions as
// cell_reduce, so the same bits.
__global__ void __launch_bounds__(128) euler_cell_warp_kernel(const szeul::EulerArgs a)
{
    const int lane = threadIdx.x & 31;
    const int cell = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (cell >= a.g.Nx * a.g.Ny) return;                        // the whole warp leaves together
    const int k0 = a.cell_off[cell], k1 = a.cell_off[cell + 1];
    double acc = 0;                                             // sum number `lane` (lanes >= N_TERMS idle)
    for (int base = k0; base < k1; base += 32) {
        double t[szeul::N_TERMS];
#pragma unroll
        for (int j = 0; j < szeul::N_TERMS; ++j) t[j] = 0;
        bool valid = false;
        if (base + lane < k1) valid = szeul::item_terms(a, a.sorted[base + lane], t);
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        const int cnt = min(32, k1 - base);
        for (int l = 0; l < cnt; ++l) {
            const bool v = (vm >> l) & 1u;
#pragma unroll
            for (int j = 0; j < szeul::N_TERMS; ++j) {
                const double x = __shfl_sync(0xffffffffu, t[j], l);
                if (lane == j && (v || j == szeul::T_M0)) acc += x;
            }
        }
    }
    double S[szeul::N_TERMS];
#pragma unroll
    for (int j = 0; j < szeul::N_TERMS; ++j) S[j] = __shfl_sync(0xffffffffu, acc, j);
    if (lane == 0) szeul::cell_finalize(a, cell, S);
}

extern "C" int sz_eulerian_data(SzContext* c, int32_t Nx, int32_t Ny, double xmin, double xmax, double ymin, double ymax, int32_t periodic,
                                const double* mass, const double* overlap_area, const double* dUi_p, const double* dVi_p, const double* stress, const double* strain,
                                double* out)
{
    if (!c) { sz_set_error("sz_eulerian_data: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_input || c->ext_mode) { sz_set_error("sz_eulerian_data: upload the floes first (single-GPU list)"); return SZ_ERR_STATE; }
    if (Nx < 1 || Ny < 1 || (i64)Nx * Ny > (1 << 24) || !mass || !out) { sz_set_error("sz_eulerian_data: bad arguments (1 <= Nx*Ny <= 2^24, mass and out required)"); return SZ_ERR_ARG; }
    if (c->prm.Nb != 0) { sz_set_error("sz_eulerian_data: boundary floes (Nb > 0, calc_eulerian_data.m:11-25) are not supported"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int n0 = c->n0, cells = Nx * Ny;
    CK(c->eu_out.ensure((size_t)szeul::N_OUT * cells + 1));
    CK(cudaMemsetAsync(c->eu_out.p, 0, (size_t)szeul::N_OUT * cells * 8, st));
    auto finish = [&]() -> int { CK(cudaMemcpyAsync(out, c->eu_out.p, (size_t)szeul::N_OUT * cells * 8, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st)); return SZ_OK; };
    if (n0 == 0) return finish();
    // ---- per-floe state the caller supplies (host or device memory); NULL = zeros
    CK(c->eu_in.ensure(12 * (size_t)n0 + 1));
    CK(cudaMemsetAsync(c->eu_in.p, 0, 12 * (size_t)n0 * 8, st));
    double* d_mass = c->eu_in.p; double* d_over = d_mass + n0; double* d_dU = d_over + n0; double* d_dV = d_dU + n0; double* d_stress = d_dV + n0; double* d_strain = d_stress + 4 * (size_t)n0;
    CK(cudaMemcpyAsync(d_mass, mass, (size_t)n0 * 8, cudaMemcpyDefault, st));
    if (overlap_area) CK(cudaMemcpyAsync(d_over, overlap_area, (size_t)n0 * 8, cudaMemcpyDefault, st));
    if (dUi_p) CK(cudaMemcpyAsync(d_dU, dUi_p, (size_t)n0 * 8, cudaMemcpyDefault, st));
    if (dVi_p) CK(cudaMemcpyAsync(d_dV, dVi_p, (size_t)n0 * 8, cudaMemcpyDefault, st));
    if (stress) CK(cudaMemcpyAsync(d_stress, stress, 4 * (size_t)n0 * 8, cudaMemcpyDefault, st));
    if (strain) CK(cudaMemcpyAsync(d_strain, strain, 4 * (size_t)n0 * 8, cudaMemcpyDefault, st));
    // ---- the list
    const size_t lcap = 4 * (size_t)n0 + 1;
    CK(c->eu_lsrc.ensure(lcap)); CK(c->eu_lx.ensure(lcap)); CK(c->eu_ly.ensure(lcap));
    CK(c->flag.ensure(2 * (size_t)n0 + 2)); CK(c->pos.ensure(2 * (size_t)n0 + 2)); CK(c->scan_tmp.ensure(scan_tmp_ints(std::max<size_t>(lcap + 2, (size_t)cells + 2))));
    CK(cudaMemsetAsync(D_CNT(eu_n1), 0, 7 * sizeof(int), st));             // eu_n1 .. eu_cap
    ++g_launches; euler_alive_kernel<<<nblk(n0, 256), 256, 0, st>>>(n0, c->alive.p, c->flag.p);
    CKS(exclusive_scan(c, c->flag.p, n0, c->pos.p, n0 + 1));
    ++g_launches; euler_list_init_kernel<<<nblk(n0, 256), 256, 0, st>>>(n0, c->flag.p, c->pos.p, c->x.p, c->y.p, c->eu_lsrc.p, c->eu_lx.p, c->eu_ly.p, D_CNT(eu_n1));
    if (periodic) {
        ++g_launches; euler_xflag_kernel<<<nblk(n0, 128), 128, 0, st>>>(n0, D_CNT(eu_n1), c->eu_lsrc.p, c->eu_lx.p, c->eu_ly.p, c->voff.p, c->vx.p, c->vy.p, xmax, ymax, c->flag.p, D_CNT(eu_yflag));
        CKS(exclusive_scan(c, c->flag.p, n0, c->pos.p, n0 + 1));
        ++g_launches; euler_xemit_kernel<<<nblk(n0, 256), 256, 0, st>>>(n0, D_CNT(eu_n1), c->flag.p, c->pos.p, c->eu_lsrc.p, c->eu_lx.p, c->eu_ly.p, xmax, D_CNT(eu_n2));
        ++g_launches; euler_yemit_kernel<<<nblk(2 * (i64)n0, 256), 256, 0, st>>>(2 * n0, D_CNT(eu_n2), D_CNT(eu_yflag), c->eu_lsrc.p, c->eu_lx.p, c->eu_ly.p, ymax, D_CNT(eu_n));
    } else {
        CK(cudaMemcpyAsync(D_CNT(eu_n), D_CNT(eu_n1), sizeof(int), cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaGetLastError());
    CKS(read_counters(c));
    const int n_list = c->h_cnt->eu_n;
    if (n_list <= 0) return finish();
    // ---- items
    szeul::EulerArgs a; memset(&a, 0, sizeof(a));
    a.g.Nx = Nx; a.g.Ny = Ny; a.g.xmin = xmin; a.g.xmax = xmax; a.g.ymin = ymin; a.g.ymax = ymax;
    a.n_list = n_list; a.lsrc = c->eu_lsrc.p; a.lx = c->eu_lx.p; a.ly = c->eu_ly.p;
    a.rmax = c->rmax.p; a.area = c->area.p; a.h = c->h.p; a.u = c->u.p; a.v = c->v.p;
    a.mass = d_mass; a.over = d_over; a.dU = d_dU; a.dV = d_dV; a.stress = d_stress; a.strain = d_strain;
    a.voff = c->voff.p; a.vx = c->vx.p; a.vy = c->vy.p; a.out = c->eu_out.p;
    CK(c->eu_icnt.ensure((size_t)n_list + 2)); CK(c->eu_ioff.ensure((size_t)n_list + 2));
    CK(c->eu_ccnt.ensure((size_t)cells + 2)); CK(c->eu_coff.ensure((size_t)cells + 2));
    ++g_launches; euler_item_count_kernel<<<nblk(n_list, 128), 128, 0, st>>>(a, c->eu_icnt.p);
    CKS(exclusive_scan(c, c->eu_icnt.p, n_list, c->eu_ioff.p, n_list + 1));
    CK(cudaGetLastError());
    int n_items = 0;
    CK(cudaMemcpyAsync(&n_items, c->eu_ioff.p + n_list, 4, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st));
    if (n_items < 0) { sz_set_error("sz_eulerian_data: more than 2^31 (cell, floe) items"); return SZ_ERR_CAPACITY; }
    if (n_items == 0) return finish();
    CK(c->eu_cell.ensure((size_t)n_items + 1)); CK(c->eu_q.ensure((size_t)n_items + 1)); CK(c->eu_area.ensure((size_t)n_items + 1)); CK(c->eu_status.ensure((size_t)n_items + 1));
    CK(c->eu_iota.ensure((size_t)n_items + 1)); CK(c->eu_sorted.ensure((size_t)n_items + 1)); CK(c->eu_keys.ensure((size_t)n_items + 1)); CK(c->eu_listL.ensure((size_t)n_items + 1));
    CK(cudaMemsetAsync(c->eu_ccnt.p, 0, ((size_t)cells + 1) * 4, st));
    ++g_launches; euler_item_fill_kernel<<<nblk(n_list, 128), 128, 0, st>>>(a, c->eu_ioff.p, c->eu_cell.p, c->eu_q.p, c->eu_ccnt.p);
    a.n_items = n_items; a.item_cell = c->eu_cell.p; a.item_q = c->eu_q.p; a.item_area = c->eu_area.p; a.item_status = c->eu_status.p;
    CK(cudaMemsetAsync(c->eu_status.p, 0, (size_t)n_items * 4, st)); CK(cudaMemsetAsync(c->eu_area.p, 0, (size_t)n_items * 8, st));
    // ---- one clip per item: class S in local memory, whatever overflows its arena in class L
    ++g_launches; sz_launch_euler_S(&a, c->eu_listL.p, D_CNT(eu_listL), st);
    CK(cudaGetLastError());
    CKS(read_counters(c));
    const int nL = c->h_cnt->eu_listL;
    if (nL > 0) {
        const int threadsL = std::min((nL + 63) / 64 * 64, 148 * 16);
        CK(c->scratchL.ensure((size_t)threadsL * sz_workspace_bytes_L()));
        ++g_launches; sz_launch_euler_L(&a, c->eu_listL.p, D_CNT(eu_listL), c->scratchL.p, threadsL, st);
        CK(cudaGetLastError());
    }
    ++g_launches; euler_status_kernel<<<nblk(n_items, 256), 256, 0, st>>>(n_items, c->eu_status.p, D_CNT(eu_fail), D_CNT(eu_cap));
    // ---- items ordered by cell (stable: ascending list position inside a cell), cell offsets, reduction
    ++g_launches; euler_iota_kernel<<<nblk(n_items, 256), 256, 0, st>>>(n_items, c->eu_iota.p);
    int end_bit = 1; while ((1 << end_bit) < cells && end_bit < 31) ++end_bit;
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->eu_cell.p, c->eu_keys.p, c->eu_iota.p, c->eu_sorted.p, n_items, 0, end_bit, st));
    CK(c->eu_tmp.ensure(tmp_bytes + 16));
    ++g_launches; CK(cub::DeviceRadixSort::SortPairs(c->eu_tmp.p, tmp_bytes, c->eu_cell.p, c->eu_keys.p, c->eu_iota.p, c->eu_sorted.p, n_items, 0, end_bit, st));
    CKS(exclusive_scan(c, c->eu_ccnt.p, cells, c->eu_coff.p, cells + 1));
    a.sorted = c->eu_sorted.p; a.cell_off = c->eu_coff.p;
    CK(cudaGetLastError());
    CKS(read_counters(c));
    if (c->h_cnt->eu_fail > 0) { sz_set_error("Clipper Error. (%d cell-floe intersection(s) of sz_eulerian_data)", c->h_cnt->eu_fail); return SZ_ERR_CLIPPER; }
    if (c->h_cnt->eu_cap > 0) { sz_set_error("sz_eulerian_data: %d outline(s) exceed the largest size class", c->h_cnt->eu_cap); return SZ_ERR_CAPACITY; }
    ++g_launches;
    if (c->opt_euler_cell_warp) euler_cell_warp_kernel<<<nblk(32 * (i64)cells, 128), 128, 0, st>>>(a);
    else euler_cell_kernel<<<nblk(cells, 128), 128, 0, st>>>(a);
    CK(cudaGetLastError());
    return finish();
}

// reorder (start, count) pools into item-major CSR
static int export_paths(SzContext* c, int n_items, const int* d_item_start, const int* d_item_np, const int* d_status, const int* d_path_vstart, const int* d_path_len,
                        int n_pool_paths, const i64* d_px, const i64* d_py, int n_pool_verts,
                        int64_t* item_path_off, int64_t* path_vert_off, int64_t* ox, int64_t* oy)
{
    std::vector<int> is(n_items), in(n_items), stt(n_items, 0), pvs(n_pool_paths), pln(n_pool_paths);
    std::vector<i64> px(n_pool_verts), py(n_pool_verts);
    cudaStream_t st = c->stream;
    if (n_items) { CK(cudaMemcpyAsync(is.data(), d_item_start, (size_t)n_items * 4, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(in.data(), d_item_np, (size_t)n_items * 4, cudaMemcpyDefault, st));
                   if (d_status) CK(cudaMemcpyAsync(stt.data(), d_status, (size_t)n_items * 4, cudaMemcpyDefault, st)); }
    if (n_pool_paths) { CK(cudaMemcpyAsync(pvs.data(), d_path_vstart, (size_t)n_pool_paths * 4, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(pln.data(), d_path_len, (size_t)n_pool_paths * 4, cudaMemcpyDefault, st)); }
    if (n_pool_verts) { CK(cudaMemcpyAsync(px.data(), d_px, (size_t)n_pool_verts * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(py.data(), d_py, (size_t)n_pool_verts * 8, cudaMemcpyDefault, st)); }
    CK(cudaStreamSynchronize(st));
    i64 np = 0, nv = 0;
    if (item_path_off) item_path_off[0] = 0;
    if (path_vert_off) path_vert_off[0] = 0;
    for (int k = 0; k < n_items; ++k) {
        const int cnt = (stt[k] == 0) ? in[k] : 0;
        for (int q = 0; q < cnt; ++q) {
            const int pp = is[k] + q;
            for (int t = 0; t < pln[pp]; ++t) { if (ox) ox[nv] = px[pvs[pp] + t]; if (oy) oy[nv] = py[pvs[pp] + t]; ++nv; }
            ++np;
            if (path_vert_off) path_vert_off[np] = nv;
        }
        if (item_path_off) item_path_off[k + 1] = np;
    }
    return SZ_OK;
}
extern "C" int sz_get_clip_polys(SzContext* c, int64_t* pair_path_off, int64_t* path_vert_off, int64_t* x, int64_t* y)
{
    NEED_STEP("sz_get_clip_polys");
    if (!c->prm.want_clip_polys) { sz_set_error("sz_get_clip_polys: the step was run without want_clip_polys"); return SZ_ERR_STATE; }
    return export_paths(c, c->n_pairs, c->poly_path_start.p, c->poly_npaths.p, c->pstatus.p, c->path_vstart.p, c->path_len.p, (int)c->summary.n_clip_paths,
                        c->pvx.p, c->pvy.p, (int)c->summary.n_clip_verts, pair_path_off, path_vert_off, x, y);
}

// ------------------------------------------------------------------------------------------------ weld / FloeSimplify pair searches
// SURVEY.md 8f row f4 (second half): the bounding-radius searches of Physical_Processes/weld.m:29-81 and
// polygon_operations/FloeSimplify.m:13-31 on the resident floes, over the same kind of cell grid as K1.
//   weld      the floes after the first Nb are binned on an Nx x Ny grid (:31-48); floe i records every floe j OF ITS BIN with
//             alive(j) && d > 1 && d < rmax(i)+rmax(j), ascending j (:66-79)
//   simplify  the same predicate for each query floe against the whole list (FloeSimplify.m:20-31)
struct SearchArgs {
    int n0, first, weld, nq; const int* qidx; GridDesc g; double rmax_max;
    const double* x; const double* y; const double* rmax; const uint8_t* alive; const int* bin;
    const int* cell_start; const int* s_idx; const double* s_x; const double* s_y; const double* s_r;
    int* cnt; const int* off; int* out;
};
// bins of weld.m:35-36,40-48 (0 = in no bin), bounding box and largest rmax of the searched floes
__global__ void search_prep_kernel(int n0, int first, int weld, int Nx, int Ny, double xmin, double xmax, double ymin, double ymax,
                                   const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ rmax, int* __restrict__ bin, Counters* c)
{
    double xmn = SZ_INF, xmx = -SZ_INF, ymn = SZ_INF, ymx = -SZ_INF, rm = 0;
    for (int i = first + blockIdx.x * blockDim.x + threadIdx.x; i < n0; i += gridDim.x * blockDim.x) {
        const double X = x[i], Y = y[i];
        if (X == X && Y == Y) { xmn = fmin(xmn, X); xmx = fmax(xmx, X); ymn = fmin(ymn, Y); ymx = fmax(ymx, Y); }
        const double r = rmax[i]; if (r > rm) rm = r;
        int b = 0;
        if (weld) {
            const double a = trunc((X - xmin) / (xmax - xmin) * Nx + 1), bb = trunc((Y - ymin) / (ymax - ymin) * Ny + 1);     // fix(...)
            if (a >= 1 && a <= Nx && bb >= 1 && bb <= Ny) b = ((int)a - 1) * Ny + (int)bb;
        }
        bin[i] = b;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        xmn = fmin(xmn, __shfl_xor_sync(0xffffffffu, xmn, d)); xmx = fmax(xmx, __shfl_xor_sync(0xffffffffu, xmx, d));
        ymn = fmin(ymn, __shfl_xor_sync(0xffffffffu, ymn, d)); ymx = fmax(ymx, __shfl_xor_sync(0xffffffffu, ymx, d));
        rm = fmax(rm, __shfl_xor_sync(0xffffffffu, rm, d));
    }
    if ((threadIdx.x & 31) == 0) {
        if (xmn <= xmx) { atomicMin(&c->bbox[0], enc_d(xmn)); atomicMax(&c->bbox[1], enc_d(xmx)); atomicMin(&c->bbox[2], enc_d(ymn)); atomicMax(&c->bbox[3], enc_d(ymx)); }
        atomicMax(&c->rmax_bits, enc_d(rm));
    }
}
__global__ void search_cell_count_kernel(int n0, int first, GridDesc g, const double* __restrict__ x, const double* __restrict__ y, int* __restrict__ cid, int* __restrict__ cell_cnt)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0) return;
    const double X = x[i], Y = y[i];
    int c = -1;
    if (X == X && Y == Y) { c = cell_coord(Y, g.y0, g.cell, g.ny) * g.nx + cell_coord(X, g.x0, g.cell, g.nx); atomicAdd(&cell_cnt[c], 1); }
    cid[i] = c;
}
__global__ void search_cell_fill_kernel(int n0, int first, const int* __restrict__ cid, const int* __restrict__ cell_start, int* __restrict__ cell_pos,
                                        const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ rmax,
                                        int* __restrict__ s_idx, double* __restrict__ s_x, double* __restrict__ s_y, double* __restrict__ s_r)
{
    const int i = first + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n0) return;
    const int c = cid[i];
    if (c < 0) return;
    const int t = cell_start[c] + atomicAdd(&cell_pos[c], 1);
    s_idx[t] = i; s_x[t] = x[i]; s_y[t] = y[i]; s_r[t] = rmax[i];
}
// one warp per query floe, like broad_kernel: lanes stride over the cell rows around the floe, ballot + popc compaction
template <bool FILL>
__global__ void __launch_bounds__(256) search_kernel(const SearchArgs a)
{
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= a.nq) return;
    const int i = a.qidx ? a.qidx[q] : a.first + q;
    const double xi = a.x[i], yi = a.y[i];
    const int bi = a.weld ? a.bin[i] : 1;
    int count = 0;
    const int off = FILL ? a.off[q] : 0;
    if (xi == xi && yi == yi && bi != 0) {
        const double ri = a.rmax[i];
        const int cxi = cell_coord(xi, a.g.x0, a.g.cell, a.g.nx), cyi = cell_coord(yi, a.g.y0, a.g.cell, a.g.ny);
        const int R = (int)((ri + a.rmax_max) / a.g.cell) + 1;
        const int cx0 = cxi - R > 0 ? cxi - R : 0, cx1 = cxi + R < a.g.nx ? cxi + R : a.g.nx - 1;
        for (int cy = (cyi - R > 0 ? cyi - R : 0); cy <= cyi + R && cy < a.g.ny; ++cy) {
            const int t0 = a.cell_start[cy * a.g.nx + cx0], t1 = a.cell_start[cy * a.g.nx + cx1 + 1];
            for (int tb = t0; tb < t1; tb += 32) {
                const int t = tb + lane;
                bool ok = false; int j = -1;
                if (t < t1) {
                    j = a.s_idx[t];
                    const double dx = xi - a.s_x[t], dy = yi - a.s_y[t];
                    const double d = sqrt(dx * dx + dy * dy);
                    ok = a.alive[j] && d > 1 && d < (ri + a.s_r[t]) && (!a.weld || a.bin[j] == bi);
                }
                const unsigned m = __ballot_sync(0xffffffffu, ok);
                if (FILL && ok) a.out[off + count + __popc(m & ((1u << lane) - 1))] = j;
                count += __popc(m);
            }
        }
    }
    if (!FILL && lane == 0) a.cnt[q] = count;
}
// ascending order inside every segment (any length): each lane ranks its elements against the whole segment
__global__ void __launch_bounds__(256) segment_rank_sort_kernel(int nseg, const int* __restrict__ off, const int* __restrict__ in, int* __restrict__ out, int base)
{
    const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (q >= nseg) return;
    const int o = off[q], c = off[q + 1] - o;
    for (int k = lane; k < c; k += 32) {
        const int v = in[o + k];
        int rank = 0;
        for (int m = 0; m < c; ++m) rank += (in[o + m] < v);
        out[o + rank] = v - base + 1;
    }
}
extern "C" int sz_pair_search(SzContext* c, int32_t mode, int32_t Nb, int32_t Nx, int32_t Ny, double xmin, double xmax, double ymin, double ymax,
                              int32_t count, const int32_t* idx, int64_t* n_partners)
{
    if (!c) { sz_set_error("sz_pair_search: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_input || c->ext_mode) { sz_set_error("sz_pair_search: upload the floes first (single-GPU list)"); return SZ_ERR_STATE; }
    const bool weld = mode == 0;
    if (mode != 0 && mode != 1) { sz_set_error("sz_pair_search: mode must be 0 (weld) or 1 (FloeSimplify)"); return SZ_ERR_ARG; }
    if (weld && (Nb < 0 || Nx < 1 || Ny < 1 || !(xmax > xmin) || !(ymax > ymin))) { sz_set_error("sz_pair_search: bad Nb / grid"); return SZ_ERR_ARG; }
    if (!weld && (count < 0 || (count > 0 && !idx))) { sz_set_error("sz_pair_search: idx missing"); return SZ_ERR_ARG; }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream; const int n0 = c->n0;
    const int first = weld ? std::min(Nb, n0) : 0;
    const int nq = weld ? n0 - first : count;
    c->have_search = false; c->se_nq = nq; c->se_np = 0; c->se_weld = weld; c->se_first = first;
    CK(c->se_bin.ensure(n0 + 1)); CK(c->se_cnt.ensure(nq + 2)); CK(c->se_off.ensure(nq + 2)); CK(c->se_q.ensure(nq + 1));
    if (!weld && nq > 0) {
        std::vector<int> q0(nq);
        cudaPointerAttributes pa; bool host = true;
        if (cudaPointerGetAttributes(&pa, idx) == cudaSuccess) host = (pa.type == cudaMemoryTypeUnregistered || pa.type == cudaMemoryTypeHost); else cudaGetLastError();
        std::vector<int> tmp(nq);
        if (host) memcpy(tmp.data(), idx, (size_t)nq * 4); else CK(cudaMemcpy(tmp.data(), idx, (size_t)nq * 4, cudaMemcpyDefault));
        for (int k = 0; k < nq; ++k) { if (tmp[k] < 1 || tmp[k] > n0) { sz_set_error("sz_pair_search: floe number %d out of range", tmp[k]); return SZ_ERR_ARG; } q0[k] = tmp[k] - 1; }
        CK(cudaMemcpyAsync(c->se_q.p, q0.data(), (size_t)nq * 4, cudaMemcpyDefault, st)); CK(cudaStreamSynchronize(st));
    }
    if (n_partners) *n_partners = 0;
    if (nq == 0 || n0 - first <= 0) { CK(cudaMemsetAsync(c->se_off.p, 0, (size_t)(nq + 1) * 4, st)); CK(cudaStreamSynchronize(st)); c->have_search = true; return SZ_OK; }
    {
        Counters init; memset(&init, 0, sizeof(init));
        init.bbox[0] = enc_d(SZ_INF); init.bbox[1] = enc_d(-SZ_INF); init.bbox[2] = enc_d(SZ_INF); init.bbox[3] = enc_d(-SZ_INF); init.rmax_bits = enc_d(0.0);
        *c->h_cnt = init;
        CK(cudaMemcpyAsync(c->d_cnt, c->h_cnt, sizeof(Counters), cudaMemcpyDefault, st));
    }
    const int ns = n0 - first;
    ++g_launches; search_prep_kernel<<<std::min(nblk(ns, 256), 148 * 8), 256, 0, st>>>(n0, first, weld ? 1 : 0, Nx, Ny, xmin, xmax, ymin, ymax, c->x.p, c->y.p, c->rmax.p, c->se_bin.p, c->d_cnt);
    CK(cudaGetLastError());
    CKS(read_counters(c));
    GridDesc g; g.x0 = g.y0 = 0; g.cell = 1; g.nx = g.ny = 1;
    const double rm = dec_d(c->h_cnt->rmax_bits);
    {
        const double bx0 = dec_d(c->h_cnt->bbox[0]), bx1 = dec_d(c->h_cnt->bbox[1]), by0 = dec_d(c->h_cnt->bbox[2]), by1 = dec_d(c->h_cnt->bbox[3]);
        if (bx0 <= bx1 && std::isfinite(bx0) && std::isfinite(bx1) && std::isfinite(by0) && std::isfinite(by1)) {
            g.x0 = bx0; g.y0 = by0;
            double cell = 2 * rm; if (!(cell > 0) || !std::isfinite(cell)) cell = 1;
            const double dens = std::sqrt(4.0 * (bx1 - bx0) * (by1 - by0) / std::max(1, ns));
            if (dens > 0 && std::isfinite(dens)) cell = std::min(cell, std::max(cell / 8, dens));
            while ((bx1 - bx0) / cell * ((by1 - by0) / cell) > 1.6e7) cell *= 2;
            g.cell = cell; g.nx = (int)((bx1 - bx0) / cell) + 1; g.ny = (int)((by1 - by0) / cell) + 1;
        }
    }
    const int ncell = g.nx * g.ny;
    CK(c->se_cid.ensure(n0 + 1)); CK(c->se_ccnt.ensure(ncell + 2)); CK(c->se_cstart.ensure(ncell + 2));
    CK(c->se_sidx.ensure(n0 + 1)); CK(c->se_sx.ensure(n0 + 1)); CK(c->se_sy.ensure(n0 + 1)); CK(c->se_sr.ensure(n0 + 1));
    CK(c->scan_tmp.ensure(scan_tmp_ints(std::max<size_t>((size_t)ncell + 2, (size_t)nq + 2))));
    CK(cudaMemsetAsync(c->se_ccnt.p, 0, (size_t)(ncell + 1) * 4, st));
    ++g_launches; search_cell_count_kernel<<<nblk(ns, 256), 256, 0, st>>>(n0, first, g, c->x.p, c->y.p, c->se_cid.p, c->se_ccnt.p);
    CKS(exclusive_scan(c, c->se_ccnt.p, ncell, c->se_cstart.p, ncell + 1));
    CK(cudaMemsetAsync(c->se_ccnt.p, 0, (size_t)(ncell + 1) * 4, st));
    ++g_launches; search_cell_fill_kernel<<<nblk(ns, 256), 256, 0, st>>>(n0, first, c->se_cid.p, c->se_cstart.p, c->se_ccnt.p, c->x.p, c->y.p, c->rmax.p, c->se_sidx.p, c->se_sx.p, c->se_sy.p, c->se_sr.p);
    SearchArgs a; memset(&a, 0, sizeof(a));
    a.n0 = n0; a.first = first; a.weld = weld ? 1 : 0; a.nq = nq; a.qidx = weld ? nullptr : c->se_q.p; a.g = g; a.rmax_max = rm;
    a.x = c->x.p; a.y = c->y.p; a.rmax = c->rmax.p; a.alive = c->alive.p; a.bin = c->se_bin.p;
    a.cell_start = c->se_cstart.p; a.s_idx = c->se_sidx.p; a.s_x = c->se_sx.p; a.s_y = c->se_sy.p; a.s_r = c->se_sr.p;
    a.cnt = c->se_cnt.p; a.off = c->se_off.p;
    ++g_launches; search_kernel<false><<<nblk(32 * (i64)nq, 256), 256, 0, st>>>(a);
    CKS(exclusive_scan(c, c->se_cnt.p, nq, c->se_off.p, nq + 1));
    CK(cudaMemcpyAsync(D_CNT(n_pairs), c->se_off.p + nq, 4, cudaMemcpyDeviceToDevice, st));
    CK(cudaGetLastError());
    CKS(read_counters(c));
    const int np = c->h_cnt->n_pairs;
    CK(c->se_tmp.ensure(np + 1)); CK(c->se_out.ensure(np + 1));
    if (np > 0) {
        a.out = c->se_tmp.p;
        ++g_launches; search_kernel<true><<<nblk(32 * (i64)nq, 256), 256, 0, st>>>(a);
        ++g_launches; segment_rank_sort_kernel<<<nblk(32 * (i64)nq, 256), 256, 0, st>>>(nq, c->se_off.p, c->se_tmp.p, c->se_out.p, first);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    c->se_np = np; c->have_search = true;
    if (n_partners) *n_partners = np;
    return SZ_OK;
}
extern "C" int sz_get_pair_search(SzContext* c, int32_t* bin, int64_t* off, int32_t* partner)
{
    if (!c) { sz_set_error("sz_get_pair_search: NULL context"); return SZ_ERR_ARG; }
    if (!c->have_search) { sz_set_error("sz_get_pair_search: no search has been run"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    const int nq = c->se_nq;
    if (bin) { if (c->se_weld) D2H(bin, c->se_bin.p + c->se_first, (size_t)nq * 4); else memset(bin, 0, (size_t)nq * 4); }
    if (off) {
        std::vector<int> tmp(nq + 1);
        CK(cudaMemcpyAsync(tmp.data(), c->se_off.p, (size_t)(nq + 1) * 4, cudaMemcpyDefault, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        for (int k = 0; k <= nq; ++k) off[k] = tmp[k];
    }
    D2H(partner, c->se_out.p, (size_t)c->se_np * 4);
    CK(cudaStreamSynchronize(c->stream));
    return SZ_OK;
}
// ------------------------------------------------------------------------------------------------ clip batch
extern "C" int sz_clip_batch(SzContext* c, int32_t count, const int32_t* method, const int64_t* soff, const int64_t* sx, const int64_t* sy,
                             const int64_t* coff, const int64_t* cx, const int64_t* cy, int64_t* n_paths, int64_t* n_verts)
{
    if (!c || count < 0 || (count > 0 && (!method || !soff || !coff || !sx || !sy || !cx || !cy))) { sz_set_error("sz_clip_batch: bad arguments"); return SZ_ERR_ARG; }
    for (int k = 0; k < count; ++k) {
        if (method[k] < 0 || method[k] > 3) { sz_set_error("sz_clip_batch: method must be 0..3 (mexclipper.cpp:206-230)"); return SZ_ERR_ARG; }
        if (soff[k + 1] < soff[k] || coff[k + 1] < coff[k]) { sz_set_error("sz_clip_batch: offsets must be non-decreasing"); return SZ_ERR_ARG; }
    }
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    c->clip_count = -1;
    const size_t ns = count ? (size_t)soff[count] : 0, nc = count ? (size_t)coff[count] : 0;
    CK(c->c_method.ensure(count + 1)); CK(c->c_status.ensure(count + 1)); CK(c->c_path_start.ensure(count + 1)); CK(c->c_npaths.ensure(count + 1));
    CK(c->c_listM.ensure(count + 1)); CK(c->c_listL.ensure(count + 1));
    CK(c->c_soff.ensure(count + 2)); CK(c->c_coff.ensure(count + 2)); CK(c->c_sx.ensure(ns + 1)); CK(c->c_sy.ensure(ns + 1)); CK(c->c_cx.ensure(nc + 1)); CK(c->c_cy.ensure(nc + 1));
    if (count > 0) {
        CK(cudaMemcpyAsync(c->c_method.p, method, (size_t)count * 4, cudaMemcpyDefault, st));
        CK(cudaMemcpyAsync(c->c_soff.p, soff, (size_t)(count + 1) * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->c_coff.p, coff, (size_t)(count + 1) * 8, cudaMemcpyDefault, st));
        if (ns) { CK(cudaMemcpyAsync(c->c_sx.p, sx, ns * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->c_sy.p, sy, ns * 8, cudaMemcpyDefault, st)); }
        if (nc) { CK(cudaMemcpyAsync(c->c_cx.p, cx, nc * 8, cudaMemcpyDefault, st)); CK(cudaMemcpyAsync(c->c_cy.p, cy, nc * 8, cudaMemcpyDefault, st)); }
    }
    CK(c->c_path_vstart.ensure((size_t)count * 2 + 64)); CK(c->c_path_len.ensure(c->c_path_vstart.cap));
    CK(c->c_pvx.ensure(ns + nc + 1024)); CK(c->c_pvy.ensure(c->c_pvx.cap));
    for (int attempt = 0; attempt < 3 && count > 0; ++attempt) {
        CK(cudaMemsetAsync(c->d_cnt, 0, sizeof(Counters), st));
        CK(cudaMemsetAsync(c->c_status.p, 0, (size_t)count * 4, st)); CK(cudaMemsetAsync(c->c_npaths.p, 0, (size_t)count * 4, st));
        ClipArgs a; memset(&a, 0, sizeof(a));
        a.count = count; a.method = c->c_method.p; a.soff = c->c_soff.p; a.sx = c->c_sx.p; a.sy = c->c_sy.p; a.coff = c->c_coff.p; a.cx = c->c_cx.p; a.cy = c->c_cy.p;
        a.status = c->c_status.p; a.item_path_start = c->c_path_start.p; a.item_npaths = c->c_npaths.p;
        a.path_vstart = c->c_path_vstart.p; a.path_len = c->c_path_len.p; a.path_cap = (int)c->c_path_vstart.cap; a.path_used = D_CNT(clip_path_used);
        a.pvx = c->c_pvx.p; a.pvy = c->c_pvy.p; a.vert_cap = (int)c->c_pvx.cap; a.vert_used = D_CNT(clip_vert_used);
        a.next_list = c->c_listM.p; a.next_count = D_CNT(clip_listM);
        ++g_launches; sz_launch_clip_S(&a, st);
        CK(cudaGetLastError());
        CKS(read_counters(c));
        if (c->h_cnt->clip_listM > 0) {
            const int threads = std::min((c->h_cnt->clip_listM + 63) / 64 * 64, 148 * 512);
            CK(c->scratchM.ensure((size_t)threads * sz_workspace_bytes_M()));
            a.list = c->c_listM.p; a.list_count = D_CNT(clip_listM); a.next_list = c->c_listL.p; a.next_count = D_CNT(clip_listL); a.scratch = c->scratchM.p; a.n_threads = threads;
            ++g_launches; sz_launch_clip_M(&a, st);
            CK(cudaGetLastError());
            CKS(read_counters(c));
            if (c->h_cnt->clip_listL > 0) {
                const int threadsL = std::min((c->h_cnt->clip_listL + 63) / 64 * 64, 148 * 32);
                CK(c->scratchL.ensure((size_t)threadsL * sz_workspace_bytes_L()));
                a.list = c->c_listL.p; a.list_count = D_CNT(clip_listL); a.next_list = nullptr; a.next_count = nullptr; a.scratch = c->scratchL.p; a.n_threads = threadsL;
                ++g_launches; sz_launch_clip_L(&a, st);
                CK(cudaGetLastError());
                CKS(read_counters(c));
            }
        }
        bool again = false;
        if ((size_t)c->h_cnt->clip_path_used > c->c_path_vstart.cap) { CK(c->c_path_vstart.ensure(c->h_cnt->clip_path_used + 64)); CK(c->c_path_len.ensure(c->c_path_vstart.cap)); again = true; }
        if ((size_t)c->h_cnt->clip_vert_used > c->c_pvx.cap) { CK(c->c_pvx.ensure(c->h_cnt->clip_vert_used + 64)); CK(c->c_pvy.ensure(c->c_pvx.cap)); again = true; }
        if (!again) break;
        if (attempt == 2) { sz_set_error("sz_clip_batch: result pools kept overflowing"); return SZ_ERR_CAPACITY; }
    }
    c->clip_count = count;
    c->clip_paths = count ? c->h_cnt->clip_path_used : 0; c->clip_verts = count ? c->h_cnt->clip_vert_used : 0;
    if (n_paths) *n_paths = c->clip_paths;
    if (n_verts) *n_verts = c->clip_verts;
    return SZ_OK;
}
extern "C" int sz_get_clip_batch(SzContext* c, int32_t* status, int64_t* item_path_off, int64_t* path_vert_off, int64_t* ox, int64_t* oy)
{
    if (!c) { sz_set_error("sz_get_clip_batch: NULL context"); return SZ_ERR_ARG; }
    if (c->clip_count < 0) { sz_set_error("sz_get_clip_batch: no clip batch has been run"); return SZ_ERR_STATE; }
    CK(cudaSetDevice(c->device));
    if (status && c->clip_count > 0) { CK(cudaMemcpyAsync(status, c->c_status.p, (size_t)c->clip_count * 4, cudaMemcpyDefault, c->stream)); CK(cudaStreamSynchronize(c->stream)); }
    return export_paths(c, c->clip_count, c->c_path_start.p, c->c_npaths.p, c->c_status.p, c->c_path_vstart.p, c->c_path_len.p, (int)c->clip_paths,
                        c->c_pvx.p, c->c_pvy.p, (int)c->clip_verts, item_path_off, path_vert_off, ox, oy);
}
