// sz_contact_mex.cpp -- thin mex gateway from MATLAB to the C ABI of include/subzero_b200.h.
//
// Modelled on the reference's own gateway private/mexclipper.cpp: one mexFunction (:83), inputs type-checked
// before use (:22-41), outputs created with mxCreate* and handed to MATLAB (:65-81), failures raised through
// mexErrMsg* (:304).  Where the reference gateway is called >= 3 times per contacting pair (polyclip.m:73), this
// one is called once per timestep:
//
//   out = sz_contact_mex(prm, soa [, bnd])
//     prm : struct of doubles  Lx Ly modulus dt Nb periodic collision  (+ optional nu mu ... to override defaults)
//     soa : struct of double column vectors x y rmax h area u v ksi alive (length N), vx vy (length V, c_alpha
//           concatenated, closed outlines) and voff (length N+1, 0-based offsets)
//     bnd : struct x y (hole vertices of floebound.poly), box_x box_y (c2_boundary), area h  -- non-periodic runs
//     out : struct  fx fy torque overlap_area (N x 1), stress (4 x N), xi yi alive kill transfer (N x 1),
//                   row_off (Next+1 x 1), rows (7 x K, one contact row per column), n_ext, n_pairs, collision_count,
//                   ghost_parent ghost_x ghost_y ghost_fx ghost_fy ghost_torque ghost_overlap_area (Next-N x 1): the
//                   periodic images Floe(N+1:Next) of floe_interactions_all.m:16-66 -- parent = 1-based index in the
//                   extended list, shifted centroid, and what :218-238 leaves in their structs
//
// Build (MATLAB):  mex -I../../include sz_contact_mex.cpp -L../_lib -lsubzero_b200
// The context (device buffers, stream) persists between calls and is released by mexAtExit.
#include "mex.h"
#include "subzero_b200.h"
#include <vector>
#include <string>
#include <cstring>

static SzContext* g_ctx = nullptr;
static void release_ctx() { if (g_ctx) { sz_destroy(g_ctx); g_ctx = nullptr; } }

static const mxArray* need_field(const mxArray* s, const char* name, size_t min_len)
{
    if (!mxIsStruct(s)) mexErrMsgIdAndTxt("subzero_b200:arg", "argument must be a struct");
    const mxArray* f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("subzero_b200:arg", "missing field '%s'", name);
    if (!mxIsDouble(f) || mxIsComplex(f)) mexErrMsgIdAndTxt("subzero_b200:arg", "field '%s' must be a real double array", name);
    if (mxGetNumberOfElements(f) < min_len) mexErrMsgIdAndTxt("subzero_b200:arg", "field '%s' is too short", name);
    return f;
}
static double scalar_field(const mxArray* s, const char* name, double dflt, bool required)
{
    const mxArray* f = mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr;
    if (!f) { if (required) mexErrMsgIdAndTxt("subzero_b200:arg", "missing parameter '%s'", name); return dflt; }
    return mxGetScalar(f);
}
static void fail(int code)
{
    // no device resource is owned by this frame: the context outlives the call, so the longjmp is safe
    mexErrMsgIdAndTxt(code == SZ_ERR_CLIPPER ? "subzero_b200:clipper" : "subzero_b200:error", "%s", sz_last_error());
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    if (nrhs < 2 || nrhs > 3) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: out = sz_contact_mex(prm, soa [, bnd])");
    if (nlhs > 1) mexErrMsgIdAndTxt("subzero_b200:arg", "one output");
    // every argument is checked before the device is touched (mexclipper.cpp:22-41 checks before it converts)
    SzParams P; sz_default_params(&P);
    const mxArray* p = prhs[0];
    P.Lx = scalar_field(p, "Lx", 0, true); P.Ly = scalar_field(p, "Ly", 0, true);
    P.modulus = scalar_field(p, "modulus", 0, true); P.dt = scalar_field(p, "dt", 0, true);
    P.Nb = (int32_t)scalar_field(p, "Nb", 0, false);
    P.pair_with_boundary_floes = scalar_field(p, "pair_with_boundary_floes", 0, false) != 0;      // opt-in, not reference behaviour
    P.periodic = scalar_field(p, "periodic", 0, true) != 0; P.collision = scalar_field(p, "collision", 1, false) != 0;
    P.nu = scalar_field(p, "nu", P.nu, false); P.mu = scalar_field(p, "mu", P.mu, false);
    P.merge_frac = scalar_field(p, "merge_frac", P.merge_frac, false); P.wall_frac = scalar_field(p, "wall_frac", P.wall_frac, false);

    const mxArray* s = prhs[1];
    const size_t n = mxGetNumberOfElements(need_field(s, "x", 0));
    SzFloesSoA F; std::memset(&F, 0, sizeof(F));
    F.n = (int32_t)n;
    F.x = mxGetPr(need_field(s, "x", n)); F.y = mxGetPr(need_field(s, "y", n)); F.rmax = mxGetPr(need_field(s, "rmax", n));
    F.h = mxGetPr(need_field(s, "h", n)); F.area = mxGetPr(need_field(s, "area", n)); F.u = mxGetPr(need_field(s, "u", n));
    F.v = mxGetPr(need_field(s, "v", n)); F.ksi = mxGetPr(need_field(s, "ksi", n));
    const double* alive_d = mxGetPr(need_field(s, "alive", n));
    const double* voff_d = mxGetPr(need_field(s, "voff", n + 1));
    std::vector<uint8_t> alive(n); std::vector<int32_t> voff(n + 1);
    for (size_t i = 0; i < n; ++i) alive[i] = alive_d[i] != 0;
    for (size_t i = 0; i <= n; ++i) voff[i] = (int32_t)voff_d[i];
    F.nverts = n ? voff[n] : 0;
    F.vx = mxGetPr(need_field(s, "vx", (size_t)F.nverts)); F.vy = mxGetPr(need_field(s, "vy", (size_t)F.nverts));
    F.alive = alive.data(); F.voff = voff.data();

    SzBoundary B; std::memset(&B, 0, sizeof(B)); const SzBoundary* pB = nullptr;
    if (nrhs == 3 && !mxIsEmpty(prhs[2])) {
        const mxArray* b = prhs[2];
        B.n = (int32_t)mxGetNumberOfElements(need_field(b, "x", 3)); B.x = mxGetPr(need_field(b, "x", 3)); B.y = mxGetPr(need_field(b, "y", (size_t)B.n));
        B.box_n = (int32_t)mxGetNumberOfElements(need_field(b, "box_x", 3)); B.box_x = mxGetPr(need_field(b, "box_x", 3)); B.box_y = mxGetPr(need_field(b, "box_y", (size_t)B.box_n));
        B.area = scalar_field(b, "area", 0, true); B.h = scalar_field(b, "h", 0, false);
        // floebound's own kinematics (floe_interactions.m:109-110 reads floe2.Ui / Vi / ksi_ice / Xi / Yi of the boundary too); zero when absent
        B.xi = scalar_field(b, "xi", 0, false); B.yi = scalar_field(b, "yi", 0, false);
        B.u = scalar_field(b, "u", 0, false); B.v = scalar_field(b, "v", 0, false); B.ksi = scalar_field(b, "ksi", 0, false);
        pB = &B;
    }
    if (!g_ctx) {
        if (sz_create(&g_ctx, 0) != SZ_OK) fail(SZ_ERR_CUDA);
        mexAtExit(release_ctx);
    }
    SzSummary S;
    int rc = sz_contact_step(g_ctx, &P, &F, pB, &S);
    if (rc != SZ_OK) fail(rc);

    const char* names[] = {"fx", "fy", "torque", "overlap_area", "stress", "xi", "yi", "alive", "kill", "transfer", "row_off", "rows", "n_ext", "n_pairs", "collision_count",
                           "ghost_parent", "ghost_x", "ghost_y", "ghost_fx", "ghost_fy", "ghost_torque", "ghost_overlap_area"};
    mxArray* out = mxCreateStructMatrix(1, 1, 22, names);
    mxArray *fx = mxCreateDoubleMatrix(n, 1, mxREAL), *fy = mxCreateDoubleMatrix(n, 1, mxREAL), *tq = mxCreateDoubleMatrix(n, 1, mxREAL), *ov = mxCreateDoubleMatrix(n, 1, mxREAL);
    mxArray *st = mxCreateDoubleMatrix(4, n, mxREAL), *xi = mxCreateDoubleMatrix(n, 1, mxREAL), *yi = mxCreateDoubleMatrix(n, 1, mxREAL);
    std::vector<uint8_t> al(n); std::vector<int32_t> kill(n), transfer(n);
    rc = sz_get_floe_outputs(g_ctx, mxGetPr(fx), mxGetPr(fy), mxGetPr(tq), mxGetPr(ov), mxGetPr(st), mxGetPr(xi), mxGetPr(yi), al.data(), kill.data(), transfer.data());
    if (rc != SZ_OK) fail(rc);
    mxArray *ma = mxCreateDoubleMatrix(n, 1, mxREAL), *mk = mxCreateDoubleMatrix(n, 1, mxREAL), *mt = mxCreateDoubleMatrix(n, 1, mxREAL);
    for (size_t i = 0; i < n; ++i) { mxGetPr(ma)[i] = al[i]; mxGetPr(mk)[i] = kill[i]; mxGetPr(mt)[i] = transfer[i]; }
    std::vector<int64_t> off((size_t)S.n + 1);
    mxArray* rows = mxCreateDoubleMatrix(7, (size_t)S.n_rows, mxREAL);
    rc = sz_get_rows(g_ctx, off.data(), mxGetPr(rows));
    if (rc != SZ_OK) fail(rc);
    mxArray* ro = mxCreateDoubleMatrix((size_t)S.n + 1, 1, mxREAL);
    for (size_t i = 0; i <= (size_t)S.n; ++i) mxGetPr(ro)[i] = (double)off[i];
    const size_t ng = (size_t)(S.n - S.n0);
    mxArray *gp = mxCreateDoubleMatrix(ng, 1, mxREAL), *gx = mxCreateDoubleMatrix(ng, 1, mxREAL), *gy = mxCreateDoubleMatrix(ng, 1, mxREAL);
    mxArray *gfx = mxCreateDoubleMatrix(ng, 1, mxREAL), *gfy = mxCreateDoubleMatrix(ng, 1, mxREAL), *gtq = mxCreateDoubleMatrix(ng, 1, mxREAL), *gov = mxCreateDoubleMatrix(ng, 1, mxREAL);
    if (ng > 0) {
        std::vector<int32_t> par(ng), fnum(ng);
        rc = sz_get_ghosts(g_ctx, par.data(), fnum.data(), mxGetPr(gx), mxGetPr(gy));
        if (rc != SZ_OK) fail(rc);
        rc = sz_get_ghost_outputs(g_ctx, mxGetPr(gfx), mxGetPr(gfy), mxGetPr(gtq), mxGetPr(gov));
        if (rc != SZ_OK) fail(rc);
        for (size_t i = 0; i < ng; ++i) mxGetPr(gp)[i] = par[i];
    }
    mxArray* vals[] = {fx, fy, tq, ov, st, xi, yi, ma, mk, mt, ro, rows, mxCreateDoubleScalar(S.n), mxCreateDoubleScalar((double)S.n_pairs), mxCreateDoubleScalar(S.collision_count),
                       gp, gx, gy, gfx, gfy, gtq, gov};
    for (int k = 0; k < 22; ++k) mxSetFieldByNumber(out, 0, k, vals[k]);
    plhs[0] = out;
}
