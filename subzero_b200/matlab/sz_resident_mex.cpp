// sz_resident_mex.cpp -- mex gateway for the device-resident form of the C ABI (include/subzero_b200.h): the floe
// state is uploaded when the topology changes, the contact step, the ocean/atmosphere forcing, the integrator of
// calc_trajectory.m and the deformation step of fracture_floe.m run on the device, and MATLAB fetches only what the
// code that stays in MATLAB needs (per-floe state every step, contact rows when corners / fracture are due).
//
// Same conventions as sz_contact_mex.cpp (modelled on the reference's private/mexclipper.cpp:13-81,83,303-304): one
// mexFunction, inputs type-checked before use, outputs created with mxCreate*, failures raised through mexErrMsg*,
// the context persists between calls and is released by mexAtExit.  First argument: a command string.
//
//   sz_resident_mex('upload', prm, soa [, bnd])        Floe -> device (structs as in sz_contact_mex)
//   s   = sz_resident_mex('step')                      one contact step; s: n_ext n_pairs n_rows collision_count ms
//   out = sz_resident_mex('floe_outputs')              fx fy torque overlap_area stress(4xN) xi yi alive kill transfer
//   r   = sz_resident_mex('rows')                      row_off (Next+1), rows (7 x K): Floe(i).interactions
//   sz_resident_mex('trajectory_init', st)             st: mass inertia [alpha dXi_p dYi_p dUi_p dVi_p dalpha_p dksi_p
//                                                      FxOA FyOA torqueOA (N x 1)] [c0x c0y (V x 1)] nz
//   sz_resident_mex('set_ocean', ocean, winds)         ocean: Xo Yo Uocn Vocn fCoriolis turn_angle; winds: u v
//   sz_resident_mex('set_points', X, Y, A)             Floe.X/.Y/.A as npts x N matrices ([Floe.X] etc.)
//   n   = sz_resident_mex('ocean_forcing', tp, doInt)  tp: dt HFo xo_min xo_max yo_min yo_max; n = floes evaluated
//   ns  = sz_resident_mex('trajectory_step', tp)       ns = floes sacked (out of the ocean grid / NaN position)
//   st  = sz_resident_mex('state')                     x y u v ksi h alive mass inertia alpha dXi_p dYi_p dUi_p dVi_p
//                                                      dalpha_p dksi_p stress(4xN) flags FxOA FyOA torqueOA strain(4xN)
//                                                      cax cay (V x 1, the rotated c_alpha pool)
//   d   = sz_resident_mex('fracture_deform', idx)      idx: floe numbers; d: changed xi yi area vert_off cx cy
//   e   = sz_resident_mex('eulerian_data', g, st)      g: Nx Ny xmin xmax ymin ymax periodic; st: mass [overlap_area dUi_p dVi_p
//                                                      (N x 1) stress strain (4 x N)]; e: the 18 Ny x Nx fields of
//                                                      calc_eulerian_data.m (u v du dv stress ... c Over Mtot area h)
//   m   = sz_resident_mex('corner_mask', idx, Nb)      idx: the selection Floe(~keep) of Subzero.m:348 as floe numbers;
//                                                      m: da_off (numel(idx)+1), da (the mask of corners.m:54-88, one
//                                                      entry per polyshape vertex of every selected floe)
//
// Build (MATLAB):  mex -I../../include sz_resident_mex.cpp -L../_lib -lsubzero_b200
#include "mex.h"
#include "subzero_b200.h"
#include <vector>
#include <string>
#include <cstring>

static SzContext* g_ctx = nullptr;
static size_t g_n = 0, g_nv = 0;       // floes and outline vertices of the last upload
static SzSummary g_sum;
static void release_ctx() { if (g_ctx) { sz_destroy(g_ctx); g_ctx = nullptr; } }

static void fail(int code)
{
    // no device resource is owned by this frame: the context outlives the call, so the longjmp is safe
    mexErrMsgIdAndTxt(code == SZ_ERR_CLIPPER ? "subzero_b200:clipper" : "subzero_b200:error", "%s", sz_last_error());
}
static void check(int rc) { if (rc != SZ_OK) fail(rc); }
static const mxArray* need_field(const mxArray* s, const char* name, size_t min_len)
{
    if (!mxIsStruct(s)) mexErrMsgIdAndTxt("subzero_b200:arg", "argument must be a struct");
    const mxArray* f = mxGetField(s, 0, name);
    if (!f) mexErrMsgIdAndTxt("subzero_b200:arg", "missing field '%s'", name);
    if (!mxIsDouble(f) || mxIsComplex(f)) mexErrMsgIdAndTxt("subzero_b200:arg", "field '%s' must be a real double array", name);
    if (mxGetNumberOfElements(f) < min_len) mexErrMsgIdAndTxt("subzero_b200:arg", "field '%s' is too short", name);
    return f;
}
static const double* opt_field(const mxArray* s, const char* name, size_t len)
{
    const mxArray* f = mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr;
    if (!f || mxIsEmpty(f)) return nullptr;
    if (!mxIsDouble(f) || mxIsComplex(f) || mxGetNumberOfElements(f) < len) mexErrMsgIdAndTxt("subzero_b200:arg", "field '%s' must be a real double array of the right length", name);
    return mxGetPr(f);
}
static double scalar_field(const mxArray* s, const char* name, double dflt, bool required)
{
    const mxArray* f = mxIsStruct(s) ? mxGetField(s, 0, name) : nullptr;
    if (!f) { if (required) mexErrMsgIdAndTxt("subzero_b200:arg", "missing parameter '%s'", name); return dflt; }
    return mxGetScalar(f);
}
static const double* need_array(const mxArray* a, size_t len, const char* what)
{
    if (!a || !mxIsDouble(a) || mxIsComplex(a) || mxGetNumberOfElements(a) < len) mexErrMsgIdAndTxt("subzero_b200:arg", "%s must be a real double array of the right size", what);
    return mxGetPr(a);
}
static mxArray* vec(size_t n) { return mxCreateDoubleMatrix(n, 1, mxREAL); }
static SzTrajectoryParams traj_params(const mxArray* tp)
{
    SzTrajectoryParams p;
    p.dt = scalar_field(tp, "dt", 0, true); p.HFo = scalar_field(tp, "HFo", 0, false);
    p.xo_min = scalar_field(tp, "xo_min", -1e300, false); p.xo_max = scalar_field(tp, "xo_max", 1e300, false);
    p.yo_min = scalar_field(tp, "yo_min", -1e300, false); p.yo_max = scalar_field(tp, "yo_max", 1e300, false);
    return p;
}

static void cmd_upload(int nrhs, const mxArray* prhs[])
{
    if (nrhs < 3) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: sz_resident_mex('upload', prm, soa [, bnd])");
    SzParams P; sz_default_params(&P);
    const mxArray* p = prhs[1];
    P.Lx = scalar_field(p, "Lx", 0, true); P.Ly = scalar_field(p, "Ly", 0, true);
    P.modulus = scalar_field(p, "modulus", 0, true); P.dt = scalar_field(p, "dt", 0, true);
    P.Nb = (int32_t)scalar_field(p, "Nb", 0, false);
    P.pair_with_boundary_floes = scalar_field(p, "pair_with_boundary_floes", 0, false) != 0;      // opt-in, not reference behaviour
    P.periodic = scalar_field(p, "periodic", 0, true) != 0; P.collision = scalar_field(p, "collision", 1, false) != 0;
    P.nu = scalar_field(p, "nu", P.nu, false); P.mu = scalar_field(p, "mu", P.mu, false);
    const mxArray* s = prhs[2];
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
    if (nrhs >= 4 && !mxIsEmpty(prhs[3])) {
        const mxArray* b = prhs[3];
        B.n = (int32_t)mxGetNumberOfElements(need_field(b, "x", 3)); B.x = mxGetPr(need_field(b, "x", 3)); B.y = mxGetPr(need_field(b, "y", (size_t)B.n));
        B.box_n = (int32_t)mxGetNumberOfElements(need_field(b, "box_x", 3)); B.box_x = mxGetPr(need_field(b, "box_x", 3)); B.box_y = mxGetPr(need_field(b, "box_y", (size_t)B.box_n));
        B.area = scalar_field(b, "area", 0, true); B.h = scalar_field(b, "h", 0, false);
        B.xi = scalar_field(b, "xi", 0, false); B.yi = scalar_field(b, "yi", 0, false);
        B.u = scalar_field(b, "u", 0, false); B.v = scalar_field(b, "v", 0, false); B.ksi = scalar_field(b, "ksi", 0, false);
        pB = &B;
    }
    check(sz_upload(g_ctx, &P, &F, pB));
    g_n = n; g_nv = (size_t)F.nverts;
}

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[])
{
    if (nrhs < 1 || !mxIsChar(prhs[0])) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: sz_resident_mex(command, ...)");
    if (nlhs > 1) mexErrMsgIdAndTxt("subzero_b200:arg", "one output");
    char cmdbuf[64];
    if (mxGetString(prhs[0], cmdbuf, sizeof(cmdbuf)) != 0) mexErrMsgIdAndTxt("subzero_b200:arg", "bad command string");
    const std::string cmd(cmdbuf);
    if (!g_ctx) {
        if (sz_create(&g_ctx, 0) != SZ_OK) fail(SZ_ERR_CUDA);
        mexAtExit(release_ctx);
    }
    const size_t n = g_n;
    if (cmd == "upload") { cmd_upload(nrhs, prhs); return; }
    if (cmd == "step") {
        check(sz_step_resident(g_ctx, &g_sum));
        const char* names[] = {"n_ext", "n_pairs", "n_rows", "collision_count", "ms"};
        mxArray* out = mxCreateStructMatrix(1, 1, 5, names);
        mxArray* vals[] = {mxCreateDoubleScalar(g_sum.n), mxCreateDoubleScalar((double)g_sum.n_pairs), mxCreateDoubleScalar((double)g_sum.n_rows),
                           mxCreateDoubleScalar(g_sum.collision_count), mxCreateDoubleScalar(g_sum.ms_device)};
        for (int k = 0; k < 5; ++k) mxSetFieldByNumber(out, 0, k, vals[k]);
        plhs[0] = out;
        return;
    }
    if (cmd == "floe_outputs") {
        const char* names[] = {"fx", "fy", "torque", "overlap_area", "stress", "xi", "yi", "alive", "kill", "transfer"};
        mxArray* out = mxCreateStructMatrix(1, 1, 10, names);
        mxArray *fx = vec(n), *fy = vec(n), *tq = vec(n), *ov = vec(n), *st = mxCreateDoubleMatrix(4, n, mxREAL), *xi = vec(n), *yi = vec(n);
        std::vector<uint8_t> al(n); std::vector<int32_t> kill(n), transfer(n);
        check(sz_get_floe_outputs(g_ctx, mxGetPr(fx), mxGetPr(fy), mxGetPr(tq), mxGetPr(ov), mxGetPr(st), mxGetPr(xi), mxGetPr(yi), al.data(), kill.data(), transfer.data()));
        mxArray *ma = vec(n), *mk = vec(n), *mt = vec(n);
        for (size_t i = 0; i < n; ++i) { mxGetPr(ma)[i] = al[i]; mxGetPr(mk)[i] = kill[i]; mxGetPr(mt)[i] = transfer[i]; }
        mxArray* vals[] = {fx, fy, tq, ov, st, xi, yi, ma, mk, mt};
        for (int k = 0; k < 10; ++k) mxSetFieldByNumber(out, 0, k, vals[k]);
        plhs[0] = out;
        return;
    }
    if (cmd == "rows") {
        const char* names[] = {"row_off", "rows"};
        mxArray* out = mxCreateStructMatrix(1, 1, 2, names);
        std::vector<int64_t> off((size_t)g_sum.n + 1);
        mxArray* rows = mxCreateDoubleMatrix(7, (size_t)g_sum.n_rows, mxREAL);
        check(sz_get_rows(g_ctx, off.data(), mxGetPr(rows)));
        mxArray* ro = vec((size_t)g_sum.n + 1);
        for (size_t i = 0; i <= (size_t)g_sum.n; ++i) mxGetPr(ro)[i] = (double)off[i];
        mxSetFieldByNumber(out, 0, 0, ro); mxSetFieldByNumber(out, 0, 1, rows);
        plhs[0] = out;
        return;
    }
    if (cmd == "trajectory_init") {
        if (nrhs < 2) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: sz_resident_mex('trajectory_init', st)");
        const mxArray* s = prhs[1];
        SzTrajectoryInit ti; std::memset(&ti, 0, sizeof(ti));
        ti.mass = mxGetPr(need_field(s, "mass", n)); ti.inertia = mxGetPr(need_field(s, "inertia", n));
        ti.alpha = opt_field(s, "alpha", n); ti.dXi_p = opt_field(s, "dXi_p", n); ti.dYi_p = opt_field(s, "dYi_p", n);
        ti.dUi_p = opt_field(s, "dUi_p", n); ti.dVi_p = opt_field(s, "dVi_p", n); ti.dalpha_p = opt_field(s, "dalpha_p", n); ti.dksi_p = opt_field(s, "dksi_p", n);
        ti.FxOA = opt_field(s, "FxOA", n); ti.FyOA = opt_field(s, "FyOA", n); ti.torqueOA = opt_field(s, "torqueOA", n);
        ti.c0x = opt_field(s, "c0x", g_nv); ti.c0y = opt_field(s, "c0y", g_nv);
        ti.nz = (int32_t)scalar_field(s, "nz", 1000, false);
        check(sz_trajectory_init(g_ctx, &ti));
        return;
    }
    if (cmd == "set_ocean") {
        if (nrhs < 3) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: sz_resident_mex('set_ocean', ocean, winds)");
        const mxArray* o = prhs[1]; const mxArray* w = prhs[2];
        SzOcean oc; std::memset(&oc, 0, sizeof(oc));
        oc.nx = (int32_t)mxGetNumberOfElements(need_field(o, "Xo", 2)); oc.ny = (int32_t)mxGetNumberOfElements(need_field(o, "Yo", 2));
        const size_t cells = (size_t)oc.nx * (size_t)oc.ny;
        oc.Xo = mxGetPr(need_field(o, "Xo", 2)); oc.Yo = mxGetPr(need_field(o, "Yo", 2));
        oc.Uocn = mxGetPr(need_field(o, "Uocn", cells)); oc.Vocn = mxGetPr(need_field(o, "Vocn", cells));      // ny x nx, column-major: as stored
        oc.Uwinds = mxGetPr(need_field(w, "u", cells)); oc.Vwinds = mxGetPr(need_field(w, "v", cells));
        oc.fCoriolis = scalar_field(o, "fCoriolis", 0, true); oc.turn_angle = scalar_field(o, "turn_angle", 0, true);
        check(sz_trajectory_set_ocean(g_ctx, &oc));
        return;
    }
    if (cmd == "set_points") {
        if (nrhs < 4) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: sz_resident_mex('set_points', X, Y, A)");
        const size_t tot = mxGetNumberOfElements(prhs[1]);
        if (n == 0 || tot % n != 0) mexErrMsgIdAndTxt("subzero_b200:arg", "X must be npts x N");
        const size_t npts = tot / n;                          // npts x N column-major = floe-major
        const double* X = need_array(prhs[1], tot, "X"); const double* Y = need_array(prhs[2], tot, "Y"); const double* A = need_array(prhs[3], tot, "A");
        std::vector<uint8_t> a8(tot);
        for (size_t k = 0; k < tot; ++k) a8[k] = A[k] != 0;
        check(sz_trajectory_set_points(g_ctx, (int32_t)npts, X, Y, a8.data()));
        return;
    }
    if (cmd == "ocean_forcing" || cmd == "trajectory_step") {
        if (nrhs < 2) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: sz_resident_mex('%s', tp ...)", cmdbuf);
        const SzTrajectoryParams tp = traj_params(prhs[1]);
        int32_t a = 0, b = 0;
        if (cmd == "ocean_forcing") check(sz_trajectory_ocean_forcing(g_ctx, &tp, nrhs >= 3 && mxGetScalar(prhs[2]) != 0, &a, &b));
        else check(sz_trajectory_step(g_ctx, &tp, &a, &b));
        plhs[0] = mxCreateDoubleScalar(a);
        return;
    }
    if (cmd == "state") {
        const char* names[] = {"x", "y", "u", "v", "ksi", "h", "alive", "mass", "inertia", "alpha", "dXi_p", "dYi_p", "dUi_p", "dVi_p", "dalpha_p", "dksi_p",
                               "stress", "flags", "FxOA", "FyOA", "torqueOA", "strain", "cax", "cay"};
        mxArray* out = mxCreateStructMatrix(1, 1, 24, names);
        mxArray* v[24];
        for (int k = 0; k < 16; ++k) v[k] = vec(n);
        v[16] = mxCreateDoubleMatrix(4, n, mxREAL); v[17] = vec(n); v[18] = vec(n); v[19] = vec(n); v[20] = vec(n); v[21] = mxCreateDoubleMatrix(4, n, mxREAL);
        v[22] = vec(g_nv); v[23] = vec(g_nv);
        std::vector<uint8_t> al(n); std::vector<int32_t> fl(n);
        check(sz_get_trajectory(g_ctx, mxGetPr(v[0]), mxGetPr(v[1]), mxGetPr(v[2]), mxGetPr(v[3]), mxGetPr(v[4]), mxGetPr(v[5]), al.data(), mxGetPr(v[7]), mxGetPr(v[8]), mxGetPr(v[9]),
                                mxGetPr(v[10]), mxGetPr(v[11]), mxGetPr(v[12]), mxGetPr(v[13]), mxGetPr(v[14]), mxGetPr(v[15]), mxGetPr(v[16]), fl.data(), mxGetPr(v[22]), mxGetPr(v[23])));
        check(sz_get_trajectory_forcing(g_ctx, mxGetPr(v[18]), mxGetPr(v[19]), mxGetPr(v[20]), mxGetPr(v[21])));
        for (size_t i = 0; i < n; ++i) { mxGetPr(v[6])[i] = al[i]; mxGetPr(v[17])[i] = fl[i]; }
        for (int k = 0; k < 24; ++k) mxSetFieldByNumber(out, 0, k, v[k]);
        plhs[0] = out;
        return;
    }
    if (cmd == "fracture_deform") {
        if (nrhs < 2) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: d = sz_resident_mex('fracture_deform', idx)");
        const size_t m = mxGetNumberOfElements(prhs[1]);
        const double* id = need_array(prhs[1], m, "idx");
        std::vector<int32_t> idx(m);
        for (size_t k = 0; k < m; ++k) idx[k] = (int32_t)id[k];
        int64_t nc = 0, nv = 0;
        check(sz_fracture_deform(g_ctx, (int32_t)m, idx.data(), &nc, &nv));
        const char* names[] = {"changed", "xi", "yi", "area", "vert_off", "cx", "cy"};
        mxArray* out = mxCreateStructMatrix(1, 1, 7, names);
        mxArray *ch = vec(m), *xi = vec(m), *yi = vec(m), *ar = vec(m), *vo = vec(m + 1), *cx = vec((size_t)nv), *cy = vec((size_t)nv);
        std::vector<uint8_t> c8(m); std::vector<int64_t> off(m + 1);
        check(sz_get_fracture_deform(g_ctx, c8.data(), mxGetPr(xi), mxGetPr(yi), mxGetPr(ar), off.data(), mxGetPr(cx), mxGetPr(cy)));
        for (size_t k = 0; k < m; ++k) mxGetPr(ch)[k] = c8[k];
        for (size_t k = 0; k <= m; ++k) mxGetPr(vo)[k] = (double)off[k];
        mxArray* vals[] = {ch, xi, yi, ar, vo, cx, cy};
        for (int k = 0; k < 7; ++k) mxSetFieldByNumber(out, 0, k, vals[k]);
        plhs[0] = out;
        return;
    }
    if (cmd == "eulerian_data") {
        if (nrhs < 3) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: e = sz_resident_mex('eulerian_data', g, st)");
        const int Nx = (int)scalar_field(prhs[1], "Nx", 0, true), Ny = (int)scalar_field(prhs[1], "Ny", 0, true);
        if (Nx < 1 || Ny < 1) mexErrMsgIdAndTxt("subzero_b200:arg", "Nx and Ny must be positive");
        const double xmin = scalar_field(prhs[1], "xmin", 0, true), xmax = scalar_field(prhs[1], "xmax", 0, true);
        const double ymin = scalar_field(prhs[1], "ymin", 0, true), ymax = scalar_field(prhs[1], "ymax", 0, true);
        const int periodic = scalar_field(prhs[1], "periodic", 0, false) != 0;
        const double* mass = mxGetPr(need_field(prhs[2], "mass", g_n));
        std::vector<double> planes((size_t)18 * Nx * Ny);
        check(sz_eulerian_data(g_ctx, Nx, Ny, xmin, xmax, ymin, ymax, periodic, mass, opt_field(prhs[2], "overlap_area", g_n), opt_field(prhs[2], "dUi_p", g_n),
                               opt_field(prhs[2], "dVi_p", g_n), opt_field(prhs[2], "stress", 4 * g_n), opt_field(prhs[2], "strain", 4 * g_n), planes.data()));
        const char* names[] = {"u", "v", "du", "dv", "stress", "stressxx", "stressyx", "stressxy", "stressyy", "strainux", "strainvx", "strainuy", "strainvy",
                               "c", "Over", "Mtot", "area", "h"};
        mxArray* out = mxCreateStructMatrix(1, 1, 18, names);
        for (int k = 0; k < 18; ++k) {
            mxArray* m = mxCreateDoubleMatrix(Ny, Nx, mxREAL);        // column-major Ny x Nx from the row-major planes
            double* d = mxGetPr(m);
            for (int jj = 0; jj < Ny; ++jj) for (int ii = 0; ii < Nx; ++ii) d[(size_t)ii * Ny + jj] = planes[((size_t)k * Ny + jj) * Nx + ii];
            mxSetFieldByNumber(out, 0, k, m);
        }
        plhs[0] = out;
        return;
    }
    if (cmd == "corner_mask") {
        if (nrhs < 2) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: m = sz_resident_mex('corner_mask', idx [, Nb])");
        const size_t m = mxGetNumberOfElements(prhs[1]);
        const double* id = need_array(prhs[1], m, "idx");
        std::vector<int32_t> idx(m);
        for (size_t k = 0; k < m; ++k) idx[k] = (int32_t)id[k];
        int32_t nb = 0;
        if (nrhs > 2) {
            if (!mxIsDouble(prhs[2]) || mxGetNumberOfElements(prhs[2]) != 1) mexErrMsgIdAndTxt("subzero_b200:arg", "Nb must be a real scalar");
            nb = (int32_t)mxGetScalar(prhs[2]);
        }
        int64_t nv = 0;
        check(sz_corner_mask(g_ctx, (int32_t)m, idx.data(), nb, &nv));
        const char* names[] = {"da_off", "da"};
        mxArray* out = mxCreateStructMatrix(1, 1, 2, names);
        mxArray *vo = vec(m + 1), *da = vec((size_t)nv);
        std::vector<uint8_t> d8((size_t)nv + 1); std::vector<int64_t> off(m + 1);
        check(sz_get_corner_mask(g_ctx, off.data(), d8.data()));
        for (size_t k = 0; k <= m; ++k) mxGetPr(vo)[k] = (double)off[k];
        for (size_t k = 0; k < (size_t)nv; ++k) mxGetPr(da)[k] = d8[k];
        mxSetFieldByNumber(out, 0, 0, vo); mxSetFieldByNumber(out, 0, 1, da);
        plhs[0] = out;
        return;
    }
    if (cmd == "weld_search" || cmd == "simplify_search") {
        // the bounding-radius pair searches of weld.m:29-81 / FloeSimplify.m:13-31 on the resident floes (sz_pair_search):
        //   w = sz_resident_mex('weld_search', struct('Nb',Nb,'Nx',Nx,'Ny',Ny,'xmin',min(x),'xmax',max(x),'ymin',min(y),'ymax',max(y)))
        //       w.bin (bin number of every floe of Floe(1+Nb:end), 0 = none), w.off, w.partner (1-based positions in that cut list)
        //   s = sz_resident_mex('simplify_search', idx)        s.off, s.partner (1-based positions in the resident list)
        const bool weld = cmd == "weld_search";
        if (nrhs < 2) mexErrMsgIdAndTxt("subzero_b200:arg", "usage: sz_resident_mex('weld_search', g) / sz_resident_mex('simplify_search', idx)");
        int64_t np = 0; size_t nq = 0;
        if (weld) {
            const int Nb = (int)scalar_field(prhs[1], "Nb", 0, false), Nx = (int)scalar_field(prhs[1], "Nx", 0, true), Ny = (int)scalar_field(prhs[1], "Ny", 0, true);
            check(sz_pair_search(g_ctx, 0, Nb, Nx, Ny, scalar_field(prhs[1], "xmin", 0, true), scalar_field(prhs[1], "xmax", 0, true), scalar_field(prhs[1], "ymin", 0, true),
                                 scalar_field(prhs[1], "ymax", 0, true), 0, nullptr, &np));
            nq = g_n > (size_t)Nb ? g_n - (size_t)Nb : 0;
        } else {
            nq = mxGetNumberOfElements(prhs[1]);
            const double* id = need_array(prhs[1], nq, "idx");
            std::vector<int32_t> idx(nq);
            for (size_t k = 0; k < nq; ++k) idx[k] = (int32_t)id[k];
            check(sz_pair_search(g_ctx, 1, 0, 1, 1, 0, 1, 0, 1, (int32_t)nq, idx.data(), &np));
        }
        std::vector<int32_t> bin(nq + 1), partner((size_t)np + 1); std::vector<int64_t> off(nq + 1);
        check(sz_get_pair_search(g_ctx, bin.data(), off.data(), partner.data()));
        const char* names[] = {"bin", "off", "partner"};
        mxArray* out = mxCreateStructMatrix(1, 1, 3, names);
        mxArray *vb = vec(nq), *vo = vec(nq + 1), *vp = vec((size_t)np);
        for (size_t k = 0; k < nq; ++k) mxGetPr(vb)[k] = bin[k];
        for (size_t k = 0; k <= nq; ++k) mxGetPr(vo)[k] = (double)off[k];
        for (size_t k = 0; k < (size_t)np; ++k) mxGetPr(vp)[k] = partner[k];
        mxSetFieldByNumber(out, 0, 0, vb); mxSetFieldByNumber(out, 0, 1, vo); mxSetFieldByNumber(out, 0, 2, vp);
        plhs[0] = out;
        return;
    }
    mexErrMsgIdAndTxt("subzero_b200:arg", "unknown command '%s'", cmdbuf);
}
