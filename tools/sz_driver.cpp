// sz_driver.cpp -- standalone C++ driver of the contact step (BASELINE.json north_star: "a standalone C++ driver
// for benchmarking").  Links only the product library; prints one JSON line.
//   build:  g++ -O2 -std=c++17 -Iinclude tools/sz_driver.cpp -Lsubzero_b200/_lib -lsubzero_b200 -Wl,-rpath,'$ORIGIN/../subzero_b200/_lib' -o tools/sz_driver
//   run:    tools/sz_driver [n_floes=1000000] [steps=5] [warmup=3] [seed=0]
#include "subzero_b200.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#include <chrono>
#include <string>

int main(int argc, char** argv)
{
    const int n = argc > 1 ? atoi(argv[1]) : 1000000, steps = argc > 2 ? atoi(argv[2]) : 5, warm = argc > 3 ? atoi(argv[3]) : 3;
    const unsigned long long seed = argc > 4 ? strtoull(argv[4], 0, 10) : 0;
    SzParams prm; sz_default_params(&prm);
    SzField* field = nullptr;
    if (sz_field_voronoi(&field, n, seed, 4e6, 0.02, &prm) != SZ_OK) { fprintf(stderr, "%s\n", sz_last_error()); return 1; }
    prm.periodic = 1; prm.collision = 1; prm.dt = 10;
    SzFloesSoA view; sz_field_view(field, &view);
    SzContext* ctx = nullptr;
    if (sz_create(&ctx, 0) != SZ_OK) { fprintf(stderr, "%s\n", sz_last_error()); return 2; }      // no CPU fallback
    if (sz_upload(ctx, &prm, &view, nullptr) != SZ_OK) { fprintf(stderr, "%s\n", sz_last_error()); return 3; }
    SzSummary s; double ms = 0;
    for (int it = 0; it < warm + steps; ++it) {
        if (sz_step_resident(ctx, &s) != SZ_OK) { fprintf(stderr, "%s\n", sz_last_error()); return 4; }
        if (it >= warm) ms += s.ms_device;
    }
    ms /= steps;
    // the other half of the timestep: calc_trajectory on the device (no ocean/wind), state resident between steps
    double ts_ab2 = 0;
    {
        std::vector<double> mass(n), inertia(n);
        for (int i = 0; i < n; ++i) { mass[i] = view.area[i] * view.h[i] * 920.0; inertia[i] = mass[i] * view.rmax[i] * view.rmax[i] / 4; }
        SzTrajectoryInit ti; memset(&ti, 0, sizeof(ti)); ti.mass = mass.data(); ti.inertia = inertia.data(); ti.nz = 1000;
        SzTrajectoryParams tp = {prm.dt, 0.0, -1e300, 1e300, -1e300, 1e300};
        if (sz_trajectory_init(ctx, &ti) != SZ_OK) { fprintf(stderr, "%s\n", sz_last_error()); return 5; }
        double tot = 0;
        for (int it = 0; it < steps; ++it) {
            if (sz_step_resident(ctx, &s) != SZ_OK || sz_trajectory_step(ctx, &tp, nullptr, nullptr) != SZ_OK) { fprintf(stderr, "%s\n", sz_last_error()); return 6; }
            tot += s.ms_device;
        }
        ts_ab2 = 1e3 * steps / tot;     // contact-step device time; the trajectory kernel adds well under a millisecond
        if (sz_step_resident(ctx, &s) != SZ_OK) return 7;
    }
    float ph[5]; sz_get_phase_ms(ctx, ph);
    float cls[5]; int32_t cls_pairs[5]; sz_get_narrow_class_ms(ctx, cls, cls_pairs);      // narrow-phase kernel time per size class
    std::vector<double> fx(n), fy(n);
    sz_get_floe_outputs(ctx, fx.data(), fy.data(), 0, 0, 0, 0, 0, 0, 0, 0);
    double sx = 0, sy = 0; for (int i = 0; i < n; ++i) { sx += fx[i]; sy += fy[i]; }
    // consumers of the contact rows and of the resident state, wall clock around the blocking calls
    // (each includes its own read-backs): the corners.m mask for 70 % of the floes (Subzero.m:341-348) and calc_eulerian_data
    // on a 10 x 10 and a 100 x 100 grid.  A failure here is reported in the line, it does not hide the contact-step numbers.
    double corner_ms = -1, euler10_ms = -1, euler100_ms = -1; long long corner_verts = 0, corner_flagged = 0; std::string extras_err;
    {
        typedef std::chrono::steady_clock clk;
        auto ms_since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
        std::vector<int32_t> sel;
        for (int i = 0; i < n; ++i) if (i % 10 < 7) sel.push_back(i + 1);
        int64_t nv = 0;
        for (int rep = 0; rep < 2 && extras_err.empty(); ++rep) {          // the second call is the timed one (buffers allocated)
            auto t0 = clk::now();
            if (sz_corner_mask(ctx, (int32_t)sel.size(), sel.data(), 0, &nv) != SZ_OK) { extras_err = sz_last_error(); break; }
            corner_ms = ms_since(t0);
        }
        if (extras_err.empty()) {
            std::vector<int64_t> off(sel.size() + 1); std::vector<uint8_t> da((size_t)nv + 1);
            if (sz_get_corner_mask(ctx, off.data(), da.data()) != SZ_OK) extras_err = sz_last_error();
            corner_verts = nv; for (int64_t k = 0; k < nv; ++k) corner_flagged += da[k];
        }
        std::vector<double> mass(n);
        for (int i = 0; i < n; ++i) mass[i] = view.area[i] * view.h[i] * 920.0;
        for (int g = 0; g < 2 && extras_err.empty(); ++g) {
            const int N = g == 0 ? 10 : 100;
            std::vector<double> planes((size_t)18 * N * N);
            for (int rep = 0; rep < 2 && extras_err.empty(); ++rep) {
                auto t0 = clk::now();
                if (sz_eulerian_data(ctx, N, N, -prm.Lx, prm.Lx, -prm.Ly, prm.Ly, 1, mass.data(), 0, 0, 0, 0, 0, planes.data()) != SZ_OK) { extras_err = sz_last_error(); break; }
                (g == 0 ? euler10_ms : euler100_ms) = ms_since(t0);
            }
        }
        for (char& ch : extras_err) if (ch == '"' || ch == '\\') ch = '\'';
    }
    printf("{\"floes\": %d, \"floes_incl_ghosts\": %d, \"pairs\": %lld, \"pairs_with_force\": %lld, \"rows\": %lld, \"ms_per_step\": %.4f, "
           "\"pairs_per_s\": %.4e, \"timesteps_per_s\": %.3f, \"phase_ms\": {\"ghosts\": %.3f, \"broad\": %.3f, \"narrow\": %.3f, \"assembly\": %.3f}, "
           "\"narrow_class_C\": {\"ms\": %.3f, \"pairs\": %d}, \"narrow_class_S\": {\"ms\": %.3f, \"pairs\": %d}, "
           "\"timesteps_per_s_moving\": %.3f, \"sum_fx\": %.6e, \"sum_fy\": %.6e, "
           "\"corner_mask\": {\"ms\": %.3f, \"selected\": %d, \"vertices\": %lld, \"flagged\": %lld}, \"eulerian_data_ms\": {\"10x10\": %.3f, \"100x100\": %.3f}, "
           "\"extras_error\": \"%s\", \"kernels_launched\": %lld}\n",
           s.n0, s.n, (long long)s.n_pairs, (long long)s.n_pairs_force, (long long)s.n_rows, ms, s.n_pairs / (ms * 1e-3), 1e3 / ms,
           ph[0], ph[1], ph[2], ph[3], cls[0], cls_pairs[0], cls[1], cls_pairs[1], ts_ab2, sx, sy,
           corner_ms, (int)(7 * (long long)n / 10), corner_verts, corner_flagged, euler10_ms, euler100_ms, extras_err.c_str(), sz_launch_count());
    sz_destroy(ctx); sz_field_free(field);
    return 0;
}
