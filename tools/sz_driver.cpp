// sz_driver.cpp -- standalone C++ driver of the contact step (BASELINE.json north_star: "a standalone C++ driver
// for benchmarking").  Links only the product library; prints one JSON line.
//   build:  g++ -O2 -std=c++17 -Iinclude tools/sz_driver.cpp -Lsubzero_b200/_lib -lsubzero_b200 -Wl,-rpath,'$ORIGIN/../subzero_b200/_lib' -o tools/sz_driver
//   run:    tools/sz_driver [n_floes=1000000] [steps=5] [warmup=3] [seed=0]
#include "subzero_b200.h"
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>

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
    printf("{\"floes\": %d, \"floes_incl_ghosts\": %d, \"pairs\": %lld, \"pairs_with_force\": %lld, \"rows\": %lld, \"ms_per_step\": %.4f, "
           "\"pairs_per_s\": %.4e, \"timesteps_per_s\": %.3f, \"phase_ms\": {\"ghosts\": %.3f, \"broad\": %.3f, \"narrow\": %.3f, \"assembly\": %.3f}, "
           "\"narrow_class_C\": {\"ms\": %.3f, \"pairs\": %d}, \"narrow_class_S\": {\"ms\": %.3f, \"pairs\": %d}, "
           "\"timesteps_per_s_moving\": %.3f, \"sum_fx\": %.6e, \"sum_fy\": %.6e, \"kernels_launched\": %lld}\n",
           s.n0, s.n, (long long)s.n_pairs, (long long)s.n_pairs_force, (long long)s.n_rows, ms, s.n_pairs / (ms * 1e-3), 1e3 / ms,
           ph[0], ph[1], ph[2], ph[3], cls[0], cls_pairs[0], cls[1], cls_pairs[1], ts_ab2, sx, sy, sz_launch_count());
    sz_destroy(ctx); sz_field_free(field);
    return 0;
}
