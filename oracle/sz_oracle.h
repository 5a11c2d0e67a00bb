/* sz_oracle.h -- TEST INFRASTRUCTURE ONLY (see sz_oracle.cpp).  C interface of the CPU checker. */
#ifndef SZ_ORACLE_H
#define SZ_ORACLE_H
#include <stdint.h>
#include "../include/subzero_b200.h"   /* struct layouts only */
#ifdef __cplusplus
extern "C" {
#endif
typedef struct SzoResult SzoResult;
/* broad_mode 0 = literal O(N^2) loop of floe_interactions_all.m:76-120; 1 = same predicate over a cell grid */
SzoResult* szo_contact_step(const SzParams* prm, const SzFloesSoA* floes, const SzBoundary* bnd, int nthreads, int broad_mode);
void szo_free(SzoResult* r);
const char* szo_error(const SzoResult* r);
void szo_summary(const SzoResult* r, SzSummary* s);
void szo_get_floe_outputs(const SzoResult* r, double* fx, double* fy, double* torque, double* overlap_area, double* stress,
                          double* xi, double* yi, uint8_t* alive, int32_t* kill, int32_t* transfer);
void szo_get_ghosts(const SzoResult* r, int32_t* parent, int32_t* floe_num, double* gx, double* gy);
void szo_get_pairs(const SzoResult* r, int32_t* pi, int32_t* pj, double* overlap_state, int32_t* n_regions, int32_t* status);
void szo_get_rows(const SzoResult* r, int64_t* row_off, double* rows);
void szo_get_clip_polys(const SzoResult* r, int64_t* pair_path_off, int64_t* path_vert_off, int64_t* x, int64_t* y);
int  szo_polyclip(const double* x1, const double* y1, int n1, const double* x2, const double* y2, int n2, int method,
                  double* ox, double* oy, int cap, int* off, int off_cap);
void szo_polyshape_area_centroid(const double* x, const double* y, int n, double* out3);
double szo_polyarea(const double* x, const double* y, int n);
int  szo_interx(const double* x1, const double* y1, int n1, const double* x2, const double* y2, int n2, double* out, int cap);
void szo_inpolygon(const double* px, const double* py, int np, const double* xv, const double* yv, int nv, uint8_t* in);
int  szo_p_poly_dist(const double* px, const double* py, int np, const double* xv, const double* yv, int nv, double* d);
int64_t szo_matlab_int64(double v);
int  szo_floe_interactions(const SzParams* prm, const double* cax, const double* cay, int n1, const double* body1,
                           const double* c2x, const double* c2y, int n2, const double* body2, int is_boundary,
                           const double* boxx, const double* boxy, int nbox,
                           double* rows_out, int rows_cap, double* overlap_state);
int  szo_hardware_threads(void);
void szo_calc_trajectory(int n, double dt, double HFo, double xo_min, double xo_max, double yo_min, double yo_max, int nz,
                         const double* cfx, const double* cfy, const double* ctq, const double* stress_now, const uint8_t* has_rows,
                         const double* area, double* x, double* y, double* u, double* v, double* ksi, double* h, uint8_t* alive,
                         double* mass, double* inertia, double* alpha, double* dXi_p, double* dYi_p, double* dUi_p, double* dVi_p,
                         double* dalpha_p, double* dksi_p, const double* FxOA, const double* FyOA, const double* torqueOA,
                         const int32_t* voff, const double* c0x, const double* c0y, double* cax, double* cay,
                         double* stress_h, int32_t* stress_count, double* stress_out, uint8_t* sacked, uint8_t* unsupported);
void szo_ocean_forcing(int n, double dt, double HFo, double xo_min, double xo_max, double yo_min, double yo_max, int do_int,
                       const uint8_t* alive, const double* x, const double* y, const double* u, const double* v, const double* ksi,
                       const double* h, const double* mass, const double* area, const double* alpha,
                       const int32_t* voff, const double* cax, const double* cay,
                       int npts, const double* PX, const double* PY, const uint8_t* PA,
                       int nx, int ny, const double* Xo, const double* Yo, const double* Uocn, const double* Vocn, const double* Uwinds, const double* Vwinds,
                       double fc, double turn_angle, double rho0, double Cd, double rho_air, double Cd_atm,
                       double* FxOA, double* FyOA, double* torqueOA, uint8_t* evaluated, uint8_t* no_points);
void szo_floe_strain(int n, const uint8_t* alive, const uint8_t* sacked, const double* area, const double* u, const double* v, const double* ksi,
                     const int32_t* voff, const double* cax, const double* cay, double* strain);
int  szo_fracture_deform(const SzFloesSoA* f, const int64_t* row_off, const double* rows, int count, const int32_t* idx,
                         uint8_t* changed, double* xi, double* yi, double* area, int64_t* vert_off, double* cx, double* cy, int64_t vcap);
int  szo_calc_eulerian_data(const SzFloesSoA* f, const double* mass, const double* overlap_area, const double* dUi_p, const double* dVi_p,
                            const double* stress, const double* strain, int Nx, int Ny, int Nb,
                            double xmin, double xmax, double ymin, double ymax, int periodic, double* out);
int  szo_corner_eligibility(const SzFloesSoA* f, const int64_t* row_off, const double* rows, int count, const int32_t* idx, int Nb,
                            double Lx, double Ly, const double* boxx, const double* boxy, int nbox,
                            int64_t* da_off, uint8_t* da, int64_t vcap);
int64_t szo_weld_search(const SzFloesSoA* f, int Nb, int Nx, int Ny, double xmin, double xmax, double ymin, double ymax,
                        int32_t* bin, int64_t* off, int32_t* partner, int64_t cap);
int64_t szo_simplify_search(const SzFloesSoA* f, int count, const int32_t* idx, int64_t* off, int32_t* partner, int64_t cap);
#ifdef __cplusplus
}
#endif
#endif
