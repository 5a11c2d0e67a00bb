/* subzero_b200.h -- C ABI of the B200 contact-force step.
 *
 * One call per timestep replaces the range floe_interactions_all.m:9-265 of the reference
 * (ghost floes, broad phase, pair loop over collisions/floe_interactions.m, mirror, torque,
 * per-floe sums, periodic wrap) plus the stress rows of calc_trajectory.m:9-13 and
 * calc_collisionNum.m:3-6.  It takes the place of the >= 3 calls per contacting pair that the
 * reference makes through its mex gateway  private/mexclipper.cpp:83 (mexFunction; boolean branch
 * :204-305), reached from polyclip.m:73.
 *
 * Conventions (modelled on the reference gateway, private/mexclipper.cpp):
 *   - plain pointers and sizes only (host or device memory, the copy direction is inferred); the caller owns every
 *     buffer it passes in or receives into
 *     (cf. :54-61 inputs copied, :65-81 outputs caller-visible arrays);
 *   - no exceptions cross the ABI; every entry point returns 0 or a negative SzStatus, and
 *     sz_last_error() returns the message (the gateway's mexErrMsgTxt strings, e.g. :304);
 *   - a context owns device buffers and streams; one host thread per context (the gateway is
 *     stateless per call, :294; here state is only a cache of allocations);
 *   - floe indices in outputs are 1-BASED positions in the extended floe list (originals, then
 *     x-ghosts, then y-ghosts), exactly the numbers the reference stores in
 *     Floe(i).interactions(:,1); the wall partner is +Inf (floe_interactions_all.m:167).
 *
 * There is no CPU fallback: every compute entry point fails with SZ_ERR_CUDA when no sm_100a
 * device is usable.
 */
#ifndef SUBZERO_B200_H
#define SUBZERO_B200_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum SzStatus {
    SZ_OK = 0,
    SZ_ERR_ARG = -1,        /* bad argument (gateway: type/shape checks, mexclipper.cpp:22-41) */
    SZ_ERR_CUDA = -2,       /* CUDA runtime error / no device */
    SZ_ERR_CLIPPER = -3,    /* "Clipper Error." (mexclipper.cpp:303-304): the sweep failed for some pair */
    SZ_ERR_CAPACITY = -4,   /* a pair exceeded the largest narrow-phase size class */
    SZ_ERR_STATE = -5       /* results requested before a step was run */
} SzStatus;

/* Physics/configuration.  Defaults (sz_default_params) are the constants hard-coded in the
 * reference; they are parameters here because the validation cases change them by editing the
 * source (README.md:113,218,242). */
typedef struct SzParams {
    double Lx, Ly;            /* max(c2_boundary(1,:)), max(c2_boundary(2,:))   floe_interactions_all.m:9-10 */
    double modulus;           /* global Modulus                                  floe_interactions_all.m:7   */
    double dt;                /*                                                 floe_interactions.m:178     */
    double nu;                /* 0.3   Poisson ratio                             floe_interactions.m:20      */
    double mu;                /* 0.2   Coulomb friction                          floe_interactions.m:21      */
    double merge_frac;        /* 0.55  overlap/area => +-Inf                     floe_interactions.m:55-58   */
    double wall_frac;         /* 0.75  boundary overlap => Inf                   floe_interactions.m:37      */
    double amin_per_vertex;   /* 100/1.75  Amin = min(N1,N2)*this                floe_interactions.m:79      */
    double vertex_match_tol;  /* 1     dist<1                                    floe_interactions.m:99      */
    double on_edge_tol;       /* 1e-8  abs(d)<1e-8                               floe_interactions.m:127     */
    double dl_min;            /* 0.1   dl<0.1 => no force                        floe_interactions.m:141     */
    double close_gap;         /* 1     norm(c(:,1)-c(:,end))>1 => close outline  floe_interactions.m:62-67   */
    double big_floe_r;        /* 1e5   r>1e5 => min(h)/min(r)                    floe_interactions.m:15      */
    double domain_area_frac;  /* 0.95                                            floe_interactions.m:54      */
    int32_t Nb;               /* number of leading topography floes (never "i" of a pair, :76) */
    int32_t periodic;         /* PERIODIC  */
    int32_t collision;        /* COLLISION */
    int32_t want_clip_polys;  /* 1: keep clip #1's int64 polygons per pair for bit-exact parity checks */
    int32_t pair_with_boundary_floes;   /* 0 (default) = the reference: the first Nb (topography) floes are never anybody's partner, because
                                 * the search only looks at j > i >= Nb+1 (floe_interactions_all.m:76,102-103).  1 = opt-in, NOT
                                 * reference behaviour: a floe i > Nb also records the topography floes j <= Nb within reach (ascending, ahead of
                                 * its other partners); the force acts on i only (no mirrored row: :196 needs partner > i), such a pair never
                                 * raises kill / transfer, and it takes no part in the ghost de-dup of :93-99,113. */
} SzParams;

/* The hot-path fields of the reference's Floe struct array (Initialize_Model/initialize_floe_values.m:12-52),
 * flattened to structure-of-arrays.  Dead floes are dropped by the caller first
 * (floe_interactions_all.m:12-13); `alive` is still honoured as in :101-103. */
typedef struct SzFloesSoA {
    int32_t n;                /* N0 */
    int64_t nverts;           /* voff[n] */
    const double* x;          /* Xi  */
    const double* y;          /* Yi  */
    const double* rmax;
    const double* h;
    const double* area;
    const double* u;          /* Ui  */
    const double* v;          /* Vi  */
    const double* ksi;        /* ksi_ice */
    const uint8_t* alive;
    const int32_t* voff;      /* [n+1] offsets into vx/vy */
    const double* vx;         /* c_alpha(1,:) -- outline about the centroid, CLOSED (first vertex repeated, :17) */
    const double* vy;         /* c_alpha(2,:) */
} SzFloesSoA;

/* The boundary "floe" of the non-periodic wall call (floe_interactions_all.m:151, Subzero.m:68-70).
 * x,y = holes(floebound.poly).Vertices as the MATLAB wrapper reads them (floe_interactions.m:31-32);
 * box_x,box_y = c2_boundary (closed, 2x5) used for the bbox guard (:54) and the centroid test (:152). */
typedef struct SzBoundary {
    const double* x; const double* y; int32_t n;
    const double* box_x; const double* box_y; int32_t box_n;
    double area, h, xi, yi, u, v, ksi;
} SzBoundary;

typedef struct SzSummary {
    int32_t n0;               /* input floes */
    int32_t n;                /* extended list: originals + x-ghosts + y-ghosts */
    int64_t n_pairs;          /* candidate pairs taken through the narrow phase */
    int64_t n_pairs_force;    /* pairs that produced contact rows */
    int64_t n_rows;           /* contact rows over all floes of the extended list (own + wall + mirrored) */
    int64_t n_clip_paths;     /* debug polygons kept (want_clip_polys) */
    int64_t n_clip_verts;
    double  collision_count;  /* calc_collisionNum over the first n0 floes */
    int32_t n_clipper_fail;   /* pairs whose sweep failed (reference would raise "Clipper Error.") */
    int32_t n_capacity_fail;  /* pairs that exceeded the largest size class */
    float   ms_device;        /* device time of the step (CUDA events), excluding host<->device copies */
    int64_t n_pairs_owned;    /* pairs whose first floe this context owns (== n_pairs unless an extended list was supplied) */
    int32_t n_kill_events;    /* entries with a kill entry (merges, floe_interactions_all.m:138-145): the cross-rank fix-up of :175-179 only runs when some rank has one */
} SzSummary;

typedef struct SzContext SzContext;

/* ---- lifetime ---- */
int  sz_create(SzContext** out, int device);            /* device = CUDA ordinal */
void sz_destroy(SzContext* ctx);
const char* sz_last_error(void);                        /* thread-local message of the last failure */
void sz_default_params(SzParams* p);
int  sz_abi_version(void);
long long sz_launch_count(void);                       /* CUDA kernels this library has launched so far (process-wide) */

/* ---- one contact step, host buffers in (the mex / ctypes entry point) ---- */
int sz_contact_step(SzContext* ctx, const SzParams* prm, const SzFloesSoA* floes,
                    const SzBoundary* bnd /* may be NULL when periodic */, SzSummary* out);

/* ---- split form for device-resident state: upload once, step many times ---- */
int sz_upload(SzContext* ctx, const SzParams* prm, const SzFloesSoA* floes, const SzBoundary* bnd);
int sz_step_resident(SzContext* ctx, SzSummary* out);
/* The same step in two halves, for callers that keep the host out of the loop (a CUDA graph of the whole step, NCCL calls
 * included; software pipelining): sz_step_enqueue launches the step on carried-over sizes (option "speculate"; fails with
 * SZ_ERR_STATE until one sz_step_resident has run) and queues the copy of the counters, without any host synchronisation;
 * sz_step_finish waits for the stream, validates the step and fills the summary.  It returns 1 (not an error) when a carried-over
 * size was too small: nothing was consumed, run sz_step_resident on the unchanged state.  With option "graph_safe" set, nothing
 * the enqueue half launches depends on host state of the call (no event timing, scans reset their own state), so the launches
 * can be captured once and replayed.  sz_add_launches credits replayed launches to sz_launch_count. */
int sz_step_enqueue(SzContext* ctx);
int sz_step_finish(SzContext* ctx, SzSummary* out);
void sz_add_launches(long long n);

/* ---- caller-supplied extended list (any host that builds its own list; the multi-GPU path below builds it on the device).
 * One record of `floes` per entry (originals and periodic images this rank owns, plus halo entries received from
 * the neighbouring slabs), ascending `gid`; x,y are the image centroids (floe_interactions_all.m:34,55).  Pairs are
 * resolved when either floe is owned; rows, sums and per-floe outputs ([n] in this mode) are produced for owned
 * entries; row partner ids and kill/transfer are global 1-based positions; kill/transfer come back per entry
 * before the serial fix-up of :175-179 (which spans ranks and is left to the caller).  Pointers may be host or
 * device memory (as for every other entry point). */
typedef struct SzExtendedList {
    const int32_t* gid;        /* 1-based position in the global extended list */
    const int32_t* floe_num;   /* FloeNums: original id, negative for images (:35,56) */
    const double*  root_x;     /* centroid of the original the entry is an image of (its own centroid for originals) */
    const double*  root_y;
    const uint8_t* owned;      /* 1: rows and sums of this entry are this context's job */
    const int32_t* parent;     /* images: 1-based LOCAL index of the parent entry when it is owned here, else 0 */
} SzExtendedList;
int sz_upload_extended(SzContext* ctx, const SzParams* prm, const SzFloesSoA* entries, const SzBoundary* bnd, const SzExtendedList* ext);

/* ---- multi-GPU slabs, device-built list: one process per GPU, each rank OWNS a set of floes -- ascending
 * global floe numbers `gid` (1-based, the numbering of the single-GPU run; any subset, e.g. a contiguous range after a sort by
 * slab) -- and keeps their state and the integrator's resident on its GPU.  Every step the rank's part of the global extended
 * floe list is rebuilt on the device from the CURRENT state, so the step is exact for floes that translate, rotate (c_alpha =
 * A_rot c0, calc_trajectory.m:221-222), thin (:75-79) and die; nothing is cached between steps.  Per step, all on the context's
 * stream (sz_set_stream) with no host synchronisation:
 *   sz_slab_prepare(meta)            image flags of :28-31,49-52 from the current outlines, x-extents, largest rmax -> this rank's
 *                                    meta record [sz_slab_meta_doubles(cap_img)] in DEVICE memory; the caller all-gathers the records
 *   sz_slab_pack(all_meta, send)     global list positions of the rank's images (rank among all ranks' images, :33-36,54-57);
 *                                    every own entry within reach (2 max rmax) of another rank's x-extent is packed for it: state,
 *                                    FloeNums, root centroid and its CURRENT outline.  send = `world` blocks of
 *                                    sz_slab_block_doubles(cap_rec, cap_vert) doubles; the caller runs one all-to-all of equal splits
 *   sz_slab_build(recv, status)      received entries sorted by global position and merged with the own ones into the resident
 *                                    extended list; status (device, 4 ints, may be NULL) = {capacity overflow, owned floes outside the rank's extent, list length, 0}
 *   sz_step_resident                 pairs with at least one owned floe (straddling pairs are resolved on both sides, which keeps
 *                                    every floe's rows bit-identical to the single-GPU run and replaces the return of partial forces)
 *   sz_trajectory_step               integrates the owned floes (sz_trajectory_init after sz_slab_upload)
 * Capacities: sz_slab_measure / sz_slab_measure_halo (host-synchronous) report the image and halo counts the current state
 * needs; the caller agrees on cap_img / cap_rec / cap_vert across ranks and calls sz_slab_configure.  A later step that exceeds
 * them raises the overflow status (and sz_step_resident's results are then invalid): measure, configure, repeat the step.
 * Walls: pass `bnd` for a non-periodic domain; every rank resolves the wall contacts of the floes it owns.
 * kill / transfer come back per owned floe BEFORE the serial fix-up of floe_interactions_all.m:175-179, which spans ranks. */
int64_t sz_slab_meta_doubles(int32_t cap_img);
int64_t sz_slab_block_doubles(int32_t cap_rec, int32_t cap_vert);
int sz_slab_upload(SzContext* ctx, const SzParams* prm, const SzFloesSoA* owned, const SzBoundary* bnd, const int32_t* gid, int32_t n_global, int32_t rank, int32_t world);
int sz_slab_measure(SzContext* ctx, double* local8 /* host: x-images, y-images of originals, y-images of x-images, x-extent of originals [2], of x-images [2], max rmax */);
int sz_slab_measure_halo(SzContext* ctx, const double* all8 /* host [world*8] */, int64_t* rec_counts /* [world] */, int64_t* vert_counts /* [world] */);
int sz_slab_configure(SzContext* ctx, int32_t cap_img, int32_t cap_rec, int32_t cap_vert);
/* the x-range [xlo, xhi) this rank's floes are expected to stay in: sz_slab_build's status word [1] counts the live owned floes
 * whose centroid left it (the caller's cue to move floes between ranks; results never depend on it).  Default: unbounded. */
int sz_slab_set_extent(SzContext* ctx, double xlo, double xhi);
int sz_slab_prepare(SzContext* ctx, double* meta_dev);
int sz_slab_pack(SzContext* ctx, const double* all_meta_dev, double* send_dev);
int sz_slab_build(SzContext* ctx, const double* recv_dev, int32_t* status_dev);
int sz_slab_get_positions(SzContext* ctx, int32_t* opos /* [n_owned] 0-based list positions */, int32_t* n_list);
/* Floe(i).interactions of the owned floes in their order (gathered on the device): row_off [n_owned + 1], rows [n_rows * 7] with room
 * for rows_cap rows; call with rows = NULL to learn n_rows first.  Partner ids are global list positions. */
int sz_slab_get_rows(SzContext* ctx, int64_t* row_off, double* rows, int64_t rows_cap, int64_t* n_rows);
int sz_slab_get_list(SzContext* ctx, int32_t* gid, int32_t* floe_num, uint8_t* owned, double* x, double* y);   /* [n_list] each, any may be NULL */
int sz_slab_get_outputs(SzContext* ctx, double* fx, double* fy, double* torque, double* overlap_area, double* stress, double* xi, double* yi,
                        uint8_t* alive, int32_t* kill, int32_t* transfer);   /* per owned floe, like sz_get_floe_outputs */

/* ---- results of the last step (caller-allocated; sizes from SzSummary) ---- */
/* per floe of the input list, each [n0] unless noted; any pointer may be NULL to skip */
int sz_get_floe_outputs(SzContext* ctx,
                        double* fx, double* fy, double* torque,   /* collision_force / collision_torque, ghosts folded in (:242-245,262-263) */
                        double* overlap_area,                      /* OverlapArea */
                        double* stress,                            /* [n0*4], row-major 2x2, calc_trajectory.m:12-13 (un-averaged) */
                        double* xi, double* yi,                    /* centroid after the periodic wrap (:267-277) */
                        uint8_t* alive,                            /* 0 where the centroid left a non-periodic domain (:152-155) */
                        int32_t* kill, int32_t* transfer);         /* 1-based ids as in :138-145,175-179; 0 = none */
/* ghost bookkeeping [n - n0]: parent = 1-based index in the extended list, floe_num = FloeNums (negative) */
int sz_get_ghosts(SzContext* ctx, int32_t* parent, int32_t* floe_num, double* gx, double* gy);
/* what the reference leaves in the ghost structs Floe(N0+1:N), [n - n0] each, any pointer may be NULL: collision_force and
 * collision_torque = column sums of the ghost's own rows (floe_interactions_all.m:218-238; before they are folded into the
 * parent, :242-245) and OverlapArea (:137,198).  The ghosts' interactions are rows row_off[n0..n] of sz_get_rows.  Needed by
 * a host that runs the reference's ridging / rafting tail, which indexes Floe(partner) with partner > N0 (:312,327,401,416). */
int sz_get_ghost_outputs(SzContext* ctx, double* fx, double* fy, double* torque, double* overlap_area);
/* candidate pairs [n_pairs], 1-based (i<j), ascending (i,j) = the order of Floe(i).potentialInteractions;
 * overlap_state: 0, +Inf or -Inf (floe_interactions.m:55-58); status: 0 ok, <0 SzStatus of that pair */
int sz_get_pairs(SzContext* ctx, int32_t* pi, int32_t* pj, double* overlap_state, int32_t* n_regions, int32_t* status);
/* contact rows in the reference's canonical order: row_off [n+1], rows [n_rows*7]
 * = [partner Fx Fy Px Py torque overlap], exactly Floe(i).interactions */
int sz_get_rows(SzContext* ctx, int64_t* row_off, double* rows);
/* debug (want_clip_polys): clip #1 result of every pair as Clipper's int64 coordinates.
 * pair_path_off [n_pairs+1] -> path_vert_off [n_clip_paths+1] -> x,y [n_clip_verts] */
int sz_get_clip_polys(SzContext* ctx, int64_t* pair_path_off, int64_t* path_vert_off, int64_t* x, int64_t* y);

/* ---- the integrator half of the timestep, calc_trajectory.m, for the branch the contact-loop
 * benchmark exercises: doInt.flag = false with the ocean/atmosphere tendencies FxOA, FyOA, torqueOA carried over (a floe
 * thinner than 0.1 m, which the reference re-forces every step (:94), is reported and the call fails unless
 * sz_trajectory_ocean_forcing evaluated it first, see below).  The state stays on the device: contact step -> trajectory step -> contact step ... with no host
 * round trip.  Single-GPU lists only.
 *   sz_trajectory_init  after sz_upload: the Floe fields the integrator owns (initialize_floe_values.m:16-24,40-47);
 *                       NULL = zeros; c0 NULL = the uploaded c_alpha (alpha_i = 0); StressH = zeros(2,2,nz), StressCount = 1
 *   sz_trajectory_step  after a contact step: consumes its collision_force/torque, stress sum, alive and wrapped
 *                       centroids (floe_interactions_all.m:279-284), advances Xi Yi alpha_i Ui Vi ksi_ice h mass
 *                       inertia_moment c_alpha in place (:36-46,67-80,170-222); a sacked floe (:89,116-117) keeps its
 *                       state and is flagged, like the caller's kill(i) = i
 *   sz_get_trajectory   any pointer may be NULL; stress = mean(StressH,3) (:20), flags bit 0 sacked, bit 1 needs ocean */
typedef struct SzTrajectoryInit {
    const double *mass, *inertia, *alpha, *dXi_p, *dYi_p, *dUi_p, *dVi_p, *dalpha_p, *dksi_p, *FxOA, *FyOA, *torqueOA;   /* [n0] */
    const double *c0x, *c0y;                                                                                            /* [nverts] */
    int32_t nz;                  /* depth of the stress history (1000 in the reference, initialize_floe_values.m:24) */
    const double* stress_h;      /* [n0][nz][4] StressH to resume from (NULL = zeros(2,2,nz)); with stress_count [n0] (NULL = 1) -- a floe that
                                    moves to another rank (or a restart) carries its history along (sz_get_stress_history) */
    const int32_t* stress_count;
} SzTrajectoryInit;
typedef struct SzTrajectoryParams { double dt, HFo, xo_min, xo_max, yo_min, yo_max; } SzTrajectoryParams;   /* HFo = mean(HFo(:)); ocean grid extent (:116) */
int sz_trajectory_init(SzContext* ctx, const SzTrajectoryInit* init);
int sz_trajectory_step(SzContext* ctx, const SzTrajectoryParams* prm, int32_t* n_sacked, int32_t* n_needs_ocean);
int sz_get_stress_history(SzContext* ctx, double* stress_h /* [n0][nz][4] */, int32_t* stress_count /* [n0] */);
int sz_get_trajectory(SzContext* ctx, double* x, double* y, double* u, double* v, double* ksi, double* h, uint8_t* alive, double* mass, double* inertia, double* alpha,
                      double* dXi_p, double* dYi_p, double* dUi_p, double* dVi_p, double* dalpha_p, double* dksi_p, double* stress, int32_t* flags, double* cax, double* cay);

/* ---- row f1, second half: the ocean / atmosphere forcing of calc_trajectory.m:94-166 (Monte-Carlo points of the floe,
 * interp2 of the ocean currents and winds, quadratic drag with turning angle, SSH-tilt and Coriolis terms, averaged
 * over the points inside the floe) and the strain rate of :224-234.
 *   sz_trajectory_set_ocean    ocean.Xo [nx], ocean.Yo [ny] (ascending), ocean.Uocn/.Vocn and winds.u/.v as MATLAB stores
 *                              them (ny x nx, column-major: element (iy, ix) at iy + ix*ny), ocean.fCoriolis,
 *                              ocean.turn_angle and the constants of calc_trajectory.m:57-64 (0 = the reference's values)
 *   sz_trajectory_set_points   Floe.X, Floe.Y, Floe.A (initialize_floe_values.m:31-33): npts points per floe, floe-major
 *   sz_trajectory_ocean_forcing  call it between the contact step and sz_trajectory_step, with the same parameters:
 *                              evaluates FxOA, FyOA, torqueOA exactly for the floes the reference would -- all live floes when
 *                              do_int != 0 (doInt.flag), else only those thinner than 0.1 m after this step's thinning (:94)
 *                              -- and makes the next sz_trajectory_step compute floe.strain when do_int != 0.  A floe
 *                              with no point inside its outline would get new random points in the reference (:100-111):
 *                              it is counted in n_no_points, flagged (bit 2 of sz_get_trajectory's flags, readable until
 *                              the next sz_trajectory_step) and keeps its forcing; the call then returns SZ_ERR_STATE
 *                              after finishing the others.
 *   sz_get_trajectory_forcing  FxOA FyOA torqueOA [n0], strain [n0][2][2] (row-major); any pointer may be NULL */
typedef struct SzOcean {
    int32_t nx, ny;
    const double *Xo, *Yo;
    const double *Uocn, *Vocn, *Uwinds, *Vwinds;
    double fCoriolis, turn_angle;
    double rho0, Cd, rho_air, Cd_atm;          /* 1027, 3e-3, 1.2, 1e-3 when 0 */
} SzOcean;
int sz_trajectory_set_ocean(SzContext* ctx, const SzOcean* ocean);
int sz_trajectory_set_points(SzContext* ctx, int32_t npts, const double* X, const double* Y, const uint8_t* A);
int sz_trajectory_ocean_forcing(SzContext* ctx, const SzTrajectoryParams* prm, int32_t do_int, int32_t* n_evaluated, int32_t* n_no_points);
int sz_get_trajectory_forcing(SzContext* ctx, double* FxOA, double* FyOA, double* torqueOA, double* strain);

/* ---- SURVEY.md 8f row f3, first consumer of the contact rows: Physical_Processes/fracture_floe.m:12-52, the deformation a
 * floe receives from its deepest contact before fracture.m splits it (the Voronoi split itself draws random points and
 * stays with the host).  floe_idx: `count` floe numbers (1-based positions in the list of the last contact step, host or
 * device memory).  For each: the floe-floe row with the largest overlap (:17-22); if its partner is an original floe
 * (:26): clip 'int' (:29), centroid of the first region and its distance to the region's outline (:34-35), the partner
 * pushed half that distance along the contact force (:36-39), clip 'dif' (:40); when more than 90 % of the area is left
 * (:44) the first region becomes the floe's outline about its new centroid (:45-48).  Uses the resident positions and
 * outlines (after sz_trajectory_step when the device integrates).  Results stay on the device until fetched:
 *   changed [count], Xi Yi area [count] (the old values for unchanged floes), and for changed floes the new c_alpha as
 *   an OPEN ring in Clipper's output order: vert_off [count + 1] into cx, cy [n_verts].  Any pointer may be NULL. */
int sz_fracture_deform(SzContext* ctx, int32_t count, const int32_t* floe_idx, int64_t* n_changed, int64_t* n_verts);
int sz_get_fracture_deform(SzContext* ctx, uint8_t* changed, double* xi, double* yi, double* area, int64_t* vert_off, double* cx, double* cy);

/* ---- SURVEY.md 8f row f3, second consumer of the contact rows: Physical_Processes/corners.m:10-88, the deterministic half
 * of the corner-grinding rule -- the mask `da` of vertices "in contact" (the random half, break1 = rand > angle/Anorm :72
 * and frac_corner, stays with the host).  floe_idx: `count` floe numbers (1-based positions in the list of the last contact
 * step, host or device memory) = the selection Floe(~keep) of Subzero.m:348; nb_skip = Nb (corners.m:54 skips the first Nb
 * entries of the SELECTION).  As written in the reference: the periodic list is rebuilt from the resident positions whether
 * or not the run is periodic (:13-51, Lx/Ly of the uploaded parameters); for every contact row with a partner in that list:
 * the vertex nearest to the contact point (dsearchn :74) and every vertex in or on the partner's outline (inpolygon :78-82);
 * for wall rows (partner Inf) every vertex outside the uploaded c2_boundary (:83-86).  Result on the device until fetched:
 *   da_off [count + 1] into da [n_verts], one byte per polyshape vertex (c_alpha without its closing duplicate, stored order). */
int sz_corner_mask(SzContext* ctx, int32_t count, const int32_t* floe_idx, int32_t nb_skip, int64_t* n_verts);
int sz_get_corner_mask(SzContext* ctx, int64_t* da_off, uint8_t* da);

/* ---- calc_eulerian_data.m, mass-weighted averages of the floe state on an Nx x Ny grid over
 * [xmin, xmax] x [ymin, ymax] (= c2_boundary's extent; Lx = xmax, Ly = ymax like :28-29).  Uses the resident floes (positions,
 * outlines, rmax, area, h, Ui, Vi, alive; after sz_trajectory_step when the device integrates) and the per-floe arrays the
 * caller passes (host or device memory, [n0]; stress and strain [n0][4] = (1,1) (1,2) (2,1) (2,2); NULL = zeros; mass is
 * required).  As written in the reference: dead floes dropped (:7-8), x images for floes poking through +-Lx when periodic
 * (:39-48), the y pass that tests the LAST floe's polygon for everybody (:56-65), rows from ymax down (:72), a cell is
 * processed when the masses of its candidate floes sum to > 0 (:113-131), Aover = area of the Clipper-exact intersection of
 * the cell with the floe (:134-147), sums over ascending list position (:149-187), largest eigenvalue of the averaged stress
 * (:170-173).  Boundary floes (Nb > 0, :11-25) are not supported (SZ_ERR_ARG).
 * out: 18 planes of Ny*Nx doubles (host or device memory), element (jj, ii) at jj*Nx + ii with jj = 0 the TOP row, in the order
 *   u v du dv stress stressxx stressyx stressxy stressyy strainux strainvx strainuy strainvy c Over Mtot area h */
int sz_eulerian_data(SzContext* ctx, int32_t Nx, int32_t Ny, double xmin, double xmax, double ymin, double ymax, int32_t periodic,
                     const double* mass, const double* overlap_area, const double* dUi_p, const double* dVi_p, const double* stress, const double* strain,
                     double* out);

/* ---- SURVEY.md 8f row f4, second half: the bounding-radius pair searches of Physical_Processes/weld.m:29-81 and
 * polygon_operations/FloeSimplify.m:13-31 on the resident floes (positions after sz_trajectory_step when the device
 * integrates), over a cell grid like the contact step's broad phase.  Predicate of both (weld.m:67, FloeSimplify.m:20):
 *   alive(j) && d > 1 && d < rmax(i) + rmax(j),   d = sqrt((Xi(i)-Xi(j))^2 + (Yi(i)-Yi(j))^2),   partners ascending j.
 * mode 0 (weld): the first Nb floes are cut off (:25); the others are binned on an Nx x Ny grid, Binx = fix((Xi-xmin)/
 *   (xmax-xmin)*Nx+1) and the same in y (:35-36; xmin..ymax = min(x), max(x), min(y), max(y) of the colon vectors of :31-32),
 *   bin number (Binx-1)*Ny + Biny (:40-48), and a floe only records partners of ITS bin (:56-81).  One query per floe of the
 *   cut list; bin [n0-Nb] = bin number, 0 = in no bin; partners are 1-based positions in the CUT list Floe(1+Nb:end) (the
 *   bin-local floeNum of :68 is the rank of that position within the bin).  count / idx unused.
 * mode 1 (FloeSimplify): `count` query floes idx (1-based positions in the resident list, host or device memory) against the
 *   whole list (FloeOld of Subzero.m:170-186); partners are 1-based positions in the resident list.  Nb / grid unused; bin = 0.
 * Results stay on the device until fetched: off [nq + 1] into partner [n_partners]; any pointer may be NULL. */
int sz_pair_search(SzContext* ctx, int32_t mode, int32_t Nb, int32_t Nx, int32_t Ny, double xmin, double xmax, double ymin, double ymax,
                   int32_t count, const int32_t* idx, int64_t* n_partners);
int sz_get_pair_search(SzContext* ctx, int32_t* bin, int64_t* off, int32_t* partner);

/* diagnostic: device time (CUDA events, ms) of the last step by phase:
 * [0] ghost floes (floe_interactions_all.m:16-66)   [1] broad phase (:68-120)
 * [2] narrow phase + force law (:125-174)           [3] mirror/torque/sums (:186-265)   [4] whole step */
int sz_get_phase_ms(SzContext* ctx, float* ms5);

/* diagnostic: device time (CUDA events on the library's stream, ms) of the narrow-phase kernel launches of the last
 * step, by size class: [0] class C (strictly convex pairs, four-edge sweep)   [1] class S   [2] class T
 * [3] class M   [4] class L; and how many pairs each class received in pairs5 (either pointer may be NULL).
 * Pairs a class declines or cannot hold are counted again in the class that re-runs them. */
int sz_get_narrow_class_ms(SzContext* ctx, float* ms5, int32_t* pairs5);

/* Makes every launch and copy of the context go to the caller's CUDA stream (a cudaStream_t; 0 or NULL restores the
 * context's own non-blocking stream -- pass cudaStreamLegacy, i.e. (cudaStream_t)1, to name the default stream).  With a caller-provided stream the library's work is ordered with the caller's own
 * kernels and collectives by the stream alone, without host synchronisation (used by the multi-GPU slab step, whose halo
 * exchange runs on the caller's stream).  Synchronises the previous stream once. */
int sz_set_stream(SzContext* ctx, void* cuda_stream);

/* run-time switches of the context (experiments and tests; the defaults are the product configuration):
 *   "convex_fast"  1 (default): strictly convex floe-floe pairs go through class C first; 0: everything through the
 *                  general sweep of class S.  Results are bit-identical either way (tests/test_gpu_parity.py).
 *   "convex_split"  0 (default): class C is one kernel; 1 (experiment, not measured yet): a sweep kernel and a force-law kernel
 *                  with the intersection polygon handed over in global memory, to halve the per-thread working set.
 *                  Results are bit-identical either way (tests/test_zzzz_experiments.py).
 *   "euler_cell_warp"  1 (default): sz_eulerian_data adds a cell's items up with a warp per cell (32 items' terms at a time,
 *                  added in list order); 0: one thread per cell.  Same order of additions, same results.
 *   "speculate"     1 (default): a step takes its list length, cell grid, pair and row capacities from the step before (with
 *                  slack), every kernel reads the true counts from device memory and the host reads the counters once, at the
 *                  end of the step; a step that outgrew a capacity, or needs a larger narrow-phase size class, is flagged on
 *                  the device and repeated with measured sizes.  0: measure every size as it is needed (four counter reads).
 *                  Results are identical either way (tests/test_gpu_parity.py).
 *   "graph_safe"    0 (default); 1: see sz_step_enqueue.
 *   "apart"         1 (default): before the narrow phase, a pair of outlines of any shape (concave, any length) whose edges are
 *                  all more than 1 mm apart and neither of which holds the other's first vertex is answered without a sweep --
 *                  the reference's clip would return nothing and the pair take the zero-force branch
 *                  (collisions/floe_interactions.m:43-44,71-74); csrc/sz_apart.cuh.  0: only bounding boxes (and the
 *                  separating-axis rule for convex pairs) answer pairs.  Results are bit-identical either way
 *                  (tests/test_gpu_parity.py, tests/test_apart.py).
 * Returns SZ_ERR_ARG for an unknown name. */
int sz_set_option(SzContext* ctx, const char* name, int32_t value);
/* counters of the context: "speculated_steps" (steps that ran on carried-over sizes), "repeated_steps" (steps that had to be
 * repeated because a carried-over size was too small), "classifier_answered" (candidate pairs of the last step that needed no
 * sweep: disjoint boxes, a separating axis, or the "apart" certificate) */
int sz_get_stat(SzContext* ctx, const char* name, int64_t* value);

/* ---- stand-alone polygon clip with the gateway's semantics (private/mexclipper.cpp:204-305):
 * `count` independent (subject, clip) pairs, one closed path each, int64 coordinates, even-odd fill,
 * method 0 dif / 1 int / 2 xor / 3 uni (mexclipper.cpp:206-230).  Runs the same device sweep as the
 * narrow phase.  sz_clip_batch computes and keeps the result on the device; sz_get_clip_batch copies
 * it out in the gateway's order (item, then Clipper's path order, :299-302). */
int sz_clip_batch(SzContext* ctx, int32_t count, const int32_t* method,
                  const int64_t* subj_off, const int64_t* sx, const int64_t* sy,   /* subj_off [count+1] */
                  const int64_t* clip_off, const int64_t* cx, const int64_t* cy,   /* clip_off [count+1] */
                  int64_t* n_paths, int64_t* n_verts);
int sz_get_clip_batch(SzContext* ctx,
                      int32_t* status,            /* [count] 0 ok / SzStatus ("Clipper Error." = SZ_ERR_CLIPPER) */
                      int64_t* item_path_off,     /* [count+1] */
                      int64_t* path_vert_off,     /* [n_paths+1] */
                      int64_t* out_x, int64_t* out_y);                              /* [n_verts] */

/* ---- synthetic input of BASELINE.json configs[4]: periodic Voronoi floe field (host utility) ---- */
typedef struct SzField SzField;   /* owns the SoA arrays */
int  sz_field_voronoi(SzField** out, int32_t n_floes, uint64_t seed, double mean_area, double inflate,
                      SzParams* prm_out /* Lx, Ly, modulus filled in */);
int  sz_field_view(const SzField* f, SzFloesSoA* view);
void sz_field_free(SzField* f);

#ifdef __cplusplus
}
#endif
#endif /* SUBZERO_B200_H */
