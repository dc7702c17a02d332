/* lbm_b200.h -- C ABI of liblbm_b200.so: the B200 (sm_100a) collide-stream-BC path of
 * LatticeBoltzmann.jl, to be bound from Julia with `ccall`.
 *
 * The reference (a pure-Julia package) has no FFI seam; its seam is multiple dispatch on
 * four generic functions called from `simulate(model, time)`
 * (src/lattice_boltzmann_model.jl:60-77).  Each entry point below names the reference
 * function(s) it replaces.  All citations are relative to /root/reference.
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative lbm_status;
 *     the message of the last error on the calling thread is lbm_last_error().
 *   - host population arrays use the reference's memory order: Julia `f[x, y, i]`
 *     column-major == C `f[i][y][x]`, Float64 (src/initial_conditions.jl:13).
 *   - x, y ranges in boundary conditions are 1-based inclusive, as `bc.xs`, `bc.ys`.
 *   - one host thread per lbm_ctx at a time; contexts are independent.
 *   - the library owns all device memory, streams and the NCCL communicator; host
 *     buffers are only touched during the call that receives them.
 *   - y-slab decomposition: with world > 1, rank r owns global rows
 *     [lbm_local_rows().y0, +ny_local) and all host arrays passed to that context are the
 *     local slab `[i][ny_local][nx]`.
 */
#ifndef LBM_B200_H
#define LBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBM_ABI_VERSION 1
#define LBM_MAX_Q 37
#define LBM_MAX_TAU 16
#define LBM_MAX_BCS 8
#define LBM_NCCL_ID_BYTES 128

typedef enum {
    LBM_OK = 0,
    LBM_ERR_INVALID = -1,     /* bad argument / descriptor */
    LBM_ERR_UNSUPPORTED = -2, /* e.g. MovingWall in a direction the reference has no method for */
    LBM_ERR_CUDA = -3,
    LBM_ERR_NCCL = -4,
    LBM_ERR_STATE = -5,       /* call not valid in the current state (e.g. stream before collide) */
    LBM_ERR_NOMEM = -6
} lbm_status;

/* src/quadratures.jl:30-38 */
typedef enum { LBM_D2Q4 = 0, LBM_D2Q5, LBM_D2Q9, LBM_D2Q13, LBM_D2Q17, LBM_D2Q21, LBM_D2Q37, LBM_NUM_LATTICES } lbm_lattice;
typedef enum { LBM_F64 = 0, LBM_F32 = 1 } lbm_dtype;
/* src/collision_models/{srt,trt,mrt}.jl */
/* LBM_ITERATIVE_INIT: IterativeInitializationCollisionModel (collision_models/iterative_initialization.jl:1-60), the
 * constant-velocity SRT operator of the Mei et al. initialisation (initial_conditions/mei_et_al.jl:11-40):
 * f_out = (1 - 1/tau) f + (1/tau) (w_i rho + w_i (css c.u0 + (css^2 (c.u0)^2 - css u0.u0) / 2)), rho = sum(f), with the
 * prescribed lattice velocity u0(x, y) given by lbm_set_velocity_field.  tau[0] = tau. */
typedef enum { LBM_SRT = 0, LBM_TRT = 1, LBM_MRT = 2, LBM_ITERATIVE_INIT = 3 } lbm_collision;
/* 0: reference operation order, no FMA contraction (bit-comparable with the oracle in Float64)
 * 1: FMA contraction + algebraic simplification allowed */
typedef enum { LBM_ARITH_EXACT = 0, LBM_ARITH_FAST = 1 } lbm_arith;
typedef enum { LBM_BC_BOUNCE_BACK = 0, LBM_BC_MOVING_WALL = 1 } lbm_bc_kind;
typedef enum { LBM_NORTH = 0, LBM_EAST = 1, LBM_SOUTH = 2, LBM_WEST = 3 } lbm_direction;

/* BounceBack(direction, xs, ys)            src/boundary_conditions/bounce_back.jl:2-6
 * MovingWall(direction, xs, ys, u, rho, T) src/boundary_conditions/moving_wall.jl:5-15
 * (MovingWall ignores xs/ys and only exists for North, moving_wall.jl:17-38.) */
typedef struct {
    int32_t kind;      /* lbm_bc_kind */
    int32_t direction; /* lbm_direction */
    int32_t x0, x1, y0, y1;
    double u[2];
    double rho;
    double T;
} lbm_bc;

/* Everything `LatticeBoltzmannModel(problem, q; collision_model, ...)` fixes
 * (src/lattice_boltzmann_model.jl:15-33). */
typedef struct {
    int32_t abi_version; /* LBM_ABI_VERSION */
    int32_t nx, ny;      /* GLOBAL grid size (problem.NX, problem.NY) */
    int32_t lattice;     /* lbm_lattice */
    int32_t dtype;       /* lbm_dtype: storage + arithmetic type on the device */
    int32_t collision;   /* lbm_collision */
    int32_t arith;       /* lbm_arith */
    int32_t ntau;
    /* SRT: tau[0]=tau (srt.jl:2).  TRT: tau[0]=tau_symmetric, tau[1]=tau_asymmetric (trt.jl:2-3).
     * MRT: tau[n-1] relaxes the n-th Hermite coefficient, ntau >= 2 when a force is set (mrt.jl:94,104). */
    double tau[LBM_MAX_TAU];
    int32_t n_bcs;
    lbm_bc bcs[LBM_MAX_BCS]; /* applied in this order after streaming (boundary_conditions.jl:13-15) */
    int32_t device;          /* CUDA device ordinal */
    int32_t rank, world;     /* y-slab decomposition; world == 1: single GPU */
    uint8_t nccl_id[LBM_NCCL_ID_BYTES]; /* from lbm_nccl_unique_id() on rank 0, same on all ranks */
} lbm_desc;

typedef struct lbm_ctx lbm_ctx;

int lbm_abi_version(void);
const char *lbm_last_error(void);

/* Built-in constant tables of a quadrature, for the host to check its own against:
 * q.abscissae / q.weights / q.speed_of_sound_squared (src/quadratures/D2Q*.jl), opposite(q, i)
 * (0-based here; src/quadratures.jl:11-19 and per-lattice overrides), the truncation order of
 * the collision equilibrium (velocity_distribution_function/quadratures.jl) and div(order(q), 2).
 * Arrays must hold LBM_MAX_Q entries; any pointer may be NULL. */
int lbm_lattice_info(int32_t lattice, int32_t *q, int32_t *cx, int32_t *cy, double *w, double *css,
                     int32_t *opposite, int32_t *eq_order, int32_t *hermite_order, int32_t *halo);

/* Rank 0 creates the id; the host broadcasts the bytes to the other ranks (any transport). */
int lbm_nccl_unique_id(uint8_t id[LBM_NCCL_ID_BYTES]);

int lbm_create(const lbm_desc *desc, lbm_ctx **out);
void lbm_destroy(lbm_ctx *ctx);
int lbm_local_rows(const lbm_ctx *ctx, int32_t *y0, int32_t *ny_local);

/* model.f_stream = f / copy(model.f_stream) / copy(model.f_collision)
 * (lattice_boltzmann_model.jl:8-9).  Host Float64 arrays [q][ny_local][nx]. */
int lbm_upload_f(lbm_ctx *ctx, const double *f);
/* model.f_collision = f: for callers that hand stream!/apply! their own post-collision array
 * (stream!(q, f, f_new), apply!(bcs, q, f_new, f_old)); call after lbm_upload_f. */
int lbm_upload_f_collision(lbm_ctx *ctx, const double *f);
int lbm_download_f(lbm_ctx *ctx, double *f);
/* The same for a block of rows [y0, y0+ny) of the local slab (host array [q][ny][nx]): lets the
 * host initialise / read grids larger than its own memory chunk by chunk (32768^2: 77 GB). */
int lbm_upload_f_rows(lbm_ctx *ctx, int32_t y0, int32_t ny, const double *f_rows);
int lbm_download_f_rows(lbm_ctx *ctx, int32_t y0, int32_t ny, double *f_rows);
/* Device-side initialize(): rows [y0, y0+ny) of f_stream := hermite_based_equilibrium!(q, rho, u, T) per node
 * (src/velocity_distribution_function/hermite.jl:10-33, called by initial_conditions/analytical_equilibrium.jl:8-17,
 * constant_density.jl:10-20 through problems/problems.jl:121-128), from host Float64 fields [ny][nx] in lattice units.
 * Moves 4 instead of Q values per node over PCIe and keeps the Hermite evaluation off the host. */
int lbm_init_equilibrium_rows(lbm_ctx *ctx, int32_t y0, int32_t ny, const double *rho, const double *ux,
                              const double *uy, const double *T);
int lbm_download_f_collision(lbm_ctx *ctx, double *f);

/* The force closure `collision_model.force(x_idx, y_idx, time)` (srt.jl:10-12,52; trt.jl:17-18,77;
 * mrt.jl:40-41,92) as data, already in lattice units (lattice_force, problems/problems.jl:103-104). */
int lbm_set_force_none(lbm_ctx *ctx);
int lbm_set_force_uniform(lbm_ctx *ctx, double fx, double fy);
/* static per-node field F[2][ny_local][nx] */
int lbm_set_force_field(lbm_ctx *ctx, const double *F);
/* LBM_ITERATIVE_INIT: the velocity u0[2][ny_local][nx] (lattice units) every node is held at -- `lattice_velocity(q,
 * problem, x, y)` of iterative_initialization.jl:26, from which the kernel evaluates the operator's `nonlinear_term`. */
int lbm_set_velocity_field(lbm_ctx *ctx, const double *u0);
/* time-dependent separable force for steps t0 .. t0+nsteps-1:
 * F_x(x, y, t) = fx_of_y[t - t0][y],  F_y(x, y, t) = fy_of_x[t - t0][x]
 * (DecayingShearFlow, problems/decaying_shear_flow.jl:131-147). */
int lbm_set_force_separable(lbm_ctx *ctx, int64_t t0, int32_t nsteps, const double *fx_of_y, const double *fy_of_x);

/* collide!(model; time): f_stream -> f_collision   (lattice_boltzmann_model.jl:84-92;
 * srt.jl:18-62, trt.jl:42-97, mrt.jl:56-118).  `step` selects the row of a separable force table. */
int lbm_collide(lbm_ctx *ctx, int64_t step, double time);
/* stream!(model): f_collision -> f_stream, fully periodic pull (lattice_boltzmann_model.jl:94-96; stream.jl:19-30) */
int lbm_stream(lbm_ctx *ctx);
/* apply_boundary_conditions!(model; time)  (lattice_boltzmann_model.jl:98-106; bounce_back.jl, moving_wall.jl) */
int lbm_apply_bcs(lbm_ctx *ctx, double time);
/* The body of `for t in time` (lattice_boltzmann_model.jl:64-67), nsteps times, fused on the
 * device: for t = t0 .. t0+nsteps-1: collide(time = t*dt) -> stream -> apply BCs. Asynchronous. */
int lbm_step(lbm_ctx *ctx, int64_t t0, int64_t nsteps, double dt);
int lbm_sync(lbm_ctx *ctx);

/* Per-node hydrodynamic fields of f_stream, host Float64 arrays [ny_local][nx], NULL = skip:
 *   rho, ux, uy : density / velocity! (moments.jl:3-19), lattice units
 *   p           : pressure(q, f, rho, u) (moments.jl:21-33; 1.0 for D2Q4/D2Q5)
 *   p_track, sxx, sxy, syy : the TrackHydrodynamicErrors pressure tr(P)/D and
 *                 deviatoric_tensor(q, tau_visc, f, rho, u) (track_hydrodynamic_errors.jl:150-181,
 *                 moments.jl:81-96), tau_visc = css * lattice_viscosity(problem). */
int lbm_moments(lbm_ctx *ctx, double tau_visc, double *rho, double *ux, double *uy, double *p,
                double *p_track, double *sxx, double *sxy, double *syy);

typedef enum {
    /* out[0] = sum of lattice u_x over local nodes, out[1] = node count, out[2] = #NaN nodes
     * (MeanVelocityStoppingCriteria, stopping_criteria.jl:17-55) */
    LBM_REDUCE_MEAN_UX = 0,
    /* out[0] = sum |u - u_old|^2, out[1] = sum |u_old|^2, then u_old := u
     * (VelocityConvergenceStoppingCriteria, stopping_criteria.jl:71-115) */
    LBM_REDUCE_VELOCITY_CHANGE = 1,
    /* out[0] = sum rho, out[1] = sum rho*(ux+uy), out[2] = sum rho*(ux^2+uy^2) in lattice units
     * (track_hydrodynamic_errors.jl:200-202 before unit scaling) */
    LBM_REDUCE_CONSERVED = 2,
    /* out[0] = sum (rho - rho_old)^2 over the local nodes, out[1] = rho at the global node (NX, NY) (0 on ranks that do
     * not own it), then rho_old := rho (rho_old starts at 0).  DensityConvergence
     * (stopping_criteria/density_convergence.jl:6-17) evaluates norm(rho - rho_old) -- but its loop `for x_idx in nx,
     * y_idx in ny` (:9) visits only the node (NX, NY), so the reference's criterion is |out[1] - previous out[1]|;
     * out[0] is the norm it evidently intended. */
    LBM_REDUCE_DENSITY_CHANGE = 3
} lbm_reduce_kind;
/* Local (per-rank) partial sums; the host adds ranks. */
int lbm_reduce(lbm_ctx *ctx, int32_t kind, double *out, int32_t n);

/* TrackHydrodynamicErrors.next! (processing_methods/track_hydrodynamic_errors.jl:114-203) without moving fields
 * to the host.  The problem's analytic fields density/velocity/pressure/deviatoric_tensor(q, problem, x, y, t)
 * (the files of src/problems) are passed in separable form -- every shipped problem's fields are sums of at most two
 * products of a function of x and a function of y:
 *     E(x, y) = c0 + a[0] x[0][x] y[0][y] + a[1] x[1][x] y[1][y]      (x[k] / y[k] == NULL: all ones)
 * expected[0..7] = rho, u_x, u_y, p, sigma_xx, sigma_xy, sigma_yx, sigma_yy (dimensionless units); x tables have nx
 * entries, y tables ny_local.  tau_visc = css * lattice_viscosity, u_max = problem.u_max (unit scaling,
 * problems.jl:110-119).  out[16] (local partial sums): (rho-e)^2, |u-e_u|^2, |e_u|^2, (p-e_p)^2, e_p^2,
 * (e_sxx-sxx)^2, e_sxx^2, (e_sxy-sxy)^2, e_sxy^2, (e_syy-syy)^2, e_syy^2, (e_syx-syx)^2, e_syx^2, rho, rho(ux+uy),
 * rho(ux^2+uy^2). */
typedef struct {
    double c0;
    double a[2];
    const double *x[2];
    const double *y[2];
} lbm_sep_field;
int lbm_reduce_errors(lbm_ctx *ctx, double tau_visc, double u_max, const lbm_sep_field expected[8], double out[16]);

/* The sums of process!(problem, q, f_in, time, stats) (src/processing_methods.jl:177-239, CompareWithAnalyticalSolution)
 * for the local slab, on the device; expected[0..3] = density, velocity x / y, pressure of the problem at `time`
 * (expected[4..7] are ignored).  out[0..11]:
 *   0 sum rho   1 sum (ux+uy) rho   2 sum (kin + T)   3 sum kin, kin = (ux^2+uy^2) rho   4 sum T, T = p / rho
 *   5 sum e_rho 6 sum e_rho (e_ux+e_uy) 7 sum (e_kin + e_T) 8 sum e_kin 9 sum e_T
 *   10 sum |u - e_u|^2   11 sum (p - e_p)^2      with u = velocity / u_max, p = pressure(q, f, rho, u) (moments.jl:31-32);
 * the caller multiplies 10, 11 by the cell area and takes the roots (processing_methods.jl:232-234, 254-255). */
int lbm_reduce_process(lbm_ctx *ctx, double u_max, const lbm_sep_field expected[8], double out[16]);

/* initialize(strategy, q, problem) (src/initial_conditions.jl:7-22) evaluated entirely on the device for the local slab:
 *   f_stream := hermite_based_equilibrium!(q, rho, u, T) [+ offeq_coef w_i [rho] dot(hermite(Val{2}, c_i, q), grad u + (grad u)')]
 * with the problem's analytic fields in the separable form of lbm_sep_field (x tables: nx entries, y tables: ny_local):
 *   rho, ux, uy, p : lattice_density / lattice_velocity / pressure (problems/problems.jl:97-106); T = p / rho as in
 *                    equilibrium(q, problem, x, y) (problems.jl:121-128) unless unit_temperature; rho := 1 if unit_density
 *                    (ConstantDensity, constant_density.jl:10-20; AnalyticalVelocityAndStress, analytical_velocity_stress.jl:5-31)
 *   grad[4]        : du_x/dx, du_x/dy, du_y/dx, du_y/dy for the off-equilibrium strategies
 *                    (analytical_offequilibrium.jl:10-87: offeq = 1 generic, offeq = 2 the TGV method with the factor rho)
 * The host moves O(nx + ny) numbers instead of Q (or 4) per node. */
typedef struct {
    lbm_sep_field rho, ux, uy, p;
    lbm_sep_field grad[4];
    int32_t unit_density, unit_temperature;
    int32_t offeq; /* 0: equilibrium only; 1: + offeq_coef w_i dot(H2_i, S); 2: + offeq_coef w_i rho dot(H2_i, S), S = grad + grad' */
    double offeq_coef;
} lbm_init_spec;
int lbm_init_analytic(lbm_ctx *ctx, const lbm_init_spec *spec);

/* Page-locked host memory for population arrays that cross the boundary often (snapshots, array-level operators): copies
 * from / to such arrays run at the full PCIe rate and asynchronously.  Plain malloc'ed arrays are accepted everywhere. */
int lbm_host_alloc(void **ptr, size_t bytes);
/* lbm_upload_f / lbm_download_f for page-locked arrays WITHOUT the final synchronisation: the copies are enqueued on the
 * context's stream.  With two contexts in flight, job k + 1's upload and job k - 1's download overlap job k's lbm_step
 * (PCIe is full duplex), so a stream of independent jobs runs at the device rate instead of copy + compute + copy.
 * `f` must stay valid (and, for uploads, unmodified) until lbm_sync or another synchronising call on the context.
 * LBM_ERR_INVALID for pageable arrays. */
int lbm_upload_f_async(lbm_ctx *ctx, const double *f);
int lbm_download_f_async(lbm_ctx *ctx, double *f);
int lbm_host_free(void *ptr);

/* TakeSnapshots.next! (src/processing_methods/take_snapshots.jl:12-29: push!(snapshots, copy(f_in))) without stalling the
 * step loop.  lbm_snapshot_begin enqueues one kernel that writes f_stream of the current state into a compact device
 * buffer (the ping-pong buffers and the fused state machine are not disturbed: no materialisation, no extra collide-only
 * launch afterwards) and a device-to-host copy on a separate copy stream, then returns; lbm_step calls that follow run
 * concurrently with the copy.  lbm_snapshot_end waits for the copy; `f` ([Q][NY_local][NX] doubles, the layout of
 * lbm_download_f) must stay valid until then.  Page-locked `f` (lbm_host_alloc) receives the copy directly; otherwise the
 * library stages through its own page-locked buffer and lbm_snapshot_end does the final host copy.  At most one snapshot
 * is in flight per context: a second lbm_snapshot_begin first completes the previous one. */
int lbm_snapshot_begin(lbm_ctx *ctx, double *f);
int lbm_snapshot_end(lbm_ctx *ctx);

/* Introspection used by bench.py / tests. */
int64_t lbm_kernel_launches(const lbm_ctx *ctx); /* kernels launched by this context so far */
/* How halos travel between y-slabs: 0 = single GPU (ghost cells written by the kernels themselves), 1 = NCCL
 * send/recv on a side stream, 2 = peer memory (the boundary-row launch stores into the neighbours' ghost rows over
 * NVLink and hand-shakes through device-side flags; chosen at lbm_create when every rank could map its
 * neighbours' buffers, lbm_set_option("p2p", 0) or LBM_P2P=0 select 1). */
int lbm_halo_path(const lbm_ctx *ctx);
int lbm_last_step_ms(lbm_ctx *ctx, float *ms);    /* CUDA-event time of the last lbm_step batch */
/* CUDA events on the library's own stream (torch.cuda.Event cannot see it): start, ...work..., stop
 * (synchronises and returns the elapsed device time). */
int lbm_timer_start(lbm_ctx *ctx);
int lbm_timer_stop(lbm_ctx *ctx, float *ms);
int lbm_set_option(lbm_ctx *ctx, const char *key, int64_t value); /* tuning knobs, see DESIGN.md */

/* ---- Batched small problems -------------------------------------------------------------------------------------
 * The reference's parameter studies are loops of `simulate(problem, q; ...)` over tiny grids: 902 500 solves of a 3 x 5
 * D2Q9 TRT Poiseuille flow, <= 5001 steps each (examples/notebooks/trt_magic_parameter.ipynb:30-103, 3 h on one thread),
 * and the 950-point diagonal of the same table in poiseuille.ipynb (cell 9).  A lbm_batch holds `nbatch` independent
 * problems of ONE shape -- the descriptor's grid, lattice, collision kind, dtype and boundary conditions -- each with
 * its own relaxation times and uniform force; lbm_batch_run advances all of them with one launch that keeps every
 * problem's populations on chip, evaluates the stop criterion there every `check_every` steps exactly as
 * TrackHydrodynamicErrors.next! does (track_hydrodynamic_errors.jl:52-59: `mod(t, 100) == 0` -> should_stop!) and
 * retires the problems that fire it.  Host population arrays are [problem][q][ny][nx] Float64 (one Julia `f[x, y, i]`
 * after the other).  Problems must fit in shared memory (LBM_ERR_UNSUPPORTED otherwise; use one lbm_ctx per problem). */
typedef struct lbm_batch lbm_batch;
typedef enum {
    LBM_BATCH_STOP_OFF = 0,
    /* MeanVelocityStoppingCriteria (stopping_criteria.jl:17-55): |mean(u_x) / previous mean - 1| < tolerance */
    LBM_BATCH_STOP_MEAN_VELOCITY = 1,
    /* VelocityConvergenceStoppingCriteria (stopping_criteria.jl:71-115): sqrt(sum |u - u_old|^2) / sum |u_old|^2 < tolerance */
    LBM_BATCH_STOP_VELOCITY_CONVERGENCE = 2
} lbm_batch_stop_kind;
typedef struct {
    int32_t kind;        /* lbm_batch_stop_kind */
    int32_t check_every; /* 100 in the reference */
    double tolerance;
} lbm_batch_stop;

/* desc: nx, ny, lattice, dtype, collision (SRT / TRT / MRT), arith, ntau, n_bcs / bcs, device; world must be 1.  desc.tau
 * is the initial value of every problem's relaxation times. */
int lbm_batch_create(const lbm_desc *desc, int32_t nbatch, lbm_batch **out);
void lbm_batch_destroy(lbm_batch *b);
/* tau[nbatch][desc.ntau]: the collision model's relaxation times per problem (TRT(tau_s, tau_a, force), trt.jl:2-3) */
int lbm_batch_set_tau(lbm_batch *b, const double *tau);
/* fxy[nbatch][2]: uniform lattice force per problem (lattice_force(problem, ...), poiseuille.jl:72-82); NULL: no force */
int lbm_batch_set_force_uniform(lbm_batch *b, const double *fxy);
/* f_stream of problems [first, first + count); resets their step counter, stop flag and criterion memory */
int lbm_batch_upload_f(lbm_batch *b, int32_t first, int32_t count, const double *f);
/* the same initial f_stream [q][ny][nx] for every problem (initialize(strategy, q, problem) of a sweep over tau) */
int lbm_batch_broadcast_f(lbm_batch *b, const double *f);
int lbm_batch_download_f(lbm_batch *b, int32_t first, int32_t count, double *f);
/* Every problem that has not stopped takes up to nsteps collide -> stream -> BC steps (lattice_boltzmann_model.jl:64-67);
 * with a stop criterion it is evaluated whenever the problem's total step count t is a multiple of check_every and the
 * problem freezes at that t when it fires.  Asynchronous.  stop == NULL: no criterion. */
int lbm_batch_run(lbm_batch *b, int64_t nsteps, const lbm_batch_stop *stop);
/* (synchronises) steps taken so far and whether the criterion has fired, per problem; either pointer may be NULL */
int lbm_batch_status(lbm_batch *b, int32_t first, int32_t count, int64_t *steps_done, int32_t *stopped);
/* lbm_reduce_errors for every problem: out[nbatch][16].  The separable tables of expected[8] are shared; coef
 * [nbatch][8][3] = (c0, a[0], a[1]) per problem and field replaces the coefficients in expected (NULL: use those). */
int lbm_batch_reduce_errors(lbm_batch *b, const double *tau_visc, const double *u_max, const lbm_sep_field expected[8],
                            const double *coef, double *out);
int lbm_batch_last_run_ms(lbm_batch *b, float *ms); /* CUDA-event time of the last lbm_batch_run */
int64_t lbm_batch_kernel_launches(const lbm_batch *b);

#ifdef __cplusplus
}
#endif
#endif /* LBM_B200_H */
