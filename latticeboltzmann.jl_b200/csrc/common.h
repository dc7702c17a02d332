// Types shared by the context code (lbm_b200.cu) and the per-lattice kernel translation units
// (kernels_inst.cu compiled once per <lattice, arithmetic mode>).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lattices.h"

namespace lbm {

// Device layout of one population buffer (all planes identical):
//   element (i, y, x), y in [-GY, nyl+GY), x in [-GX, nx+GX)  ->  base[i*plane + y*pitch + x]
// where `base` already points at (0, 0, 0).  Ghost cells hold periodic images (x always; y when
// world == 1) or the neighbouring slab's rows (world > 1), so the pull never wraps.

struct BCd {  // device view of one boundary condition (global, 1-based inclusive ranges)
    int kind, dir;
    int x0, x1, y0, y1;
    double ax, ay;  // MovingWall: equilibrium_coefficient(Val{1}) = rho_w * u_w (moving_wall.jl:20)
};

// slots of the per-context flag block used by the peer-memory halo protocol (unsigned long long each)
enum {
    P2P_EPOCH_FROM_DOWN = 0,  // last epoch published by the down neighbour
    P2P_EPOCH_FROM_UP = 1,    // ... by the up neighbour
    P2P_TIMEOUT = 2,          // != 0: a wait gave up (protocol error); reported as LBM_ERR_STATE
    P2P_CTA_COUNT = 3,        // CTAs of the running boundary launch that have finished
    P2P_BATCH_FROM_DOWN = 4,  // lbm_step batch tokens
    P2P_BATCH_FROM_UP = 5,
    P2P_EPOCH_BASE = 6,       // epoch of the state a replayed CUDA graph starts from (set before each replay)
    P2P_MAGIC = 7,            // written at create time; read through the mapping by the neighbours as a self-check
    P2P_NFLAGS = 8
};

template <typename T>
struct KParams {
    const T *src;  // pulled / read buffer
    T *dst;        // written buffer
    const T *aux;  // f_old (post-collision) for the standalone BC kernel
    // per-plane base pointers so that a node address is ONE 32-bit multiply-add per population:
    //   srcn[i] = src + i*plane                       (unshifted)
    //   srcp[i] = src + i*plane - c_y[i]*pitch - c_x[i] (pull source of population i)
    //   dstp[i] = dst + i*plane
    // indexed with the node offset n = y*pitch + x (< 2^31, checked by lbm_create)
    const T *srcn[LBM_MAX_Q];
    const T *srcp[LBM_MAX_Q];
    T *dstp[LBM_MAX_Q];
    long long pitch, plane;
    int nx, nyl;   // local interior size
    int y0g, nyg;  // first global row of this slab, global row count
    // launched rows r -> local row: r < row_an ? row_a0 + r : (r - row_an < row_bn ? row_b0 + r - row_an : row_c0 + r - row_an - row_bn)
    int row_a0, row_an, row_b0, row_bn, row_c0, nrows;
    int st_mode;   // tuning builds: cache operator of the population stores (0 default, 1 .cs, 2 .cg, 3 .wt)
    int pf_rows;   // > 0: L2-prefetch distance in rows for the fused pull (0 = off)
    int p2p_rows;  // P2P launches: the first p2p_rows launched rows are the slab's edge rows -- only the CTAs that start
                   // inside them wait for the neighbours' epoch and publish this one (nrows: every CTA, as in a launch of
                   // edge rows only)
    int wrap_y;    // 1: this slab is the whole periodic domain in y -> kernels write the y images too
    // collision constants (already converted to T):
    //   SRT: c[0] = 1 - 1/tau, c[1] = 1/tau, shift = tau
    //   TRT: c[0] = -(1/tau_s), c[1] = 1/tau_a, shift = tau_a
    //   MRT: c[2n] = 1 - 1/tau_n, c[2n+1] = 1/tau_n (n = 2..4), k[n] = css^n / n!, shift = tau_2
    T c[12];
    T kn[5];
    T shift;
    int mrt_skip[5];  // tau_n == 1 -> a_coll[n] = a_eq[n] exactly; skip the projection of f
    // forcing
    int force_mode;  // 0 none, 1 uniform, 2 field, 3 separable table
    T fx, fy;
    const T *field;   // [2][nyl][nx]
    const T *sep_fx;  // [nsteps][nyl]
    const T *sep_fy;  // [nsteps][nx]
    long long sep_t0;
    // peer-memory halo exchange (world > 1 with the neighbours' buffers mapped, see kernels_inst.cu): only read by
    // the P2P instances of the fused kernel
    T *peer_up, *peer_dn;               // (plane 0, row 0, x 0) of the neighbours' buffer with the same role as dst
    long long plane_up, plane_dn;       // their plane strides (slabs may differ by one row; the pitch is common)
    int nyl_dn;                         // rows of the down neighbour (its top ghost rows start there)
    unsigned long long *flags;          // local flag block, slots P2P_*
    unsigned long long *flag_at_up;     // up neighbour's flags[P2P_EPOCH_FROM_DOWN]  (I am its `down`)
    unsigned long long *flag_at_dn;     // down neighbour's flags[P2P_EPOCH_FROM_UP]  (I am its `up`)
    unsigned long long epoch;           // sequence number of the state this launch produces ...
    const unsigned long long *epoch_base;  // ... plus *epoch_base when set (launches replayed from a CUDA graph)
    // boundary conditions
    int nbc;
    int bc_sides;  // bit d set: some boundary condition faces direction d (lbm_direction) -> only those edges take the BC path
    BCd bc[LBM_MAX_BCS];
};

struct MomentsOut {  // device Float64 arrays [nyl][nx] or nullptr
    double *rho, *ux, *uy, *p, *p_track, *sxx, *sxy, *syy;
    double tau_visc;
};

struct ReduceArgs {
    int kind;
    double *partials;  // [nblocks][4]
    double *u_old;     // [2][nyl][nx] for LBM_REDUCE_VELOCITY_CHANGE
    double *out;       // [4] device
    int nblocks;       // capacity of `partials` (the launcher replaces it by the grid size)
    int rows_per_cta;  // set by the launcher
    long long zero;    // 0 (run-time constant for after_load)
};

// Expected (analytic) fields for the on-device error norms, in separable form
//   E_f(x, y) = c0[f] + a[f][0] X[f][0](x) Y[f][0](y) + a[f][1] X[f][1](x) Y[f][1](y)
// f = rho, ux, uy, p, sxx, sxy, syx, syy.  tab holds, per field and term, X (nx values) then Y (nyl values).
struct ErrorArgs {
    double c0[8];
    double a[8][2];
    const double *tab;   // [8][2][nx + nyl]
    double tau_visc;     // css * lattice_viscosity
    double u_max;        // dimensionless_velocity / dimensionless_stress scaling
    double *partials;    // [nblocks][16]
    double *out;         // [16] device
    int nblocks;         // capacity of `partials` (the launcher replaces it by the grid size)
    int rows_per_cta;    // set by the launcher
    const double *tx[16], *ty[16];  // per slot: X table (nx entries), Y table (nyl entries) -- set by the launcher from tab
    unsigned mask;       // bit 2 f + k: term k of field f has a non-zero coefficient (set by the launcher)
    int mode;            // 0: TrackHydrodynamicErrors sums, 1: the sums of process! (CompareWithAnalyticalSolution)
};

// ----------------------------------------------------------------------------------------------
// Batched small problems (lbm_batch_*): B independent problems of one shape, each with its own relaxation times and
// uniform force, advanced by ONE persistent launch that keeps every problem's populations in shared memory for the
// whole run and evaluates the stop criterion on chip (batch.cuh).
// ----------------------------------------------------------------------------------------------
template <typename T>
struct BatchConsts {  // per problem; the same constants make_params derives from the relaxation times
    T c[12];
    T kn[5];
    T shift;
    T fx, fy;
    int mrt_skip[5];
    int forced;
};

// device-side names of lbm_batch_stop_kind (include/lbm_b200.h)
enum { LBM_BATCH_STOP_NONE = 0, LBM_BATCH_STOP_MEAN_UX = 1, LBM_BATCH_STOP_VELOCITY_CHANGE = 2 };

struct BatchParams {
    int nx, nyg, nb;       // problem shape (nyg = NY: the BC helpers read p.nyg), number of problems
    void *f;               // T [nb][Q][NY][NX]: f_stream of every problem
    const void *consts;    // BatchConsts<T> [nb]
    double *crit;          // stop-criterion state: [nb][2 N] previous velocity field (x-major pairs) or [nb][..][0] previous mean
    long long *steps_done; // [nb] lattice steps taken so far (the reference's t)
    int *stopped;          // [nb] the criterion fired at steps_done
    long long nsteps;      // advance every running problem by at most this many steps
    int stop_kind, check_every;
    double tol;
    int has_mw;            // some boundary condition is a moving wall (additive term table in shared memory)
    int nbc, bc_sides;
    BCd bc[LBM_MAX_BCS];
};

struct BatchErrorArgs {
    const void *f;            // T [nb][Q][NY][NX]
    int nx, ny, nb;
    const double *tau_visc;   // [nb]
    const double *u_max;      // [nb]
    const double *coef;       // [nb][8][3]: c0, a0, a1 of every expected field
    const double *tab;        // [8][2][nx + ny] shared separable tables (X then Y per field and term)
    double *out;              // [nb][16]
};

// ----------------------------------------------------------------------------------------------
// Persistent multi-step kernel (persist.cuh): one cooperative launch takes `nsteps` fused steps of a slab; the CTAs
// are co-resident, each owns a contiguous range of nodes and synchronises only with the CTAs that own the rows within
// the stencil's reach (and, on the slab edges, with the neighbouring GPUs through the peer-memory epoch flags).
// ----------------------------------------------------------------------------------------------
struct PersistArgs {
    int nsteps;
    long long step0;               // lattice step index of the first step (force tables are indexed by it)
    unsigned long long *done;      // [grid] per-CTA count of completed steps of this launch (zeroed before the launch)
    unsigned long long *edge_count;  // [2] CTAs that have finished the bottom / top boundary rows of the running step
    unsigned long long *error;     // != 0: a wait gave up
    unsigned long long epoch0;     // peer-memory epoch of the state the launch starts from
    long long npc;                 // nodes per CTA
    int nctas;                     // CTAs that own nodes
    int n_bot, n_top;              // CTAs owning nodes of the bottom / top H rows
};

// Device-side initialize(): analytic fields in separable form (same layout as ErrorArgs): rho, ux, uy, p (lattice units),
// du_x/dx, du_x/dy, du_y/dx, du_y/dy.
struct InitArgs {
    double c0[8];
    double a[8][2];
    const double *tab;   // [8][2][nx + nyl]
    int unit_density, unit_temperature;
    int offeq;           // 0: equilibrium only, 1: + coef w_i dot(H2_i, grad + grad'), 2: the same times rho
    double coef;
};

// Launchers exported by one kernels_inst.cu instance.
struct Ops {
    int lattice, arith;
    // collide (+ optional fused pull of the previous step's stream + BCs)
    void (*step64)(int cm, bool pull, const KParams<double> &p, long long step, int variant, cudaStream_t s);
    void (*step32)(int cm, bool pull, const KParams<float> &p, long long step, int variant, cudaStream_t s);
    // fused pull + collide of boundary rows that also pushes them into the neighbours' ghost rows (peer memory)
    void (*step64_p2p)(int cm, const KParams<double> &p, long long step, int variant, cudaStream_t s);
    void (*step32_p2p)(int cm, const KParams<float> &p, long long step, int variant, cudaStream_t s);
    void (*p2p_barrier)(unsigned long long *flags, unsigned long long *at_up, unsigned long long *at_dn, unsigned long long token, cudaStream_t s);
    void (*p2p_wait_epoch)(unsigned long long *flags, unsigned long long need, cudaStream_t s);
    void (*p2p_set_base)(unsigned long long *flags, unsigned long long base, cudaStream_t s);
    // periodic pull (+ BCs when p.nbc > 0) without collision
    void (*stream64)(const KParams<double> &p, cudaStream_t s);
    void (*stream32)(const KParams<float> &p, cudaStream_t s);
    // standalone apply!(bcs, q, f_new = dst, f_old = aux)
    void (*bcs64)(const KParams<double> &p, cudaStream_t s);
    void (*bcs32)(const KParams<float> &p, cudaStream_t s);
    // ghost refresh of `dst` (periodic images)
    void (*ghosts64)(const KParams<double> &p, cudaStream_t s);
    void (*ghosts32)(const KParams<float> &p, cudaStream_t s);
    void (*moments64)(bool pull, const KParams<double> &p, const MomentsOut &m, cudaStream_t s);
    void (*moments32)(bool pull, const KParams<float> &p, const MomentsOut &m, cudaStream_t s);
    void (*reduce64)(bool pull, const KParams<double> &p, const ReduceArgs &r, cudaStream_t s);
    void (*reduce32)(bool pull, const KParams<float> &p, const ReduceArgs &r, cudaStream_t s);
    void (*errors64)(bool pull, const KParams<double> &p, const ErrorArgs &e, cudaStream_t s);
    void (*errors32)(bool pull, const KParams<float> &p, const ErrorArgs &e, cudaStream_t s);
    // host f64 [q][nyl][nx] staging <-> device storage conversion (f32 stores f - w)
    void (*import32)(const KParams<float> &p, const double *staging, int plane_idx, cudaStream_t s);
    void (*export32)(const KParams<float> &p, double *staging, int plane_idx, cudaStream_t s);
    // device-side hermite_based_equilibrium! from host-provided (rho, ux, uy, T) rows
    void (*init_eq64)(const KParams<double> &p, const double *rho, const double *ux, const double *uy, const double *T, cudaStream_t s);
    void (*init_eq32)(const KParams<float> &p, const double *rho, const double *ux, const double *uy, const double *T, cudaStream_t s);
    void (*init_analytic64)(const KParams<double> &p, const InitArgs &ia, cudaStream_t s);
    void (*init_analytic32)(const KParams<float> &p, const InitArgs &ia, cudaStream_t s);
    // f_stream of the current state as compact Float64 [Q][nyl][nx] (pull: the state holds post-collision populations)
    void (*snapshot64)(bool pull, const KParams<double> &p, double *out, cudaStream_t s);
    void (*snapshot32)(bool pull, const KParams<float> &p, double *out, cudaStream_t s);
    // TMA-staged fused pull step (tma.cuh); tmap: CUtensorMap of the source buffer; returns -1 when not available
    int (*step_tma64)(int cm, const KParams<double> &p, const void *tmap, long long step, int gx, int gy, int cfg, cudaStream_t s);
    int (*step_tma32)(int cm, const KParams<float> &p, const void *tmap, long long step, int gx, int gy, int cfg, cudaStream_t s);
    // persistent multi-step kernel: co-resident grid (CTAs, threads) for this <collision model, dtype>, 0 CTAs if the
    // device cannot launch cooperatively; pa: src = current state, pb: the two buffers swapped
    void (*persist_grid64)(int cm, bool p2p, int *ctas, int *threads);
    void (*persist_grid32)(int cm, bool p2p, int *ctas, int *threads);
    int (*persist64)(int cm, bool p2p, const KParams<double> &pa, const KParams<double> &pb, const PersistArgs &a, int ctas, int threads, cudaStream_t s);
    int (*persist32)(int cm, bool p2p, const KParams<float> &pa, const KParams<float> &pb, const PersistArgs &a, int ctas, int threads, cudaStream_t s);
    // batched small problems: returns 0, or -1 when one problem does not fit in shared memory
    int (*batch64)(int cm, const BatchParams &p, cudaStream_t s);
    int (*batch32)(int cm, const BatchParams &p, cudaStream_t s);
    void (*batch_errors64)(const BatchErrorArgs &e, cudaStream_t s);
    void (*batch_errors32)(const BatchErrorArgs &e, cudaStream_t s);
    int (*init_constants)();  // uploads the __constant__ lattice tables on the current device
};

const Ops *get_ops(int lattice, int arith);

}  // namespace lbm
