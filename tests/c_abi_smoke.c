/* C consumer of include/lbm_b200.h -- what a non-Python binding (the Julia `ccall` shim, a C host) sees.
 *
 *   c_abi_smoke layout
 *       prints `struct.field offset size` for every field of the ABI structs and `sizeof struct N` lines;
 *       tests/test_abi.py compares them with the ctypes mirror (lbm/_abi.py) and with the Julia struct definitions --
 *       catches header <-> binding drift that a size check cannot.
 *   c_abi_smoke run <liblbm_b200.so> <f0.bin> <want.bin> <nx> <ny> <nsteps> <lattice> <collision> <tau0> <tau1> <fx> <fy> <walls>
 *       fills lbm_desc from C, uploads f0 ([q][ny][nx] doubles), steps, downloads and compares bit for bit with want
 *       (written by the oracle); also drives a 3-problem lbm_batch through the same steps.  Exit code 0 = identical.
 *
 * The library is dlopen'ed and every entry point is bound through the header's own prototype
 * (__typeof__), so a signature change in the header or the library breaks this file at compile time or at run time.
 */
#include <dlfcn.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/lbm_b200.h"

#define FIELD(S, f) printf("%s.%s %zu %zu\n", #S, #f, offsetof(S, f), sizeof(((S *)0)->f))

static int layout(void) {
    FIELD(lbm_bc, kind); FIELD(lbm_bc, direction); FIELD(lbm_bc, x0); FIELD(lbm_bc, x1); FIELD(lbm_bc, y0); FIELD(lbm_bc, y1);
    FIELD(lbm_bc, u); FIELD(lbm_bc, rho); FIELD(lbm_bc, T);
    FIELD(lbm_desc, abi_version); FIELD(lbm_desc, nx); FIELD(lbm_desc, ny); FIELD(lbm_desc, lattice); FIELD(lbm_desc, dtype);
    FIELD(lbm_desc, collision); FIELD(lbm_desc, arith); FIELD(lbm_desc, ntau); FIELD(lbm_desc, tau); FIELD(lbm_desc, n_bcs);
    FIELD(lbm_desc, bcs); FIELD(lbm_desc, device); FIELD(lbm_desc, rank); FIELD(lbm_desc, world); FIELD(lbm_desc, nccl_id);
    FIELD(lbm_sep_field, c0); FIELD(lbm_sep_field, a); FIELD(lbm_sep_field, x); FIELD(lbm_sep_field, y);
    FIELD(lbm_batch_stop, kind); FIELD(lbm_batch_stop, check_every); FIELD(lbm_batch_stop, tolerance);
    printf("sizeof lbm_bc %zu\nsizeof lbm_desc %zu\nsizeof lbm_sep_field %zu\nsizeof lbm_batch_stop %zu\n", sizeof(lbm_bc),
           sizeof(lbm_desc), sizeof(lbm_sep_field), sizeof(lbm_batch_stop));
    printf("const LBM_ABI_VERSION %d\nconst LBM_MAX_Q %d\nconst LBM_MAX_TAU %d\nconst LBM_MAX_BCS %d\nconst LBM_NCCL_ID_BYTES %d\n",
           LBM_ABI_VERSION, LBM_MAX_Q, LBM_MAX_TAU, LBM_MAX_BCS, LBM_NCCL_ID_BYTES);
    printf("enum LBM_D2Q37 %d\nenum LBM_F32 %d\nenum LBM_ITERATIVE_INIT %d\nenum LBM_ARITH_FAST %d\nenum LBM_BC_MOVING_WALL %d\n"
           "enum LBM_WEST %d\nenum LBM_REDUCE_DENSITY_CHANGE %d\nenum LBM_BATCH_STOP_VELOCITY_CONVERGENCE %d\nenum LBM_ERR_NOMEM %d\n",
           LBM_D2Q37, LBM_F32, LBM_ITERATIVE_INIT, LBM_ARITH_FAST, LBM_BC_MOVING_WALL, LBM_WEST, LBM_REDUCE_DENSITY_CHANGE,
           LBM_BATCH_STOP_VELOCITY_CONVERGENCE, LBM_ERR_NOMEM);
    return 0;
}

#define BIND(name)                                                        \
    __typeof__(name) *p_##name = (__typeof__(name) *)dlsym(lib, #name);   \
    if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }
#define CHECK(call)                                                                            \
    do {                                                                                       \
        int rc_ = (call);                                                                      \
        if (rc_ != 0) { fprintf(stderr, "%s -> %d: %s\n", #call, rc_, p_lbm_last_error()); return 3; } \
    } while (0)

static double *read_doubles(const char *path, size_t n) {
    FILE *fp = fopen(path, "rb");
    if (!fp) { perror(path); return NULL; }
    double *buf = (double *)malloc(n * sizeof(double));
    if (fread(buf, sizeof(double), n, fp) != n) { fprintf(stderr, "%s: short read\n", path); fclose(fp); free(buf); return NULL; }
    fclose(fp);
    return buf;
}

int main(int argc, char **argv) {
    if (argc >= 2 && !strcmp(argv[1], "layout")) return layout();
    if (argc != 15 || strcmp(argv[1], "run")) { fprintf(stderr, "usage: see the header comment\n"); return 1; }
    void *lib = dlopen(argv[2], RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
    BIND(lbm_abi_version) BIND(lbm_last_error) BIND(lbm_lattice_info) BIND(lbm_create) BIND(lbm_destroy) BIND(lbm_local_rows)
    BIND(lbm_upload_f) BIND(lbm_download_f) BIND(lbm_set_force_none) BIND(lbm_set_force_uniform) BIND(lbm_step) BIND(lbm_sync)
    BIND(lbm_reduce) BIND(lbm_kernel_launches) BIND(lbm_halo_path)
    BIND(lbm_batch_create) BIND(lbm_batch_destroy) BIND(lbm_batch_set_tau) BIND(lbm_batch_set_force_uniform)
    BIND(lbm_batch_broadcast_f) BIND(lbm_batch_download_f) BIND(lbm_batch_run) BIND(lbm_batch_status)
    if (p_lbm_abi_version() != LBM_ABI_VERSION) { fprintf(stderr, "ABI version %d != header %d\n", p_lbm_abi_version(), LBM_ABI_VERSION); return 2; }

    const int nx = atoi(argv[5]), ny = atoi(argv[6]), nsteps = atoi(argv[7]), lattice = atoi(argv[8]), collision = atoi(argv[9]);
    const double tau0 = atof(argv[10]), tau1 = atof(argv[11]), fx = atof(argv[12]), fy = atof(argv[13]);
    const int walls = atoi(argv[14]);
    int32_t Q = 0;
    CHECK(p_lbm_lattice_info(lattice, &Q, NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL));
    const size_t n = (size_t)Q * nx * ny;
    double *f0 = read_doubles(argv[3], n), *want = read_doubles(argv[4], n), *got = (double *)malloc(n * sizeof(double));
    if (!f0 || !want || !got) return 2;

    lbm_desc d;
    memset(&d, 0, sizeof(d));
    d.abi_version = LBM_ABI_VERSION;
    d.nx = nx; d.ny = ny; d.lattice = lattice; d.dtype = LBM_F64; d.collision = collision; d.arith = LBM_ARITH_EXACT;
    d.ntau = collision == LBM_TRT ? 2 : 1;
    d.tau[0] = tau0; d.tau[1] = tau1;
    if (walls) {  /* Poiseuille: bounce-back North + South over the whole width */
        d.n_bcs = 2;
        d.bcs[0].kind = LBM_BC_BOUNCE_BACK; d.bcs[0].direction = LBM_NORTH;
        d.bcs[1].kind = LBM_BC_BOUNCE_BACK; d.bcs[1].direction = LBM_SOUTH;
        for (int b = 0; b < 2; ++b) { d.bcs[b].x0 = 1; d.bcs[b].x1 = nx; d.bcs[b].y0 = 1; d.bcs[b].y1 = ny; d.bcs[b].rho = 1.0; d.bcs[b].T = 1.0; }
    }
    d.device = 0; d.rank = 0; d.world = 1;

    lbm_ctx *ctx = NULL;
    CHECK(p_lbm_create(&d, &ctx));
    int32_t y0 = -1, nyl = -1;
    CHECK(p_lbm_local_rows(ctx, &y0, &nyl));
    if (y0 != 0 || nyl != ny || p_lbm_halo_path(ctx) != 0) { fprintf(stderr, "single-GPU context reports rows %d+%d\n", y0, nyl); return 4; }
    if (fx != 0.0 || fy != 0.0) CHECK(p_lbm_set_force_uniform(ctx, fx, fy));
    else CHECK(p_lbm_set_force_none(ctx));
    CHECK(p_lbm_upload_f(ctx, f0));
    CHECK(p_lbm_step(ctx, 0, nsteps / 2, 1.0));
    CHECK(p_lbm_step(ctx, nsteps / 2, nsteps - nsteps / 2, 1.0));
    CHECK(p_lbm_sync(ctx));
    CHECK(p_lbm_download_f(ctx, got));
    double sums[4] = {0, 0, 0, 0};
    CHECK(p_lbm_reduce(ctx, LBM_REDUCE_CONSERVED, sums, 4));
    const long long launches = (long long)p_lbm_kernel_launches(ctx);
    p_lbm_destroy(ctx);
    size_t bad = 0;
    double maxd = 0;
    for (size_t k = 0; k < n; ++k) {
        if (memcmp(&got[k], &want[k], sizeof(double)) != 0) ++bad;
        const double dd = got[k] > want[k] ? got[k] - want[k] : want[k] - got[k];
        if (dd > maxd) maxd = dd;
    }
    double mass = 0;
    for (size_t k = 0; k < n; ++k) mass += want[k];
    printf("ctx: %zu of %zu values differ, max abs diff %.3e, mass %.15g vs %.15g, kernel launches %lld\n", bad, n, maxd, sums[0], mass, launches);
    if (bad || launches < 1) return 5; /* a persistent launch covers many steps */
    if (!(sums[0] > mass * (1 - 1e-12) && sums[0] < mass * (1 + 1e-12))) return 6;

    /* the same solve three times over as a batch (every problem must reproduce `want`) */
    lbm_batch *batch = NULL;
    CHECK(p_lbm_batch_create(&d, 3, &batch));
    double taus[6] = {tau0, tau1, tau0, tau1, tau0, tau1}, taus1[3] = {tau0, tau0, tau0}, forces[6] = {fx, fy, fx, fy, fx, fy};
    CHECK(p_lbm_batch_set_tau(batch, d.ntau == 2 ? taus : taus1));
    CHECK(p_lbm_batch_set_force_uniform(batch, (fx != 0.0 || fy != 0.0) ? forces : NULL));
    CHECK(p_lbm_batch_broadcast_f(batch, f0));
    {   /* the batch kernel keeps every problem in shared memory: a problem too large for that is refused, not run slowly */
        int rc_ = p_lbm_batch_run(batch, nsteps, NULL);
        if (rc_ == LBM_ERR_UNSUPPORTED && (size_t)nx * ny * 9 * sizeof(double) > 200000) {
            printf("batch: skipped (%d x %d does not fit on chip: %s)\n", nx, ny, p_lbm_last_error());
            p_lbm_batch_destroy(batch);
            free(f0); free(want); free(got);
            return 0;
        }
        if (rc_ != 0) { fprintf(stderr, "lbm_batch_run -> %d: %s\n", rc_, p_lbm_last_error()); return 3; }
    }
    int64_t steps_done[3] = {0, 0, 0};
    int32_t stopped[3] = {1, 1, 1};
    CHECK(p_lbm_batch_status(batch, 0, 3, steps_done, stopped));
    double *gotb = (double *)malloc(3 * n * sizeof(double));
    CHECK(p_lbm_batch_download_f(batch, 0, 3, gotb));
    p_lbm_batch_destroy(batch);
    size_t badb = 0;
    for (int p = 0; p < 3; ++p) {
        if (steps_done[p] != nsteps || stopped[p] != 0) { fprintf(stderr, "batch problem %d: steps %lld stopped %d\n", p, (long long)steps_done[p], stopped[p]); return 7; }
        for (size_t k = 0; k < n; ++k) badb += memcmp(&gotb[p * n + k], &want[k], sizeof(double)) != 0;
    }
    printf("batch: %zu of %zu values differ\n", badb, 3 * n);
    free(f0); free(want); free(got); free(gotb);
    return badb ? 8 : 0;
}
