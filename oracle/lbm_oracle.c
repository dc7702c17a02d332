/* ORACLE -- C restatement of LatticeBoltzmann.jl's collide -> stream -> BC step.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Built by oracle/Makefile into
 * oracle/_build/liblbm_oracle.so; loaded only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * Same arithmetic, in the same floating-point operation order, as
 * oracle/lbm_oracle.py (which is pinned against the reference's golden table);
 * compile with -ffp-contract=off so no FMA is formed.  tests/test_oracle_c.py
 * checks the two agree bit-for-bit (SRT/TRT) / to 1e-15 (MRT).
 * Layout: f[i][y][x] (== Julia f[x,y,i] column-major).
 *
 * Reference lines restated (under /root/reference/src):
 *   collide  : collision_models/srt.jl:18-62, trt.jl:42-97, mrt.jl:56-118
 *   moments  : velocity_distribution_function/moments.jl:3-19
 *   feq      : velocity_distribution_function/maxwell_boltzmann_equilibrium.jl:12-66,
 *              velocity_distribution_function/quadratures.jl:3-159
 *   hermite  : hermite_polynomials.jl:47-82;  a_eq: velocity_distribution_function/hermite.jl:37-77
 *   stream   : stream.jl:19-30,69-74
 *   BCs      : boundary_conditions/bounce_back.jl:8-72, moving_wall.jl:17-38
 *   loop     : lattice_boltzmann_model.jl:64-67
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define QMAX 37
#define NMAX 4

typedef struct {
    int Q;
    int cx[QMAX], cy[QMAX], opp[QMAX];
    double w[QMAX];
    double css;
    int eq_order; /* 1..4 */
    int N;        /* hermite orders */
    /* H[n][i][t]: hermite(Val{n}, c_i, q), t = column-major flat index (first index fastest) */
    double H[NMAX + 1][QMAX][16];
} lattice_t;

typedef struct {
    int model;        /* 0 SRT, 1 TRT, 2 MRT */
    double tau[16];   /* SRT: tau[0]; TRT: tau_s, tau_a; MRT: tau_n (1-based n -> tau[n-1]) */
    int force_mode;   /* 0 none, 1 uniform (fx,fy), 2 field F[2][ny][nx] */
    double fx, fy;
    const double *field;
} collision_t;

typedef struct {
    int kind;      /* 0 bounce-back, 1 moving wall */
    int dir;       /* 0 N, 1 E, 2 S, 3 W */
    int x0, x1, y0, y1; /* 1-based inclusive */
    double ux, uy, rho;
} bc_t;

static int delta(int a, int b) { return a == b; }

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* hermite(Val{n}, xi, q): hermite_polynomials.jl:47-82 */
void oracle_lattice_init(lattice_t *L, int Q, const int *cx, const int *cy, const double *w, double css,
                         const int *opp, int eq_order, int N) {
    memset(L, 0, sizeof(*L));
    L->Q = Q; L->css = css; L->eq_order = eq_order; L->N = N;
    double cs = 1 / css;
    for (int i = 0; i < Q; ++i) {
        L->cx[i] = cx[i]; L->cy[i] = cy[i]; L->opp[i] = opp[i]; L->w[i] = w[i];
        double xi[2] = {(double)cx[i], (double)cy[i]};
        L->H[1][i][0] = xi[0]; L->H[1][i][1] = xi[1];
        for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a)
            L->H[2][i][a + 2 * b] = xi[b] * xi[a] - cs * delta(a, b);
        for (int c = 0; c < 2; ++c) for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a)
            L->H[3][i][a + 2 * b + 4 * c] = xi[c] * xi[b] * xi[a]
                - cs * (xi[a] * delta(b, c) + xi[b] * delta(a, c) + xi[c] * delta(a, b));
        for (int e = 0; e < 2; ++e) for (int c = 0; c < 2; ++c) for (int b = 0; b < 2; ++b) for (int a = 0; a < 2; ++a)
            L->H[4][i][a + 2 * b + 4 * c + 8 * e] = xi[e] * xi[c] * xi[b] * xi[a]
                - cs * (xi[a] * xi[b] * delta(c, e) + xi[a] * xi[c] * delta(b, e) + xi[a] * xi[e] * delta(b, c)
                        + xi[b] * xi[c] * delta(a, e) + xi[b] * xi[e] * delta(a, c) + xi[c] * xi[e] * delta(a, b))
                + (cs * cs) * (delta(a, b) * delta(c, e) + delta(a, c) * delta(b, e) + delta(a, e) * delta(b, c));
    }
}

size_t oracle_lattice_size(void) { return sizeof(lattice_t); }

static inline double pow4(double x) { double x2 = x * x; return x2 * x2; }

/* equilibrium!(q, rho, u, 1.0, feq) */
static inline void feq_collision(const lattice_t *L, double rho, double ux, double uy, double *feq) {
    const double cs = L->css;
    const double u2 = (0.0 + ux * ux) + uy * uy;
    for (int i = 0; i < L->Q; ++i) {
        double udx = (double)L->cx[i] * ux + (double)L->cy[i] * uy;
        double a1 = cs * udx;
        double poly = 1.0 + a1;
        if (L->eq_order >= 2) {
            double a2 = ((cs * cs) * (udx * udx) + 0.0) + (-cs) * u2;
            poly = poly + (1.0 / 2) * a2;
        }
        if (L->eq_order >= 3) {
            double a3 = (cs * udx) * ((((cs * cs) * (udx * udx)) - (3 * cs) * u2) + 0.0);
            poly = poly + (1.0 / 6) * a3;
        }
        if (L->eq_order >= 4) {
            double cs3 = cs * cs * cs;
            double a4 = ((pow4(cs) * pow4(udx)) - ((6 * cs3) * u2) * (udx * udx)) + (3 * (cs * cs)) * (u2 * u2);
            poly = poly + (1.0 / 24) * a4;
        }
        feq[i] = (rho * L->w[i]) * poly;
    }
}

static inline void collide_node(const lattice_t *L, const collision_t *cm, const double *f, double Fx, double Fy,
                                double *out) {
    const int Q = L->Q;
    double rho = f[0];
    for (int i = 1; i < Q; ++i) rho = rho + f[i];
    double ux = 0.0 * f[0], uy = 0.0 * f[0];
    for (int i = 0; i < Q; ++i) ux = ux + f[i] * (double)L->cx[i];
    ux = ux / rho;
    for (int i = 0; i < Q; ++i) uy = uy + f[i] * (double)L->cy[i];
    uy = uy / rho;
    double feq[QMAX];
    if (cm->model == 0) {
        double tau = cm->tau[0];
        if (cm->force_mode) { ux = ux + tau * Fx; uy = uy + tau * Fy; }
        feq_collision(L, rho, ux, uy, feq);
        double a = (1 - 1 / tau), b = (1 / tau);
        for (int i = 0; i < Q; ++i) out[i] = a * f[i] + b * feq[i];
    } else if (cm->model == 1) {
        double ts = cm->tau[0], ta = cm->tau[1];
        if (cm->force_mode) { ux = ux + ta * Fx; uy = uy + ta * Fy; }
        feq_collision(L, rho, ux, uy, feq);
        double ws = -(1 / ts), wa = (1 / ta);
        for (int i = 0; i < Q; ++i) {
            int o = L->opp[i];
            double feq_s = 0.5 * (feq[i] + feq[o]);
            double feq_a = 0.5 * (feq[i] - feq[o]);
            double f_s = 0.5 * (f[i] + f[o]);
            double f_a = 0.5 * (f[i] - f[o]);
            out[i] = f[i] + (ws * (f_s - feq_s) - wa * (f_a - feq_a));
        }
    } else {
        const double cs = L->css;
        const int N = L->N;
        static const double fact[5] = {1, 1, 2, 6, 24};
        if (cm->force_mode) { ux = ux + cm->tau[1] * Fx; uy = uy + cm->tau[1] * Fy; }
        double u[2] = {ux, uy};
        double a_coll[NMAX + 1][16];
        double csn[NMAX + 1];
        for (int n = 2; n <= N; ++n) {
            int nt = 1 << n;
            double tn = cm->tau[n - 1];
            csn[n] = pow(cs, (double)n); /* cs^n with runtime n -> pow, mrt.jl:112 */
            for (int t = 0; t < nt; ++t) {
                /* a_eq (T = 1): rho * u[a]*u[b]*... (left-assoc), hermite.jl:45-77 */
                double prod = u[t & 1];
                for (int k = 1; k < n; ++k) prod = prod * u[(t >> k) & 1];
                double a_eq = rho * (prod + 0.0);
                double a_f = f[0] * L->H[n][0][t];
                for (int i = 1; i < Q; ++i) a_f = a_f + f[i] * L->H[n][i][t];
                a_coll[n][t] = (1 - 1 / tn) * a_f + (1 / tn) * a_eq;
            }
        }
        for (int i = 0; i < Q; ++i) {
            double first = (cs * rho) * (ux * L->H[1][i][0] + uy * L->H[1][i][1]);
            double acc = rho + first;
            if (N >= 2) {
                double hs = 0;
                for (int n = 2; n <= N; ++n) {
                    int nt = 1 << n;
                    double dot = a_coll[n][0] * L->H[n][i][0];
                    for (int t = 1; t < nt; ++t) dot = dot + a_coll[n][t] * L->H[n][i][t];
                    double term = csn[n] * dot / fact[n];
                    hs = (n == 2) ? term : hs + term;
                }
                acc = acc + hs;
            }
            out[i] = L->w[i] * acc;
        }
    }
}

void oracle_collide(const lattice_t *L, const collision_t *cm, int nx, int ny, const double *fin, double *fout) {
    const int Q = L->Q;
    const size_t plane = (size_t)nx * ny;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; ++y) {
        double f[QMAX], out[QMAX];
        for (int x = 0; x < nx; ++x) {
            size_t n = (size_t)y * nx + x;
            for (int i = 0; i < Q; ++i) f[i] = fin[i * plane + n];
            double Fx = cm->fx, Fy = cm->fy;
            if (cm->force_mode == 2) { Fx = cm->field[n]; Fy = cm->field[plane + n]; }
            collide_node(L, cm, f, Fx, Fy, out);
            for (int i = 0; i < Q; ++i) fout[i * plane + n] = out[i];
        }
    }
}

static inline int mod1m(int a, int n) { /* 0-based periodic index */
    int r = a % n;
    return r < 0 ? r + n : r;
}

void oracle_stream(const lattice_t *L, int nx, int ny, const double *fin, double *fout) {
    const size_t plane = (size_t)nx * ny;
#pragma omp parallel for schedule(static) collapse(2)
    for (int i = 0; i < L->Q; ++i)
        for (int y = 0; y < ny; ++y) {
            int ys = mod1m(y - L->cy[i], ny);
            const double *src = fin + i * plane + (size_t)ys * nx;
            double *dst = fout + i * plane + (size_t)y * nx;
            int cx = L->cx[i];
            if (nx > 8) {
                int s = mod1m(-cx, nx); /* dst[x] = src[(x - cx) mod nx] */
                memcpy(dst, src + s, (size_t)(nx - s) * sizeof(double));
                memcpy(dst + (nx - s), src, (size_t)s * sizeof(double));
            } else {
                for (int x = 0; x < nx; ++x) dst[x] = src[mod1m(x - cx, nx)];
            }
        }
}

void oracle_apply_bcs(const lattice_t *L, int nbc, const bc_t *bcs, int nx, int ny, double *fnew, const double *fold) {
    const size_t plane = (size_t)nx * ny;
    const int Q = L->Q;
    for (int b = 0; b < nbc; ++b) {
        const bc_t *bc = &bcs[b];
        for (int i = 0; i < Q; ++i) {
            int o = L->opp[i];
            double add = 0.0;
            int x0 = bc->x0, x1 = bc->x1, y0 = bc->y0, y1 = bc->y1;
            if (bc->kind == 1) {
                double a1 = L->w[i] * L->css * ((bc->rho * bc->ux) * (double)L->cx[i] + (bc->rho * bc->uy) * (double)L->cy[i]);
                add = 2 * a1;
                x0 = 1; x1 = nx; y0 = 1; y1 = ny; /* moving_wall.jl:28,32 ignores xs/ys */
            }
            if (bc->dir == 0 || bc->dir == 2) {
                for (int y = y0; y <= y1; ++y) {
                    if (bc->dir == 0 ? (y + L->cy[o] <= ny) : (y + L->cy[o] > 0)) continue;
                    for (int x = x0; x <= x1; ++x) {
                        size_t n = (size_t)(y - 1) * nx + (x - 1);
                        fnew[i * plane + n] = bc->kind == 1 ? fold[o * plane + n] + add : fold[o * plane + n];
                    }
                }
            } else {
                for (int x = x0; x <= x1; ++x) {
                    if (bc->dir == 1 ? (x + L->cx[o] <= nx) : (x + L->cx[o] > 0)) continue;
                    for (int y = y0; y <= y1; ++y) {
                        size_t n = (size_t)(y - 1) * nx + (x - 1);
                        fnew[i * plane + n] = fold[o * plane + n];
                    }
                }
            }
        }
    }
}

/* nsteps x { collide(f_stream -> f_coll); stream(f_coll -> f_stream); apply!(bcs, f_stream, f_coll) }
 * uniform/static force only (time-dependent forces are driven step by step from Python). */
void oracle_steps(const lattice_t *L, const collision_t *cm, int nbc, const bc_t *bcs, int nx, int ny,
                  double *f_stream, double *f_coll, int nsteps) {
    for (int s = 0; s < nsteps; ++s) {
        oracle_collide(L, cm, nx, ny, f_stream, f_coll);
        oracle_stream(L, nx, ny, f_coll, f_stream);
        oracle_apply_bcs(L, nbc, bcs, nx, ny, f_stream, f_coll);
    }
}
