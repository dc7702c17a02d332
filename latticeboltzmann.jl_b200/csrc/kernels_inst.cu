// Hand-written sm_100a kernels of the collide -> stream -> boundary-condition path, compiled once
// per <LBM_LATTICE, LBM_FAST>:
//   LBM_FAST=0  ("exact"): built with -fmad=false; every expression keeps the operation order of
//               the reference (and of oracle/), so Float64 SRT/TRT results are bit-identical to
//               the oracle.
//   LBM_FAST=1  ("fast"):  same source, FMA contraction on.
// Reference lines restated here (under /root/reference/src):
//   density/velocity!  velocity_distribution_function/moments.jl:3-19
//   equilibrium!       velocity_distribution_function/maxwell_boltzmann_equilibrium.jl:12-66,
//                      velocity_distribution_function/quadratures.jl:3-159
//   collide!           collision_models/srt.jl:18-62, trt.jl:42-97, mrt.jl:56-118
//   stream!            stream.jl:19-30,69-74
//   apply!             boundary_conditions/bounce_back.jl:8-72, moving_wall.jl:17-38
//   diagnostics        moments.jl:21-96, processing_methods/track_hydrodynamic_errors.jl:134-186,
//                      processing_methods/stopping_criteria/stopping_criteria.jl:17-115
#include <type_traits>
#include <utility>
#include "common.h"

#ifndef LBM_LATTICE
#error "compile with -DLBM_LATTICE=<lbm_lattice value> -DLBM_FAST=<0|1>"
#endif
#ifndef LBM_FAST
#define LBM_FAST 0
#endif
#define LBM_CAT2(a, b, c) a##b##_##c
#define LBM_CAT(a, b, c) LBM_CAT2(a, b, c)
#define LBM_NS LBM_CAT(inst_, LBM_LATTICE, LBM_FAST)

namespace lbm {
namespace LBM_NS {

using L = Lat<LBM_LATTICE>;
constexpr int Q = L::Q;
constexpr int H = L::H;
constexpr int NH = L::N;

// ------------------------------------------------------------------------------------------
// lattice constants in constant memory (operands of DFMA/FFMA come straight from the bank)
// ------------------------------------------------------------------------------------------
template <typename T>
struct LatConst {
    T w[Q];
    T css, cs_inv;
    T H2[Q][3];  // hermite(Val{2}, c_i, q): xx, xy, yy          (hermite_polynomials.jl:47-51)
    T H3[Q][4];  // xxx, xxy, xyy, yyy                            (:52-61)
    T H4[Q][5];  // xxxx, xxxy, xxyy, xyyy, yyyy                  (:63-82)
};
__constant__ LatConst<double> c_lat64;
__constant__ LatConst<float> c_lat32;

template <typename T> __device__ __forceinline__ const LatConst<T> &LC();
template <> __device__ __forceinline__ const LatConst<double> &LC<double>() { return c_lat64; }
template <> __device__ __forceinline__ const LatConst<float> &LC<float>() { return c_lat32; }

static double hermite_entry(int n, const int *idx, const int *xi, double cs) {
    auto d = [](int a, int b) { return a == b ? 1 : 0; };
    if (n == 2) { int a = idx[0], b = idx[1]; return (double)(xi[b] * xi[a]) - cs * d(a, b); }
    if (n == 3) {
        int a = idx[0], b = idx[1], c = idx[2];
        return (double)(xi[c] * xi[b] * xi[a]) - cs * (double)(xi[a] * d(b, c) + xi[b] * d(a, c) + xi[c] * d(a, b));
    }
    int a = idx[0], b = idx[1], c = idx[2], e = idx[3];
    return (double)(xi[e] * xi[c] * xi[b] * xi[a])
           - cs * (double)(xi[a] * xi[b] * d(c, e) + xi[a] * xi[c] * d(b, e) + xi[a] * xi[e] * d(b, c)
                           + xi[b] * xi[c] * d(a, e) + xi[b] * xi[e] * d(a, c) + xi[c] * xi[e] * d(a, b))
           + (cs * cs) * (double)(d(a, b) * d(c, e) + d(a, c) * d(b, e) + d(a, e) * d(b, c));
}

static int init_constants() {
    LatticeInfo li;
    lattice_info(LBM_LATTICE, li);
    LatConst<double> h;
    LatConst<float> hf;
    h.css = li.css;
    h.cs_inv = 1 / li.css;
    for (int i = 0; i < Q; ++i) {
        h.w[i] = li.w[i];
        const int xi[2] = {li.cx[i], li.cy[i]};
        for (int k = 0; k <= 2; ++k) { int idx[2] = {0, 0}; for (int j = 0; j < k; ++j) idx[1 - j] = 1; h.H2[i][k] = hermite_entry(2, idx, xi, h.cs_inv); }
        for (int k = 0; k <= 3; ++k) { int idx[3] = {0, 0, 0}; for (int j = 0; j < k; ++j) idx[2 - j] = 1; h.H3[i][k] = hermite_entry(3, idx, xi, h.cs_inv); }
        for (int k = 0; k <= 4; ++k) { int idx[4] = {0, 0, 0, 0}; for (int j = 0; j < k; ++j) idx[3 - j] = 1; h.H4[i][k] = hermite_entry(4, idx, xi, h.cs_inv); }
    }
    hf.css = (float)h.css; hf.cs_inv = (float)h.cs_inv;
    for (int i = 0; i < Q; ++i) {
        hf.w[i] = (float)h.w[i];
        for (int k = 0; k < 3; ++k) hf.H2[i][k] = (float)h.H2[i][k];
        for (int k = 0; k < 4; ++k) hf.H3[i][k] = (float)h.H3[i][k];
        for (int k = 0; k < 5; ++k) hf.H4[i][k] = (float)h.H4[i][k];
    }
    if (cudaMemcpyToSymbol(c_lat64, &h, sizeof(h)) != cudaSuccess) return -1;
    if (cudaMemcpyToSymbol(c_lat32, &hf, sizeof(hf)) != cudaSuccess) return -1;
    return 0;
}

// ------------------------------------------------------------------------------------------
// compile-time loops over populations
// ------------------------------------------------------------------------------------------
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}

// c * u for a compile-time integer c, with the trivial cases folded (value-identical to the
// reference's Float64(c) * u up to the sign of zero).
template <int C, typename T>
__device__ __forceinline__ T cmul(T u) {
    if constexpr (C == 0) return T(0);
    else if constexpr (C == 1) return u;
    else if constexpr (C == -1) return -u;
    else return T(C) * u;
}
template <int CX, int CY, typename T>
__device__ __forceinline__ T cdot(T ux, T uy) {  // abscissae[1,i]*u[1] + abscissae[2,i]*u[2]
    if constexpr (CX == 0 && CY == 0) return T(0);
    else if constexpr (CY == 0) return cmul<CX>(ux);
    else if constexpr (CX == 0) return cmul<CY>(uy);
    else return cmul<CX>(ux) + cmul<CY>(uy);
}

// ------------------------------------------------------------------------------------------
// boundary-condition resolution for the fused pull
// ------------------------------------------------------------------------------------------
// Which boundary condition (last in list order wins, boundary_conditions.jl:13-15) overwrites
// population i at the 1-based global node (x1, y1)?  (cxo, cyo) = c[opposite(i)].  -1: none.
// (P: KParams<T> or BatchParams -- anything with nbc, bc[], nx, nyg)
template <class P>
__device__ __noinline__ int resolve_bc(const P &p, int x1, int y1, int cxo, int cyo) {
    for (int b = p.nbc - 1; b >= 0; --b) {
        const BCd &bc = p.bc[b];
        bool hit;
        switch (bc.dir) {
        case LBM_NORTH: hit = (y1 + cyo > p.nyg); break;  // bounce_back.jl:15, moving_wall.jl:29
        case LBM_SOUTH: hit = (y1 + cyo <= 0); break;     // bounce_back.jl:31
        case LBM_EAST: hit = (x1 + cxo > p.nx); break;    // bounce_back.jl:48
        default: hit = (x1 + cxo <= 0); break;            // bounce_back.jl:64
        }
        if (!hit) continue;
        if (bc.kind == LBM_BC_MOVING_WALL) return b;  // ignores xs/ys (moving_wall.jl:28,32)
        if (x1 >= bc.x0 && x1 <= bc.x1 && y1 >= bc.y0 && y1 <= bc.y1) return b;
    }
    return -1;
}

template <class P>
__device__ __forceinline__ bool near_wall(const P &p, int x, int yg) {
    const int m = p.bc_sides;
    return m != 0 && (((m >> LBM_WEST) & 1 && x < H) || ((m >> LBM_EAST) & 1 && x >= p.nx - H) ||
                      ((m >> LBM_SOUTH) & 1 && yg < H) || ((m >> LBM_NORTH) & 1 && yg >= p.nyg - H));
}

// f_new[x,y,i] = f_old[x,y,opp(i)] (+ 2 a_1 for a moving wall), a_1 = w_i * css * dot(rho_w u_w, c_i)
template <int I, typename T, class P>
__device__ __forceinline__ T bounced(const P &p, int b, T f_opp) {
    if (p.bc[b].kind == LBM_BC_MOVING_WALL) {
        const LatConst<T> &c = LC<T>();
        const T ax = T(p.bc[b].ax), ay = T(p.bc[b].ay);
        const T a1 = (c.w[I] * c.css) * (ax * T(L::cx(I)) + ay * T(L::cy(I)));
        return f_opp + 2 * a1;
    }
    return f_opp;
}

// ------------------------------------------------------------------------------------------
// node load / store
// ------------------------------------------------------------------------------------------
// PULL = false: f[i] = src[i, y, x]
// PULL = true : f[i] = (stream + BCs)(src)[i, y, x]  i.e. src[i, y - c_y, x - c_x] unless a boundary
//               condition overwrites it with src[opp(i), y, x].
// COHERENT: read through L2 (ld.global.cg) instead of the non-coherent path -- the persistent kernel re-reads addresses
//           that other CTAs rewrote since this SM last cached them.
template <bool COHERENT, typename T>
__device__ __forceinline__ T ld_pop(const T *q) {
    if constexpr (COHERENT) return __ldcg(q);
    else return __ldg(q);
}

template <typename T, bool PULL, bool COHERENT = false>
__device__ __forceinline__ void load_node(const KParams<T> &p, int x, int y, T (&f)[Q]) {
    const unsigned n = (unsigned)y * (unsigned)p.pitch + (unsigned)x;
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        if constexpr (PULL) f[i] = ld_pop<COHERENT>(p.srcp[i] + n);
        else f[i] = ld_pop<COHERENT>(p.srcn[i] + n);
    });
    if constexpr (PULL) {
        const int yg = p.y0g + y;
        if (near_wall(p, x, yg)) {
            static_for<0, Q>([&](auto I) {
                constexpr int i = decltype(I)::value;
                constexpr int o = L::opp(i);
                if constexpr (i != o) {
                    const int b = resolve_bc(p, x + 1, yg + 1, L::cx(o), L::cy(o));
                    if (b >= 0) f[i] = bounced<i>(p, b, ld_pop<COHERENT>(p.srcn[o] + n));
                }
            });
        }
    }
}

// After a node has been stored: copy it to its periodic images inside the ghost frame (edge
// threads only; all planes share the offsets).  The values are read back from dst (the thread's own
// stores, L1/L2 hits) so that the collision outputs need not stay in registers.
template <typename T>
__device__ __forceinline__ void store_images(const KParams<T> &p, int x, int y) {
    const bool ex = (x < H) || (x >= p.nx - H);
    const bool ey = p.wrap_y && ((y < H) || (y >= p.nyl - H));
    if (ex || ey) {
        T *d = p.dst + (long long)y * p.pitch + x;
        const int kx_lo = -((x + H) / p.nx), kx_hi = (p.nx + H - 1 - x) / p.nx;
        int ky_lo = 0, ky_hi = 0;
        if (p.wrap_y) { ky_lo = -((y + H) / p.nyl); ky_hi = (p.nyl + H - 1 - y) / p.nyl; }
        for (int ky = ky_lo; ky <= ky_hi; ++ky)
            for (int kx = kx_lo; kx <= kx_hi; ++kx) {
                if (kx == 0 && ky == 0) continue;
                T *g = d + (long long)(ky * p.nyl) * p.pitch + kx * p.nx;
#pragma unroll 1
                for (int i = 0; i < Q; ++i) g[i * p.plane] = d[i * p.plane];
            }
    }
}

template <typename T, bool GHOSTS>
__device__ __forceinline__ void store_node(const KParams<T> &p, int x, int y, T (&out)[Q]) {
    const unsigned n = (unsigned)y * (unsigned)p.pitch + (unsigned)x;
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        p.dstp[i][n] = out[i];
    });
    if constexpr (GHOSTS) store_images(p, x, y);
}

template <typename T>
__device__ __forceinline__ int launched_row(const KParams<T> &p, int r) {
    if (r < p.row_an) return p.row_a0 + r;
    r -= p.row_an;
    return r < p.row_bn ? p.row_b0 + r : p.row_c0 + (r - p.row_bn);
}

// ------------------------------------------------------------------------------------------
// moments and equilibrium
// ------------------------------------------------------------------------------------------
// Float32 contexts store the deviation g_i = f_i - w_i ("shifted populations") and do all
// arithmetic on deviations: rho = 1 + sum(g), j = sum(g c) (sum(w c) = 0), feq_i - w_i =
// w_i (drho + rho P_i(u)) with P the equilibrium polynomial without its leading 1.  This keeps
// ~7 significant digits on the O(1e-4) velocities instead of losing them against f ~ w.
template <typename T>
struct Shifted { static constexpr bool value = std::is_same<T, float>::value; };

template <typename T>
__device__ __forceinline__ void rho_u(const T (&f)[Q], T &rho, T &ux, T &uy, T &drho) {
    // density = sum(f) (left fold, moments.jl:3); velocity! (moments.jl:5-19)
    drho = f[0];
    static_for<1, Q>([&](auto I) { drho = drho + f[decltype(I)::value]; });
    if constexpr (Shifted<T>::value) rho = T(1) + drho;
    else rho = drho;
    T jx = T(0), jy = T(0);
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        if constexpr (L::cx(i) != 0) jx = jx + cmul<L::cx(i)>(f[i]);
    });
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        if constexpr (L::cy(i) != 0) jy = jy + cmul<L::cy(i)>(f[i]);
    });
    ux = jx / rho;
    uy = jy / rho;
}

template <typename T>
__device__ __forceinline__ T pow4(T x) { const T x2 = x * x; return x2 * x2; }

// _equilibrium(q, rho, w_i, u.c_i, u.u, T = 1, ...) for compile-time i.
// Shifted storage: returns feq_i - w_i = w_i (drho + rho (poly - 1)).
template <int I, typename T>
__device__ __forceinline__ T feq_i(T rho, T ux, T uy, T u2, T drho) {
    const LatConst<T> &c = LC<T>();
    const T cs = c.css;
    constexpr bool REST = (L::cx(I) == 0 && L::cy(I) == 0);
    constexpr bool SH = Shifted<T>::value;
    const T udx = cdot<L::cx(I), L::cy(I)>(ux, uy);
    T poly;
    if constexpr (REST) poly = SH ? T(0) : T(1);
    else if constexpr (SH) poly = cs * udx;
    else poly = T(1) + cs * udx;
    if constexpr (L::EQ_ORDER >= 2) {
        T a2;
        if constexpr (REST) a2 = (-cs) * u2;
        else a2 = (cs * cs) * (udx * udx) + (-cs) * u2;
        poly = poly + T(0.5) * a2;
    }
    if constexpr (L::EQ_ORDER >= 3 && !REST) {
        const T a3 = (cs * udx) * ((cs * cs) * (udx * udx) - (3 * cs) * u2);
        poly = poly + T(1.0 / 6) * a3;
    }
    if constexpr (L::EQ_ORDER >= 4) {
        T a4;
        if constexpr (REST) a4 = (3 * (cs * cs)) * (u2 * u2);
        else a4 = ((pow4(cs) * pow4(udx)) - ((6 * (cs * cs * cs)) * u2) * (udx * udx)) + (3 * (cs * cs)) * (u2 * u2);
        poly = poly + T(1.0 / 24) * a4;
    }
    if constexpr (SH) return c.w[I] * (drho + rho * poly);
    else return (rho * c.w[I]) * poly;
}

// fast mode only: symmetric / antisymmetric parts of the pair (i, opp(i)) directly --
//   feq_s = rho w (1 + a2/2 + a4/24),  feq_a = rho w (a1 + a3/6)   (a1, a3 odd in c.u; a2, a4 even)
// (shifted storage: feq_s - w = w (drho + rho (a2/2 + a4/24))).  Same polynomial as feq_i, fewer
// operations, different rounding -- which is why the exact mode does not use it.
template <int I, typename T>
__device__ __forceinline__ void feq_sym_asym(T rho, T ux, T uy, T u2, T drho, T &es, T &ea) {
    const LatConst<T> &c = LC<T>();
    const T cs = c.css;
    constexpr bool SH = Shifted<T>::value;
    const T udx = cdot<L::cx(I), L::cy(I)>(ux, uy);
    const T a1 = cs * udx;
    T even = SH ? T(0) : T(1);
    T odd = a1;
    if constexpr (L::EQ_ORDER >= 2) even = even + T(0.5) * (a1 * a1 - cs * u2);
    if constexpr (L::EQ_ORDER >= 3) odd = odd + T(1.0 / 6) * (a1 * (a1 * a1 - (3 * cs) * u2));
    if constexpr (L::EQ_ORDER >= 4) {
        const T b = cs * u2;
        even = even + T(1.0 / 24) * ((a1 * a1) * (a1 * a1 - 6 * b) + 3 * (b * b));
    }
    const T rw = rho * c.w[I];
    if constexpr (SH) es = c.w[I] * drho + rw * even;
    else es = rw * even;
    ea = rw * odd;
}

// ------------------------------------------------------------------------------------------
// collision operators (Float64 / raw populations)
// ------------------------------------------------------------------------------------------
// `emit(I, value)` receives each post-collision population as soon as it is known (it is stored
// right away, so no second Q-sized array is live).
// (P: KParams<T> or BatchConsts<T> -- anything with c[], kn[], shift, mrt_skip[])
template <int CM, typename T, class P, class Emit>
__device__ __forceinline__ void collide_node(const P &p, const T (&f)[Q], bool forced, T Fx, T Fy, Emit &&emit) {
    T rho, ux, uy, drho;
    rho_u(f, rho, ux, uy, drho);
    if constexpr (CM == LBM_ITERATIVE_INIT) {
        // iterative_initialization.jl:42-60: the node keeps the prescribed velocity u0 = (Fx, Fy); only rho comes from f.
        // nonlinear_term (:21-35) = w_i rho_0 (a_H_1 + a_H_2 / 2), rho_0 = 1, is evaluated here instead of being stored.
        const LatConst<T> &c = LC<T>();
        const T cs = c.css;
        const T u2 = Fx * Fx + Fy * Fy;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            const T udx = cdot<L::cx(i), L::cy(i)>(Fx, Fy);
            const T a1 = cs * udx;
            const T a2 = (cs * cs) * (udx * udx) - cs * u2;
            const T nl = c.w[i] * (a1 + a2 / T(2));
            const T feq = c.w[i] * (Shifted<T>::value ? drho : rho) + nl;  // shifted storage: feq - w
            emit(I, p.c[0] * f[i] + p.c[1] * feq);
        });
        return;
    }
    if (forced) {  // equilibrium velocity shift u + tau F (srt.jl:54, trt.jl:79, mrt.jl:94)
        ux = ux + p.shift * Fx;
        uy = uy + p.shift * Fy;
    }
    if constexpr (CM == LBM_SRT) {
        const T u2 = ux * ux + uy * uy;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            emit(I, p.c[0] * f[i] + p.c[1] * feq_i<i>(rho, ux, uy, u2, drho));  // srt.jl:58
        });
    } else if constexpr (CM == LBM_TRT) {
        // trt.jl:82-94, processed per pair (i, opp(i)) so that only f (not a second array of feq)
        // stays live.  Bit-identical to the per-population form: f_s, feq_s are symmetric in
        // (i, o) and f_a, feq_a change sign exactly.
        const T u2 = ux * ux + uy * uy;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            constexpr int o = L::opp(i);
            if constexpr (i == o) {
                const T fe = feq_i<i>(rho, ux, uy, u2, drho);
                const T feq_s = T(0.5) * (fe + fe), feq_a = T(0.5) * (fe - fe);
                const T f_s = T(0.5) * (f[i] + f[i]), f_a = T(0.5) * (f[i] - f[i]);
                emit(I, f[i] + (p.c[0] * (f_s - feq_s) - p.c[1] * (f_a - feq_a)));
            } else if constexpr (i < o) {
                T feq_s, feq_a;
                if constexpr (LBM_FAST) {
                    feq_sym_asym<i>(rho, ux, uy, u2, drho, feq_s, feq_a);
                } else {
                    const T fi = feq_i<i>(rho, ux, uy, u2, drho), fo = feq_i<o>(rho, ux, uy, u2, drho);
                    feq_s = T(0.5) * (fi + fo);
                    feq_a = T(0.5) * (fi - fo);
                }
                const T f_s = T(0.5) * (f[i] + f[o]), f_a = T(0.5) * (f[i] - f[o]);
                const T s_part = p.c[0] * (f_s - feq_s), a_part = p.c[1] * (f_a - feq_a);
                emit(I, f[i] + (s_part - a_part));
                emit(std::integral_constant<int, o>{}, f[o] + (s_part + a_part));
            }
        });
    } else {
        // regularised MRT (mrt.jl:56-118) with the symmetric tensors reduced to their unique
        // components (a^(n) has n+1 of them, multiplicity C(n,k)) and the populations folded per opposite
        // pair: H_n(-c) = (-1)^n H_n(c), so the even orders only need f_i + f_opp and the odd order f_i - f_opp --
        // half the multiply-adds of the projection and of the reconstruction (the kernel is FMA-bound on the
        // wide lattices: profiles/r02/ncu_summary.md).
        const LatConst<T> &c = LC<T>();
        T a2[3] = {T(0), T(0), T(0)}, a3[4] = {T(0), T(0), T(0), T(0)}, a4[5] = {T(0), T(0), T(0), T(0), T(0)};
        const bool do2 = NH >= 2 && !p.mrt_skip[2], do3 = NH >= 3 && !p.mrt_skip[3], do4 = NH >= 4 && !p.mrt_skip[4];
        if (do2 || do3 || do4) {
            static_for<0, Q>([&](auto I) {
                constexpr int i = decltype(I)::value;
                constexpr int o = L::opp(i);
                if constexpr (i <= o) {
                    const T fs = (i == o) ? f[i] : f[i] + f[o];
                    if constexpr (NH >= 2) {
                        if (do2) {
#pragma unroll
                            for (int k = 0; k < 3; ++k) a2[k] = a2[k] + fs * c.H2[i][k];
                        }
                    }
                    if constexpr (NH >= 3 && i != o) {  // H_3(0) = 0
                        if (do3) {
                            const T fa = f[i] - f[o];
#pragma unroll
                            for (int k = 0; k < 4; ++k) a3[k] = a3[k] + fa * c.H3[i][k];
                        }
                    }
                    if constexpr (NH >= 4) {
                        if (do4) {
#pragma unroll
                            for (int k = 0; k < 5; ++k) a4[k] = a4[k] + fs * c.H4[i][k];
                        }
                    }
                }
            });
        }
        if constexpr (NH >= 2) {
            const T e[3] = {rho * (ux * ux), rho * (ux * uy), rho * (uy * uy)};
            const T m[3] = {T(1), T(2), T(1)};
#pragma unroll
            for (int k = 0; k < 3; ++k) a2[k] = (p.kn[2] * m[k]) * (p.c[4] * a2[k] + p.c[5] * e[k]);
        }
        if constexpr (NH >= 3) {
            const T e[4] = {rho * (ux * ux * ux), rho * (ux * ux * uy), rho * (ux * uy * uy), rho * (uy * uy * uy)};
            const T m[4] = {T(1), T(3), T(3), T(1)};
#pragma unroll
            for (int k = 0; k < 4; ++k) a3[k] = (p.kn[3] * m[k]) * (p.c[6] * a3[k] + p.c[7] * e[k]);
        }
        if constexpr (NH >= 4) {
            const T x2 = ux * ux, y2 = uy * uy;
            const T e[5] = {rho * (x2 * x2), rho * (x2 * ux * uy), rho * (x2 * y2), rho * (ux * uy * y2), rho * (y2 * y2)};
            const T m[5] = {T(1), T(4), T(6), T(4), T(1)};
#pragma unroll
            for (int k = 0; k < 5; ++k) a4[k] = (p.kn[4] * m[k]) * (p.c[8] * a4[k] + p.c[9] * e[k]);
        }
        const T csrho = c.css * rho;
        static_for<0, Q>([&](auto I) {  // mrt.jl:107-114, per pair: out_i = w (even + odd), out_opp = w (even - odd)
            constexpr int i = decltype(I)::value;
            constexpr int o = L::opp(i);
            if constexpr (i <= o) {
                // shifted storage: sum(w H_n) = 0 for n <= N, so a_f is the projection of g and out - w = w (drho + ...)
                T even = Shifted<T>::value ? drho : rho;
                T odd = csrho * cdot<L::cx(i), L::cy(i)>(ux, uy);
                if constexpr (NH >= 2) {
                    T hs = a2[0] * c.H2[i][0];
                    hs = hs + a2[1] * c.H2[i][1];
                    hs = hs + a2[2] * c.H2[i][2];
                    if constexpr (NH >= 4) {
#pragma unroll
                        for (int k = 0; k < 5; ++k) hs = hs + a4[k] * c.H4[i][k];
                    }
                    even = even + hs;
                }
                if constexpr (NH >= 3 && i != o) {
                    T ho = a3[0] * c.H3[i][0];
#pragma unroll
                    for (int k = 1; k < 4; ++k) ho = ho + a3[k] * c.H3[i][k];
                    odd = odd + ho;
                }
                if constexpr (i == o) {
                    emit(I, c.w[i] * (even + odd));
                } else {
                    emit(I, c.w[i] * (even + odd));
                    emit(std::integral_constant<int, o>{}, c.w[i] * (even - odd));
                }
            }
        });
    }
}

template <typename T>
__device__ __forceinline__ bool load_force(const KParams<T> &p, int x, int y, long long step, T &Fx, T &Fy) {
    Fx = T(0); Fy = T(0);
    switch (p.force_mode) {
    case 0: return false;
    case 1: Fx = p.fx; Fy = p.fy; return true;
    case 2: {
        const long long n = (long long)y * p.nx + x;
        Fx = __ldg(p.field + n);
        Fy = __ldg(p.field + (long long)p.nyl * p.nx + n);
        return true;
    }
    default: {
        const long long s = step - p.sep_t0;
        Fx = __ldg(p.sep_fx + s * p.nyl + y);
        Fy = __ldg(p.sep_fy + s * p.nx + x);
        return true;
    }
    }
}

// ------------------------------------------------------------------------------------------
// peer-memory halo exchange (device side; y-slabs, world > 1, after the peers' buffers were mapped)
// ------------------------------------------------------------------------------------------
// The launch that produces a slab's 2H boundary rows writes them straight into the neighbours' ghost rows
// over NVLink and hand-shakes through flags in peer memory: no NCCL call and no copy kernel in the step
// loop.  Protocol (KParams::flags): the launch producing state `epoch`
//   1. waits until both neighbours have published epoch - 1: their boundary rows of that state are then in
//      my ghost rows, and their boundary launch no longer reads the ghost rows (of the other buffer, which
//      is the one with the role of my dst) that I am about to overwrite;
//   2. computes its rows, stores them locally and into the neighbours' ghost rows;
//   3. the last CTA to finish publishes `epoch` to both neighbours (release at system scope).
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin (one thread) until flags[a] >= need and flags[b] >= need.  Bounded: a protocol error sets
// flags[P2P_TIMEOUT] (reported by the host as LBM_ERR_STATE) instead of hanging the GPU.
__device__ __noinline__ void p2p_spin(unsigned long long *flags, int a, int b, unsigned long long need, unsigned long long limit_ns) {
    if (ld_acquire_sys(flags + a) >= need && ld_acquire_sys(flags + b) >= need) return;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flags + a) < need || ld_acquire_sys(flags + b) < need) {
        if (globaltimer_ns() - t0 > limit_ns) { atomicMax(flags + P2P_TIMEOUT, need ? need : 1ULL); break; }
        __nanosleep(100);
    }
}

template <typename T>
__device__ __forceinline__ unsigned long long p2p_epoch(const KParams<T> &p) {
    return p.epoch + (p.epoch_base ? *p.epoch_base : 0ULL);
}

template <typename T>
__device__ __forceinline__ void p2p_wait(const KParams<T> &p) {
    if (threadIdx.x == 0 && threadIdx.y == 0) p2p_spin(p.flags, P2P_EPOCH_FROM_DOWN, P2P_EPOCH_FROM_UP, p2p_epoch(p) - 1, 10000000000ULL);
    __syncthreads();
}

// After a boundary node has been stored locally: copy the populations the neighbour will pull (c_y > 0 at
// least as large as the distance to my top edge go up, c_y < 0 go down) and their x images into its ghost rows.
template <typename T>
__device__ __noinline__ void p2p_store(const KParams<T> &p, int x, int y) {
    const T *d = p.dst + (long long)y * p.pitch + x;
    const int kx_lo = -((x + H) / p.nx), kx_hi = (p.nx + H - 1 - x) / p.nx;
    if (y >= p.nyl - H && p.peer_up) {  // my top rows are the up neighbour's bottom ghost rows -H .. -1
        const int dist = p.nyl - y;
        T *g = p.peer_up + (long long)(y - p.nyl) * p.pitch + x;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if constexpr (L::cy(i) > 0) {
                if (L::cy(i) >= dist) {
                    const T v = d[i * p.plane];
                    for (int kx = kx_lo; kx <= kx_hi; ++kx) g[i * p.plane_up + kx * p.nx] = v;
                }
            }
        });
    }
    if (y < H && p.peer_dn) {  // my bottom rows are the down neighbour's top ghost rows nyl_dn .. nyl_dn + H - 1
        const int dist = y + 1;
        T *g = p.peer_dn + (long long)(p.nyl_dn + y) * p.pitch + x;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if constexpr (L::cy(i) < 0) {
                if (-L::cy(i) >= dist) {
                    const T v = d[i * p.plane];
                    for (int kx = kx_lo; kx <= kx_hi; ++kx) g[i * p.plane_dn + kx * p.nx] = v;
                }
            }
        });
    }
}

// End of a producing launch: every thread makes its (peer) stores visible system-wide, the last CTA to
// finish publishes the epoch to both neighbours.
template <typename T>
__device__ __forceinline__ void p2p_signal(const KParams<T> &p, unsigned long long total) {
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        if (atomicAdd(p.flags + P2P_CTA_COUNT, 1ULL) == total - 1) {
            p.flags[P2P_CTA_COUNT] = 0;
            __threadfence_system();
            // a launch whose own wait gave up computed with stale ghost rows: do not publish, so the failure
            // propagates (the neighbours time out too) instead of spreading wrong rows
            if (*(volatile unsigned long long *)(p.flags + P2P_TIMEOUT) == 0ULL) {
                const unsigned long long e = p2p_epoch(p);
                st_release_sys(p.flag_at_up, e);
                st_release_sys(p.flag_at_dn, e);
            }
        }
    }
}

// Stream-ordered host-side hooks (one thread).
// k_p2p_barrier: start of an lbm_step batch -- tell both neighbours that everything this rank enqueued before
// the batch has finished, and wait for the same from them (after it, nobody reads or writes ghost rows
// outside the protocol above).  k_p2p_wait_epoch: end of a batch -- the last state's halos have landed.
__global__ void k_p2p_barrier(unsigned long long *flags, unsigned long long *at_up, unsigned long long *at_dn, unsigned long long token) {
    __threadfence_system();
    st_release_sys(at_up, token);
    st_release_sys(at_dn, token);
    p2p_spin(flags, P2P_BATCH_FROM_DOWN, P2P_BATCH_FROM_UP, token, 120000000000ULL);
}
__global__ void k_p2p_wait_epoch(unsigned long long *flags, unsigned long long need) {
    p2p_spin(flags, P2P_EPOCH_FROM_DOWN, P2P_EPOCH_FROM_UP, need, 10000000000ULL);
}
__global__ void k_p2p_set_base(unsigned long long *flags, unsigned long long base) { flags[P2P_EPOCH_BASE] = base; }

// ------------------------------------------------------------------------------------------
// K1/K2: collide, optionally fused with the pull (stream + BCs) of the previous step
// ------------------------------------------------------------------------------------------
// MINB: minimum resident CTAs per SM asked of the register allocator; NPT: consecutive rows handled
// by one thread (all NPT*Q loads are issued before the first use -> more bytes in flight per thread,
// which is what the 4-byte populations need to cover the HBM latency).
// P2P: the launch takes part in the peer-memory halo protocol above (whole CTAs stay alive for it).
template <int CM, typename T, bool PULL, int MINB = 1, int NPT = 1, bool P2P = false>
__global__ void __launch_bounds__(256, MINB) k_step(const __grid_constant__ KParams<T> p, long long step) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    // P2P: CTAs that start inside the slab's edge rows (the first p2p_rows launched rows) take part in the halo protocol;
    // the others (interior rows of a merged launch) neither wait nor publish
    bool edge_cta = false;
    if constexpr (P2P) {
        edge_cta = (int)(blockIdx.y * blockDim.y) * NPT < p.p2p_rows;
        if (edge_cta) p2p_wait(p);
    } else if (x >= p.nx) return;
    for (int r0 = (!P2P || x < p.nx) ? (blockIdx.y * blockDim.y + threadIdx.y) * NPT : p.nrows; r0 < p.nrows; r0 += gridDim.y * blockDim.y * NPT) {
        T f[NPT][Q];
#pragma unroll
        for (int j = 0; j < NPT; ++j)
            if (r0 + j < p.nrows) load_node<T, PULL>(p, x, launched_row(p, r0 + j), f[j]);
        if constexpr (PULL) {
            // L2 prefetch of the row a later wave of CTAs will pull (option "prefetch" = distance in rows; 0 = off).  ONE warp
            // per CTA asks for the CTA's 128-byte lines of every population pf_rows rows ahead: Q * (CTA width * sizeof(T) /
            // 128) prefetches spread over 32 lanes -- a handful of instructions for one warp instead of Q for every warp
            // (profiles/r02/ncu_summary.md: the per-warp form added 14 % instructions; Float32 lost more than it gained).
            if (p.pf_rows > 0 && blockDim.y == 1 && threadIdx.x < 32) {
                const int yp = launched_row(p, r0) + p.pf_rows;
                if (yp < p.nyl) {
                    constexpr int EPL = 128 / (int)sizeof(T);                  // elements per line
                    const int lines = ((int)blockDim.x + EPL - 1) / EPL;       // lines of one population under this CTA
                    const unsigned base = (unsigned)yp * (unsigned)p.pitch + blockIdx.x * blockDim.x;
                    for (int k = threadIdx.x; k < Q * lines; k += 32) {
                        const int i = k / lines, l = k - i * lines;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.srcp[i] + base + (unsigned)(l * EPL)));
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NPT; ++j)
            if (r0 + j < p.nrows) {
                const int y = launched_row(p, r0 + j);
                T Fx, Fy;
                const bool forced = load_force(p, x, y, step, Fx, Fy);
                const unsigned n = (unsigned)y * (unsigned)p.pitch + (unsigned)x;
#ifdef LBM_TUNE
                // store-policy experiment (option "store"): 1 = st.cs (evict first), 2 = st.cg, 3 = st.wt
                collide_node<CM, T>(p, f[j], forced, Fx, Fy, [&](auto I, T v) {
                    T *q = p.dstp[decltype(I)::value] + n;
                    if (p.st_mode == 1) __stcs(q, v);
                    else if (p.st_mode == 2) __stcg(q, v);
                    else if (p.st_mode == 3) __stwt(q, v);
                    else *q = v;
                });
#else
                collide_node<CM, T>(p, f[j], forced, Fx, Fy,
                                    [&](auto I, T v) { p.dstp[decltype(I)::value][n] = v; });
#endif
                store_images(p, x, y);
                if constexpr (P2P) {
                    if (y < H || y >= p.nyl - H) p2p_store(p, x, y);
                }
            }
    }
    if constexpr (P2P) {
        if (edge_cta) {
            const int rows_per_cta = (int)blockDim.y * NPT;
            int edge_y = (p.p2p_rows + rows_per_cta - 1) / rows_per_cta;
            if (edge_y > (int)gridDim.y) edge_y = (int)gridDim.y;
            p2p_signal(p, (unsigned long long)gridDim.x * (unsigned long long)edge_y);
        }
    }
}

// K3: stream (+BC) only:  dst interior = pull(src)
template <typename T>
__global__ void __launch_bounds__(256) k_stream(const __grid_constant__ KParams<T> p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    for (int r = blockIdx.y * blockDim.y + threadIdx.y; r < p.nrows; r += gridDim.y * blockDim.y) {
        const int y = launched_row(p, r);
        T f[Q];
        load_node<T, true>(p, x, y, f);
        store_node<T, false>(p, x, y, f);
    }
}

// K4: standalone apply!(bcs, q, f_new = dst, f_old = aux): overwrite in place
template <typename T>
__global__ void __launch_bounds__(256) k_bcs(const __grid_constant__ KParams<T> p) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    for (int r = blockIdx.y * blockDim.y + threadIdx.y; r < p.nrows; r += gridDim.y * blockDim.y) {
        const int y = launched_row(p, r);
        const int yg = p.y0g + y;
        if (!near_wall(p, x, yg)) continue;
        const long long n = (long long)y * p.pitch + x;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            constexpr int o = L::opp(i);
            if constexpr (i != o) {
                const int b = resolve_bc(p, x + 1, yg + 1, L::cx(o), L::cy(o));
                if (b >= 0) p.dst[i * p.plane + n] = bounced<i>(p, b, p.aux[o * p.plane + n]);
            }
        });
    }
}

// ghost refresh: every ghost cell of dst := its periodic source (x always, y when wrap_y)
template <typename T>
__global__ void __launch_bounds__(256) k_ghosts(const __grid_constant__ KParams<T> p) {
    const int W = p.nx + 2 * H;
    const int xx = blockIdx.x * blockDim.x + threadIdx.x - H;  // [-H, nx+H)
    if (xx >= p.nx + H) return;
    const int ylo = p.wrap_y ? -H : 0, yhi = p.wrap_y ? p.nyl + H : p.nyl;
    for (int yy = ylo + blockIdx.y * blockDim.y + threadIdx.y; yy < yhi; yy += gridDim.y * blockDim.y) {
        const bool gx = (xx < 0 || xx >= p.nx), gy = (yy < 0 || yy >= p.nyl);
        if (!gx && !gy) continue;
        int xs = xx % p.nx; if (xs < 0) xs += p.nx;
        int ys = yy;
        if (gy) { ys = yy % p.nyl; if (ys < 0) ys += p.nyl; }
        T *d = p.dst + (long long)yy * p.pitch + xx;
        const T *s = p.dst + (long long)ys * p.pitch + xs;
        for (int i = 0; i < Q; ++i) d[i * p.plane] = s[i * p.plane];
    }
    (void)W;
}

// ------------------------------------------------------------------------------------------
// K5: diagnostics
// ------------------------------------------------------------------------------------------
template <typename T>
struct NodeDiag { double rho, ux, uy, axx, axy, ayy; };

// rho, u and a_bar_2 = sum(f[idx] * hermite(Val{2}, c_idx, q)) (moments.jl:27-28, 90-92; left folds) of one node's
// populations held in registers (Float32 storage: deviations, widened and shifted back first)
template <typename T>
__device__ __forceinline__ void fields_of(const T (&f)[Q], double &rho, double &ux, double &uy, double &axx, double &axy, double &ayy) {
    double g[Q];
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        if constexpr (Shifted<T>::value) g[i] = (double)f[i] + c_lat64.w[i];
        else g[i] = (double)f[i];
    });
    const LatConst<double> &c = c_lat64;
#if LBM_FAST
    {   // density = sum(f), velocity! = sum(f c) / rho (moments.jl:3-19) with ONE reciprocal instead of two divisions
        rho = g[0];
        static_for<1, Q>([&](auto I) { rho = rho + g[decltype(I)::value]; });
        double jx = 0, jy = 0;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if constexpr (L::cx(i) != 0) jx = jx + cmul<L::cx(i)>(g[i]);
            if constexpr (L::cy(i) != 0) jy = jy + cmul<L::cy(i)>(g[i]);
        });
        const double inv = 1.0 / rho;
        ux = jx * inv; uy = jy * inv;
        // a_bar_2 = sum(f H2(c)) with H2(c) = c c - delta / css (hermite_polynomials.jl:47-51): integer-coefficient sums
        // (adds for |c| = 1) and one multiply-add by 1 / css, instead of 3 Q multiply-adds by table entries
        double pxx = 0, pxy = 0, pyy = 0;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if constexpr (L::cx(i) * L::cx(i) != 0) pxx = pxx + cmul<L::cx(i) * L::cx(i)>(g[i]);
            if constexpr (L::cx(i) * L::cy(i) != 0) pxy = pxy + cmul<L::cx(i) * L::cy(i)>(g[i]);
            if constexpr (L::cy(i) * L::cy(i) != 0) pyy = pyy + cmul<L::cy(i) * L::cy(i)>(g[i]);
        });
        axx = pxx - c.cs_inv * rho; axy = pxy; ayy = pyy - c.cs_inv * rho;
        return;
    }
#else
    double drho_unused;
    rho_u<double>(g, rho, ux, uy, drho_unused);
#endif
    axx = g[0] * c.H2[0][0]; axy = g[0] * c.H2[0][1]; ayy = g[0] * c.H2[0][2];
    static_for<1, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        axx = axx + g[i] * c.H2[i][0];
        axy = axy + g[i] * c.H2[i][1];
        ayy = ayy + g[i] * c.H2[i][2];
    });
}

template <typename T, bool PULL>
__device__ __forceinline__ void node_fields(const KParams<T> &p, int x, int y, double &rho, double &ux, double &uy,
                                            double &axx, double &axy, double &ayy) {
    T f[Q];
    load_node<T, PULL>(p, x, y, f);
    fields_of<T>(f, rho, ux, uy, axx, axy, ayy);
}

template <typename T, bool PULL>
__global__ void __launch_bounds__(256) k_moments(const __grid_constant__ KParams<T> p, const MomentsOut m) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < p.nyl; y += gridDim.y * blockDim.y) {
        double rho, ux, uy, axx, axy, ayy;
        node_fields<T, PULL>(p, x, y, rho, ux, uy, axx, axy, ayy);
        const long long n = (long long)y * p.nx + x;
        if (m.rho) m.rho[n] = rho;
        if (m.ux) m.ux[n] = ux;
        if (m.uy) m.uy[n] = uy;
        if (m.p) {  // moments.jl:31-32
            if constexpr (L::UNIT_PRESSURE) m.p[n] = 1.0;
            else m.p[n] = ((axx + ayy) - rho * ((ux * ux + uy * uy) - 2)) / 2;
        }
        if (m.p_track || m.sxx || m.sxy || m.syy) {
            const double tau = m.tau_visc;
            const double exx = rho * (ux * ux + 0.0), exy = rho * (ux * uy), eyy = rho * (uy * uy + 0.0);
            const double den = 1 + 1 / (2 * tau);
            if (m.p_track) {  // track_hydrodynamic_errors.jl:153,173,181
                const double bxx = (axx + (1 / (2 * tau)) * exx) / den;
                const double byy = (ayy + (1 / (2 * tau)) * eyy) / den;
                const double Pxx = bxx - rho * (ux * ux - 1);
                const double Pyy = byy - rho * (uy * uy - 1);
                m.p_track[n] = (Pxx + Pyy) / 2;
            }
            // deviatoric_tensor(q, tau, f, rho, u)  moments.jl:81-96
            const double sxx = (axx - exx) / den, sxy = (axy - exy) / den, syy = (ayy - eyy) / den;
            const double tr = (sxx + syy) / 2;
            if (m.sxx) m.sxx[n] = sxx - tr;
            if (m.sxy) m.sxy[n] = sxy;
            if (m.syy) m.syy[n] = syy - tr;
        }
    }
}

// TakeSnapshots (take_snapshots.jl:12-29): f_stream of the current state as compact Float64 [Q][nyl][nx] (the host array's
// memory order) WITHOUT touching the ping-pong buffers: from post-collision populations the stream + BC step happens in
// the load, exactly as for the diagnostics.  The D2H copy of `out` then runs on a copy stream under the next steps.
template <typename T, bool PULL>
__global__ void __launch_bounds__(256) k_snapshot(const __grid_constant__ KParams<T> p, double *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    const long long N = (long long)p.nyl * p.nx;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < p.nyl; y += gridDim.y * blockDim.y) {
        T f[Q];
        load_node<T, PULL>(p, x, y, f);
        const long long n = (long long)y * p.nx + x;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if constexpr (Shifted<T>::value) out[i * N + n] = (double)f[i] + c_lat64.w[i];
            else out[i * N + n] = (double)f[i];
        });
    }
}

// ------------------------------------------------------------------------------------------
// Reducing diagnostics.  Geometry (launch_reduce / launch_errors): the CTA covers blockDim.x columns and `rows_per_cta`
// consecutive rows; a thread keeps its column and walks the CTA's rows (stride blockDim.y) accumulating in registers,
// then warp shuffles -> shared memory -> one partial per CTA and sum, folded by k_final_sum in a fixed order
// (deterministic: no atomics, the result depends only on the grid size).  rows_per_cta is a handful of rows, so a 4096^2
// grid runs thousands of CTAs and the loads of successive rows overlap (the loop has no aliasing stores).
// ------------------------------------------------------------------------------------------
template <int NS>
__device__ __forceinline__ void cta_partials(const double (&acc)[NS], double *__restrict__ partials) {
    __shared__ double sm[NS][8];
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) sm[k][warp] = v;
    }
    __syncthreads();
    if (tid < NS) {
        const int nw = (blockDim.x * blockDim.y + 31) >> 5;
        double v = 0;
        for (int w = 0; w < nw; ++w) v += sm[tid][w];
        partials[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * NS + tid] = v;
    }
}

// `v` with a numerically void data dependence on `loaded` (zero == 0 at run time, opaque to the compiler).  Used where a
// value is stored to the address another value was just loaded from: without the dependence the store is issued while the
// load of the same sector is still in flight, and B200 serialises that -- the velocity-change reduction ran at 1.4-1.9 ms
// per 4096^2 instead of 0.28 ms (tools/microbench/rmw_reduce.cu, profiles/r02/microbench_rmw_reduce.jsonl).
__device__ __forceinline__ double after_load(double v, double loaded, long long zero) {
    return __longlong_as_double(__double_as_longlong(v) | (__double_as_longlong(loaded) & zero));
}

// up to 4 per-node quantities (stop criteria, conserved sums); KIND is a compile-time lbm_reduce_kind
template <typename T, bool PULL, int KIND>
__global__ void __launch_bounds__(256, 3) k_reduce(const __grid_constant__ KParams<T> p, const ReduceArgs ra) {
    double acc[4] = {0, 0, 0, 0};
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * ra.rows_per_cta, y1 = min(p.nyl, y0 + ra.rows_per_cta);
    double *__restrict__ old = ra.u_old;
    const long long N = (long long)p.nyl * p.nx;
    if (x < p.nx) {
#pragma unroll 2
        for (int y = y0 + threadIdx.y; y < y1; y += blockDim.y) {
            double rho, ux, uy, axx, axy, ayy;
            node_fields<T, PULL>(p, x, y, rho, ux, uy, axx, axy, ayy);
            const long long n = (long long)y * p.nx + x;
            if constexpr (KIND == LBM_REDUCE_MEAN_UX) {
                acc[0] += ux; acc[1] += 1.0; acc[2] += (ux != ux) ? 1.0 : 0.0;
            } else if constexpr (KIND == LBM_REDUCE_VELOCITY_CHANGE) {
                const double ox = old[n], oy = old[N + n];
                acc[0] += ((ux - ox) * (ux - ox) + (uy - oy) * (uy - oy));
                acc[1] += ox * ox + oy * oy;
                old[n] = after_load(ux, ox, ra.zero); old[N + n] = after_load(uy, oy, ra.zero);
            } else if constexpr (KIND == LBM_REDUCE_DENSITY_CHANGE) {
                const double o = old[n];
                acc[0] += (rho - o) * (rho - o);
                if (x == p.nx - 1 && p.y0g + y == p.nyg - 1) acc[1] += rho;
                old[n] = after_load(rho, o, ra.zero);
            } else {
                acc[0] += rho; acc[1] += rho * (ux + uy); acc[2] += rho * (ux * ux + uy * uy);
            }
        }
    }
    cta_partials<4>(acc, ra.partials);
}

// Division by a launch-wide constant.  exact mode divides like the reference; fast mode multiplies by the reciprocal
// computed once per thread (<= 1 ulp per operation, a Float64 division is ~10 FP64-pipe instructions).
struct InvConst {
    double d, inv;
    __device__ __forceinline__ explicit InvConst(double d_) : d(d_), inv(1.0 / d_) {}
    __device__ __forceinline__ double div(double x) const {
#if LBM_FAST
        return x * inv;
#else
        return x / d;
#endif
    }
};

// One node's contribution to the 16 sums of TrackHydrodynamicErrors.next! (track_hydrodynamic_errors.jl:134-203):
// rho, u, p = tr(P)/D, sigma as in k_moments, scaled to dimensionless units, against the expected fields e[8] =
// rho, ux, uy, p, sxx, sxy, syx, syy.
template <bool EXPECTED_SQUARES>
__device__ __forceinline__ void error_terms(double rho, double ux, double uy, double axx, double axy, double ayy, double half_inv_tau,
                                            const InvConst &den, const InvConst &u_max, double fac, const double (&e)[8],
                                            double (&acc)[16]) {
    const double exx = rho * (ux * ux), exy = rho * (ux * uy), eyy = rho * (uy * uy);
    const double bxx = den.div(axx + half_inv_tau * exx), byy = den.div(ayy + half_inv_tau * eyy);
    const double pr = ((bxx - rho * (ux * ux - 1)) + (byy - rho * (uy * uy - 1))) / 2;
    double sxx = den.div(axx - exx), sxy = den.div(axy - exy), syy = den.div(ayy - eyy);
    const double tr = (sxx + syy) / 2;
    sxx = (sxx - tr) * fac; syy = (syy - tr) * fac; sxy = sxy * fac;
    const double vx = u_max.div(ux), vy = u_max.div(uy);
    acc[0] += (rho - e[0]) * (rho - e[0]);
    acc[1] += (vx - e[1]) * (vx - e[1]) + (vy - e[2]) * (vy - e[2]);
    acc[3] += (pr - e[3]) * (pr - e[3]);
    acc[5] += (e[4] - sxx) * (e[4] - sxx);
    acc[7] += (e[5] - sxy) * (e[5] - sxy);
    acc[9] += (e[7] - syy) * (e[7] - syy);
    acc[11] += (e[6] - sxy) * (e[6] - sxy);
    if constexpr (EXPECTED_SQUARES) {  // sums of the expected fields alone: separable, so the large-grid path leaves them to the host
        acc[2] += e[1] * e[1] + e[2] * e[2];
        acc[4] += e[3] * e[3];
        acc[6] += e[4] * e[4]; acc[8] += e[5] * e[5]; acc[10] += e[7] * e[7]; acc[12] += e[6] * e[6];
    }
    acc[13] += rho; acc[14] += rho * (vx + vy); acc[15] += rho * (vx * vx + vy * vy);
}

// One node's contribution to the sums of process! (processing_methods.jl:177-239, CompareWithAnalyticalSolution):
//   0 rho  1 (ux+uy) rho  2 kin + T  3 kin = (ux^2+uy^2) rho  4 T = p / rho  (p = pressure(q, f, rho, u), moments.jl:31-32)
//   5 e_rho  6 e_rho (e_ux+e_uy)  7 e_kin + e_T  8 e_kin = e_ux^2+e_uy^2  9 e_T = e_p / e_rho
//   10 (ux-e_ux)^2 + (uy-e_uy)^2  11 (p-e_p)^2      (u dimensionless = u / u_max; the host applies the cell area)
__device__ __forceinline__ void process_terms(double rho, double ux, double uy, double axx, double ayy, const InvConst &u_max,
                                              const double (&e)[8], double (&acc)[16]) {
    double pr;
    if constexpr (L::UNIT_PRESSURE) pr = 1.0;
    else pr = ((axx + ayy) - rho * ((ux * ux + uy * uy) - 2)) / 2;
    const double T = pr / rho;
    const double vx = u_max.div(ux), vy = u_max.div(uy);
    const double kin = (vx * vx + vy * vy) * rho;
    acc[0] += rho; acc[1] += (vx + vy) * rho; acc[2] += kin + T; acc[3] += kin; acc[4] += T;
    const double eT = e[3] / e[0], ekin = e[1] * e[1] + e[2] * e[2];
    acc[5] += e[0]; acc[6] += e[0] * (e[1] + e[2]); acc[7] += ekin + eT; acc[8] += ekin; acc[9] += eT;
    acc[10] += (vx - e[1]) * (vx - e[1]) + (vy - e[2]) * (vy - e[2]);
    acc[11] += (pr - e[3]) * (pr - e[3]);
}

// TrackHydrodynamicErrors.next! (track_hydrodynamic_errors.jl:114-203) / process! entirely on the device: per node
// rho, u, p = tr(P)/D, sigma as in k_moments, scaled to dimensionless units, compared with the problem's
// analytic fields given in separable form; 16 sums, deterministic two-stage reduction (geometry: see k_reduce).
//   0 (rho-e)^2  1 |u-e|^2  2 |e_u|^2  3 (p-e)^2  4 e_p^2  5 (e_sxx-sxx)^2  6 e_sxx^2  7 (e_sxy-sxy)^2  8 e_sxy^2
//   9 (e_syy-syy)^2  10 e_syy^2  11 (e_syx-syx)^2  12 e_syx^2  13 rho  14 rho (ux+uy)  15 rho (ux^2+uy^2)
// MODE 0: those sums, MODE 1: the sums of process!.  Terms whose coefficient is zero (most fields of most problems have
// one term or none) are skipped by a launch-uniform branch, so their tables are never read.
template <typename T, bool PULL, int MODE, int MINB>
__global__ void __launch_bounds__(256, MINB) k_errors(const __grid_constant__ KParams<T> p, const __grid_constant__ ErrorArgs ea) {
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0;
    const double half_inv_tau = 1 / (2 * ea.tau_visc);
    const InvConst den(1 + 1 / (2 * ea.tau_visc)), u_max(ea.u_max);
    const double fac = 1 / (ea.u_max * ea.u_max);
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * ea.rows_per_cta, y1 = min(p.nyl, y0 + ea.rows_per_cta);
    const unsigned mask = ea.mask;
    if (x < p.nx) {
        // slot s = 2 f + k of the tables: X_s = ea.tx[s], Y_s = ea.ty[s] (pointers in the parameter block, so that an
        // address is one multiply-add like the population loads)
#pragma unroll 1
        for (int y = y0 + threadIdx.y; y < y1; y += blockDim.y) {
            double rho, ux, uy, axx, axy, ayy;
            node_fields<T, PULL>(p, x, y, rho, ux, uy, axx, axy, ayy);
            double e[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                double v = ea.c0[f];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int s = 2 * f + k;
                    if (mask & (1u << s)) v = v + ea.a[f][k] * (__ldg(ea.tx[s] + (unsigned)x) * __ldg(ea.ty[s] + (unsigned)y));
                }
                e[f] = v;
            }
            if constexpr (MODE == 0) error_terms<false>(rho, ux, uy, axx, axy, ayy, half_inv_tau, den, u_max, fac, e, acc);
            else process_terms(rho, ux, uy, axx, ayy, u_max, e, acc);
        }
    }
    cta_partials<16>(acc, ea.partials);
}

// Second stage of the deterministic reductions: CTA k folds sum k.  Thread t adds partials t, t + 256, ... in order, then a
// fixed shuffle / shared-memory tree (same result on every run).
__global__ void __launch_bounds__(256) k_final_sum(const double *__restrict__ partials, int nblocks, int nsum, double *__restrict__ out) {
    const int k = blockIdx.x, t = threadIdx.x;
    double v = 0;
    for (int b = t; b < nblocks; b += 256) v += partials[(size_t)b * nsum + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __shared__ double sm[8];
    if ((t & 31) == 0) sm[t >> 5] = v;
    __syncthreads();
    if (t == 0) {
        double r = sm[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) r += sm[w];
        out[k] = r;
    }
}

// K7: device-side initialisation.  f_i = hermite_based_equilibrium!(q, rho, u, T)
// (velocity_distribution_function/hermite.jl:10-33) from per-node fields [ny][nx] (Float64):
//   f_i = w_i (rho + sum_{n=1..N} <a_eq^(n), H_n(c_i)> / (n! (1/css)^n)),  a_eq = equilibrium_coefficient(Val{n}, q, rho, u, T)
// including the (T - 1) terms and the Val{4} delta-index quirk (hermite.jl:69,71), evaluated over the full
// (non-symmetric) index set.  Writes rows [0, p.nyl) of dst (the caller offsets dst / sets nyl).
// one node: f_i (or f_i - w_i for Float32 storage) + extra_i -> dst
template <typename T, class Extra>
__device__ __forceinline__ void init_eq_node(const KParams<T> &p, int x, int y, double rho, double u0, double u1, double Tm1, Extra &&extra) {
    const LatConst<double> &c = c_lat64;
    const double u[2] = {u0, u1};
    const double cs = c.cs_inv;
    // a_eq^(n)[t], t = bit string of the indices (bit k = k-th index)
    double a2[4], a3[8], a4[16];
    auto d = [](int a, int b) { return a == b ? 1.0 : 0.0; };
    for (int t = 0; t < 4; ++t) {
        const int a = t & 1, b = (t >> 1) & 1;
        a2[t] = rho * (u[a] * u[b] + cs * Tm1 * d(a, b));
    }
    for (int t = 0; t < 8; ++t) {
        const int a = t & 1, b = (t >> 1) & 1, e = (t >> 2) & 1;
        a3[t] = rho * (u[a] * u[b] * u[e] + cs * Tm1 * (u[a] * d(b, e) + u[b] * d(a, e) + u[e] * d(a, b)));
    }
    for (int t = 0; t < 16; ++t) {
        const int a = t & 1, b = (t >> 1) & 1, e = (t >> 2) & 1, g = (t >> 3) & 1;
        a4[t] = rho * (u[a] * u[b] * u[e] * u[g]
                       + cs * Tm1 * (u[a] * u[b] * d(e, g) + u[a] * u[e] * d(b, g) + u[a] * u[g] * d(b, g)
                                     + u[b] * u[e] * d(a, g) + u[b] * u[g] * d(a, g) + u[e] * u[g] * d(a, b))
                       + cs * cs * Tm1 * Tm1 * (d(a, b) * d(e, g) + d(a, e) * d(b, g) + d(a, g) * d(b, e)));
    }
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        double s = (rho * u[0] * L::cx(i) + rho * u[1] * L::cy(i)) / cs;
        if constexpr (NH >= 2) {
            double dot = 0;
            for (int t = 0; t < 4; ++t) dot += a2[t] * c.H2[i][__popc(t)];
            s += dot / (2 * cs * cs);
        }
        if constexpr (NH >= 3) {
            double dot = 0;
            for (int t = 0; t < 8; ++t) dot += a3[t] * c.H3[i][__popc(t)];
            s += dot / (6 * cs * cs * cs);
        }
        if constexpr (NH >= 4) {
            double dot = 0;
            for (int t = 0; t < 16; ++t) dot += a4[t] * c.H4[i][__popc(t)];
            s += dot / (24 * cs * cs * cs * cs);
        }
        const long long m = i * p.plane + (long long)y * p.pitch + x;
        const double ex = extra(I);
        if constexpr (Shifted<T>::value) p.dst[m] = (T)(c.w[i] * ((rho - 1) + s) + ex);
        else p.dst[m] = (T)(c.w[i] * (rho + s) + ex);
    });
}

template <typename T>
__global__ void __launch_bounds__(256) k_init_eq(const __grid_constant__ KParams<T> p, const double *rho_, const double *ux_,
                                                 const double *uy_, const double *T_) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < p.nyl; y += gridDim.y * blockDim.y) {
        const long long n = (long long)y * p.nx + x;
        init_eq_node<T>(p, x, y, rho_[n], ux_[n], uy_[n], T_[n] - 1, [](auto) { return 0.0; });
    }
}

// K7b: initialize(strategy, q, problem) entirely on the device.  The problem's analytic fields arrive in separable form
// (sums of <= 2 products X(x) Y(y), as for the error norms): lattice density, velocity, pressure (T = p / rho,
// problems.jl:97-106, 121-128) and the velocity gradient for the strategies that add the off-equilibrium part
//   f_i += coef w_i [rho] dot(hermite(Val{2}, c_i, q), grad u + (grad u)')
// (analytical_offequilibrium.jl:10-87, analytical_velocity_stress.jl:5-31).  The host moves O(NX + NY) numbers.
template <typename T>
__global__ void __launch_bounds__(256) k_init_analytic(const __grid_constant__ KParams<T> p, const __grid_constant__ InitArgs ia) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    const int W = p.nx + p.nyl;
    const LatConst<double> &c = c_lat64;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < p.nyl; y += gridDim.y * blockDim.y) {
        double e[8];
#pragma unroll
        for (int f = 0; f < 8; ++f) {
            const double *t0 = ia.tab + (size_t)(2 * f) * W, *t1 = t0 + W;
            e[f] = ia.c0[f] + ia.a[f][0] * (__ldg(t0 + x) * __ldg(t0 + p.nx + y)) + ia.a[f][1] * (__ldg(t1 + x) * __ldg(t1 + p.nx + y));
        }
        const double rho = ia.unit_density ? 1.0 : e[0];
        const double Tm1 = ia.unit_temperature ? 0.0 : e[3] / e[0] - 1;
        const double sxx = e[4] + e[4], sxy = e[5] + e[6], syy = e[7] + e[7];
        const double kk = ia.offeq == 0 ? 0.0 : (ia.offeq == 2 ? ia.coef * rho : ia.coef);
        init_eq_node<T>(p, x, y, rho, e[1], e[2], Tm1, [&](auto I) {
            constexpr int i = decltype(I)::value;
            if (ia.offeq == 0) return 0.0;
            // dot over the full 2 x 2 index set: H11 S11 + H12 S12 + H21 S21 + H22 S22
            const double dot = ((c.H2[i][0] * sxx + c.H2[i][1] * sxy) + c.H2[i][1] * sxy) + c.H2[i][2] * syy;
            return (c.w[i] * kk) * dot;
        });
    }
}

// Float32 storage <-> host Float64 planes: g = (float)(f - w), f = (double)g + w
__global__ void __launch_bounds__(256) k_import32(const __grid_constant__ KParams<float> p, const double *stage, int i) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < p.nyl; y += gridDim.y * blockDim.y)
        p.dst[i * p.plane + (long long)y * p.pitch + x] = (float)(stage[(long long)y * p.nx + x] - c_lat64.w[i]);
}
__global__ void __launch_bounds__(256) k_export32(const __grid_constant__ KParams<float> p, double *stage, int i) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.nx) return;
    for (int y = blockIdx.y * blockDim.y + threadIdx.y; y < p.nyl; y += gridDim.y * blockDim.y)
        stage[(long long)y * p.nx + x] = (double)p.src[i * p.plane + (long long)y * p.pitch + x] + c_lat64.w[i];
}

// ------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------
static inline void pick_block(int nx, dim3 &block) {
    int bx = 32;
    while (bx < 256 && bx < nx) bx <<= 1;
    block = dim3(bx, 256 / bx, 1);
}

template <typename T>
static inline dim3 grid_for(const KParams<T> &p, const dim3 &block, int rows, int cols) {
    unsigned gy = (rows + block.y - 1) / block.y;
    if (gy > 65535u) gy = 65535u;
    if (gy == 0) gy = 1;
    return dim3((cols + block.x - 1) / block.x, gy, 1);
}

#include "packed_f32.cuh"
#include "tma.cuh"

// Launch configuration of the fused kernel per <collision model, dtype>: {MINB, NPT, CTA threads},
// chosen from tools/sweep.py runs on B200 (profiles/r01_sweep_summary.md).
struct StepCfg { int minb, npt, threads; };
template <int CM, typename T>
constexpr StepCfg step_cfg() {
    if (std::is_same<T, double>::value) {
        if (Q <= 13) return {4, 1, 256};
        if (Q <= 17) return {3, 1, 256};
        return {2, 1, 128};
    }
    if (Q <= 9) return {6, 1, 256};
    if (Q <= 13) return {4, 2, 128};
    if (Q <= 17) return {5, 1, 128};
    if (Q <= 25) return {4, 1, 128};
    return {3, 1, 128};
}

template <int CM, typename T, bool PULL, int MINB, int NPT, bool P2P = false>
static void launch_step_cfg(const KParams<T> &p, long long step, int threads, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    if (threads == 128 && block.x >= 128) block = dim3(128, 1, 1);
    dim3 grid = grid_for(p, block, (p.nrows + NPT - 1) / NPT, p.nx);
    k_step<CM, T, PULL, MINB, NPT, P2P><<<grid, block, 0, s>>>(p, step);
}

constexpr int X2_MINB = (Q <= 13) ? 4 : ((Q <= 17) ? 3 : 2);

// P2P = true: the fused (pull) launch that also pushes its boundary rows into the neighbours' ghost rows
// (same kernel choice and arithmetic as the plain launch, so results do not depend on the exchange path).
template <typename T, bool P2P>
static void launch_step_impl(int cm, bool pull, const KParams<T> &p, long long step, int variant, cudaStream_t s) {
    if (p.nrows <= 0) return;
    if constexpr (LBM_FAST && std::is_same<T, float>::value) {
        // Float32 fast mode: packed two-nodes-per-thread kernel (variant 99 forces the scalar one)
        // (measured faster for Q <= 13; the wide lattices run out of registers with 64-bit pairs -- variant 98 forces it)
        if (((Q <= 13 && variant != 99) || variant == 98) && (cm == LBM_SRT || cm == LBM_TRT || cm == LBM_MRT) && p.nx % 2 == 0 && p.nx >= 2) {
#define LBM_X2(CM_)                                                                 \
    {                                                                               \
        if constexpr (P2P) launch_step_x2<CM_, true, X2_MINB, true>(p, step, s);    \
        else if (pull) launch_step_x2<CM_, true, X2_MINB>(p, step, s);              \
        else launch_step_x2<CM_, false, X2_MINB>(p, step, s);                       \
    }
            if (cm == LBM_SRT) LBM_X2(LBM_SRT)
            else if (cm == LBM_TRT) LBM_X2(LBM_TRT)
            else LBM_X2(LBM_MRT)
#undef LBM_X2
            return;
        }
    }
    if constexpr (P2P) {
#define LBM_LAUNCH_P2P(CM)                                                                                        \
    {                                                                                                             \
        constexpr StepCfg c = step_cfg<CM, T>();                                                                  \
        launch_step_cfg<CM, T, true, c.minb, c.npt, true>(p, step, c.threads, s);                                 \
    }
        switch (cm) {
        case LBM_SRT: LBM_LAUNCH_P2P(LBM_SRT) break;
        case LBM_TRT: LBM_LAUNCH_P2P(LBM_TRT) break;
        case LBM_MRT: LBM_LAUNCH_P2P(LBM_MRT) break;
        default: LBM_LAUNCH_P2P(LBM_ITERATIVE_INIT) break;
        }
#undef LBM_LAUNCH_P2P
        return;
    }
#ifdef LBM_TUNE
    // variant = MINB + 10 * log2(NPT) + 100 * (128-thread CTAs); 0 = production configuration
    if (pull && variant > 0) {
        const int threads = variant >= 100 ? 128 : 256, nc = (variant % 100) / 10, mb = variant % 10;
#define LBM_T3(CM, MB, NP) if constexpr (NP * Q <= 52) if (cm == CM && mb == MB && nc == (NP == 1 ? 0 : (NP == 2 ? 1 : 2))) { launch_step_cfg<CM, T, true, MB, NP>(p, step, threads, s); return; }
#define LBM_T2(CM, MB) LBM_T3(CM, MB, 1) LBM_T3(CM, MB, 2) LBM_T3(CM, MB, 4)
#define LBM_T1(CM) LBM_T2(CM, 1) LBM_T2(CM, 2) LBM_T2(CM, 3) LBM_T2(CM, 4) LBM_T2(CM, 5) LBM_T2(CM, 6) LBM_T2(CM, 8)
        LBM_T1(LBM_SRT) LBM_T1(LBM_TRT) LBM_T1(LBM_MRT)
    }
#endif
#define LBM_LAUNCH(CM)                                                                                            \
    {                                                                                                             \
        constexpr StepCfg c = step_cfg<CM, T>();                                                                  \
        if (pull) launch_step_cfg<CM, T, true, c.minb, c.npt>(p, step, c.threads, s);                             \
        else launch_step_cfg<CM, T, false, c.minb, c.npt>(p, step, c.threads, s);                                 \
    }
    switch (cm) {
    case LBM_SRT: LBM_LAUNCH(LBM_SRT) break;
    case LBM_TRT: LBM_LAUNCH(LBM_TRT) break;
    case LBM_MRT: LBM_LAUNCH(LBM_MRT) break;
    default: LBM_LAUNCH(LBM_ITERATIVE_INIT) break;
    }
#undef LBM_LAUNCH
}

template <typename T>
static void launch_step(int cm, bool pull, const KParams<T> &p, long long step, int variant, cudaStream_t s) {
    launch_step_impl<T, false>(cm, pull, p, step, variant, s);
}
template <typename T>
static void launch_step_p2p(int cm, const KParams<T> &p, long long step, int variant, cudaStream_t s) {
    launch_step_impl<T, true>(cm, true, p, step, variant, s);
}
static void launch_p2p_barrier(unsigned long long *flags, unsigned long long *at_up, unsigned long long *at_dn,
                               unsigned long long token, cudaStream_t s) {
    k_p2p_barrier<<<1, 1, 0, s>>>(flags, at_up, at_dn, token);
}
static void launch_p2p_wait_epoch(unsigned long long *flags, unsigned long long need, cudaStream_t s) {
    k_p2p_wait_epoch<<<1, 1, 0, s>>>(flags, need);
}
static void launch_p2p_set_base(unsigned long long *flags, unsigned long long base, cudaStream_t s) {
    k_p2p_set_base<<<1, 1, 0, s>>>(flags, base);
}

template <typename T>
static void launch_stream(const KParams<T> &p, cudaStream_t s) {
    if (p.nrows <= 0) return;
    dim3 block; pick_block(p.nx, block);
    k_stream<T><<<grid_for(p, block, p.nrows, p.nx), block, 0, s>>>(p);
}
template <typename T>
static void launch_bcs(const KParams<T> &p, cudaStream_t s) {
    if (p.nrows <= 0 || p.nbc == 0) return;
    dim3 block; pick_block(p.nx, block);
    k_bcs<T><<<grid_for(p, block, p.nrows, p.nx), block, 0, s>>>(p);
}
template <typename T>
static void launch_ghosts(const KParams<T> &p, cudaStream_t s) {
    dim3 block; pick_block(p.nx + 2 * H, block);
    k_ghosts<T><<<grid_for(p, block, p.nyl + 2 * H, p.nx + 2 * H), block, 0, s>>>(p);
}
template <typename T>
static void launch_moments(bool pull, const KParams<T> &p, const MomentsOut &m, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    dim3 grid = grid_for(p, block, p.nyl, p.nx);
    if (pull) k_moments<T, true><<<grid, block, 0, s>>>(p, m);
    else k_moments<T, false><<<grid, block, 0, s>>>(p, m);
}
template <typename T>
static void launch_snapshot(bool pull, const KParams<T> &p, double *out, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    dim3 grid = grid_for(p, block, p.nyl, p.nx);
    if (pull) k_snapshot<T, true><<<grid, block, 0, s>>>(p, out);
    else k_snapshot<T, false><<<grid, block, 0, s>>>(p, out);
}
// Geometry of the reducing kernels: rows per CTA such that (1) at most `cap` partials are written, (2) a thread walks up to
// 16 rows when the grid is large (fewer, down to 1, when that would leave SMs without CTAs).
static inline dim3 reduce_grid(int nx, int nyl, const dim3 &block, int cap, int *rows_per_cta) {
    const long long gx = (nx + block.x - 1) / block.x, gy_nat = (nyl + block.y - 1) / block.y;
    long long rpt = (gx * gy_nat) / (148 * 8);          // rows per thread that still leave ~8 CTAs per SM
    rpt = rpt < 1 ? 1 : (rpt > 16 ? 16 : rpt);
    while (gx * ((gy_nat + rpt - 1) / rpt) > cap && rpt < gy_nat) ++rpt;  // (gx <= cap is checked by the callers)
    *rows_per_cta = (int)(rpt * block.y);
    return dim3((unsigned)gx, (unsigned)((gy_nat + rpt - 1) / rpt), 1);
}

template <typename T, int KIND>
static void launch_reduce_kind(bool pull, const KParams<T> &p, const ReduceArgs &ra, const dim3 &grid, const dim3 &block, cudaStream_t s) {
    if (pull) k_reduce<T, true, KIND><<<grid, block, 0, s>>>(p, ra);
    else k_reduce<T, false, KIND><<<grid, block, 0, s>>>(p, ra);
}
template <typename T>
static void launch_reduce(bool pull, const KParams<T> &p, const ReduceArgs &r, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    ReduceArgs ra = r;
    const dim3 grid = reduce_grid(p.nx, p.nyl, block, r.nblocks, &ra.rows_per_cta);
    ra.nblocks = grid.x * grid.y;
    ra.zero = 0;
    switch (ra.kind) {
    case LBM_REDUCE_MEAN_UX: launch_reduce_kind<T, LBM_REDUCE_MEAN_UX>(pull, p, ra, grid, block, s); break;
    case LBM_REDUCE_VELOCITY_CHANGE: launch_reduce_kind<T, LBM_REDUCE_VELOCITY_CHANGE>(pull, p, ra, grid, block, s); break;
    case LBM_REDUCE_DENSITY_CHANGE: launch_reduce_kind<T, LBM_REDUCE_DENSITY_CHANGE>(pull, p, ra, grid, block, s); break;
    default: launch_reduce_kind<T, LBM_REDUCE_CONSERVED>(pull, p, ra, grid, block, s); break;
    }
    k_final_sum<<<4, 256, 0, s>>>(ra.partials, ra.nblocks, 4, ra.out);
}

template <typename T>
static void launch_errors(bool pull, const KParams<T> &p, const ErrorArgs &e, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    ErrorArgs ea = e;
    const dim3 grid = reduce_grid(p.nx, p.nyl, block, e.nblocks, &ea.rows_per_cta);
    ea.nblocks = grid.x * grid.y;
    ea.mask = 0;
    const size_t W = (size_t)p.nx + (size_t)p.nyl;
    for (int f = 0; f < 8; ++f)
        for (int k = 0; k < 2; ++k) {
            const int sl = 2 * f + k;
            if (ea.a[f][k] != 0.0) ea.mask |= 1u << sl;
            ea.tx[sl] = ea.tab + (size_t)sl * W;
            ea.ty[sl] = ea.tab + (size_t)sl * W + p.nx;
        }
    // CTAs per SM (measured on B200, 4096^2 D2Q9 f64, profiles/r02/diag_D2Q9_minb*.json): the 10 accumulators + stresses of
    // TrackHydrodynamicErrors want 80 registers (3 CTAs: 0.49 ms, 4 CTAs with spills: 0.53 ms); the lighter process! sums
    // run best at 64 registers (4 CTAs: 0.28 ms, 3 CTAs: 0.30 ms); two rows in flight per thread at 128 registers (2 CTAs)
    // measured 0.54 ms.  LBM_ERRORS_MINB=3|4 overrides both (tuning hook).
    static const int forced = [] { const char *e = getenv("LBM_ERRORS_MINB"); return e ? atoi(e) : 0; }();
#define LBM_KE(MODE_, MINB_)                                                        \
    {                                                                               \
        if (pull) k_errors<T, true, MODE_, MINB_><<<grid, block, 0, s>>>(p, ea);    \
        else k_errors<T, false, MODE_, MINB_><<<grid, block, 0, s>>>(p, ea);        \
    }
    if (ea.mode == 0) { if (forced == 4) LBM_KE(0, 4) else LBM_KE(0, 3) }
    else { if (forced == 3) LBM_KE(1, 3) else LBM_KE(1, 4) }
#undef LBM_KE
    k_final_sum<<<16, 256, 0, s>>>(ea.partials, ea.nblocks, 16, ea.out);
}

static void launch_import32(const KParams<float> &p, const double *stage, int i, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    k_import32<<<grid_for(p, block, p.nyl, p.nx), block, 0, s>>>(p, stage, i);
}
static void launch_export32(const KParams<float> &p, double *stage, int i, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    k_export32<<<grid_for(p, block, p.nyl, p.nx), block, 0, s>>>(p, stage, i);
}

template <typename T>
static void launch_init_eq(const KParams<T> &p, const double *rho, const double *ux, const double *uy, const double *Tm, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    k_init_eq<T><<<grid_for(p, block, p.nyl, p.nx), block, 0, s>>>(p, rho, ux, uy, Tm);
}

template <typename T>
static void launch_init_analytic(const KParams<T> &p, const InitArgs &ia, cudaStream_t s) {
    dim3 block; pick_block(p.nx, block);
    k_init_analytic<T><<<grid_for(p, block, p.nyl, p.nx), block, 0, s>>>(p, ia);
}

#include "batch.cuh"
#include "persist.cuh"

static const Ops ops = {
    LBM_LATTICE, LBM_FAST,
    &launch_step<double>, &launch_step<float>,
    &launch_step_p2p<double>, &launch_step_p2p<float>,
    &launch_p2p_barrier, &launch_p2p_wait_epoch, &launch_p2p_set_base,
    &launch_stream<double>, &launch_stream<float>,
    &launch_bcs<double>, &launch_bcs<float>,
    &launch_ghosts<double>, &launch_ghosts<float>,
    &launch_moments<double>, &launch_moments<float>,
    &launch_reduce<double>, &launch_reduce<float>,
    &launch_errors<double>, &launch_errors<float>,
    &launch_import32, &launch_export32,
    &launch_init_eq<double>, &launch_init_eq<float>,
    &launch_init_analytic<double>, &launch_init_analytic<float>,
    &launch_snapshot<double>, &launch_snapshot<float>,
    &launch_step_tma<double>, &launch_step_tma<float>,
    &persist_grid<double>, &persist_grid<float>,
    &launch_persist<double>, &launch_persist<float>,
    &launch_batch<double>, &launch_batch<float>,
    &launch_batch_errors<double>, &launch_batch_errors<float>,
    &init_constants,
};

}  // namespace LBM_NS

#define LBM_GETTER2(a, b) get_ops_##a##_##b
#define LBM_GETTER(a, b) LBM_GETTER2(a, b)
const Ops *LBM_GETTER(LBM_LATTICE, LBM_FAST)() { return &LBM_NS::ops; }

}  // namespace lbm
