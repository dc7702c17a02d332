// Batched small problems: B independent problems of one shape (NX x NY nodes, same lattice / collision kind / boundary
// conditions), each with its own relaxation times and uniform force.  Included by kernels_inst.cu inside namespace
// lbm::LBM_NS.
//
// The reference's parameter studies are loops of `simulate(problem, q; ...)` over tiny grids -- 902 500 solves of a 3 x 5
// Poiseuille flow, up to 5001 steps each, in examples/notebooks/trt_magic_parameter.ipynb:30-103 (3 h on one thread).  One
// context per solve is pure launch latency on a GPU.  Here ONE launch advances every problem of the batch: each problem's
// populations live in shared memory for the whole run (ping-pong of post-collision populations, the pull of
// stream! + apply! folded into a precomputed source-index table), the stop criterion
// (stopping_criteria.jl:17-115, evaluated every `check_every` steps as TrackHydrodynamicErrors.next! does,
// track_hydrodynamic_errors.jl:52-59) runs on chip in the reference's summation order, and a problem that fires it
// writes its f_stream back to HBM and retires.  HBM traffic: one read and one write of the populations per RUN instead
// of per step; the kernel is bound by the FP64 pipe.
//
// Thread mapping: problems of at most 32 nodes share a warp (floor(32 / N) problems per warp, one lane per node, the
// per-step barrier is __syncwarp); larger problems get one CTA each (threads stride over the nodes, __syncthreads).
#pragma once

struct BatchLayout {
    size_t off_mw;      // moving-wall additive table T [Q][N] (only when a MovingWall is present)
    size_t off_prob;    // first per-problem block
    size_t prob_bytes;  // red double[2N] | crit double[2N] | A T[Q N] | B T[Q N] | ctrl int[4]
    size_t total;
};
template <typename T>
__host__ __device__ inline BatchLayout batch_layout(int N, int ppc, bool has_mw) {
    BatchLayout l;
    size_t o = ((size_t)Q * N * sizeof(unsigned short) + 15) & ~(size_t)15;
    l.off_mw = o;
    if (has_mw) o += ((size_t)Q * N * sizeof(T) + 15) & ~(size_t)15;
    l.off_prob = o;
    l.prob_bytes = (4 * (size_t)N * sizeof(double) + 2 * (size_t)Q * N * sizeof(T) + 16 + 15) & ~(size_t)15;
    l.total = o + (size_t)ppc * l.prob_bytes;
    return l;
}

__device__ __forceinline__ int wrap_index(int v, int n) {  // mod1 of stream_periodically_to (stream.jl:69-74), 0-based
    v %= n;
    return v < 0 ? v + n : v;
}

template <int CM, typename T, bool WARP, int MINB>
__global__ void __launch_bounds__(256, MINB) k_batch(const __grid_constant__ BatchParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int N = P.nx * P.nyg;
    const int tid = threadIdx.x;
    const int group = WARP ? 32 / N : 1;
    const int ppc = WARP ? (blockDim.x >> 5) * group : 1;
    const BatchLayout lay = batch_layout<T>(N, ppc, P.has_mw != 0);
    unsigned short *idx = reinterpret_cast<unsigned short *>(smem);
    T *mw = P.has_mw ? reinterpret_cast<T *>(smem + lay.off_mw) : nullptr;

    // Source of population i at node n after stream! + apply!: the periodic pull source, unless a boundary condition
    // overwrites it with the node's own opposite population (+ 2 a_1 for a moving wall) -- same resolution as load_node.
    for (int e = tid; e < Q * N; e += blockDim.x) {
        const int i = e / N, n = e - i * N, x = n % P.nx, y = n / P.nx;
        const int o = L::opp(i);
        int src = -1;
        T add = T(0);
        if (i != o && near_wall(P, x, y)) {
            const int b = resolve_bc(P, x + 1, y + 1, L::cx(o), L::cy(o));
            if (b >= 0) {
                src = o * N + n;
                if (P.bc[b].kind == LBM_BC_MOVING_WALL) {  // bounced<>: f_opp + 2 a_1 (moving_wall.jl:20-24,33)
                    const LatConst<T> &c = LC<T>();
                    const T ax = T(P.bc[b].ax), ay = T(P.bc[b].ay);
                    const T a1 = (c.w[i] * c.css) * (ax * T(L::cx(i)) + ay * T(L::cy(i)));
                    add = 2 * a1;
                }
            }
        }
        if (src < 0) src = i * N + wrap_index(y - L::cy(i), P.nyg) * P.nx + wrap_index(x - L::cx(i), P.nx);
        idx[e] = (unsigned short)src;
        if (mw) mw[e] = add;
    }

    // thread -> (problem slot in this CTA, node range)
    int ps, n0, n1, nstride;
    bool leader;
    if (WARP) {
        const int lane = tid & 31, slot = lane / N, node = lane - slot * N;
        const bool valid = slot < group;
        ps = (tid >> 5) * group + (valid ? slot : 0);
        n0 = node; n1 = valid ? node + 1 : 0; nstride = 32;
        leader = valid && node == 0;
    } else {
        ps = 0; n0 = tid; n1 = N; nstride = blockDim.x;
        leader = tid == 0;
    }
    const long long b = (long long)blockIdx.x * ppc + ps;
    unsigned char *pb = smem + lay.off_prob + (size_t)ps * lay.prob_bytes;
    double *red = reinterpret_cast<double *>(pb);
    double *crit = red + 2 * N;
    T *src = reinterpret_cast<T *>(crit + 2 * N);
    T *dst = src + Q * N;
    int *ctrl = reinterpret_cast<int *>(dst + Q * N);

    bool active = b < P.nb && (!WARP || n1 > n0);
    if (active && P.stopped[b]) active = false;
    BatchConsts<T> k;
    T *gf = nullptr;
    double *gcrit = nullptr;
    long long t = 0, t_end = 0;
    int to_check = 0;  // steps until t is the next multiple of check_every
    if (active) {
        k = reinterpret_cast<const BatchConsts<T> *>(P.consts)[b];
        gf = reinterpret_cast<T *>(P.f) + b * (long long)(Q * N);
        gcrit = P.crit ? P.crit + b * (long long)(2 * N) : nullptr;
        t = P.steps_done[b];
        t_end = t + P.nsteps;
        to_check = P.check_every - (int)(t % P.check_every);
        for (int n = n0; n < n1; n += nstride) {
            static_for<0, Q>([&](auto I) { constexpr int i = decltype(I)::value; src[i * N + n] = gf[i * N + n]; });
            if (P.stop_kind == LBM_BATCH_STOP_VELOCITY_CHANGE) { crit[2 * n] = gcrit[2 * n]; crit[2 * n + 1] = gcrit[2 * n + 1]; }
        }
        if (P.stop_kind == LBM_BATCH_STOP_MEAN_UX && leader) crit[0] = gcrit[0];
    }
    __syncthreads();  // tables + initial state

    auto bar = [] { if (WARP) __syncwarp(); else __syncthreads(); };
    bool first = true;  // src holds f_stream itself (no pull yet); afterwards post-collision populations
    // f[] := f_stream at node n = (stream! + apply!)(post-collision populations in src)
    auto pull = [&](int n, T(&f)[Q]) {
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if (first) {
                f[i] = src[i * N + n];
            } else {
                T v = src[idx[i * N + n]];
                if (mw) { const T m = mw[i * N + n]; if (m != T(0)) v = v + m; }
                f[i] = v;
            }
        });
    };

    for (;;) {
        if (WARP) { if (!__any_sync(0xffffffffu, active)) break; }
        else if (!active) break;
        const bool at_end = active && t == t_end;
        // next!(.., t): `mod(t, 100) == 0` -> should_stop! (track_hydrodynamic_errors.jl:52-59); t = steps taken so far
        // (a run that starts on a multiple does not repeat the check the previous run ended with)
        const bool check = active && P.stop_kind != LBM_BATCH_STOP_NONE && to_check == 0;
        if (to_check == 0) to_check = P.check_every;
        const bool any_check = WARP ? __any_sync(0xffffffffu, check) : check;
        bool stop_now = false;
        if (any_check) {
            if (check) {
                for (int n = n0; n < n1; n += nstride) {
                    T f[Q];
                    pull(n, f);
                    double g[Q], rho, ux, uy, dr;
                    static_for<0, Q>([&](auto I) {
                        constexpr int i = decltype(I)::value;
                        if constexpr (Shifted<T>::value) g[i] = (double)f[i] + c_lat64.w[i];
                        else g[i] = (double)f[i];
                    });
                    rho_u<double>(g, rho, ux, uy, dr);
                    if (P.stop_kind == LBM_BATCH_STOP_VELOCITY_CHANGE) {  // stopping_criteria.jl:94-99
                        const double ox = crit[2 * n], oy = crit[2 * n + 1];
                        red[n] = ((ux - ox) * (ux - ox) + (uy - oy) * (uy - oy));
                        red[N + n] = ox * ox + oy * oy;
                        crit[2 * n] = ux; crit[2 * n + 1] = uy;
                    } else {
                        red[n] = ux;  // stopping_criteria.jl:37
                    }
                }
            }
            bar();
            if (check && leader) {
                // the reference's loop order: for x_idx in 1:Nx, y_idx in 1:Ny (left folds)
                int stop = 0;
                if (P.stop_kind == LBM_BATCH_STOP_VELOCITY_CHANGE) {
                    double err = 0.0, old_norm = 0.0;
                    for (int x = 0; x < P.nx; ++x)
                        for (int y = 0; y < P.nyg; ++y) { err += red[y * P.nx + x]; old_norm += red[N + y * P.nx + x]; }
                    const double converged = sqrt(err) / old_norm;  // the denominator is not sqrt'ed (:101)
                    stop = (converged < P.tol) || (converged != converged);
                } else {
                    double u_mean = 0.0;
                    for (int x = 0; x < P.nx; ++x)
                        for (int y = 0; y < P.nyg; ++y) u_mean += red[y * P.nx + x];
                    u_mean /= (double)N;
                    const double converged = fabs(u_mean / crit[0] - 1);
                    stop = (converged < P.tol) || (u_mean != u_mean);
                    if (!stop) crit[0] = u_mean;  // :52
                }
                ctrl[0] = stop;
            }
            bar();
            if (check) stop_now = ctrl[0] != 0;
        }
        if (active && (stop_now || at_end)) {  // retire: f_stream back to HBM
            for (int n = n0; n < n1; n += nstride) {
                T f[Q];
                pull(n, f);
                static_for<0, Q>([&](auto I) { constexpr int i = decltype(I)::value; gf[i * N + n] = f[i]; });
                if (P.stop_kind == LBM_BATCH_STOP_VELOCITY_CHANGE) { gcrit[2 * n] = crit[2 * n]; gcrit[2 * n + 1] = crit[2 * n + 1]; }
            }
            if (leader) {
                if (P.stop_kind == LBM_BATCH_STOP_MEAN_UX) gcrit[0] = crit[0];
                P.steps_done[b] = t;
                if (stop_now) P.stopped[b] = 1;
            }
            active = false;
        }
        if (active) {  // collide!(time = t dt) of f_stream -> post-collision populations in dst
            for (int n = n0; n < n1; n += nstride) {
                T f[Q];
                pull(n, f);
                collide_node<CM, T>(k, f, k.forced != 0, k.fx, k.fy, [&](auto I, T v) { dst[decltype(I)::value * N + n] = v; });
            }
            ++t;
            --to_check;
        }
        bar();
        T *tmp = src; src = dst; dst = tmp;
        first = false;
    }
}

// TrackHydrodynamicErrors.next! (track_hydrodynamic_errors.jl:114-203) for every problem of a batch: one warp per problem,
// lanes stride over the nodes in the reference's loop order (x outer, y inner) and the 16 sums are folded over the lanes in
// order -- for problems of at most 32 nodes that is exactly the reference's left fold.
template <typename T>
__global__ void __launch_bounds__(256) k_batch_errors(const __grid_constant__ BatchErrorArgs ea) {
    __shared__ double sm[8][16][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + w;
    if (b >= ea.nb) return;  // whole warps leave; only __syncwarp below
    const int N = ea.nx * ea.ny, W = ea.nx + ea.ny;
    const T *gf = reinterpret_cast<const T *>(ea.f) + b * (long long)(Q * N);
    const double tau = ea.tau_visc[b], um = ea.u_max[b];
    const double *cf = ea.coef + b * 24;
    const double half_inv_tau = 1 / (2 * tau), fac = 1 / (um * um);
    const InvConst den(1 + 1 / (2 * tau)), u_max(um);
    double acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0;
    for (int kk = lane; kk < N; kk += 32) {
        const int x = kk / ea.ny, y = kk - x * ea.ny, n = y * ea.nx + x;
        T f[Q];
        static_for<0, Q>([&](auto I) { constexpr int i = decltype(I)::value; f[i] = gf[i * N + n]; });
        double rho, ux, uy, axx, axy, ayy;
        fields_of<T>(f, rho, ux, uy, axx, axy, ayy);
        double e[8];
#pragma unroll
        for (int fi = 0; fi < 8; ++fi) {
            const double *t0 = ea.tab + (size_t)(2 * fi) * W, *t1 = t0 + W;
            e[fi] = cf[3 * fi] + cf[3 * fi + 1] * (__ldg(t0 + x) * __ldg(t0 + ea.nx + y)) + cf[3 * fi + 2] * (__ldg(t1 + x) * __ldg(t1 + ea.nx + y));
        }
        error_terms<true>(rho, ux, uy, axx, axy, ayy, half_inv_tau, den, u_max, fac, e, acc);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) sm[w][j][lane] = acc[j];
    __syncwarp();
    if (lane < 16) {
        double v = 0;
        for (int l = 0; l < 32; ++l) v += sm[w][lane][l];
        ea.out[b * 16 + lane] = v;
    }
}

constexpr int BATCH_MINB = Q <= 13 ? 2 : 1;
static const size_t BATCH_MAX_SMEM = 227 * 1024;

template <int CM, typename T, bool WARP>
static int launch_batch_cm(const BatchParams &p, int threads, int grid, size_t smem, cudaStream_t s) {
    auto kern = k_batch<CM, T, WARP, BATCH_MINB>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -2;
    kern<<<grid, threads, smem, s>>>(p);
    return 0;
}

template <typename T>
static int launch_batch(int cm, const BatchParams &p, cudaStream_t s) {
    const int N = p.nx * p.nyg;
    if ((long long)Q * N > 65535 || p.nb <= 0) return -1;
    const bool warp = N <= 32;
    int threads = warp ? 256 : ((N + 31) / 32 * 32 < 256 ? (N + 31) / 32 * 32 : 256);
    int ppc = warp ? (threads / 32) * (32 / N) : 1;
    BatchLayout lay = batch_layout<T>(N, ppc, p.has_mw != 0);
    while (warp && lay.total > BATCH_MAX_SMEM && threads > 32) {
        threads >>= 1;
        ppc = (threads / 32) * (32 / N);
        lay = batch_layout<T>(N, ppc, p.has_mw != 0);
    }
    if (lay.total > BATCH_MAX_SMEM) return -1;
    const int grid = (int)((p.nb + ppc - 1) / ppc);
#define LBM_BATCH(CM) return warp ? launch_batch_cm<CM, T, true>(p, threads, grid, lay.total, s) : launch_batch_cm<CM, T, false>(p, threads, grid, lay.total, s);
    switch (cm) {
    case LBM_SRT: LBM_BATCH(LBM_SRT)
    case LBM_TRT: LBM_BATCH(LBM_TRT)
    case LBM_MRT: LBM_BATCH(LBM_MRT)
    default: return -1;
    }
#undef LBM_BATCH
}

template <typename T>
static void launch_batch_errors(const BatchErrorArgs &e, cudaStream_t s) {
    const int warps = 8;
    k_batch_errors<T><<<(e.nb + warps - 1) / warps, warps * 32, 0, s>>>(e);
}
