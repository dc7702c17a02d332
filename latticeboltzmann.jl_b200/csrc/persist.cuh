// Persistent multi-step variant of the fused collide+stream kernel for slabs that are launch-bound (a 1024^2 D2Q9 step
// is 23 us of HBM time; two kernel boundaries per step -- boundary-row launch + interior launch, fork/join events -- cost
// 4..8 us more, which is what held C3 strong scaling at 79.5 % on 8 GPUs in round 1).  Included by kernels_inst.cu inside
// namespace lbm::LBM_NS.
//
// ONE cooperative launch takes `nsteps` steps.  The CTAs are co-resident (grid = SMs x occupancy); CTA k owns the nodes
// [k npc, (k+1) npc) of the slab in row-major order (perfectly balanced, coalesced row segments) for the whole launch.
// There is no grid-wide barrier: before step s a CTA only waits until the CTAs that own rows within the stencil's reach
// (+-H rows, circular when the slab is the whole periodic domain) have finished step s-1 -- per-CTA step counters,
// st.release.gpu / ld.acquire.gpu.  That one condition covers both hazards of the ping-pong buffers: their rows of
// step s-1 are complete (RAW), and they no longer read the buffer this CTA is about to overwrite (WAR).
// Across GPUs the slab edges use the SAME protocol with the peer-memory epoch flags of the P2P launch (kernels_inst.cu):
// the CTAs that own the H boundary rows of an edge do those rows first, store them locally and into the neighbour's
// ghost rows over NVLink, and publish the epoch mid-step -- the neighbour needs it only at the start of its next step,
// so the NVLink latency is off the critical path.
// Populations are read with ld.global.cg (L2): lines cached in L1 two steps earlier would be stale.
#pragma once

__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// warp 0: wait until done[j] >= need for every j in [lo, hi] (lanes poll different CTAs).  Bounded; a give-up (or one
// seen elsewhere) sets *error and lets every later wait fall through, so the launch ends instead of hanging the GPU.
__device__ __noinline__ void persist_wait(const PersistArgs &a, int lo, int hi, unsigned long long need) {
    for (int j = lo + (int)(threadIdx.x & 31); j <= hi; j += 32) {
        if (ld_acquire_gpu(a.done + j) >= need) continue;
        const unsigned long long t0 = globaltimer_ns();
        while (ld_acquire_gpu(a.done + j) < need) {
            if (*(volatile unsigned long long *)a.error != 0ULL) break;
            if (globaltimer_ns() - t0 > 10000000000ULL) { atomicMax(a.error, need + 1); break; }
            __nanosleep(32);
        }
    }
}

__device__ __noinline__ void persist_wait_peer(const PersistArgs &a, unsigned long long *flags, int slot, unsigned long long need) {
    if (ld_acquire_sys(flags + slot) >= need) return;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(flags + slot) < need) {
        if (*(volatile unsigned long long *)a.error != 0ULL) break;
        if (globaltimer_ns() - t0 > 10000000000ULL) { atomicMax(a.error, need + 1); atomicMax(flags + P2P_TIMEOUT, need ? need : 1ULL); break; }
        __nanosleep(64);
    }
}

// nodes [lo, hi) of the slab (row-major) for one step: pull + BC + collide + store (+ ghost images, + peer rows)
template <int CM, typename T, bool P2P, int THREADS>
__device__ __forceinline__ void persist_range(const KParams<T> &p, unsigned lo, unsigned hi, long long step) {
    for (unsigned idx = lo + threadIdx.x; idx < hi; idx += THREADS) {
        const int y = (int)(idx / (unsigned)p.nx), x = (int)(idx - (unsigned)y * (unsigned)p.nx);
        T f[Q];
        load_node<T, true, true>(p, x, y, f);
        T Fx, Fy;
        const bool forced = load_force(p, x, y, step, Fx, Fy);
        const unsigned n = (unsigned)y * (unsigned)p.pitch + (unsigned)x;
        collide_node<CM, T>(p, f, forced, Fx, Fy, [&](auto I, T v) { p.dstp[decltype(I)::value][n] = v; });
        store_images(p, x, y);
        if constexpr (P2P) {
            if (y < H || y >= p.nyl - H) p2p_store(p, x, y);
        }
    }
}

template <int CM, typename T, bool P2P, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB * (256 / THREADS)) k_persist(const __grid_constant__ KParams<T> pa, const __grid_constant__ KParams<T> pb,
                                                           const __grid_constant__ PersistArgs a) {
    const int k = blockIdx.x, tid = threadIdx.x;
    const int nx = pa.nx, nyl = pa.nyl;
    const long long N = (long long)nx * nyl;
    long long lo64 = (long long)k * a.npc, hi64 = lo64 + a.npc;
    if (lo64 > N) lo64 = N;
    if (hi64 > N) hi64 = N;
    const unsigned n_lo = (unsigned)lo64, n_hi = (unsigned)hi64;
    const bool owns = n_lo < n_hi;
    // boundary-row sub-ranges (peer exchange) and the rest
    const unsigned bot_end = P2P ? (unsigned)((long long)H * nx) : 0u;          // rows [0, H)
    const unsigned top_beg = P2P ? (unsigned)((long long)(nyl - H) * nx) : (unsigned)N;  // rows [nyl - H, nyl)
    const bool has_bot = P2P && owns && n_lo < bot_end;
    const bool has_top = P2P && owns && n_hi > top_beg;
    const unsigned mid_lo = n_lo > bot_end ? n_lo : bot_end, mid_hi = n_hi < top_beg ? n_hi : top_beg;

    // CTAs this one synchronises with: owners of the rows within +-H of its own (at most three index ranges)
    int r_lo[5], r_hi[5], nr = 0;
    if (owns) {
        const int y_lo = (int)(n_lo / (unsigned)nx), y_hi = (int)((n_hi - 1) / (unsigned)nx);
        auto owners = [&](int ya, int yb) {  // rows [ya, yb] -> CTA index range
            r_lo[nr] = (int)(((long long)ya * nx) / a.npc);
            r_hi[nr] = (int)((((long long)yb + 1) * nx - 1) / a.npc);
            ++nr;
        };
        if (pa.wrap_y && nyl <= 2 * H + 1) {
            r_lo[0] = 0; r_hi[0] = a.nctas - 1; nr = 1;
        } else {
            owners(y_lo - H < 0 ? 0 : y_lo - H, y_hi + H > nyl - 1 ? nyl - 1 : y_hi + H);
            if (pa.wrap_y && y_lo - H < 0) owners(nyl + (y_lo - H), nyl - 1);
            if (pa.wrap_y && y_hi + H > nyl - 1) owners(0, y_hi + H - nyl);
        }
        // the CTAs sharing an edge's boundary rows count their arrivals (edge_count): keep them within one step of each other
        if (has_bot && a.n_bot > 1) { r_lo[nr] = 0; r_hi[nr] = a.n_bot - 1; ++nr; }
        if (has_top && a.n_top > 1) { r_lo[nr] = a.nctas - a.n_top; r_hi[nr] = a.nctas - 1; ++nr; }
    }

    for (int s = 0; s < a.nsteps; ++s) {
        const KParams<T> &p = (s & 1) ? pb : pa;
        const long long step = a.step0 + s;
        if (owns) {
            if (tid < 32) {
                for (int r = 0; r < nr; ++r) persist_wait(a, r_lo[r], r_hi[r], (unsigned long long)s);
                if constexpr (P2P) {
                    // the neighbour's boundary rows of the previous state are in my ghost rows, and it no longer reads the
                    // ghost rows (of the buffer with the role of my dst) that this step overwrites
                    if (tid == 0 && has_bot) persist_wait_peer(a, p.flags, P2P_EPOCH_FROM_DOWN, a.epoch0 + (unsigned long long)s);
                    if (tid == 0 && has_top) persist_wait_peer(a, p.flags, P2P_EPOCH_FROM_UP, a.epoch0 + (unsigned long long)s);
                }
            }
            __syncthreads();
            if constexpr (P2P) {
                if (has_bot || has_top) {
                    if (has_bot) persist_range<CM, T, true, THREADS>(p, n_lo, n_hi < bot_end ? n_hi : bot_end, step);
                    if (has_top) persist_range<CM, T, true, THREADS>(p, n_lo > top_beg ? n_lo : top_beg, n_hi, step);
                    __threadfence_system();
                    __syncthreads();
                    if (tid == 0) {
                        const unsigned long long e = a.epoch0 + (unsigned long long)s + 1;
                        const bool ok = *(volatile unsigned long long *)a.error == 0ULL;
                        if (has_bot && atomicAdd(a.edge_count + 0, 1ULL) == (unsigned long long)a.n_bot - 1) {
                            a.edge_count[0] = 0;
                            __threadfence_system();
                            if (ok) st_release_sys(p.flag_at_dn, e);
                        }
                        if (has_top && atomicAdd(a.edge_count + 1, 1ULL) == (unsigned long long)a.n_top - 1) {
                            a.edge_count[1] = 0;
                            __threadfence_system();
                            if (ok) st_release_sys(p.flag_at_up, e);
                        }
                    }
                }
            }
            if (mid_lo < mid_hi) persist_range<CM, T, false, THREADS>(p, mid_lo, mid_hi, step);
            __threadfence();
            __syncthreads();
            if (tid == 0) st_release_gpu(a.done + k, (unsigned long long)s + 1);
        }
    }
}

template <int CM, typename T, bool P2P>
struct PersistKernel {
    static constexpr StepCfg cfg = step_cfg<CM, T>();
    // MINB is stated per 256 threads, as for k_step (__launch_bounds__(256, MINB) whatever the CTA size)
    static constexpr int THREADS = cfg.threads, MINB = cfg.minb;
    static auto get() {
        if constexpr (THREADS == 128) return k_persist<CM, T, P2P, 128, MINB>;
        else return k_persist<CM, T, P2P, 256, MINB>;
    }
};

template <int CM, typename T, bool P2P>
static void persist_grid_cm(int *ctas, int *threads) {
    auto kern = PersistKernel<CM, T, P2P>::get();
    int dev = 0, sms = 0, coop = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    *threads = PersistKernel<CM, T, P2P>::THREADS;
    if (!coop || cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, *threads, 0) != cudaSuccess) { cudaGetLastError(); *ctas = 0; return; }
    *ctas = sms * per_sm;
}

template <typename T>
static void persist_grid(int cm, bool p2p, int *ctas, int *threads) {
#define LBM_PG(CM) { if (p2p) persist_grid_cm<CM, T, true>(ctas, threads); else persist_grid_cm<CM, T, false>(ctas, threads); return; }
    switch (cm) {
    case LBM_SRT: LBM_PG(LBM_SRT)
    case LBM_TRT: LBM_PG(LBM_TRT)
    case LBM_MRT: LBM_PG(LBM_MRT)
    default: *ctas = 0; *threads = 0; return;
    }
#undef LBM_PG
}

template <int CM, typename T, bool P2P>
static int launch_persist_cm(const KParams<T> &pa, const KParams<T> &pb, const PersistArgs &a, int ctas, int threads, cudaStream_t s) {
    auto kern = PersistKernel<CM, T, P2P>::get();
    void *args[3] = {(void *)&pa, (void *)&pb, (void *)&a};
    return cudaLaunchCooperativeKernel((const void *)kern, dim3(ctas), dim3(threads), args, 0, s) == cudaSuccess ? 0 : -1;
}

template <typename T>
static int launch_persist(int cm, bool p2p, const KParams<T> &pa, const KParams<T> &pb, const PersistArgs &a, int ctas, int threads,
                          cudaStream_t s) {
#define LBM_PL(CM) return p2p ? launch_persist_cm<CM, T, true>(pa, pb, a, ctas, threads, s) : launch_persist_cm<CM, T, false>(pa, pb, a, ctas, threads, s);
    switch (cm) {
    case LBM_SRT: LBM_PL(LBM_SRT)
    case LBM_TRT: LBM_PL(LBM_TRT)
    case LBM_MRT: LBM_PL(LBM_MRT)
    default: return -1;
    }
#undef LBM_PL
}
