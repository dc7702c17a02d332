// TMA-staged variant of the fused pull + BC + collide kernel (included by kernels_inst.cu inside namespace lbm::LBM_NS).
//
// The register kernel (k_step) lets every warp issue its own Q loads and then waits for them: with Q = 17..37 live
// populations per thread only a few warps fit on an SM, and on the Float32 wide lattices neither DRAM nor the issue slots
// saturate (profiles/r02/ncu_summary.md: 66-69 % DRAM, 1.1-1.9 eligible warps per scheduler).  Here the loads are taken
// off the warps: one thread per CTA issues, per tile of TW x 1 nodes, Q bulk tensor copies (cp.async.bulk.tensor.3d) into a
// STAGES-deep shared-memory ring paced by mbarriers, and the CTA is persistent over tiles -- memory-level parallelism
// comes from the ring, not from occupancy.  The y part of the pull shift is the box coordinate (row y - c_y of population
// i); the x part cannot be: the global address of a box must be 16-byte aligned (a box starting at x0 - c_x raises
// "illegal instruction", measured), so every box is TW + 8 elements wide, starts TMA_PAD = 4 elements left of the tile
// (16 / 32 bytes: aligned because the ghost frame is 128 bytes wide) and the shift happens in the shared-memory read:
// f_i = row_i[tid + TMA_PAD - c_x].  Consumers read conflict-free (consecutive threads, consecutive words), release the
// stage, then do the boundary fix-up, the collision and the coalesced stores exactly as k_step does (same device
// functions, same arithmetic, bit-identical results).
//
// MEASURED (B200, round 2, profiles/r02/tma_sweep_v2.jsonl; fraction of the HBM peak, register kernel -> best TMA
// configuration): D2Q9 f64 0.99 -> 0.78, f32 0.92 -> 0.82; D2Q17 f64 0.95 -> 0.72, f32 0.85 -> 0.75; D2Q37 f64 0.94 -> 0.94,
// f32 0.87 -> 0.67.  It is slower or equal everywhere: with one-row boxes of 544 / 1088 bytes the copy engine handles
// Q small requests per 128 nodes, every tile costs a CTA-wide barrier plus Q single-thread issues, and the rings of the
// wide lattices (D2Q37 f32: 23.7 KB per stage) push the occupancy below what the arithmetic phase needs -- the register
// kernel's plain coalesced loads from the ghost-framed layout were already the better fit.  The path is therefore OFF
// unless lbm_set_option("tma", 1) asks for it; tests keep it bit-identical to the register kernel.
#pragma once
#include <cuda.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LBM_TMA_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LBM_TMA_DONE_%=;\n"
        "bra LBM_TMA_WAIT_%=;\n"
        "LBM_TMA_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned long long *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"((unsigned long long)tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

constexpr int TMA_PAD = 4;                      // >= max |c_x| (3), and 16 bytes for Float32
constexpr int TMA_BOXW = 128 + 2 * TMA_PAD;     // elements per box (the tensor map's box: lbm_b200.cu make_tensor_maps)
template <typename T>
constexpr int tma_row_stride() { return (int)((TMA_BOXW * sizeof(T) + 127) / 128 * 128 / sizeof(T)); }  // 128-byte aligned rows

// tm: the source buffer as a 3-D tensor (pitch, rows incl. ghosts, Q) of T from the start of the allocation; (gx, gy):
// ghost columns / rows in front of node (0, 0); ntx: tiles per row.
template <int CM, typename T, int TW, int STAGES, int MINB>
__global__ void __launch_bounds__(TW, MINB) k_step_tma(const __grid_constant__ KParams<T> p, const __grid_constant__ CUtensorMap tm,
                                                       long long step, int gx, int gy, int ntx) {
    extern __shared__ __align__(128) unsigned char tma_smem[];
    static_assert(TW == 128, "the tensor map's box is 128 + 8 elements wide");
    constexpr int RS = tma_row_stride<T>();
    T *tiles = reinterpret_cast<T *>(tma_smem);  // [STAGES][Q][RS]
    __shared__ __align__(8) unsigned long long full[STAGES];
    const int tid = threadIdx.x;
    const long long ntiles = (long long)ntx * p.nrows;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](long long t, int s) {
        const int r = (int)(t / ntx), tx = (int)(t - (long long)r * ntx);
        const int y = launched_row(p, r);
        mbar_expect_tx(&full[s], (unsigned)(Q * TMA_BOXW * sizeof(T)));
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            tma_load_3d(tiles + ((size_t)s * Q + i) * RS, &tm, tx * TW + gx - TMA_PAD, y + gy - L::cy(i), i, &full[s]);
        });
    };
    if (tid == 0)
        for (int k = 0; k < STAGES; ++k) {
            const long long t = blockIdx.x + (long long)k * gridDim.x;
            if (t < ntiles) issue(t, k);
        }
    int s = 0;
    unsigned phase = 0;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int r = (int)(t / ntx), tx = (int)(t - (long long)r * ntx);
        const int y = launched_row(p, r), x = tx * TW + tid;
        mbar_wait(&full[s], phase);
        T f[Q];
        const T *mine = tiles + (size_t)s * Q * RS + tid + TMA_PAD;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            f[i] = mine[i * RS - L::cx(i)];
        });
        __syncthreads();  // every thread holds its populations: the stage can be refilled
        if (tid == 0) {
            const long long tn = t + (long long)STAGES * gridDim.x;
            if (tn < ntiles) issue(tn, s);
        }
        if (x < p.nx) {
            const unsigned n = (unsigned)y * (unsigned)p.pitch + (unsigned)x;
            const int yg = p.y0g + y;
            if (near_wall(p, x, yg)) {
                static_for<0, Q>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr int o = L::opp(i);
                    if constexpr (i != o) {
                        const int b = resolve_bc(p, x + 1, yg + 1, L::cx(o), L::cy(o));
                        if (b >= 0) f[i] = bounced<i>(p, b, __ldg(p.srcn[o] + n));
                    }
                });
            }
            T Fx, Fy;
            const bool forced = load_force(p, x, y, step, Fx, Fy);
            collide_node<CM, T>(p, f, forced, Fx, Fy, [&](auto I, T v) { p.dstp[decltype(I)::value][n] = v; });
            store_images(p, x, y);
        }
        if (++s == STAGES) { s = 0; phase ^= 1u; }
    }
}

// cfg = 100 * (TW / 128) + 10 * STAGES + MINB (CTAs of TW threads per SM)  (0: the default for this <lattice, dtype>)
template <int CM, typename T, int TW, int STAGES, int MINB>
static int launch_tma_cfg(const KParams<T> &p, const CUtensorMap &tm, long long step, int gx, int gy, cudaStream_t s) {
    auto kern = k_step_tma<CM, T, TW, STAGES, MINB>;
    const size_t smem = (size_t)STAGES * Q * tma_row_stride<T>() * sizeof(T);
    if (smem > 227 * 1024) return -1;
    // per device: opt in to the shared-memory size once, remember how many CTAs are co-resident
    static int ctas_of[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return -1;
    if (ctas_of[dev] == 0) {
        int sms = 0, per_sm = 0;
        ctas_of[dev] = -1;
        if (smem > 48 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return -1; }
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TW, smem) != cudaSuccess || per_sm < 1) { cudaGetLastError(); return -1; }
        ctas_of[dev] = sms * per_sm;
    }
    const int ctas = ctas_of[dev];
    if (ctas < 0) return -1;
    const int ntx = (p.nx + TW - 1) / TW;
    const long long ntiles = (long long)ntx * p.nrows;
    const int grid = (int)(ntiles < ctas ? ntiles : ctas);
    if (grid <= 0) return 0;
    kern<<<grid, TW, smem, s>>>(p, tm, step, gx, gy, ntx);
    return 0;
}

template <int CM, typename T>
static int launch_tma_cm(const KParams<T> &p, const CUtensorMap *tm, long long step, int gx, int gy, int cfg, cudaStream_t s) {
    constexpr int ES = (int)sizeof(T);
    constexpr size_t STAGE = (size_t)Q * tma_row_stride<T>() * ES;  // bytes per stage
    if (cfg == 0) cfg = (2 * STAGE * 4 <= 200 * 1024) ? 124 : ((2 * STAGE * 3 <= 200 * 1024) ? 123 : 122);
#define LBM_TMA_CASE(TW_, ST_, MB_) \
    if (cfg == (TW_ / 128) * 100 + ST_ * 10 + MB_) { if constexpr (ST_ * STAGE * MB_ <= 227 * 1024) return launch_tma_cfg<CM, T, TW_, ST_, MB_>(p, *tm, step, gx, gy, s); else return -1; }
    LBM_TMA_CASE(128, 2, 4) LBM_TMA_CASE(128, 2, 3) LBM_TMA_CASE(128, 2, 2)
#ifdef LBM_TUNE
    LBM_TMA_CASE(128, 2, 5) LBM_TMA_CASE(128, 2, 6) LBM_TMA_CASE(128, 3, 2) LBM_TMA_CASE(128, 3, 3) LBM_TMA_CASE(128, 3, 4) LBM_TMA_CASE(128, 4, 2) LBM_TMA_CASE(128, 4, 3)
#endif
#undef LBM_TMA_CASE
    return -1;
}

// -> 0 launched, -1 not available for this configuration (the caller falls back to k_step)
template <typename T>
static int launch_step_tma(int cm, const KParams<T> &p, const void *tmap, long long step, int gx, int gy, int cfg, cudaStream_t s) {
    if (p.nrows <= 0) return 0;
    const CUtensorMap *tm = reinterpret_cast<const CUtensorMap *>(tmap);
    switch (cm) {
    case LBM_SRT: return launch_tma_cm<LBM_SRT, T>(p, tm, step, gx, gy, cfg, s);
    case LBM_TRT: return launch_tma_cm<LBM_TRT, T>(p, tm, step, gx, gy, cfg, s);
    case LBM_MRT: return launch_tma_cm<LBM_MRT, T>(p, tm, step, gx, gy, cfg, s);
    default: return -1;
    }
}
