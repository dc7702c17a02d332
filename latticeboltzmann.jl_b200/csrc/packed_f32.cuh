// Blackwell packed-FP32 variant of the fused collide+stream kernel (Float32 contexts, fast mode,
// SRT, TRT and regularised MRT).  Included by kernels_inst.cu inside namespace lbm::LBM_NS.
//
// One thread owns TWO x-adjacent nodes: populations travel as 64-bit pairs (LDG.64 / STG.64 on the
// 8-byte aligned rows) and all collision arithmetic runs on the sm_100 packed pipes (FFMA2 / FADD2 /
// FMUL2 via fma/add/sub/mul.rn.f32x2), halving both the memory and the math instruction count per
// node -- the scalar Float32 kernel is issue-bound, not bandwidth-bound (profiles/r01_ncu_summary.md).
// Pull sources with odd c_x are not 8-byte aligned: each lane loads the aligned pair next to it and
// takes the missing element from its neighbour lane (one SHFL; the warp-edge lane does one scalar load).
#pragma once

struct f2 {
    unsigned long long v;
    __device__ __forceinline__ f2() {}
    __device__ __forceinline__ f2(float s) { v = ((unsigned long long)__float_as_uint(s) << 32) | __float_as_uint(s); }
    __device__ __forceinline__ f2(float lo, float hi) { v = ((unsigned long long)__float_as_uint(hi) << 32) | __float_as_uint(lo); }
    __device__ __forceinline__ float lo() const { return __uint_as_float((unsigned)v); }
    __device__ __forceinline__ float hi() const { return __uint_as_float((unsigned)(v >> 32)); }
};
__device__ __forceinline__ f2 operator+(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator-(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 operator*(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

// acc + C * u for a compile-time integer C, given u and the precomputed 2u, 3u
template <int C>
__device__ __forceinline__ f2 add_c(f2 acc, f2 u1, f2 u2, f2 u3) {
    if constexpr (C == 0) return acc;
    else if constexpr (C == 1) return acc + u1;
    else if constexpr (C == -1) return acc - u1;
    else if constexpr (C == 2) return acc + u2;
    else if constexpr (C == -2) return acc - u2;
    else if constexpr (C == 3) return acc + u3;
    else return acc - u3;
}
// C * u (C != 0)
template <int C>
__device__ __forceinline__ f2 mul_c(f2 u1, f2 u2, f2 u3) {
    if constexpr (C == 1) return u1;
    else if constexpr (C == 2) return u2;
    else if constexpr (C == 3) return u3;
    else return f2(0.0f) - (C == -1 ? u1 : (C == -2 ? u2 : u3));
}
template <int CX, int CY>
__device__ __forceinline__ f2 cdot2(f2 ux, f2 ux2, f2 ux3, f2 uy, f2 uy2, f2 uy3) {
    if constexpr (CX != 0) return add_c<CY>(mul_c<CX>(ux, ux2, ux3), uy, uy2, uy3);
    else return mul_c<CY>(uy, uy2, uy3);
}

struct X2Consts {  // loop-invariant broadcast constants
    f2 cs, half, sixth, t24, m3, m6, eighth, c0, c1;
};

// even / odd parts of the equilibrium polynomial (without the leading 1) for direction I:
//   even = a1^2/2 - b/2 [+ (a1^2 (a1^2 - 6b))/24 + b^2/8],  odd = a1 [+ a1 (a1^2 - 3b)/6],  a1 = css c.u, b = css u.u
template <int I>
__device__ __forceinline__ void poly_even_odd(const X2Consts &k, f2 a1, f2 b, f2 hb_e4, f2 &even, f2 &odd) {
    const f2 s = a1 * a1;
    odd = a1;
    even = hb_e4;  // -b/2 (+ b^2/8)
    if constexpr (L::EQ_ORDER >= 2) even = fma2(k.half, s, even);
    if constexpr (L::EQ_ORDER >= 3) odd = fma2(a1 * k.sixth, fma2(k.m3, b, s), odd);
    if constexpr (L::EQ_ORDER >= 4) even = fma2(s * k.t24, fma2(k.m6, b, s), even);
}

// Loads the two nodes (x0, x0+1) of row y: g[i] = pulled (stream + BC) deviation pair of population i.
template <bool PULL>
__device__ __forceinline__ void load_pair(const KParams<float> &p, int x0, int y, bool valid, f2 (&g)[Q]) {
    const unsigned n = (unsigned)y * (unsigned)p.pitch + (unsigned)x0;
    const int lane = threadIdx.x & 31;
    // pass 1: issue EVERY global load (aligned pairs + the warp-edge scalars) before anything
    // consumes one, so a row costs a single memory round trip
    float2 v[Q];
    float edge[Q];
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int cx = PULL ? L::cx(i) : 0;
        edge[i] = 0.0f;
        if constexpr ((cx & 1) == 0) {  // aligned pair (srcp already carries the -c_x shift)
            v[i] = __ldg(reinterpret_cast<const float2 *>((PULL ? p.srcp[i] : p.srcn[i]) + n));
        } else if constexpr (cx > 0) {
            // need [x0-cx, x0-cx+1]: hi of the aligned pair one lane to the left, lo of mine
            v[i] = __ldg(reinterpret_cast<const float2 *>(p.srcp[i] + n + 1));
            if (lane == 0) edge[i] = __ldg(p.srcp[i] + n);
        } else {
            // need [x0-cx, x0-cx+1] with x0-cx odd: hi of the aligned pair at x0-cx-1, lo of the right neighbour's
            v[i] = __ldg(reinterpret_cast<const float2 *>(p.srcp[i] + n - 1));
            if (lane == 31) edge[i] = __ldg(p.srcp[i] + n + 1);
        }
    });
    // pass 2: neighbour-lane exchange
    static_for<0, Q>([&](auto I) {
        constexpr int i = decltype(I)::value;
        constexpr int cx = PULL ? L::cx(i) : 0;
        if constexpr ((cx & 1) == 0) {
            g[i] = f2(v[i].x, v[i].y);
        } else if constexpr (cx > 0) {
            const float left = __shfl_up_sync(0xffffffffu, v[i].y, 1);
            g[i] = f2(lane == 0 ? edge[i] : left, v[i].x);
        } else {
            const float right = __shfl_down_sync(0xffffffffu, v[i].x, 1);
            g[i] = f2(v[i].y, lane == 31 ? edge[i] : right);
        }
    });
    if constexpr (PULL) {
        const int yg = p.y0g + y;
        if (valid && (near_wall(p, x0, yg) || near_wall(p, x0 + 1, yg))) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int x = x0 + e;
                if (!near_wall(p, x, yg)) continue;
                static_for<0, Q>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr int o = L::opp(i);
                    if constexpr (i != o) {
                        const int b = resolve_bc(p, x + 1, yg + 1, L::cx(o), L::cy(o));
                        if (b >= 0) {
                            const float v = bounced<i>(p, b, __ldg(p.srcn[o] + n + e));
                            g[i] = e == 0 ? f2(v, g[i].hi()) : f2(g[i].lo(), v);
                        }
                    }
                });
            }
        }
    }
}

template <int CM, bool PULL, int MINB, bool P2P = false>
__global__ void __launch_bounds__(256, MINB) k_step_x2(const __grid_constant__ KParams<float> p, long long step) {
    const int x0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    // Whole warps stay alive for the shuffles; out-of-range lanes work on x = 0 and skip the store.
    // x = 0 is the periodic image of x = nx, which is exactly what the last valid lane must receive
    // from its right neighbour (the first out-of-range lane), so no special case is needed there.
    const bool valid = x0 < p.nx;
    const int xl = valid ? x0 : 0;
    bool edge_cta = false;  // see k_step: only CTAs that start inside the edge rows take part in the halo protocol
    if constexpr (P2P) {
        edge_cta = (int)(blockIdx.y * blockDim.y) < p.p2p_rows;
        if (edge_cta) p2p_wait(p);
    }
    const LatConst<float> &c = c_lat32;
    X2Consts k;
    k.cs = f2(c.css); k.half = f2(0.5f); k.sixth = f2(1.0f / 6); k.t24 = f2(1.0f / 24); k.m3 = f2(-3.0f);
    k.m6 = f2(-6.0f); k.eighth = f2(0.125f); k.c0 = f2(p.c[0]); k.c1 = f2(p.c[1]);
    for (int r = blockIdx.y * blockDim.y + threadIdx.y; r < p.nrows; r += gridDim.y * blockDim.y) {
        const int y = launched_row(p, r);
        f2 g[Q];
        load_pair<PULL>(p, xl, y, valid, g);
        // moments of the deviations: rho = 1 + sum g, j = sum g c
        f2 drho = g[0];
        static_for<1, Q>([&](auto I) { drho = drho + g[decltype(I)::value]; });
        f2 jx(0.0f), jy(0.0f);
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            if constexpr (L::cx(i) == 1) jx = jx + g[i];
            else if constexpr (L::cx(i) == -1) jx = jx - g[i];
            else if constexpr (L::cx(i) != 0) jx = fma2(f2((float)L::cx(i)), g[i], jx);
            if constexpr (L::cy(i) == 1) jy = jy + g[i];
            else if constexpr (L::cy(i) == -1) jy = jy - g[i];
            else if constexpr (L::cy(i) != 0) jy = fma2(f2((float)L::cy(i)), g[i], jy);
        });
        const f2 rho = f2(1.0f) + drho;
        const f2 inv(rcp_approx(rho.lo()), rcp_approx(rho.hi()));
        f2 ux = jx * inv, uy = jy * inv;
        {
            float Fx0, Fy0, Fx1, Fy1;
            const bool forced = load_force(p, xl, y, step, Fx0, Fy0);
            if (forced) {
                load_force(p, xl + 1 < p.nx ? xl + 1 : xl, y, step, Fx1, Fy1);
                ux = fma2(f2(p.shift), f2(Fx0, Fx1), ux);
                uy = fma2(f2(p.shift), f2(Fy0, Fy1), uy);
            }
        }
        const unsigned n = (unsigned)y * (unsigned)p.pitch + (unsigned)xl;
        auto emit = [&](auto I, f2 v) {
            if (valid) *reinterpret_cast<float2 *>(p.dstp[decltype(I)::value] + n) = make_float2(v.lo(), v.hi());
        };
        if constexpr (CM == LBM_MRT) {
            // regularised MRT (mrt.jl:56-118) on node pairs: the arithmetic of collide_node<LBM_MRT> (symmetric tensors
            // reduced to their unique components, populations folded per opposite pair) in packed form
            f2 a2[3] = {f2(0.0f), f2(0.0f), f2(0.0f)}, a3[4] = {f2(0.0f), f2(0.0f), f2(0.0f), f2(0.0f)};
            f2 a4[5] = {f2(0.0f), f2(0.0f), f2(0.0f), f2(0.0f), f2(0.0f)};
            const bool do2 = NH >= 2 && !p.mrt_skip[2], do3 = NH >= 3 && !p.mrt_skip[3], do4 = NH >= 4 && !p.mrt_skip[4];
            if (do2 || do3 || do4) {
                static_for<0, Q>([&](auto I) {
                    constexpr int i = decltype(I)::value;
                    constexpr int o = L::opp(i);
                    if constexpr (i <= o) {
                        f2 fs = g[i];
                        if constexpr (i != o) fs = g[i] + g[o];
                        if constexpr (NH >= 2) {
                            if (do2) {
#pragma unroll
                                for (int q = 0; q < 3; ++q) a2[q] = fma2(fs, f2(c.H2[i][q]), a2[q]);
                            }
                        }
                        if constexpr (NH >= 3 && i != o) {
                            if (do3) {
                                const f2 fa = g[i] - g[o];
#pragma unroll
                                for (int q = 0; q < 4; ++q) a3[q] = fma2(fa, f2(c.H3[i][q]), a3[q]);
                            }
                        }
                        if constexpr (NH >= 4) {
                            if (do4) {
#pragma unroll
                                for (int q = 0; q < 5; ++q) a4[q] = fma2(fs, f2(c.H4[i][q]), a4[q]);
                            }
                        }
                    }
                });
            }
            const f2 xx = ux * ux, xy = ux * uy, yy = uy * uy;
            if constexpr (NH >= 2) {
                const f2 e[3] = {rho * xx, rho * xy, rho * yy};
                const float m[3] = {1.0f, 2.0f, 1.0f};
#pragma unroll
                for (int q = 0; q < 3; ++q) a2[q] = f2(p.kn[2] * m[q]) * fma2(f2(p.c[4]), a2[q], f2(p.c[5]) * e[q]);
            }
            if constexpr (NH >= 3) {
                const f2 e[4] = {rho * (xx * ux), rho * (xx * uy), rho * (xy * uy), rho * (yy * uy)};
                const float m[4] = {1.0f, 3.0f, 3.0f, 1.0f};
#pragma unroll
                for (int q = 0; q < 4; ++q) a3[q] = f2(p.kn[3] * m[q]) * fma2(f2(p.c[6]), a3[q], f2(p.c[7]) * e[q]);
            }
            if constexpr (NH >= 4) {
                const f2 e[5] = {rho * (xx * xx), rho * (xx * xy), rho * (xx * yy), rho * (xy * yy), rho * (yy * yy)};
                const float m[5] = {1.0f, 4.0f, 6.0f, 4.0f, 1.0f};
#pragma unroll
                for (int q = 0; q < 5; ++q) a4[q] = f2(p.kn[4] * m[q]) * fma2(f2(p.c[8]), a4[q], f2(p.c[9]) * e[q]);
            }
            const f2 csrho = k.cs * rho;
            const f2 ux2 = ux + ux, ux3 = ux2 + ux, uy2 = uy + uy, uy3 = uy2 + uy;
            static_for<0, Q>([&](auto I) {
                constexpr int i = decltype(I)::value;
                constexpr int o = L::opp(i);
                if constexpr (i <= o) {
                    const f2 w(c.w[i]);
                    f2 even = drho;  // deviations: sum(w H_n) = 0 for n <= N
                    if constexpr (NH >= 2) {
                        f2 hs = a2[0] * f2(c.H2[i][0]);
                        hs = fma2(a2[1], f2(c.H2[i][1]), hs);
                        hs = fma2(a2[2], f2(c.H2[i][2]), hs);
                        if constexpr (NH >= 4) {
#pragma unroll
                            for (int q = 0; q < 5; ++q) hs = fma2(a4[q], f2(c.H4[i][q]), hs);
                        }
                        even = even + hs;
                    }
                    if constexpr (i == o) {
                        emit(I, w * even);
                    } else {
                        f2 odd = csrho * cdot2<L::cx(i), L::cy(i)>(ux, ux2, ux3, uy, uy2, uy3);
                        if constexpr (NH >= 3) {
                            f2 ho = a3[0] * f2(c.H3[i][0]);
#pragma unroll
                            for (int q = 1; q < 4; ++q) ho = fma2(a3[q], f2(c.H3[i][q]), ho);
                            odd = odd + ho;
                        }
                        emit(I, w * (even + odd));
                        emit(std::integral_constant<int, o>{}, w * (even - odd));
                    }
                }
            });
        } else {
        const f2 b = k.cs * fma2(ux, ux, uy * uy);  // css u.u
        f2 hb_e4(0.0f);  // u-only part of the even polynomial: -b/2 (order >= 2) + b^2/8 (order 4)
        if constexpr (L::EQ_ORDER >= 2) hb_e4 = f2(0.0f) - k.half * b;
        if constexpr (L::EQ_ORDER >= 4) hb_e4 = fma2(k.eighth * b, b, hb_e4);
        const f2 ax = k.cs * ux, ay = k.cs * uy;  // css u, so that a1 = c . (css u)
        const f2 ax2 = ax + ax, ax3 = ax2 + ax, ay2 = ay + ay, ay3 = ay2 + ay;
        static_for<0, Q>([&](auto I) {
            constexpr int i = decltype(I)::value;
            constexpr int o = L::opp(i);
            const f2 w(c.w[i]);
            if constexpr (i == o) {  // rest population: odd part vanishes
                f2 even = hb_e4;
                const f2 geq = w * fma2(rho, even, drho);
                if constexpr (CM == LBM_SRT) emit(I, fma2(k.c0, g[i], k.c1 * geq));
                else emit(I, fma2(k.c0, g[i] - geq, g[i]));
            } else if constexpr (i < o) {
                const f2 a1 = cdot2<L::cx(i), L::cy(i)>(ax, ax2, ax3, ay, ay2, ay3);
                f2 even, odd;
                poly_even_odd<i>(k, a1, b, hb_e4, even, odd);
                const f2 es = w * fma2(rho, even, drho);  // symmetric part of feq - w
                const f2 ea = (w * rho) * odd;            // antisymmetric part
                if constexpr (CM == LBM_SRT) {
                    // srt.jl:58 on deviations: g' = (1 - 1/tau) g + (1/tau)(feq - w), feq_i/o = es +- ea
                    emit(I, fma2(k.c0, g[i], k.c1 * (es + ea)));
                    emit(std::integral_constant<int, o>{}, fma2(k.c0, g[o], k.c1 * (es - ea)));
                } else {
                    // trt.jl:82-94 with c0 = -1/tau_s, c1 = 1/tau_a
                    const f2 f_s = k.half * (g[i] + g[o]), f_a = k.half * (g[i] - g[o]);
                    const f2 s_part = k.c0 * (f_s - es), a_part = k.c1 * (f_a - ea);
                    emit(I, g[i] + (s_part - a_part));
                    emit(std::integral_constant<int, o>{}, g[o] + (s_part + a_part));
                }
            }
        });
        }  // SRT / TRT
        if (valid) {
            store_images(p, xl, y);
            store_images(p, xl + 1, y);
            if constexpr (P2P) {
                if (y < H || y >= p.nyl - H) {
                    p2p_store(p, xl, y);
                    p2p_store(p, xl + 1, y);
                }
            }
        }
    }
    if constexpr (P2P) {
        if (edge_cta) {
            int edge_y = (p.p2p_rows + (int)blockDim.y - 1) / (int)blockDim.y;
            if (edge_y > (int)gridDim.y) edge_y = (int)gridDim.y;
            p2p_signal(p, (unsigned long long)gridDim.x * (unsigned long long)edge_y);
        }
    }
}

template <int CM, bool PULL, int MINB, bool P2P = false>
static void launch_step_x2(const KParams<float> &p, long long step, cudaStream_t s) {
    const int pairs = p.nx / 2;
    dim3 block; pick_block(pairs, block);
    dim3 grid = grid_for(p, block, p.nrows, pairs);
    k_step_x2<CM, PULL, MINB, P2P><<<grid, block, 0, s>>>(p, step);
}
