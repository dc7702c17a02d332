// Micro-benchmark behind the layout of the velocity-change reduction (k_reduce<..., LBM_REDUCE_VELOCITY_CHANGE>):
// a kernel that streams Q planes, reads the previous velocity and writes the new one.  Variants:
//   0  two planes old[n], old[N + n] updated in place (N = nx * ny: a power-of-two distance at 4096^2)
//   1  the same with the second plane padded by 4 KB + 128 B
//   2  interleaved pairs old[2 n], old[2 n + 1] (one 16-byte access)
//   3  planes, read from `old`, written to a second buffer (no read-modify-write of a line)
//   4  as 0 without the stores (read only)        5  as 0 without the loads of old (write only)
//   6  no `old` at all: the Q planes themselves are rewritten in place (what an AA-pattern / in-place streaming step does)
//   7..11  as 0 with cache operators: ld.lu | ld.cv | st.cs | st.wt | ld.cg + st.cg
//   12  as 0, but the stored value carries a (numerically void) data dependence on the loaded one, so the store cannot
//       be issued while the load of the same sector is still in flight          13  the same for the interleaved layout
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rmw_reduce rmw_reduce.cu ; run: ./rmw_reduce [n]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int Q = 9;
struct P { const double *f[Q]; double *old, *neu; long long N, off2; int nx, ny, rows; double *part; long long zero; };
__device__ __forceinline__ double after(double v, double loaded, long long zero) {
    return __longlong_as_double(__double_as_longlong(v) | (__double_as_longlong(loaded) & zero));
}
template <int V> __device__ __forceinline__ double ldv(const double *a) {
    double v;
    if (V == 7) asm volatile("ld.global.lu.f64 %0, [%1];" : "=d"(v) : "l"(a));
    else if (V == 8) asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(a));
    else if (V == 11) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(a));
    else v = *a;
    return v;
}
template <int V> __device__ __forceinline__ void stv(double *a, double v) {
    if (V == 9) asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(a), "d"(v));
    else if (V == 10) asm volatile("st.global.wt.f64 [%0], %1;" ::"l"(a), "d"(v));
    else if (V == 11) asm volatile("st.global.cg.f64 [%0], %1;" ::"l"(a), "d"(v));
    else *a = v;
}
template <int V>
__global__ void __launch_bounds__(256, 3) k(const __grid_constant__ P p) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y0 = blockIdx.y * p.rows, y1 = min(p.ny, y0 + p.rows);
    double a0 = 0, a1 = 0;
    if (x < p.nx)
        for (int y = y0; y < y1; ++y) {
            const long long n = (long long)y * p.nx + x;
            double f[Q], rho = 0, jx = 0, jy = 0;
#pragma unroll
            for (int i = 0; i < Q; ++i) f[i] = (V == 6) ? p.f[i][n] : __ldg(p.f[i] + n);
#pragma unroll
            for (int i = 0; i < Q; ++i) { rho += f[i]; jx += f[i] * (i % 3 - 1); jy += f[i] * (i / 3 - 1); }
            const double ux = jx / rho, uy = jy / rho;
            double ox = 0, oy = 0;
            if (V == 2 || V == 13) { const double2 o = reinterpret_cast<const double2 *>(p.old)[n]; ox = o.x; oy = o.y; }
            else if (V != 5 && V != 6) { ox = ldv<V>(p.old + n); oy = ldv<V>(p.old + p.off2 + n); }
            a0 += (ux - ox) * (ux - ox) + (uy - oy) * (uy - oy);
            a1 += ox * ox + oy * oy;
            if (V == 2) reinterpret_cast<double2 *>(p.old)[n] = make_double2(ux, uy);
            else if (V == 13) reinterpret_cast<double2 *>(p.old)[n] = make_double2(after(ux, ox, p.zero), after(uy, oy, p.zero));
            else if (V == 12) { p.old[n] = after(ux, ox, p.zero); p.old[p.off2 + n] = after(uy, oy, p.zero); }
            else if (V == 3) { p.neu[n] = ux; p.neu[p.off2 + n] = uy; }
            else if (V == 6) {
#pragma unroll
                for (int i = 0; i < Q; ++i) const_cast<double *>(p.f[i])[n] = f[i] + 1e-9 * ux;
            } else if (V != 4) { stv<V>(p.old + n, ux); stv<V>(p.old + p.off2 + n, uy); }
        }
    for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_down_sync(~0u, a0, o); a1 += __shfl_down_sync(~0u, a1, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(p.part, a0); atomicAdd(p.part + 1, a1); }
}
template <int V> float run(const P &p, dim3 g) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) k<V><<<g, 256>>>(p);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) k<V><<<g, 256>>>(p);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 10;
}
int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 4096;
    const long long N = (long long)n * n;
    P p; p.zero = 0; p.nx = n; p.ny = n; p.N = N;
    double *fbuf; cudaMalloc(&fbuf, (N + 1024) * Q * 8); { double one = 1.0; (void)one; cudaMemset(fbuf, 0x3f, (N + 1024) * Q * 8); }
    for (int i = 0; i < Q; ++i) p.f[i] = fbuf + i * (N + 1024);
    cudaMalloc(&p.old, (2 * N + 1024) * 8); cudaMalloc(&p.neu, (2 * N + 1024) * 8); cudaMalloc(&p.part, 16);
    cudaMemset(p.old, 0, (2 * N + 1024) * 8); cudaMemset(p.neu, 0, (2 * N + 1024) * 8);
    for (int rows : {4, 16}) {
        p.rows = rows;
        dim3 g((n + 255) / 256, (n + rows - 1) / rows);
        float t[14];
        p.off2 = N; t[0] = run<0>(p, g);
        p.off2 = N + 528; t[1] = run<1>(p, g);
        p.off2 = N; t[2] = run<2>(p, g); t[3] = run<3>(p, g); t[4] = run<4>(p, g); t[5] = run<5>(p, g);
        t[6] = run<6>(p, g); t[7] = run<7>(p, g); t[8] = run<8>(p, g); t[9] = run<9>(p, g); t[10] = run<10>(p, g); t[11] = run<11>(p, g); t[12] = run<12>(p, g); t[13] = run<13>(p, g);
        printf("{\"n\": %d, \"rows_per_cta\": %d, \"ms\": {\"planes_in_place\": %.4f, \"planes_padded\": %.4f, \"interleaved\": %.4f, \"planes_two_buffers\": %.4f, \"read_only\": %.4f, \"write_only\": %.4f, "
               "\"q_planes_in_place\": %.4f, \"ld_lu\": %.4f, \"ld_cv\": %.4f, \"st_cs\": %.4f, \"st_wt\": %.4f, \"ld_cg_st_cg\": %.4f, \"planes_dependent_store\": %.4f, \"interleaved_dependent_store\": %.4f}}\n",
               n, rows, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], t[12], t[13]);
    }
    return 0;
}
