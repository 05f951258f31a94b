// Dependent-chain latencies of the fp64 building blocks of the Jacobi / Householder steps (cycles, one warp).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
    __shared__ double sm[1024];
    int lane = threadIdx.x & 31;
    sm[threadIdx.x] = seed + threadIdx.x;
    __syncthreads();
    double x = seed + lane * 1e-3, y = 1.000001;
    long long t0, t1;
    int idx = 0;
#define TIME(name, N, ...) { t0 = clock64(); _Pragma("unroll 1") for (int i = 0; i < N; ++i) { __VA_ARGS__; } t1 = clock64(); if (threadIdx.x == 0) cyc[idx] = (t1 - t0) / N; idx++; }
    TIME("dfma", 256, x = x * y + 1e-9)
    TIME("dadd", 256, x = x + y)
    TIME("shfl64", 256, x += __shfl_xor_sync(0xffffffffu, x, 16))
    TIME("warp_sum", 64, { double v = x; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); x = v * 1e-3 + 1.0; })
    TIME("sqrt", 128, x = sqrt(x + 2.0))
    TIME("rsqrt", 128, x = rsqrt(x + 2.0))
    TIME("div", 128, x = 1.0 / (x + 2.0))
    TIME("lds", 256, x += sm[(lane + (int)x) & 1023])
    TIME("syncthreads", 256, __syncthreads())
    TIME("rot_old", 64, { double saa = x, sbb = x + 1.0, sab = 0.3 * x; double d = sbb - saa; double tg = 2.0 * sab / (d + copysign(sqrt(d * d + 4.0 * sab * sab), d)); double cs = rsqrt(1.0 + tg * tg); x = cs * tg + 1.0; })
    TIME("rot_new", 64, { double saa = x, sbb = x + 1.0, sab = 0.3 * x; double d = sbb - saa; double rh = rsqrt(d * d + 4.0 * sab * sab); double c2 = 0.5 + 0.5 * fabs(d) * rh; double rc = rsqrt(c2); x = copysign(sab * rh * rc, d * sab) + 1.0; })
    TIME("house", 64, { double alpha = x, s = 0.5 * x; double beta = -copysign(sqrt(alpha * alpha + s), alpha); double tj = (beta - alpha) / beta; double scale = 1.0 / (alpha - beta); x = tj + scale + 2.0; })
    out[threadIdx.x] = x;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024); cudaMallocManaged(&cyc, 8 * 64);
    const char* names[] = {"dfma", "dadd", "shfl64+dadd", "warp_sum(5 stages)+fma", "sqrt", "rsqrt", "div", "lds+dadd", "syncthreads", "rot_old", "rot_new", "householder scalars"};
    for (int threads : {32, 256, 512}) {
        k<<<1, threads>>>(out, cyc, 1.5);
        cudaDeviceSynchronize();
        printf("threads=%d:", threads);
        for (int i = 0; i < 12; ++i) printf(" %s=%lld", names[i], cyc[i]);
        printf("\n");
    }
    return 0;
}
