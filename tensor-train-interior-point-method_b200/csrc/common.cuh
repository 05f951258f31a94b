// Shared device-side building blocks of libttipm_b200 (sm_100a).
//
//  * fp64 tensor-core (DMMA, mma.sync.m8n8k4.f64) warp tiles
//  * tgemm: a block-cooperative "tensor-contraction GEMM" whose three axes are
//    composite indices described by AxisMap stride pairs, so every contraction of
//    the AMEn hot path (reference src/tt_als.py:190-265, cy_src/lgmres_cy.pyx:126-153)
//    runs through one routine without materialised transposes
//  * a grid barrier for the persistent cooperative kernels
//
// The same sources also compile as plain C++ with -DTTIPM_EMU (tests/emu): every
// CUDA thread becomes an OS thread, which lets the CPU-only test tier execute the
// real kernel code on tiny shapes.  The product library never defines TTIPM_EMU.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#ifdef TTIPM_EMU
#include "emu.h"
#else
#include <cuda_runtime.h>
#define TT_DEV __device__ __forceinline__
#define TT_DEVFN __device__
#define TT_DEVM __device__ __forceinline__
#define TT_HD __host__ __device__ __forceinline__
#define TT_GLOBAL __global__
#define TT_SMEM_DECL(name) extern __shared__ __align__(16) unsigned char name[]
typedef cudaStream_t tt_stream_t;
#endif

#define TT_MAX_THREADS 256
#define TT_WARP 32

namespace ttipm {

// ---------------------------------------------------------------------------
// error reporting for the C ABI
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

// ---------------------------------------------------------------------------
// composite-axis addressing:  off(i) = (i / n1) * s0 + (i % n1) * s1
// ---------------------------------------------------------------------------
struct AxisMap {
    int n1;
    int s0;
    int s1;
};
#define TT_AX_BIG (1 << 30)
TT_DEV AxisMap ax1(int stride) { return AxisMap{TT_AX_BIG, 0, stride}; }
TT_DEV AxisMap ax2(int n1, int s0, int s1) { return AxisMap{n1, s0, s1}; }
TT_DEV int axoff(const AxisMap& a, int i) { return (i / a.n1) * a.s0 + (i % a.n1) * a.s1; }

// ---------------------------------------------------------------------------
// DMMA 8x8x4:  D(8x8) += A(8x4, row) * B(4x8, col)
//   a : A[lane/4][lane%4]       b : B[lane%4][lane/4]
//   c0,c1 : C[lane/4][2*(lane%4) + {0,1}]
// ---------------------------------------------------------------------------
#ifndef TTIPM_EMU
TT_DEV void dmma884(double a, double b, double& c0, double& c1) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
TT_DEV double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
TT_DEV unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
TT_DEV double ld_cg(const double* p) { return __ldcg(p); }
TT_DEV int ld_cg_i(const int* p) { return __ldcg(p); }
#endif

// ---------------------------------------------------------------------------
// block-wide sum (deterministic order), scratch: >= 32 doubles of shared memory
// ---------------------------------------------------------------------------
TT_DEV double block_sum(double v, double* scratch) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) scratch[wid] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += scratch[w];
    __syncthreads();
    return t;
}

// ---------------------------------------------------------------------------
// grid barrier (cooperative launch guarantees co-residency).  `counter` is a
// zero-initialised device word; `epoch` is a per-thread running target.
// ---------------------------------------------------------------------------
TT_DEV void grid_sync(unsigned* counter, unsigned& epoch) {
    if (gridDim.x == 1) {
        __syncthreads();
        return;
    }
    __syncthreads();
    epoch += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < epoch) {
        }
        __threadfence();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------
// Programmatic dependent launch (sm_90+): every kernel starts with pdl_entry().  launch_dependents lets the NEXT kernel
// of the stream be scheduled while this one still runs (its CTAs then sit in griddepcontrol.wait); wait returns once
// every kernel this one depends on has completed and its memory is visible, so no kernel touches global memory before
// its predecessors are done.  The solve is a chain of ~1700 us-sized dependent launches per KKT system: overlapping
// the launch latency of one kernel with the execution of the previous one is worth ~1-2 us per launch.  Both
// instructions are no-ops for launches without the programmatic-serialization attribute.
// ---------------------------------------------------------------------------
#ifndef TTIPM_EMU
TT_DEV void pdl_entry() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#else
TT_DEV void pdl_entry() {}
#endif
int pdl_enabled();      // api.cu: process-wide switch (ttipm_use_pdl)

// ---------------------------------------------------------------------------
// thread-block cluster barrier (hardware; release/acquire at cluster scope, so global and distributed shared
// memory written before the barrier is visible to every CTA of the cluster after it).  All threads of every
// CTA of the cluster must call it.
// ---------------------------------------------------------------------------
#ifndef TTIPM_EMU
TT_DEV void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
#else
TT_DEV void cluster_sync_all() { emu_cluster_sync_impl(); }
#endif

// ---------------------------------------------------------------------------
// 1-D bulk asynchronous copies (TMA engine, SASS UBLKCP) global -> shared with an mbarrier that counts the
// bytes landed (SASS SYNCS).  Addresses and sizes are multiples of 16 bytes.  One thread issues, every thread
// that reads the data waits on the barrier's phase parity.  Under the CPU emulator the issuing thread copies
// synchronously and the wait is a block barrier.
// ---------------------------------------------------------------------------
#ifndef TTIPM_EMU
TT_DEV unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
TT_DEV void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
TT_DEV void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
TT_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
TT_DEV void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// orders generic-proxy accesses (ld/st) against async-proxy accesses (bulk copies) of the same memory
TT_DEV void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#else
TT_DEV void mbar_init(unsigned long long*, unsigned) {}
TT_DEV void mbar_expect_tx(unsigned long long*, unsigned) {}
TT_DEV void mbar_wait(unsigned long long*, unsigned) { __syncthreads(); }
TT_DEV void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long*) { memcpy(dst, src, bytes); }
TT_DEV void fence_proxy_async() {}
#endif

// ---------------------------------------------------------------------------
// tgemm:  C(m, n) = sum_k A(m, k) * B(k, n)   for m < M, n < N, k < K
//   A element = A[axoff(aM, m) + axoff(aK, k)],  B element = B[axoff(bK, k) + axoff(bN, n)]
//   store(m, n, value) is called exactly once per output element.
//   offs: shared-memory scratch of at least (M + 2*K + N) ints.
// All threads of the block must call it; it synchronises the block on entry and exit.
// ---------------------------------------------------------------------------
template <int MI, int NI, class Store>
TT_DEV void tgemm_tiles(int M, int N, int K, const double* __restrict__ A, const double* __restrict__ B,
                        const int* oAM, const int* oAK, const int* oBK, const int* oBN, Store store) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int TMs = 8 * MI, TNs = 8 * NI;
    const int tm = (M + TMs - 1) / TMs, tn = (N + TNs - 1) / TNs;
    for (int tile = wid; tile < tm * tn; tile += nw) {
        const int m0 = (tile / tn) * TMs, n0 = (tile % tn) * TNs;
        double acc[MI][NI][2];
        int am[MI], bn[NI];
        bool amok[MI], bnok[NI];
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int m = m0 + 8 * i + g;
            amok[i] = m < M;
            am[i] = amok[i] ? oAM[m] : 0;
#pragma unroll
            for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int n = n0 + 8 * j + g;
            bnok[j] = n < N;
            bn[j] = bnok[j] ? oBN[n] : 0;
        }
        for (int k0 = 0; k0 < K; k0 += 4) {
            const int k = k0 + t;
            const bool kok = k < K;
            const int ak = kok ? oAK[k] : 0, bk = kok ? oBK[k] : 0;
            double a[MI], b[NI];
#pragma unroll
            for (int i = 0; i < MI; ++i) a[i] = (kok && amok[i]) ? A[am[i] + ak] : 0.0;
#pragma unroll
            for (int j = 0; j < NI; ++j) b[j] = (kok && bnok[j]) ? B[bk + bn[j]] : 0.0;
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < NI; ++j) dmma884(a[i], b[j], acc[i][j][0], acc[i][j][1]);
        }
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int m = m0 + 8 * i + g;
            if (m < M) {
#pragma unroll
                for (int j = 0; j < NI; ++j) {
                    const int n = n0 + 8 * j + 2 * t;
                    if (n < N) store(m, n, acc[i][j][0]);
                    if (n + 1 < N) store(m, n + 1, acc[i][j][1]);
                }
            }
        }
    }
}

template <class Store>
TT_DEV void tgemm(int M, int N, int K, const double* __restrict__ A, AxisMap aM, AxisMap aK,
                  const double* __restrict__ B, AxisMap bK, AxisMap bN, Store store, int* offs) {
    int* oAM = offs;
    int* oAK = oAM + M;
    int* oBK = oAK + K;
    int* oBN = oBK + K;
    __syncthreads();
    for (int i = threadIdx.x; i < M; i += blockDim.x) oAM[i] = axoff(aM, i);
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        oAK[i] = axoff(aK, i);
        oBK[i] = axoff(bK, i);
    }
    for (int i = threadIdx.x; i < N; i += blockDim.x) oBN[i] = axoff(bN, i);
    __syncthreads();
    const int nw = blockDim.x >> 5;
    // pick the warp tile so that every warp gets work and skinny axes are not over-padded
    const bool wideM = M > 8 && ((M + 15) / 16) * ((N + 7) / 8) >= nw;
    const bool wideN = N > 8 && ((M + 7) / 8) * ((N + 15) / 16) >= nw;
    if (wideM && wideN && ((M + 15) / 16) * ((N + 15) / 16) >= nw)
        tgemm_tiles<2, 2>(M, N, K, A, B, oAM, oAK, oBK, oBN, store);
    else if (wideM)
        tgemm_tiles<2, 1>(M, N, K, A, B, oAM, oAK, oBK, oBN, store);
    else if (wideN)
        tgemm_tiles<1, 2>(M, N, K, A, B, oAM, oAK, oBK, oBN, store);
    else
        tgemm_tiles<1, 1>(M, N, K, A, B, oAM, oAK, oBK, oBN, store);
    __syncthreads();
}

// One row of a memory-bound copy kernel: dst[j] = op(src[j]) for j = tx, tx + tw, ... < n, four independent loads
// in flight per thread before the first store (dst and src never alias; src == nullptr writes zeros).
// op: 0 copy, 1 multiply by s, 2 divide by s.
TT_DEV void row_stream(double* __restrict__ dst, const double* __restrict__ src, int n, int tx, int tw, int op, double s) {
    if (!src) {
        for (int j = tx; j < n; j += tw) dst[j] = 0.0;
        return;
    }
    for (int j = tx; j < n; j += 4 * tw) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = j + u * tw < n ? src[j + u * tw] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (j + u * tw < n) dst[j + u * tw] = op == 0 ? v[u] : (op == 1 ? v[u] * s : v[u] / s);
    }
}

TT_HD int align_up(int v, int a) { return (v + a - 1) / a * a; }
TT_HD int imin(int a, int b) { return a < b ? a : b; }
TT_HD int imax(int a, int b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------
// launch helper: every kernel takes ONE parameter struct by value
// ---------------------------------------------------------------------------
#ifndef TTIPM_EMU
// largest shared-memory opt-in of the device, cached
static inline int smem_optin_max() {
    static int optin_max = 0;
    if (!optin_max) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&optin_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
    return optin_max;
}
#endif

// Can `grid_x` CTAs (grid = ONE thread-block cluster along x, times grid_y independent clusters) be launched?
template <class P>
bool cluster_launch_possible(void (*kern)(P), int grid_x, int grid_y, dim3 block, size_t smem) {
#ifdef TTIPM_EMU
    (void)kern; (void)grid_x; (void)grid_y; (void)block; (void)smem;
    return true;
#else
    if (grid_x > 16) return false;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_max() > (int)smem ? smem_optin_max() : (int)smem);
    if (grid_x > 8 && cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid_x, grid_y);
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = grid_x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return n >= 1;
#endif
}

// launch with grid.x CTAs forming one thread-block cluster (grid.y independent clusters)
template <class P>
int launch_kernel_cluster(const char* name, void (*kern)(P), dim3 grid, dim3 block, size_t smem, tt_stream_t st,
                          const P& params) {
#ifdef TTIPM_EMU
    (void)name;
    (void)st;
    ::emu::launch(grid, block, smem, true, [&]() { kern(params); });
    return 0;
#else
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = grid.x;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, params);
    if (e != cudaSuccess) {
        set_error("%s: cluster launch (%u CTAs) failed: %s", name, grid.x, cudaGetErrorString(e));
        return 3;
    }
    return 0;
#endif
}

template <class P>
int launch_kernel(const char* name, void (*kern)(P), dim3 grid, dim3 block, size_t smem, tt_stream_t st, bool coop,
                  const P& params) {
#ifdef TTIPM_EMU
    (void)name;
    (void)st;
    ::emu::launch(grid, block, smem, coop, [&]() { kern(params); });
    return 0;
#else
    if (smem > 48 * 1024) {
        // opt in to the device maximum (never to `smem` itself: a later, larger launch of the same kernel would find the
        // attribute lowered and its occupancy query would report zero resident CTAs)
        static int optin_max = 0;
        if (!optin_max) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&optin_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        }
        cudaError_t e = cudaFuncSetAttribute((const void*)kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             optin_max > (int)smem ? optin_max : (int)smem);
        if (e != cudaSuccess) {
            set_error("%s: cannot opt in to %zu B of shared memory: %s", name, smem, cudaGetErrorString(e));
            return 2;
        }
    }
    if (coop) {
        void* args[] = {(void*)&params};
        cudaError_t e = cudaLaunchCooperativeKernel((const void*)kern, grid, block, args, smem, st);
        if (e != cudaSuccess) {
            set_error("%s: cooperative launch failed: %s", name, cudaGetErrorString(e));
            return 3;
        }
        return 0;
    }
    if (pdl_enabled()) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = grid;
        cfg.blockDim = block;
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, params);
        if (e != cudaSuccess) {
            set_error("%s: launch failed: %s", name, cudaGetErrorString(e));
            return 3;
        }
        return 0;
    }
    kern<<<grid, block, smem, st>>>(params);
    return check_launch(name);
#endif
}

}  // namespace ttipm
