// TEST INFRASTRUCTURE ONLY (-DTTIPM_EMU, built by tests/emu/build_emu.py).
// A minimal CUDA-on-CPU shim: each CUDA thread of a launch runs as an OS thread,
// __syncthreads / warp shuffles / DMMA are rendezvous points.  It exists so the
// CPU-only test tier can execute the real kernel sources on tiny shapes before
// GPU time is spent; the product library (libttipm_b200.so) never includes it.
#pragma once
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

#define TT_DEV static inline
#define TT_DEVFN static
#define TT_DEVM inline
#define TT_HD static inline
#define TT_GLOBAL static
#define __launch_bounds__(...)
#define TT_SMEM_DECL(name) unsigned char* name = ::emu::cta->smem
typedef void* tt_stream_t;
typedef int cudaError_t;

namespace emu {
struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
    Dim3() {}
    Dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct Cta {
    pthread_barrier_t bar;
    std::vector<pthread_barrier_t> wbar;
    std::vector<double> xa, xb;   // per-warp exchange buffers (32 doubles each)
    unsigned char* smem = nullptr;
};
extern thread_local Dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
extern thread_local Cta* cta;

inline void run_cta(Dim3 grid, Dim3 block, size_t smem, Dim3 bidx, const std::function<void()>& body,
                    std::vector<std::thread>& pool, std::vector<Cta*>& ctas) {
    const unsigned nt = block.x, nw = (nt + 31) / 32;
    Cta* c = new Cta();
    pthread_barrier_init(&c->bar, nullptr, nt);
    c->wbar.resize(nw);
    for (unsigned w = 0; w < nw; ++w) {
        unsigned cnt = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
        pthread_barrier_init(&c->wbar[w], nullptr, cnt);
    }
    c->xa.assign(nw * 32, 0.0);
    c->xb.assign(nw * 32, 0.0);
    c->smem = (unsigned char*)aligned_alloc(64, ((smem + 63) / 64 + 1) * 64);
    memset(c->smem, 0xCD, smem);   // poison: uninitialised shared memory reads show up
    ctas.push_back(c);
    for (unsigned t = 0; t < nt; ++t) {
        pool.emplace_back([=, &body]() {
            threadIdx_ = Dim3(t);
            blockIdx_ = bidx;
            blockDim_ = block;
            gridDim_ = grid;
            cta = c;
            body();
        });
    }
}

inline void free_ctas(std::vector<Cta*>& ctas) {
    for (Cta* c : ctas) {
        pthread_barrier_destroy(&c->bar);
        for (auto& b : c->wbar) pthread_barrier_destroy(&b);
        free(c->smem);
        delete c;
    }
    ctas.clear();
}

// concurrent = all CTAs alive at once (needed by kernels that use grid_sync)
inline void launch(Dim3 grid, Dim3 block, size_t smem, bool concurrent, const std::function<void()>& body) {
    std::vector<std::thread> pool;
    std::vector<Cta*> ctas;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                run_cta(grid, block, smem, Dim3(bx, by, bz), body, pool, ctas);
                if (!concurrent) {
                    for (auto& th : pool) th.join();
                    pool.clear();
                    free_ctas(ctas);
                }
            }
    for (auto& th : pool) th.join();
    free_ctas(ctas);
}
}  // namespace emu
typedef ::emu::Dim3 dim3;

#define threadIdx (::emu::threadIdx_)
#define blockIdx (::emu::blockIdx_)
#define blockDim (::emu::blockDim_)
#define gridDim (::emu::gridDim_)

static inline void __syncthreads() { pthread_barrier_wait(&::emu::cta->bar); }
static inline void __syncwarp() { pthread_barrier_wait(&::emu::cta->wbar[threadIdx.x >> 5]); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

namespace ttipm {
static inline double warp_xchg(double v, int src_lane) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* xa = &::emu::cta->xa[w * 32];
    xa[lane] = v;
    __syncwarp();
    double r = xa[src_lane];
    __syncwarp();
    return r;
}
static inline double warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += warp_xchg(v, (threadIdx.x & 31) ^ o);
    return v;
}
static inline void dmma884(double a, double b, double& c0, double& c1) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* xa = &::emu::cta->xa[w * 32];
    double* xb = &::emu::cta->xb[w * 32];
    xa[lane] = a;
    xb[lane] = b;
    __syncwarp();
    const int g = lane >> 2, t = lane & 3;
    for (int k = 0; k < 4; ++k) {
        c0 += xa[g * 4 + k] * xb[(2 * t) * 4 + k];
        c1 += xa[g * 4 + k] * xb[(2 * t + 1) * 4 + k];
    }
    __syncwarp();
}
static inline unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v = __atomic_load_n(p, __ATOMIC_ACQUIRE);
    if (v == 0xFFFFFFFFu) std::this_thread::yield();
    std::this_thread::yield();
    return v;
}
static inline double ld_cg(const double* p) { return *(const volatile double*)p; }
static inline int ld_cg_i(const int* p) { return *(const volatile int*)p; }
}  // namespace ttipm
