// TEST INFRASTRUCTURE ONLY (-DTTIPM_EMU, built by tests/emu/build_emu.py).
// A minimal CUDA-on-CPU shim: each CUDA thread of a launch runs as an OS thread,
// __syncthreads / warp shuffles / DMMA are rendezvous points.  It exists so the
// CPU-only test tier can execute the real kernel sources on tiny shapes before
// GPU time is spent; the product library (libttipm_b200.so) never includes it.
#pragma once
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
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

// Each CUDA thread is a user-level fiber (ucontext); one OS thread runs all fibers of a CTA round-robin.
// Barriers are counters that waiting fibers poll between yields, so any mix of block- and warp-level
// rendezvous works.  CTAs of a cooperative launch run on separate OS threads.
namespace emu {
struct Dim3 {
    unsigned x = 1, y = 1, z = 1;
    Dim3() {}
    Dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct Barrier {
    unsigned n = 0, count = 0, gen = 0;
};
struct ClusterBar {          // one per thread-block cluster of a launch (CTAs sharing blockIdx.y/z)
    std::atomic<unsigned> count{0}, gen{0};
    unsigned n = 1;
};
struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false;
};
struct Cta {
    Barrier bar;
    std::vector<Barrier> wbar;
    std::vector<double> xa, xb;   // per-warp exchange buffers (32 doubles each)
    unsigned char* smem = nullptr;
    std::vector<Fiber> fibers;
    ucontext_t sched;
    int current = 0;
    Dim3 grid, block, bidx;
    const std::function<void()>* body = nullptr;
    ClusterBar* cbar = nullptr;
};
extern thread_local Dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
extern thread_local Cta* cta;

inline void yield_fiber() {
    Cta* c = cta;
    swapcontext(&c->fibers[c->current].ctx, &c->sched);
}
inline void barrier_wait(Barrier& b) {
    const unsigned gen = b.gen;
    if (++b.count == b.n) {
        b.count = 0;
        ++b.gen;
        return;
    }
    while (b.gen == gen) yield_fiber();
}
inline void fiber_entry() {
    Cta* c = cta;
    (*c->body)();
    c->fibers[c->current].done = true;
    swapcontext(&c->fibers[c->current].ctx, &c->sched);
}

inline void run_cta(Dim3 grid, Dim3 block, size_t smem, Dim3 bidx, const std::function<void()>& body,
                    ClusterBar* cbar = nullptr) {
    const unsigned nt = block.x, nw = (nt + 31) / 32;
    const size_t stack_bytes = 256 * 1024;
    Cta c;
    c.bar.n = nt;
    c.wbar.resize(nw);
    for (unsigned w = 0; w < nw; ++w) c.wbar[w].n = (w + 1) * 32 <= nt ? 32 : nt - w * 32;
    c.xa.assign(nw * 32, 0.0);
    c.xb.assign(nw * 32, 0.0);
    c.smem = (unsigned char*)aligned_alloc(64, ((smem + 63) / 64 + 1) * 64);
    memset(c.smem, 0xCD, smem);   // poison: uninitialised shared memory reads show up
    c.grid = grid; c.block = block; c.bidx = bidx; c.body = &body;
    c.cbar = cbar;
    c.fibers.resize(nt);
    cta = &c;
    blockIdx_ = bidx; blockDim_ = block; gridDim_ = grid;
    for (unsigned t = 0; t < nt; ++t) {
        Fiber& f = c.fibers[t];
        f.stack = (char*)malloc(stack_bytes);
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = stack_bytes;
        f.ctx.uc_link = &c.sched;
        makecontext(&f.ctx, (void (*)())fiber_entry, 0);
    }
    unsigned alive = nt;
    while (alive) {
        for (unsigned t = 0; t < nt; ++t) {
            Fiber& f = c.fibers[t];
            if (f.done) continue;
            c.current = (int)t;
            threadIdx_ = Dim3(t);
            swapcontext(&c.sched, &f.ctx);
            if (f.done) --alive;
        }
    }
    for (auto& f : c.fibers) free(f.stack);
    free(c.smem);
    cta = nullptr;
}

// concurrent = all CTAs alive at once (needed by kernels that use grid_sync)
inline void launch(Dim3 grid, Dim3 block, size_t smem, bool concurrent, const std::function<void()>& body) {
    std::vector<std::thread> pool;
    std::vector<ClusterBar> cbars(grid.y * grid.z);
    for (auto& cb : cbars) cb.n = grid.x;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                ClusterBar* cb = &cbars[bz * grid.y + by];
                if (concurrent && grid.x * grid.y * grid.z > 1)
                    pool.emplace_back([=, &body]() { run_cta(grid, block, smem, Dim3(bx, by, bz), body, cb); });
                else
                    run_cta(grid, block, smem, Dim3(bx, by, bz), body, cb);
            }
    for (auto& th : pool) th.join();
}
}  // namespace emu
typedef ::emu::Dim3 dim3;

#define threadIdx (::emu::threadIdx_)
#define blockIdx (::emu::blockIdx_)
#define blockDim (::emu::blockDim_)
#define gridDim (::emu::gridDim_)

static inline void __syncthreads() { ::emu::barrier_wait(::emu::cta->bar); }
// thread-block cluster barrier: the CTAs sharing blockIdx.y/z of a concurrent launch (a lone CTA: block barrier)
static inline void emu_cluster_sync_impl() {
    ::emu::barrier_wait(::emu::cta->bar);
    ::emu::ClusterBar* cb = ::emu::cta->cbar;
    if (cb && cb->n > 1 && ::emu::threadIdx_.x == 0) {
        const unsigned gen = cb->gen.load();
        if (cb->count.fetch_add(1) + 1 == cb->n) {
            cb->count.store(0);
            cb->gen.fetch_add(1);
        } else {
            while (cb->gen.load() == gen) std::this_thread::yield();
        }
    }
    ::emu::barrier_wait(::emu::cta->bar);
}
static inline void __syncwarp() { ::emu::barrier_wait(::emu::cta->wbar[::emu::cta->current >> 5]); }
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
    std::this_thread::yield();
    return v;
}
static inline double ld_cg(const double* p) { return *(const volatile double*)p; }
static inline int ld_cg_i(const int* p) { return *(const volatile int*)p; }
}  // namespace ttipm
