// Error reporting, device queries and small host utilities of the C ABI.
#include <stdlib.h>
#include <string.h>
#include "api_util.h"

#ifdef TTIPM_EMU
namespace emu {
thread_local Dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;
thread_local Cta* cta = nullptr;
}  // namespace emu
#endif

namespace ttipm {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
#ifndef TTIPM_EMU
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return 6;
    }
#else
    (void)what;
#endif
    return 0;
}

// keep freed blocks in the stream-ordered pool across synchronisations (the default trims it to zero, which turns every
// scratch allocation after a sync into a fresh cudaMalloc)
void pool_keep_freed_blocks() {
#ifndef TTIPM_EMU
    static bool pool_ready = false;
    if (pool_ready) return;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_ready = true;
#endif
}

DevInfo dev_info() {
    static DevInfo cached = {0, 0, 0};
    if (cached.sms) return cached;
#ifdef TTIPM_EMU
    const char* s = getenv("TTIPM_EMU_SMS");
    cached.sms = s ? atoi(s) : 2;
    cached.smem_optin = 227 * 1024;
#else
    int dev = 0;
    cudaGetDevice(&dev);
    cached.dev = dev;
    cudaDeviceGetAttribute(&cached.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&cached.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
#endif
    return cached;
}

#ifndef TTIPM_EMU
}  // namespace ttipm
#include <cublas_v2.h>
#include <cusolverDn.h>
namespace ttipm {
// library handles are expensive to create (tens of ms): one pair per process, re-bound to the caller's stream
void* blas_handle(tt_stream_t st) {
    static cublasHandle_t h = nullptr;
    if (!h && cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) {
        h = nullptr;
        return nullptr;
    }
    cublasSetStream(h, st);
    return (void*)h;
}
void* solver_handle(tt_stream_t st) {
    static cusolverDnHandle_t h = nullptr;
    if (!h && cusolverDnCreate(&h) != CUSOLVER_STATUS_SUCCESS) {
        h = nullptr;
        return nullptr;
    }
    cusolverDnSetStream(h, st);
    return (void*)h;
}
#else
void* blas_handle(tt_stream_t) { return nullptr; }
void* solver_handle(tt_stream_t) { return nullptr; }
#endif

static int g_use_pdl = -1;
int pdl_enabled() {
    if (g_use_pdl < 0) {
        const char* e = getenv("TTIPM_PDL");
        g_use_pdl = (e && e[0] == '0') ? 0 : 1;
    }
    return g_use_pdl;
}

int check_bound_device() {
#ifndef TTIPM_EMU
    const DevInfo di = dev_info();
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != di.dev)
        return fail(6, "libttipm_b200 is bound to CUDA device %d (first use) but the current device is %d: use one process per GPU",
                    di.dev, dev);
#endif
    return 0;
}

int block_threads() {
#ifdef TTIPM_EMU
    static int bt = 0;
    if (!bt) {
        const char* s = getenv("TTIPM_EMU_THREADS");
        bt = s ? atoi(s) : 64;
    }
    return bt;
#else
    return TT_MAX_THREADS;
#endif
}

int dev_memset(void* p, int v, size_t bytes, tt_stream_t st) {
#ifdef TTIPM_EMU
    (void)st;
    memset(p, v, bytes);
    return 0;
#else
    return cudaMemsetAsync(p, v, bytes, st) != cudaSuccess;
#endif
}

int dev_copy(void* dst, const void* src, size_t bytes, tt_stream_t st) {
#ifdef TTIPM_EMU
    (void)st;
    memmove(dst, src, bytes);
    return 0;
#else
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess;
#endif
}

}  // namespace ttipm

extern "C" int ttipm_use_pdl(int on) {
    const int old = ttipm::pdl_enabled();
    if (on >= 0) ttipm::g_use_pdl = on ? 1 : 0;
    return old;
}
extern "C" int ttipm_abi_version(void) { return TTIPM_ABI_VERSION; }
extern "C" const char* ttipm_last_error(void) { return ttipm::g_err; }
extern "C" int ttipm_device_info(int* sm_count, int* smem_optin_bytes) {
    ttipm::DevInfo d = ttipm::dev_info();
    if (sm_count) *sm_count = d.sms;
    if (smem_optin_bytes) *smem_optin_bytes = d.smem_optin;
    return d.sms > 0 ? 0 : 1;
}
