// Host-side helpers shared by the C-ABI entry points.
#pragma once
#include <limits.h>
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include "../../include/ttipm.h"
#include "common.cuh"
#include "matvec.cuh"

namespace ttipm {

struct DevInfo {
    int sms;
    int smem_optin;
    int dev;       // ordinal of the device the process-wide state (handles, opt-ins, staging, this cache) is bound to
};
DevInfo dev_info();
// The library keeps process-wide state bound to ONE device (one process per GPU, as torch.distributed launches it):
// 0 if the calling thread's current device is that device, else an error through fail().
int check_bound_device();
// process-wide cuBLAS / cuSOLVER handles (cublasHandle_t / cusolverDnHandle_t as void*), re-bound to `st`; nullptr on failure
void* blas_handle(tt_stream_t st);
void* solver_handle(tt_stream_t st);
void pool_keep_freed_blocks();
int block_threads();
int dev_memset(void* p, int v, size_t bytes, tt_stream_t st);
int dev_copy(void* dst, const void* src, size_t bytes, tt_stream_t st);
int fail(int code, const char* fmt, ...);

static inline bool fits_int(int64_t v) { return v >= INT_MIN && v <= INT_MAX; }

static inline int convert_term(const ttipm_term& in, MvTerm& out) {
    out.P1 = in.P1; out.A = in.A; out.P2 = in.P2;
    for (int i = 0; i < 3; ++i) {
        if (!fits_int(in.p1_strides[i]) || !fits_int(in.p2_strides[i])) return 1;
        out.p1s[i] = (int)in.p1_strides[i];
        out.p2s[i] = (int)in.p2_strides[i];
    }
    for (int i = 0; i < 4; ++i) {
        if (!fits_int(in.a_strides[i])) return 1;
        out.as_[i] = (int)in.a_strides[i];
    }
    out.s = in.s; out.S = in.S; out.in_blk = in.in_block; out.out_blk = in.out_block; out.alpha = in.alpha;
    return 0;
}

// large-rank path of K1 (cgemm.cu)
bool mv_big_possible(int nterms, int l, int L, int nm, int nb_out, bool want_sumsq);
bool mv_big_wanted(const MvTerm* t, int nterms, int l, int L, int r, int R, int nm, int nb_out, int nbatch);
int mv_big(const MvTerm* t, int nterms, int l, int L, int r, int R, int nm, int nb_out, const double* x, long x_bs,
           long x_rs, long x_ns, long x_batch, double* y, long y_bs, long y_rs, long y_ns, long y_batch, double y_scale,
           const double* sub, double sub_scale, double* sumsq, int nbatch, tt_stream_t st);

// large-rank path of K2 (cgemm.cu)
struct PhiTermLite {
    const double* Phi;
    const double* A;
    double* out;
    int as_[4];
    int s, S;
};
bool phi_big_wanted(const PhiTermLite* t, int nterms, int forward, int ul, int uL, int vr, int vR, int nm);
int phi_big(const PhiTermLite* t, int nterms, int forward, const double* U, int ul, int uL, const double* V, int vr, int vR,
            int nm, tt_stream_t st);

}  // namespace ttipm
