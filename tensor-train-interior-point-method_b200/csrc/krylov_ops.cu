// Vector kernels of the HOST-DRIVEN LGMRES (amen_driver.cu: host_lgmres) -- the Krylov solve of local KKT blocks whose
// reduced-operator matvec does not tile well inside the persistent kernel k_lgmres (left ranks of ~130 next to operator
// ranks of ~30 at graphm_3: the first intermediate of one output column fills the shared memory, the persistent kernel
// is left with 16-48 busy CTAs that re-read P1 per column).  There the matvec runs through ttipm_block_matvec (grouped
// contraction GEMMs over the whole machine) and the Krylov vector algebra through the three kernels below; the small
// dense recurrences (Givens rotations, back substitution) run on the host, one synchronising read-back per inner step.
// Same algorithm and the same arithmetic order as k_lgmres (reference cy_src/lgmres_cy.pyx:203-510 over PETSc LGMRES):
// classical Gram-Schmidt in one pass, per-chunk partial sums added in a fixed order.
#include "api_util.h"

namespace ttipm {

#define KR_MAX_VECS 112        // restart <= 100 (+ the new vector, + augmentation vectors in a combination)

struct CgsParams {
    const double* V;     // nvec basis vectors, stride ldv
    long ldv;
    int nvec;
    double* w;           // the new vector (nv)
    long nv;
    int chunk;           // elements per CTA
    double* partials;    // [gridDim.x][nvec] partial dots
    int nparts;
    double* hout;        // nvec coefficients (written by CTA 0 of the update kernel)
    double* sumsq;       // [gridDim.x] partial ||w||^2 after the update
};

// partials[g][i] = sum over chunk g of V_i[e] w[e]: the chunk of w is staged in shared memory, one warp per vector
TT_GLOBAL void k_cgs_dots(const CgsParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* ws = (double*)smem_raw;
    const long e0 = (long)blockIdx.x * p.chunk;
    const int len = (int)(e0 < p.nv ? (p.nv - e0 < p.chunk ? p.nv - e0 : p.chunk) : 0);
    for (int e = threadIdx.x; e < len; e += blockDim.x) ws[e] = p.w[e0 + e];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = wid; i < p.nvec; i += nw) {
        const double* vi = p.V + (long)i * p.ldv + e0;
        double d0 = 0.0, d1 = 0.0;
        int e = lane;
        for (; e + 32 < len; e += 64) {
            d0 += vi[e] * ws[e];
            d1 += vi[e + 32] * ws[e + 32];
        }
        if (e < len) d0 += vi[e] * ws[e];
        const double d = warp_sum(d0 + d1);
        if (lane == 0) p.partials[(long)blockIdx.x * p.nvec + i] = d;
    }
}

// h_i = sum_g partials[g][i] (fixed order);  w <- w - sum_i h_i V_i over the CTA's chunk;  sumsq[g] = partial ||w||^2
TT_GLOBAL void k_cgs_update(const CgsParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* hs = (double*)smem_raw;                   // nvec
    double* scr = hs + KR_MAX_VECS;
    for (int i = threadIdx.x; i < p.nvec; i += blockDim.x) {
        double t = 0.0;
        for (int g = 0; g < p.nparts; ++g) t += p.partials[(long)g * p.nvec + i];
        hs[i] = t;
        if (blockIdx.x == 0) p.hout[i] = t;
    }
    __syncthreads();
    const long e0 = (long)blockIdx.x * p.chunk;
    const long e1 = e0 + p.chunk < p.nv ? e0 + p.chunk : p.nv;
    double s2 = 0.0;
    for (long e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
        double w = p.w[e];
        for (int i = 0; i < p.nvec; ++i) w -= hs[i] * p.V[(long)i * p.ldv + e];
        p.w[e] = w;
        s2 += w * w;
    }
    s2 = block_sum(s2, scr);
    if (threadIdx.x == 0) p.sumsq[blockIdx.x] = s2;
}

struct LincombParams {
    const double* v[KR_MAX_VECS];
    double c[KR_MAX_VECS];
    int nvec;
    const double* base;  // may be NULL
    double beta, scale;
    double* out;
    long nv;
};
// out = beta * base + scale * sum_i c_i v_i   (terms added in index order)
TT_GLOBAL void k_lincomb(const LincombParams p) {
    pdl_entry();
    const long stride = (long)gridDim.x * blockDim.x;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < p.nv; e += stride) {
        double u = 0.0;
        for (int i = 0; i < p.nvec; ++i) u += p.c[i] * p.v[i][e];
        u *= p.scale;
        if (p.base) u += p.beta * p.base[e];
        p.out[e] = u;
    }
}

static int kr_parts(long nv) {
    long g = (nv + 511) / 512;
    if (g < 1) g = 1;
    if (g > 128) g = 128;
    return (int)g;
}

}  // namespace ttipm

using namespace ttipm;

extern "C" int ttipm_cgs_parts(int64_t nv) { return kr_parts((long)nv); }

extern "C" int ttipm_cgs_project(const double* V, int64_t ldv, int nvec, double* w, int64_t nv, double* partials,
                                 double* h_out, double* sumsq, void* stream) {
    if (nvec < 1 || nvec > KR_MAX_VECS || nv < 1) return fail(1, "cgs_project: bad sizes (nvec %d, nv %lld)", nvec, (long long)nv);
    CgsParams p;
    p.V = V; p.ldv = (long)ldv; p.nvec = nvec; p.w = w; p.nv = (long)nv;
    p.nparts = kr_parts((long)nv);
    p.chunk = (int)(((long)nv + p.nparts - 1) / p.nparts);
    p.partials = partials; p.hout = h_out; p.sumsq = sumsq;
    const int bt = block_threads();
    int rc = launch_kernel("k_cgs_dots", k_cgs_dots, dim3(p.nparts), dim3(bt), (size_t)p.chunk * 8, (tt_stream_t)stream, false, p);
    if (rc) return rc;
    return launch_kernel("k_cgs_update", k_cgs_update, dim3(p.nparts), dim3(bt), (KR_MAX_VECS + 40) * 8, (tt_stream_t)stream,
                         false, p);
}

extern "C" int ttipm_lincomb(int nvec, const double* const* vecs, const double* coefs, double scale, const double* base,
                             double beta, double* out, int64_t nv, void* stream) {
    if (nvec < 0 || nvec > KR_MAX_VECS || nv < 1) return fail(1, "lincomb: bad sizes (nvec %d)", nvec);
    LincombParams p;
    for (int i = 0; i < nvec; ++i) {
        p.v[i] = vecs[i];
        p.c[i] = coefs[i];
    }
    p.nvec = nvec; p.base = base; p.beta = beta; p.scale = scale; p.out = out; p.nv = (long)nv;
    const int bt = block_threads();
    long g = ((long)nv + bt - 1) / bt;
    if (g > 296) g = 296;
    return launch_kernel("k_lincomb", k_lincomb, dim3((unsigned)g), dim3(bt), 0, (tt_stream_t)stream, false, p);
}
