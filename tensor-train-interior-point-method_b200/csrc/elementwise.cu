// Memory-bound helpers of the sweep: permutations with per-block scaling, block norms, fused
// elementwise combinations with squared-norm partials, truncation residual norms.
#include "api_util.h"

namespace ttipm {

#define EW_MAX_PARTS 256

// ---------------------------------------------------------------------------------------
struct PermParams {
    const double* in;
    double* out;
    int od[4];        // output dims
    long is_[4];      // input stride for each OUTPUT axis
    const double* scale;
    int axis, mode;   // mode 0 none, 1 multiply, 2 divide by scale[o_axis]
    long total;
    int tw, tw_shift; // > 0: row-tiled variant (last axis in place), tw = threads along the last axis (power of two)
};
TT_GLOBAL void k_permute4(const PermParams p) {
    pdl_entry();
    if (p.tw > 0) {
        // last axis contiguous on both sides (every permutation of the sweep): tw threads run along it, the
        // remaining threads of the CTA take further rows; one index decode per row, independent loads in flight
        const int tx = threadIdx.x & (p.tw - 1), ty = threadIdx.x >> p.tw_shift, rows_per = blockDim.x >> p.tw_shift;
        const long rows = (long)p.od[0] * p.od[1] * p.od[2];
        const int inner = p.od[3];
        for (long row = (long)blockIdx.x * rows_per + ty; row < rows; row += (long)gridDim.x * rows_per) {
            long t = row;
            const int o2 = (int)(t % p.od[2]); t /= p.od[2];
            const int o1 = (int)(t % p.od[1]); t /= p.od[1];
            const int o0 = (int)t;
            const double* src = p.in + o0 * p.is_[0] + o1 * p.is_[1] + o2 * p.is_[2];
            double* dst = p.out + row * inner;
            if (p.mode && p.axis == 3) {
                for (int j = tx; j < inner; j += p.tw) dst[j] = p.mode == 1 ? src[j] * p.scale[j] : src[j] / p.scale[j];
            } else {
                // (a division stays a division: same rounding as the reference's `/ scales`)
                const double sv = p.mode ? p.scale[p.axis == 0 ? o0 : p.axis == 1 ? o1 : o2] : 1.0;
                row_stream(dst, src, inner, tx, p.tw, p.mode, sv);
            }
        }
        return;
    }
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += stride) {
        long t = i;
        const int o3 = (int)(t % p.od[3]); t /= p.od[3];
        const int o2 = (int)(t % p.od[2]); t /= p.od[2];
        const int o1 = (int)(t % p.od[1]); t /= p.od[1];
        const int o0 = (int)t;
        double v = p.in[o0 * p.is_[0] + o1 * p.is_[1] + o2 * p.is_[2] + o3 * p.is_[3]];
        if (p.mode) {
            const int idx = p.axis == 0 ? o0 : p.axis == 1 ? o1 : p.axis == 2 ? o2 : o3;
            const double s = p.scale[idx];
            v = p.mode == 1 ? v * s : v / s;
        }
        p.out[i] = v;
    }
}

// ---------------------------------------------------------------------------------------
struct NormParams {
    const double* x;
    double* out;
    int r, b, inner;
    double floor_;
};
// one CTA per block j: out[j] = max(sqrt(sum_{rho, i} x[rho, j, i]^2), floor)
TT_GLOBAL void k_block_norms(const NormParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* scr = (double*)smem_raw;
    const int j = blockIdx.x;
    double s = 0.0;
    const long tot = (long)p.r * p.inner;
    for (long i = threadIdx.x; i < tot; i += blockDim.x) {
        const long rho = i / p.inner, q = i % p.inner;
        const double v = p.x[(rho * p.b + j) * p.inner + q];
        s += v * v;
    }
    s = block_sum(s, scr);
    if (threadIdx.x == 0) p.out[j] = fmax(sqrt(s), p.floor_);
}

// ---------------------------------------------------------------------------------------
struct EwParams {
    int rows, inner;
    double alpha, beta, gamma;
    const double* a; const double* b; const double* c; const double* w;
    long a_rs, b_rs, c_rs, w_rs, o_rs;
    double* out;
    double* sumsq;
    int tw, tw_shift;
};
// out = w .* (alpha a + beta b) + gamma c   over a (rows x inner) panel set; optional sum of squares
TT_DEV double ew_value(const EwParams& p, long ia, long ib, long ic, long iw) {
    double v = p.alpha * p.a[ia];
    if (p.b) v += p.beta * p.b[ib];
    if (p.w) v *= p.w[iw];
    if (p.c) v += p.gamma * p.c[ic];
    return v;
}
TT_GLOBAL void k_ewise(const EwParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* scr = (double*)smem_raw;
    double ss = 0.0;
    if (p.rows == 1) {
        // contiguous operands: flat grid-stride loop, W independent elements per thread and trip (W = 8 for the
        // reduction-only form, whose grid is capped by the 256 partial-sum slots; 4 otherwise)
        const long tot = p.inner, stride = (long)gridDim.x * blockDim.x;
        long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
        if (p.sumsq) {
            for (; i + 7 * stride < tot; i += 8 * stride) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = ew_value(p, i + u * stride, i + u * stride, i + u * stride, i + u * stride);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (p.out) p.out[i + u * stride] = v[u];
                    ss += v[u] * v[u];
                }
            }
        } else {
            for (; i + 3 * stride < tot; i += 4 * stride) {
                double v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = ew_value(p, i + u * stride, i + u * stride, i + u * stride, i + u * stride);
#pragma unroll
                for (int u = 0; u < 4; ++u) p.out[i + u * stride] = v[u];
            }
        }
        for (; i < tot; i += stride) {
            const double v = ew_value(p, i, i, i, i);
            if (p.out) p.out[i] = v;
            ss += v * v;
        }
    } else {
        // strided panels (x[:, j] slices): tw threads along a row, the rest of the CTA on further rows
        const int tx = threadIdx.x & (p.tw - 1), ty = threadIdx.x >> p.tw_shift, rows_per = blockDim.x >> p.tw_shift;
        for (long row = (long)blockIdx.x * rows_per + ty; row < p.rows; row += (long)gridDim.x * rows_per) {
#pragma unroll 4
            for (int q = tx; q < p.inner; q += p.tw) {
                const double v = ew_value(p, row * p.a_rs + q, row * p.b_rs + q, row * p.c_rs + q, row * p.w_rs + q);
                if (p.out) p.out[row * p.o_rs + q] = v;
                ss += v * v;
            }
        }
    }
    if (p.sumsq) {
        ss = block_sum(ss, scr);
        if (threadIdx.x == 0) p.sumsq[blockIdx.x] = ss;
    }
}

// ---------------------------------------------------------------------------------------
struct TruncParams {
    const double* base;   // len
    const double* Y;      // q x len
    double* out;          // q x EW_MAX_PARTS partial sums of squares
    int q;
    long len;
};
// running residual res_j = base - sum_{i >= j} Y_i, j = q-1 .. 0 ; out[j][cta] = partial ||res_j||^2
// (reference src/tt_als.py:338-345 / :466-471 evaluated for every candidate rank at once)
TT_GLOBAL void k_trunc_resnorms(const TruncParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* acc = (double*)smem_raw;   // q x nwarps
    const int nw = blockDim.x >> 5, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < p.q * nw; i += blockDim.x) acc[i] = 0.0;
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    const long start = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long niter = (p.len + stride - 1) / stride;      // uniform trip count (warp reductions inside)
    for (long it = 0; it < niter; ++it) {
        const long e = start + it * stride;
        const bool ok = e < p.len;
        double v = ok ? p.base[e] : 0.0;
        for (int j = p.q - 1; j >= 0; --j) {
            if (ok) v -= p.Y[(long)j * p.len + e];
            const double s = warp_sum(v * v);
            if (lane == 0) acc[j * nw + wid] += s;
        }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < p.q; j += blockDim.x) {
        double t = 0.0;
        for (int w = 0; w < nw; ++w) t += acc[j * nw + w];
        p.out[(long)j * EW_MAX_PARTS + blockIdx.x] = t;
    }
}

}  // namespace ttipm

using namespace ttipm;

// threads along a row of `inner` elements: the smallest power of two >= min(inner, block) (at least 1)
static void row_tile(int inner, int block, int& tw, int& shift) {
    tw = 1; shift = 0;
    while (tw < inner && tw < block) { tw <<= 1; ++shift; }
}

extern "C" int ttipm_permute4(const double* in, const int32_t* in_dims, const int32_t* perm, double* out,
                              const double* scale, int scale_axis, int scale_mode, void* stream) {
    PermParams p;
    long istr[4];
    istr[3] = 1;
    for (int k = 2; k >= 0; --k) istr[k] = istr[k + 1] * in_dims[k + 1];
    int seen = 0;
    p.total = 1;
    for (int k = 0; k < 4; ++k) {
        if (perm[k] < 0 || perm[k] > 3) return fail(1, "permute4: bad perm");
        seen |= 1 << perm[k];
        p.od[k] = in_dims[perm[k]];
        p.is_[k] = istr[perm[k]];
        p.total *= p.od[k];
    }
    if (seen != 15) return fail(1, "permute4: perm is not a permutation");
    if (p.total == 0) return 0;
    p.in = in; p.out = out; p.scale = scale; p.axis = scale_axis; p.mode = scale ? scale_mode : 0;
    const int bt = block_threads();
    long blocks = (p.total + bt - 1) / bt;
    p.tw = 0; p.tw_shift = 0;
    if (perm[3] == 3) {
        row_tile(p.od[3], bt, p.tw, p.tw_shift);
        const long rows = p.total / p.od[3], rows_per = bt >> p.tw_shift;
        blocks = (rows + rows_per - 1) / rows_per;
    }
    const long cap = (long)dev_info().sms * 8;
    if (blocks > cap) blocks = cap;
    return launch_kernel("k_permute4", k_permute4, dim3((unsigned)blocks), dim3(bt), 0, (tt_stream_t)stream, false, p);
}

extern "C" int ttipm_block_norms(const double* x, int r, int b, int inner, double floor_, double* out, void* stream) {
    NormParams p{x, out, r, b, inner, floor_};
    return launch_kernel("k_block_norms", k_block_norms, dim3(b), dim3(block_threads()), 40 * 8, (tt_stream_t)stream,
                         false, p);
}

extern "C" int ttipm_ewise(int rows, int inner, double alpha, const double* a, int64_t a_rs, double beta,
                           const double* b, int64_t b_rs, double gamma, const double* c, int64_t c_rs, const double* w,
                           int64_t w_rs, double* out, int64_t out_rs, double* sumsq, void* stream) {
    if (rows < 1 || inner < 1) return fail(1, "ewise: bad dims");
    EwParams p;
    p.rows = rows; p.inner = inner; p.alpha = alpha; p.beta = beta; p.gamma = gamma;
    p.a = a; p.b = b; p.c = c; p.w = w; p.a_rs = a_rs; p.b_rs = b_rs; p.c_rs = c_rs; p.w_rs = w_rs;
    p.out = out; p.o_rs = out_rs; p.sumsq = sumsq;
    const int bt = block_threads();
    long blocks = ((long)rows * inner + bt - 1) / bt;
    row_tile(inner, bt, p.tw, p.tw_shift);
    if (rows > 1) {
        const long rows_per = bt >> p.tw_shift;
        blocks = (rows + rows_per - 1) / rows_per;
    } else {
        blocks = sumsq ? (blocks + 7) / 8 : (blocks + 3) / 4;       // elements per thread and trip
    }
    // the sum-of-squares partials have EW_MAX_PARTS slots; without them fill the machine
    const long cap = sumsq ? EW_MAX_PARTS : (long)dev_info().sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    tt_stream_t st = (tt_stream_t)stream;
    if (sumsq && dev_memset(sumsq, 0, sizeof(double) * EW_MAX_PARTS, st)) return fail(5, "ewise: memset failed");
    return launch_kernel("k_ewise", k_ewise, dim3((unsigned)blocks), dim3(bt), 40 * 8, st, false, p);
}

extern "C" int ttipm_trunc_resnorms(const double* base, const double* Y, int q, int64_t len, double* out,
                                    void* stream) {
    if (q < 1 || len < 1) return fail(1, "trunc_resnorms: bad dims");
    TruncParams p{base, Y, out, q, (long)len};
    const int bt = block_threads();
    long blocks = (len + bt - 1) / bt;
    if (blocks > EW_MAX_PARTS) blocks = EW_MAX_PARTS;
    tt_stream_t st = (tt_stream_t)stream;
    if (dev_memset(out, 0, sizeof(double) * (size_t)q * EW_MAX_PARTS, st)) return fail(5, "trunc: memset failed");
    const size_t smem = sizeof(double) * (size_t)q * (bt / 32);
    return launch_kernel("k_trunc_resnorms", k_trunc_resnorms, dim3((unsigned)blocks), dim3(bt), smem, st, false, p);
}
