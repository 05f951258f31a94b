// Memory-bound TT primitives: block-diagonal core assembly (tt_add), Kronecker / diagonal embeddings,
// row / column scaling.  One thread per output element, coalesced along the last (rank) axis.
#include "api_util.h"

namespace ttipm {

struct BdParams {
    const double* a;
    const double* b;
    double* out;
    int ra, Ra, rb, Rb, n, mode;   // mode 0 first (concat last axis), 1 middle (block diagonal), 2 last (concat first axis)
    long total;
    int tw, tw_shift;              // threads along the rank axis (power of two)
};
// reference cy_src/tt_ops_cy.pyx:229-258
TT_GLOBAL void k_block_diag(const BdParams p) {
    pdl_entry();
    // rows = (i, q) of the output core, tw threads along the rank axis j; one decode per row, coalesced segments
    const int Ro = p.mode == 2 ? p.Ra : p.Ra + p.Rb;
    const int ro = p.mode == 0 ? p.ra : p.ra + p.rb;
    const int tx = threadIdx.x & (p.tw - 1), ty = threadIdx.x >> p.tw_shift, rows_per = blockDim.x >> p.tw_shift;
    const long rows = (long)ro * p.n;
    for (long row = (long)blockIdx.x * rows_per + ty; row < rows; row += (long)gridDim.x * rows_per) {
        const int i = (int)(row / p.n), q = (int)(row % p.n);
        double* dst = p.out + row * Ro;
        // source segments of this output row: [0, Ra) from a (or zero), [Ra, Ra + Rb) from b (or zero)
        const double* sa = nullptr;
        const double* sb = nullptr;
        if (p.mode == 0) {
            sa = p.a + ((long)i * p.n + q) * p.Ra;
            sb = p.b + ((long)i * p.n + q) * p.Rb;
        } else if (p.mode == 2) {
            sa = i < p.ra ? p.a + ((long)i * p.n + q) * p.Ra : p.b + ((long)(i - p.ra) * p.n + q) * p.Rb;
        } else {
            if (i < p.ra) sa = p.a + ((long)i * p.n + q) * p.Ra;
            else sb = p.b + ((long)(i - p.ra) * p.n + q) * p.Rb;
        }
        if (p.mode == 2) {
            row_stream(dst, sa, Ro, tx, p.tw, 0, 1.0);
        } else {
            row_stream(dst, sa, p.Ra, tx, p.tw, 0, 1.0);
            row_stream(dst + p.Ra, sb, p.Rb, tx, p.tw, 0, 1.0);
        }
    }
}

struct EmbedParams {
    const double* in;
    double* out;
    int r, R, q, mode;   // mode 0: I (x) M, 1: M (x) I  (in (r,2,2,R) -> out (r,4,4,R));  2: diag (in (r,q,R) -> out (r,q,q,R))
    long total;
    int tw, tw_shift;
};
// reference src/tt_ops.py:360-375, :312-316
TT_GLOBAL void k_embed(const EmbedParams p) {
    pdl_entry();
    const int Q = p.mode == 2 ? p.q : 4;
    const int tx = threadIdx.x & (p.tw - 1), ty = threadIdx.x >> p.tw_shift, rows_per = blockDim.x >> p.tw_shift;
    const long rows = (long)p.r * Q * Q;
    for (long rw = (long)blockIdx.x * rows_per + ty; rw < rows; rw += (long)gridDim.x * rows_per) {
        const int col = (int)(rw % Q), row = (int)((rw / Q) % Q), rr = (int)(rw / ((long)Q * Q));
        const double* src = nullptr;            // the input row this output row copies, or none (zeros)
        if (p.mode == 2) {
            if (row == col) src = p.in + ((long)rr * Q + row) * p.R;
        } else {
            const int m = row >> 1, i = row & 1, n = col >> 1, j = col & 1;
            if (p.mode == 0) {
                if (m == n) src = p.in + (((long)rr * 2 + i) * 2 + j) * p.R;
            } else {
                if (i == j) src = p.in + (((long)rr * 2 + m) * 2 + n) * p.R;
            }
        }
        double* dst = p.out + rw * p.R;
        row_stream(dst, src, p.R, tx, p.tw, 0, 1.0);
    }
}

struct Scale2Params {
    const double* in;
    const double* s;
    double* out;
    long in_rs, in_cs;
    int rows, cols, axis, divide;   // axis 0: by row index, 1: by column index
};
TT_GLOBAL void k_scale2d(const Scale2Params p) {
    pdl_entry();
    const long total = (long)p.rows * p.cols, stride = (long)gridDim.x * blockDim.x;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const int i = (int)(e / p.cols), j = (int)(e % p.cols);
        const double sv = p.s ? p.s[p.axis == 0 ? i : j] : 1.0;
        double v = p.in[i * p.in_rs + j * p.in_cs];
        if (p.divide) v = sv != 0.0 ? v / sv : v;
        else v *= sv;
        p.out[e] = v;
    }
}

static void row_tile2(int inner, int block, int& tw, int& shift) {
    tw = 1; shift = 0;
    while (tw < inner && tw < block) { tw <<= 1; ++shift; }
}
// grid of the row-tiled kernels: rows / (rows per CTA pass), capped at 8 CTAs per SM
static unsigned grid_rows(long rows, int shift) {
    const long rows_per = block_threads() >> shift;
    long blocks = (rows + rows_per - 1) / rows_per;
    const long cap = (long)dev_info().sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static unsigned grid_for(long total) {
    const int bt = block_threads();
    long blocks = (total + bt - 1) / bt;
    const long cap = (long)dev_info().sms * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace ttipm

using namespace ttipm;

extern "C" int ttipm_block_diag(const double* a, const double* b, double* out, int ra, int Ra, int rb, int Rb, int n,
                                int mode, void* stream) {
    if (mode < 0 || mode > 2) return fail(1, "block_diag: mode %d", mode);
    if ((mode == 0 && ra != rb) || (mode == 2 && Ra != Rb)) return fail(1, "block_diag: boundary ranks differ");
    BdParams p{a, b, out, ra, Ra, rb, Rb, n, mode, 0, 1, 0};
    const long ro = mode == 0 ? ra : ra + rb, Ro = mode == 2 ? Ra : Ra + Rb;
    p.total = ro * n * Ro;
    row_tile2((int)Ro, block_threads(), p.tw, p.tw_shift);
    return launch_kernel("k_block_diag", k_block_diag, dim3(grid_rows(ro * n, p.tw_shift)), dim3(block_threads()), 0,
                         (tt_stream_t)stream, false, p);
}

extern "C" int ttipm_embed(const double* in, double* out, int r, int R, int q, int mode, void* stream) {
    if (mode < 0 || mode > 2) return fail(1, "embed: mode %d", mode);
    EmbedParams p{in, out, r, R, q, mode, 0, 1, 0};
    const long Q = mode == 2 ? q : 4;
    p.total = (long)r * Q * Q * R;
    row_tile2(R, block_threads(), p.tw, p.tw_shift);
    return launch_kernel("k_embed", k_embed, dim3(grid_rows((long)r * Q * Q, p.tw_shift)), dim3(block_threads()), 0,
                         (tt_stream_t)stream, false, p);
}

extern "C" int ttipm_scale2d(const double* in, int64_t in_rs, int64_t in_cs, int rows, int cols, const double* s,
                             int axis, int divide, double* out, void* stream) {
    Scale2Params p{in, s, out, (long)in_rs, (long)in_cs, rows, cols, axis, divide};
    return launch_kernel("k_scale2d", k_scale2d, dim3(grid_for((long)rows * cols)), dim3(block_threads()), 0,
                         (tt_stream_t)stream, false, p);
}

// ---------------------------------------------------------------------------------------
// Fused inner-product chain <a, b> of two tensor trains (reference cy_src/tt_ops_cy.pyx:506-520): the running
// (R1 x R2) matrix and the intermediate T stay in shared memory of ONE CTA for the whole chain -- one launch instead of
// two GEMM launches per core.  Per core, in the reference's order:
//   T[j, (n, b)]  = sum_i res[i, j] c1[i, n, b]          (tensordot(result, core1, axes = ([0], [0])))
//   res'[b, e]    = sum_(j, n) T[j, (n, b)] c2[j, n, e]  (tensordot(temp, core2, axes = ([0, 1], [0, 1])))
// Used for small trains (intermediate of a core <= 2048 doubles: bond ranks up to ~20 at mode size 4, ~11 at mode size
// 16); larger trains keep the GEMM chain, which is faster there.
namespace ttipm {

#define TT_INNER_MAX_CORES 40
struct InnerParams {
    const double* a[TT_INNER_MAX_CORES];
    const double* b[TT_INNER_MAX_CORES];
    int ra[TT_INNER_MAX_CORES + 1], rb[TT_INNER_MAX_CORES + 1], nm[TT_INNER_MAX_CORES];
    int d;
    int oT;              // offset of T behind the two res buffers (doubles)
    int res_cap;         // doubles per res buffer
    double* out;
};
TT_GLOBAL void k_tt_inner(const InnerParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    double* res = smem;
    double* nxt = smem + p.res_cap;
    double* T = smem + p.oT;
    if (threadIdx.x == 0) res[0] = 1.0;
    __syncthreads();
    for (int k = 0; k < p.d; ++k) {
        const int r1 = p.ra[k], R1 = p.ra[k + 1], r2 = p.rb[k], R2 = p.rb[k + 1], nn = p.nm[k];
        const double* c1 = p.a[k];
        const double* c2 = p.b[k];
        const int nT = r2 * nn * R1;
        for (int o = threadIdx.x; o < nT; o += blockDim.x) {
            const int j = o / (nn * R1), q = o % (nn * R1);            // q = (n, b)
            double acc = 0.0;
            for (int i = 0; i < r1; ++i) acc += res[i * r2 + j] * c1[(long)i * nn * R1 + q];
            T[o] = acc;
        }
        __syncthreads();
        for (int o = threadIdx.x; o < R1 * R2; o += blockDim.x) {
            const int bb = o / R2, e = o % R2;
            double acc = 0.0;
            for (int j = 0; j < r2; ++j)
                for (int n = 0; n < nn; ++n)
                    acc += T[(j * nn + n) * R1 + bb] * c2[((long)j * nn + n) * R2 + e];
            nxt[o] = acc;
        }
        __syncthreads();
        double* t = res;
        res = nxt;
        nxt = t;
    }
    if (threadIdx.x == 0) p.out[0] = res[0];
}

}  // namespace ttipm

// a[k]: (ra[k], nm[k], ra[k+1]) contiguous device cores, b[k] likewise; out: one device double.
// Returns 0, an error code, or -1 when the chain does not fit the kernel (caller falls back to the GEMM chain).
extern "C" int ttipm_tt_inner_chain(int d, const double* const* a, const double* const* b, const int32_t* ra,
                                    const int32_t* rb, const int32_t* nm, double* out, void* stream) {
    using namespace ttipm;
    if (d < 1) return fail(1, "tt_inner_chain: empty train");
    if (d > TT_INNER_MAX_CORES) return -1;
    InnerParams p;
    long res_cap = 1, t_cap = 1;
    for (int k = 0; k < d; ++k) {
        p.a[k] = a[k]; p.b[k] = b[k]; p.ra[k] = ra[k]; p.rb[k] = rb[k]; p.nm[k] = nm[k];
        res_cap = std::max<long>(res_cap, (long)ra[k + 1] * rb[k + 1]);
        t_cap = std::max<long>(t_cap, (long)rb[k] * nm[k] * ra[k + 1]);
    }
    p.ra[d] = ra[d]; p.rb[d] = rb[d];
    if (ra[0] != 1 || rb[0] != 1 || ra[d] != 1 || rb[d] != 1) return fail(1, "tt_inner_chain: boundary ranks must be 1");
    res_cap += res_cap & 1;
    const long bytes = (2 * res_cap + t_cap + 2) * 8;
    DevInfo di = dev_info();
    // one CTA runs the chain as dependent FMA loops: measured on B200 (tools/bench_inner.py) 36 us per call at d = 5, bond rank 8
    // (~100 us as a GEMM chain) but 909 us at d = 13, bond rank 32 (intermediate of 4096 doubles per core) against ~250 us for
    // the GEMM chain -- the fused form is for the small trains of the IPM iterates only
    if (bytes > di.smem_optin - 1024 || t_cap > 2048) return -1;
    p.d = d; p.res_cap = (int)res_cap; p.oT = (int)(2 * res_cap); p.out = out;
    return launch_kernel("k_tt_inner", k_tt_inner, dim3(1), dim3(block_threads()), (size_t)bytes, (tt_stream_t)stream, false, p);
}
