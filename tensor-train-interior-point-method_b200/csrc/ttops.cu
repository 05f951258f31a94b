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
};
// reference cy_src/tt_ops_cy.pyx:229-258
TT_GLOBAL void k_block_diag(const BdParams p) {
    const int ro = p.mode == 0 ? p.ra : p.ra + p.rb;
    const int Ro = p.mode == 2 ? p.Ra : p.Ra + p.Rb;
    (void)ro;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += stride) {
        const int j = (int)(e % Ro);
        const int q = (int)((e / Ro) % p.n);
        const int i = (int)(e / ((long)Ro * p.n));
        double v = 0.0;
        if (p.mode == 0) {
            v = j < p.Ra ? p.a[((long)i * p.n + q) * p.Ra + j] : p.b[((long)i * p.n + q) * p.Rb + j - p.Ra];
        } else if (p.mode == 2) {
            v = i < p.ra ? p.a[((long)i * p.n + q) * p.Ra + j] : p.b[((long)(i - p.ra) * p.n + q) * p.Rb + j];
        } else {
            if (i < p.ra && j < p.Ra) v = p.a[((long)i * p.n + q) * p.Ra + j];
            else if (i >= p.ra && j >= p.Ra) v = p.b[((long)(i - p.ra) * p.n + q) * p.Rb + j - p.Ra];
        }
        p.out[e] = v;
    }
}

struct EmbedParams {
    const double* in;
    double* out;
    int r, R, q, mode;   // mode 0: I (x) M, 1: M (x) I  (in (r,2,2,R) -> out (r,4,4,R));  2: diag (in (r,q,R) -> out (r,q,q,R))
    long total;
};
// reference src/tt_ops.py:360-375, :312-316
TT_GLOBAL void k_embed(const EmbedParams p) {
    const long stride = (long)gridDim.x * blockDim.x;
    const int Q = p.mode == 2 ? p.q : 4;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < p.total; e += stride) {
        const int Rr = (int)(e % p.R);
        const int col = (int)((e / p.R) % Q);
        const int row = (int)((e / ((long)p.R * Q)) % Q);
        const int rr = (int)(e / ((long)p.R * Q * Q));
        double v = 0.0;
        if (p.mode == 2) {
            if (row == col) v = p.in[((long)rr * Q + row) * p.R + Rr];
        } else {
            const int m = row >> 1, i = row & 1, n = col >> 1, j = col & 1;
            if (p.mode == 0) {
                if (m == n) v = p.in[(((long)rr * 2 + i) * 2 + j) * p.R + Rr];
            } else {
                if (i == j) v = p.in[(((long)rr * 2 + m) * 2 + n) * p.R + Rr];
            }
        }
        p.out[e] = v;
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
    BdParams p{a, b, out, ra, Ra, rb, Rb, n, mode, 0};
    const long ro = mode == 0 ? ra : ra + rb, Ro = mode == 2 ? Ra : Ra + Rb;
    p.total = ro * n * Ro;
    return launch_kernel("k_block_diag", k_block_diag, dim3(grid_for(p.total)), dim3(block_threads()), 0,
                         (tt_stream_t)stream, false, p);
}

extern "C" int ttipm_embed(const double* in, double* out, int r, int R, int q, int mode, void* stream) {
    if (mode < 0 || mode > 2) return fail(1, "embed: mode %d", mode);
    EmbedParams p{in, out, r, R, q, mode, 0};
    const long Q = mode == 2 ? q : 4;
    p.total = (long)r * Q * Q * R;
    return launch_kernel("k_embed", k_embed, dim3(grid_for(p.total)), dim3(block_threads()), 0, (tt_stream_t)stream,
                         false, p);
}

extern "C" int ttipm_scale2d(const double* in, int64_t in_rs, int64_t in_cs, int rows, int cols, const double* s,
                             int axis, int divide, double* out, void* stream) {
    Scale2Params p{in, s, out, (long)in_rs, (long)in_cs, rows, cols, axis, divide};
    return launch_kernel("k_scale2d", k_scale2d, dim3(grid_for((long)rows * cols)), dim3(block_threads()), 0,
                         (tt_stream_t)stream, false, p);
}
