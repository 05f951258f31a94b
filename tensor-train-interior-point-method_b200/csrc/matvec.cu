// K1 / K4 kernels: local block matvec, projected diagonal, dense local operator.
#include "matvec.cuh"
#include "api_util.h"

namespace ttipm {

struct MvParams {
    MvGeom g;
    int nterms;
    MvTerm t[TTIPM_MAX_TERMS];
    const double* x;
    long x_bs, x_rs, x_ns, x_batch;
    double* y;
    long y_bs, y_rs, y_ns, y_batch;
    double y_scale, sub_scale;
    const double* sub;
    double* sumsq;
    int nb_out;
};

// grid.x = nb_out * ntiles, grid.y = batch
TT_GLOBAL void __launch_bounds__(TT_MAX_THREADS) k_block_matvec(const MvParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    const MvGeom& g = p.g;
    const int out_blk = blockIdx.x / g.ntiles, tile = blockIdx.x % g.ntiles, batch = blockIdx.y;
    const int L0 = tile * g.Lt, Ltc = imin(g.Lt, g.L - L0);
    mv_zero_tile(g, smem);
    const double* xb = p.x + (long)batch * p.x_batch;
    for (int it = 0; it < p.nterms; ++it) {
        if (p.t[it].out_blk != out_blk) continue;
        mv_accumulate_term(p.t[it], xb + (long)p.t[it].in_blk * p.x_bs, (int)p.x_rs, (int)p.x_ns, g, L0, Ltc, smem);
    }
    // epilogue: Ys[lam, (mu, lt)] -> y[out_blk][lam, mu, L0 + lt]
    const double* Ys = smem + g.oYs;
    double* yb = p.y + (long)batch * p.y_batch + (long)out_blk * p.y_bs;
    const double* sb = p.sub ? p.sub + (long)out_blk * p.y_bs : nullptr;
    double ss = 0.0;
    const int cols = g.nm * Ltc;
    for (int i = threadIdx.x; i < g.l * cols; i += blockDim.x) {
        const int lam = i / cols, c = i % cols, mu = c / Ltc, lt = c % Ltc;
        const long o = (long)lam * p.y_rs + mu * p.y_ns + L0 + lt;
        double v = p.y_scale * Ys[lam * g.ldY + mu * Ltc + lt];
        if (sb) v += p.sub_scale * sb[o];
        yb[o] = v;
        ss += v * v;
    }
    if (p.sumsq) {
        double* scratch = smem + g.oOffs;   // offs are dead here
        __syncthreads();
        const double tot = block_sum(ss, scratch);
        double* slots = p.sumsq + (long)batch * p.nb_out * g.L + out_blk * g.L;
        if (threadIdx.x == 0) slots[tile] = tot;
        // slots beyond the tile count (L per output block) are cleared by the last tile instead of a memset node
        if (tile == g.ntiles - 1)
            for (int i = g.ntiles + threadIdx.x; i < g.L; i += blockDim.x) slots[i] = 0.0;
    }
}

// diag[l, m, L] = sum_{s,S} P1[l,s,l] A[s,m,m,S] P2[L,S,L]
struct DiagParams {
    MvTerm t;
    int l, L, nm, invert;
    double* out;
};
TT_GLOBAL void k_local_diag(const DiagParams p) {
    pdl_entry();
    const int total = p.l * p.nm * p.L;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int lam = i / (p.nm * p.L), mu = (i / p.L) % p.nm, Lam = i % p.L;
        double acc = 0.0;
        for (int sg = 0; sg < p.t.s; ++sg) {
            const double p1 = p.t.P1[lam * p.t.p1s[0] + sg * p.t.p1s[1] + lam * p.t.p1s[2]];
            double inner = 0.0;
            for (int sp = 0; sp < p.t.S; ++sp)
                inner += p.t.A[sg * p.t.as_[0] + mu * p.t.as_[1] + mu * p.t.as_[2] + sp * p.t.as_[3]] *
                         p.t.P2[Lam * p.t.p2s[0] + sp * p.t.p2s[1] + Lam * p.t.p2s[2]];
            acc += p1 * inner;
        }
        p.out[i] = p.invert ? 1.0 / acc : acc;
    }
}

// dense[(lam,mu,Lam),(rho,nu,Rho)] = sum_{sg,sp} P1[lam,sg,rho] A[sg,mu,nu,sp] P2[Lam,sp,Rho]
// one CTA per (lam, rho) pair: W[sg?]..: first W[(mu,nu),sp] = sum_sg P1[lam,sg,rho] A[sg,mu,nu,sp],
// then out[(mu,Lam),(nu,Rho)] = sum_sp W[(mu,nu),sp] P2[Lam,sp,Rho].
struct DenseParams {
    MvTerm t;
    int l, L, r, R, nm;
    double* out;
};
TT_GLOBAL void k_local_dense(const DenseParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* W = (double*)smem_raw;   // nm*nm*S
    const int lam = blockIdx.x / p.r, rho = blockIdx.x % p.r;
    const int nm = p.nm, S = p.t.S;
    for (int i = threadIdx.x; i < nm * nm * S; i += blockDim.x) {
        const int mu = i / (nm * S), nu = (i / S) % nm, sp = i % S;
        double acc = 0.0;
        for (int sg = 0; sg < p.t.s; ++sg)
            acc += p.t.P1[lam * p.t.p1s[0] + sg * p.t.p1s[1] + rho * p.t.p1s[2]] *
                   p.t.A[sg * p.t.as_[0] + mu * p.t.as_[1] + nu * p.t.as_[2] + sp * p.t.as_[3]];
        W[i] = acc;
    }
    __syncthreads();
    const long ncol = (long)p.r * nm * p.R;
    const int per = nm * p.L * nm * p.R;
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
        const int Rho = i % p.R, nu = (i / p.R) % nm, Lam = (i / (p.R * nm)) % p.L, mu = i / (p.R * nm * p.L);
        double acc = 0.0;
        for (int sp = 0; sp < S; ++sp)
            acc += W[(mu * nm + nu) * S + sp] * p.t.P2[Lam * p.t.p2s[0] + sp * p.t.p2s[1] + Rho * p.t.p2s[2]];
        const long row = ((long)lam * nm + mu) * p.L + Lam, col = ((long)rho * nm + nu) * p.R + Rho;
        p.out[row * ncol + col] = acc;
    }
}

// sum of squares of a stored y (l, nb, nm, L through strides) into slot 0 of the batch entry's partial sums: the
// fallback of the grouped path when an output block has more tiles than sum-of-squares slots (a large left rank next to
// a right rank of 1 or 2 at the ends of the train)
struct SumsqParams {
    const double* y;
    long y_bs, y_rs, y_ns, y_batch;
    int l, L, nm, nb;
    double* sumsq;
    long sumsq_batch;
};
TT_GLOBAL void k_sumsq_strided(const SumsqParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* scr = (double*)smem_raw;
    const double* y = p.y + (long)blockIdx.x * p.y_batch;
    const long total = (long)p.l * p.nb * p.nm * p.L;
    double acc = 0.0;
    for (long e = threadIdx.x; e < total; e += blockDim.x) {
        long t = e;
        const int Lam = (int)(t % p.L); t /= p.L;
        const int mu = (int)(t % p.nm); t /= p.nm;
        const int i = (int)(t % p.nb); const int lam = (int)(t / p.nb);
        const double v = y[i * p.y_bs + lam * p.y_rs + mu * p.y_ns + Lam];
        acc += v * v;
    }
    acc = block_sum(acc, scr);
    if (threadIdx.x == 0) p.sumsq[(long)blockIdx.x * p.sumsq_batch] = acc;
}

}  // namespace ttipm

using namespace ttipm;

static double g_starved_min_flops = 5e7;

extern "C" int ttipm_block_matvec(const ttipm_term* terms, int nterms, int l, int L, int r, int R, int nmode,
                                  int nb_out, const double* x, int64_t x_block_stride, int64_t x_row_stride,
                                  int64_t x_mode_stride, int64_t x_batch_stride, double* y, int64_t y_block_stride,
                                  int64_t y_row_stride, int64_t y_mode_stride, int64_t y_batch_stride, double y_scale,
                                  const double* sub, double sub_scale, double* sumsq, int nbatch, void* stream) {
    if (nterms < 0 || nterms > TTIPM_MAX_TERMS) return fail(1, "block_matvec: nterms=%d out of range", nterms);
    if (l < 1 || L < 1 || r < 1 || R < 1 || nmode < 1 || nb_out < 1 || nbatch < 1)
        return fail(1, "block_matvec: bad dims l=%d L=%d r=%d R=%d n=%d nb=%d batch=%d", l, L, r, R, nmode, nb_out, nbatch);
    MvParams p;
    int smax = 1, Smax = 1;
    for (int i = 0; i < nterms; ++i) {
        if (convert_term(terms[i], p.t[i])) return fail(1, "block_matvec: term %d stride overflow", i);
        if (terms[i].out_block < 0 || terms[i].out_block >= nb_out)
            return fail(1, "block_matvec: term %d out_block %d", i, terms[i].out_block);
        if (terms[i].s > smax) smax = terms[i].s;
        if (terms[i].S > Smax) Smax = terms[i].S;
    }
    p.nterms = nterms;
    if (!fits_int(x_row_stride) || !fits_int(x_mode_stride) || !fits_int(y_row_stride) || !fits_int(y_mode_stride))
        return fail(1, "block_matvec: x / y strides too large");
    // the grouped-GEMM form writes one sum-of-squares slot per output tile: the others are cleared up front (the fused
    // kernel clears its unused slots itself -- no memset node in the launch chain of the common case)
    auto clear_sumsq = [&]() {
        return sumsq && dev_memset(sumsq, 0, sizeof(double) * (size_t)nbatch * nb_out * L, (tt_stream_t)stream);
    };
    // large local blocks: every stage of the chain is a machine-filling GEMM (cgemm.cu)
    if (mv_big_wanted(p.t, nterms, l, L, r, R, nmode, nb_out, nbatch)) {
        if (clear_sumsq()) return fail(5, "memset failed");
        return mv_big(p.t, nterms, l, L, r, R, nmode, nb_out, x, x_block_stride, x_row_stride, x_mode_stride, x_batch_stride,
                      y, y_block_stride, y_row_stride, y_mode_stride, y_batch_stride, y_scale, sub, sub_scale, sumsq, nbatch,
                      (tt_stream_t)stream);
    }
    DevInfo di = dev_info();
    if (mv_plan(p.g, l, L, r, R, nmode, smax, Smax, nb_out, (di.sms * 2 + nbatch - 1) / nbatch, di.smem_optin)) {
        if (clear_sumsq()) return fail(5, "memset failed");
        // the fused kernel's intermediates do not fit shared memory: the grouped-GEMM path has no such limit
        if (mv_big_possible(nterms, l, L, nmode, nb_out, sumsq != nullptr))
            return mv_big(p.t, nterms, l, L, r, R, nmode, nb_out, x, x_block_stride, x_row_stride, x_mode_stride,
                          x_batch_stride, y, y_block_stride, y_row_stride, y_mode_stride, y_batch_stride, y_scale, sub,
                          sub_scale, sumsq, nbatch, (tt_stream_t)stream);
        if (mv_big_possible(nterms, l, L, nmode, nb_out, false)) {
            // more output tiles than sum-of-squares slots: product without the fused norm, then one pass over y
            int rc = mv_big(p.t, nterms, l, L, r, R, nmode, nb_out, x, x_block_stride, x_row_stride, x_mode_stride,
                            x_batch_stride, y, y_block_stride, y_row_stride, y_mode_stride, y_batch_stride, y_scale, sub,
                            sub_scale, nullptr, nbatch, (tt_stream_t)stream);
            if (rc) return rc;
            SumsqParams sp{y, (long)y_block_stride, (long)y_row_stride, (long)y_mode_stride, (long)y_batch_stride, l, L, nmode,
                           nb_out, sumsq, (long)nb_out * L};
            return launch_kernel("k_sumsq_strided", k_sumsq_strided, dim3(nbatch), dim3(block_threads()), 40 * 8,
                                 (tt_stream_t)stream, false, sp);
        }
        return fail(4, "block_matvec: shape l=%d L=%d r=%d R=%d s=%d S=%d needs %d B shared memory (> %d)", l, L, r,
                    R, smax, Smax, p.g.smem_bytes, di.smem_optin);
    }
    {
        // A starved fused launch: the fused kernel owns (output block, L-tile) items, so a block with few output columns
        // (tall-left local blocks of graphm_3: l = 130-190 next to L = 16) runs on a handful of CTAs that each stream the
        // whole left interface.  From 5e7 flop on the three machine-wide launches of the grouped form are faster there
        // (measured on B200, 300 Krylov steps at (l, L) = (130, 16): 80.7 -> 61.0 ms; blocks with L >= 64 keep the fused form).
        double flops = 0.0;
        for (int q = 0; q < nterms; ++q)
            flops += 2.0 * r * nmode * R * L * p.t[q].S + 2.0 * r * L * p.t[q].s * nmode * nmode * p.t[q].S +
                     2.0 * l * nmode * L * r * p.t[q].s;
        if ((long)nb_out * p.g.ntiles * nbatch * 3 < di.sms && flops * nbatch >= g_starved_min_flops &&
            mv_big_possible(nterms, l, L, nmode, nb_out, sumsq != nullptr)) {
            if (clear_sumsq()) return fail(5, "memset failed");
            return mv_big(p.t, nterms, l, L, r, R, nmode, nb_out, x, x_block_stride, x_row_stride, x_mode_stride,
                          x_batch_stride, y, y_block_stride, y_row_stride, y_mode_stride, y_batch_stride, y_scale, sub,
                          sub_scale, sumsq, nbatch, (tt_stream_t)stream);
        }
    }
    p.x = x; p.x_bs = x_block_stride; p.x_rs = x_row_stride; p.x_ns = x_mode_stride; p.x_batch = x_batch_stride;
    p.y = y; p.y_bs = y_block_stride; p.y_rs = y_row_stride; p.y_ns = y_mode_stride; p.y_batch = y_batch_stride;
    p.y_scale = y_scale; p.sub_scale = sub_scale;
    p.sub = sub; p.sumsq = sumsq; p.nb_out = nb_out;
    tt_stream_t st = (tt_stream_t)stream;
    return launch_kernel("k_block_matvec", k_block_matvec, dim3(nb_out * p.g.ntiles, nbatch), dim3(block_threads()),
                         p.g.smem_bytes, st, false, p);
}

extern "C" int ttipm_local_diag(const ttipm_term* term, int l, int L, int nmode, int invert, double* out,
                                void* stream) {
    DiagParams p;
    if (convert_term(*term, p.t)) return fail(1, "local_diag: stride overflow");
    p.l = l; p.L = L; p.nm = nmode; p.invert = invert; p.out = out;
    int total = l * nmode * L, bt = block_threads();
    return launch_kernel("k_local_diag", k_local_diag, dim3((total + bt - 1) / bt), dim3(bt), 0,
                         (tt_stream_t)stream, false, p);
}

extern "C" int ttipm_local_dense(const ttipm_term* term, int l, int L, int r, int R, int nmode, double* out,
                                 void* stream) {
    DenseParams p;
    if (convert_term(*term, p.t)) return fail(1, "local_dense: stride overflow");
    p.l = l; p.L = L; p.r = r; p.R = R; p.nm = nmode; p.out = out;
    return launch_kernel("k_local_dense", k_local_dense, dim3(l * r), dim3(block_threads()),
                         sizeof(double) * nmode * nmode * term->S, (tt_stream_t)stream, false, p);
}
