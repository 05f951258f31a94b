// K1: the AMEn local block matvec  y[l,m,L] = sum P1[l,s,r] A[s,m,n,S] P2[L,S,R] x[r,n,R]
// (reference src/tt_als.py:190-238, cy_src/lgmres_cy.pyx:126-153) as a fused 3-stage DMMA
// chain per (output block, L-tile) work item.  Operands are addressed through strides, so
// the transposed apply ('lsr,smnS,LSR,lmL->rnR') and the permuted residual-interface
// variants ('snmS', 'RSL', 'rsl') need no materialised transposes.
#pragma once
#include "common.cuh"

namespace ttipm {

struct MvTerm {
    const double* P1;   // logical (l, s, r) through p1s
    const double* A;    // logical (s, m, n, S) through as_
    const double* P2;   // logical (L, S, R) through p2s
    int p1s[3];
    int as_[4];
    int p2s[3];
    int s, S;
    int in_blk, out_blk;
    double alpha;
};

// geometry + shared-memory carve-up of one work item; filled on the host by mv_plan()
struct MvGeom {
    int l, L, r, R, nm;     // output ranks (l, L), input ranks (r, R), mode size (m = n)
    int Lt, ntiles;         // L-tile width, number of tiles
    int smax, Smax;
    int ld1, ld2, ldA, ldY; // leading dimensions of T1, T2, As, Ys
    int oT1, oT2, oAs, oYs, oOffs;   // offsets in doubles from the shared base (oOffs too)
    int stageA;             // operator core staged in shared memory (As); 0: stage 2 reads it through its strides (L2)
    int sc;                 // operator-rank chunk of stages 2 / 3: T2 holds sc of the s slices at a time (sc = smax: all)
    int smem_bytes;
};

static inline int mv_pad(int v, int rem, int mod) {   // smallest w >= v with w % mod == rem
    int w = v;
    while (w % mod != rem) ++w;
    return w;
}

// Host: choose the tile width and lay out shared memory.  target_ctas is the number of
// co-resident CTAs one wants to fill (148 SMs on B200); returns 0 or an error.
static inline int mv_plan(MvGeom& g, int l, int L, int r, int R, int nm, int smax, int Smax, int nb_out,
                          int target_ctas, int smem_limit) {
    g.l = l; g.L = L; g.r = r; g.R = R; g.nm = nm; g.smax = smax; g.Smax = Smax;
    int Lt = (L * nb_out + target_ctas - 1) / target_ctas;
    if (Lt < 1) Lt = 1;
    if (Lt > L) Lt = L;
    // The staged operator core As is (nm Smax) x (smax nm) doubles: at operator ranks of ~36 (graphm_3, late IPM
    // iterations) it alone exceeds the shared memory of an SM.  Above a quarter of the limit stage 2 reads the core
    // through its strides instead (a few 100 KB, L2 resident).
    {
        const int ldA0 = mv_pad(smax * nm, 0, 2) + ((smax * nm) % 16 == 0 ? 8 : 0);
        g.stageA = (long)nm * Smax * ldA0 * 8 <= smem_limit / 4 ? 1 : 0;
    }
    int sc = smax;
    for (;;) {
        g.Lt = Lt;
        g.sc = sc;
        g.ntiles = (L + Lt - 1) / Lt;
        g.ld1 = mv_pad(nm * Smax, 4, 8);
        g.ld2 = mv_pad(nm * Lt, 0, 2) + ((nm * Lt) % 16 == 0 ? 8 : 0);
        g.ldA = mv_pad(smax * nm, 0, 2) + ((smax * nm) % 16 == 0 ? 8 : 0);
        g.ldY = nm * Lt;
        int nT1 = r * Lt * g.ld1, nT2 = sc * r * g.ld2, nAs = g.stageA ? nm * Smax * g.ldA : 0, nYs = l * g.ldY;
        int m1 = r * nm, m2 = r * Lt, m3 = l;
        int k1 = R, k2 = nm * Smax, k3 = smax * r;
        int n1 = Lt * Smax, n2 = smax * nm, n3 = nm * Lt;
        int mM = m1 > m2 ? m1 : m2; if (m3 > mM) mM = m3;
        int mK = k1 > k2 ? k1 : k2; if (k3 > mK) mK = k3;
        int mN = n1 > n2 ? n1 : n2; if (n3 > mN) mN = n3;
        int nOffs = (mM + 2 * mK + mN + 1) / 2;     // ints, counted in doubles
        g.oT1 = 0;
        g.oT2 = g.oT1 + nT1;
        g.oAs = g.oT2 + nT2;
        g.oYs = g.oAs + nAs;
        g.oOffs = g.oYs + nYs;
        g.smem_bytes = (g.oOffs + nOffs + 40) * 8;
        if (g.smem_bytes <= smem_limit) return 0;
        if (Lt > 1) {
            Lt = Lt / 2;
            continue;
        }
        // one output column per tile and still too large (left ranks of ~130 with operator ranks of ~25 at graphm_3):
        // stages 2 and 3 run over chunks of the operator rank, T2 holds one chunk
        if (sc == 1) return 1;
        sc = (sc + 1) / 2;
    }
}

#if defined(__CUDACC__) || defined(TTIPM_EMU)
// Small local blocks (the small / medium regime: r, R <= ~8, operator ranks <= ~6): each stage has only a few hundred
// outputs with inner dimensions of 4..24, where the DMMA tile machinery of tgemm (offset tables, three block barriers
// and a dependent tensor-core chain per stage) costs ~1.3 us per stage against ~0.1 us of arithmetic.  Here every
// thread computes whole output elements with plain FMAs straight from the strided operands; one barrier per stage.
// Same contraction order as the reference's 3-GEMM chain (cy_src/lgmres_cy.pyx:146-153).
TT_DEV bool mv_term_is_small(const MvTerm& t, const MvGeom& g, int Ltc) {
    const int n1 = g.r * Ltc * g.nm * t.S, n2 = t.s * g.r * g.nm * Ltc, n3 = g.l * g.nm * Ltc;
    const int lim = 4 * (int)blockDim.x;
    return n1 <= lim && n2 <= lim && n3 <= lim && g.R <= 32 && g.nm * t.S <= 64 && t.s * g.r <= 64 && g.sc >= t.s;
}
TT_DEV void mv_accumulate_term_small(const MvTerm& t, const double* __restrict__ x_blk, int x_rs, int x_ns, const MvGeom& g,
                                     int L0, int Ltc, double* smem) {
    double* T1 = smem + g.oT1;
    double* T2 = smem + g.oT2;
    double* Ys = smem + g.oYs;
    const int r = g.r, R = g.R, nm = g.nm, l = g.l, s = t.s, S = t.S;
    const int ld1 = g.ld1, ld2 = g.ld2, ldY = g.ldY;
    __syncthreads();                         // T1 / T2 of the previous term are no longer read
    // stage 1: T1[(rho, lt), (nu, sig')] = sum_P x[rho, nu, P] P2[L0 + lt, sig', P]
    for (int i = threadIdx.x; i < r * Ltc * nm * S; i += blockDim.x) {
        const int sp = i % S, nu = (i / S) % nm, lt = (i / (S * nm)) % Ltc, rho = i / (S * nm * Ltc);
        const double* xp = x_blk + (long)rho * x_rs + (long)nu * x_ns;
        const double* pp = t.P2 + (long)(L0 + lt) * t.p2s[0] + (long)sp * t.p2s[1];
        double acc = 0.0;
        for (int P = 0; P < R; ++P) acc += xp[P] * pp[(long)P * t.p2s[2]];
        T1[(rho * Ltc + lt) * ld1 + nu * S + sp] = acc;
    }
    __syncthreads();
    // stage 2: T2[(sig, rho), (mu, lt)] = sum_(nu, sig') T1[(rho, lt), (nu, sig')] A[sig, mu, nu, sig']
    for (int i = threadIdx.x; i < s * r * nm * Ltc; i += blockDim.x) {
        const int lt = i % Ltc, mu = (i / Ltc) % nm, rho = (i / (Ltc * nm)) % r, sg = i / (Ltc * nm * r);
        const double* tp = T1 + (rho * Ltc + lt) * ld1;
        const double* ap = t.A + (long)sg * t.as_[0] + (long)mu * t.as_[1];
        double acc = 0.0;
        for (int nu = 0; nu < nm; ++nu)
            for (int sp = 0; sp < S; ++sp) acc += tp[nu * S + sp] * ap[(long)nu * t.as_[2] + (long)sp * t.as_[3]];
        T2[(sg * r + rho) * ld2 + mu * Ltc + lt] = acc;
    }
    __syncthreads();
    // stage 3: Ys[lam, (mu, lt)] += alpha sum_(sig, rho) P1[lam, sig, rho] T2[(sig, rho), (mu, lt)]
    const double alpha = t.alpha;
    for (int i = threadIdx.x; i < l * nm * Ltc; i += blockDim.x) {
        const int c = i % (nm * Ltc), lam = i / (nm * Ltc);
        const double* pp = t.P1 + (long)lam * t.p1s[0];
        double acc = 0.0;
        for (int sg = 0; sg < s; ++sg)
            for (int rho = 0; rho < r; ++rho) acc += pp[(long)sg * t.p1s[1] + (long)rho * t.p1s[2]] * T2[(sg * r + rho) * ld2 + c];
        Ys[lam * ldY + c] += alpha * acc;
    }
    __syncthreads();
}

// Accumulate alpha * (term applied to x_blk) for L-tile [L0, L0+Ltc) into Ys (l x ldY, column (m, Lt)).
//   x_blk: element (rho, nu, Rho) at x_blk[rho * x_rs + nu * x_ns + Rho]
TT_DEV void mv_accumulate_term(const MvTerm& t, const double* __restrict__ x_blk, int x_rs, int x_ns, const MvGeom& g,
                               int L0, int Ltc, double* smem) {
    double* T1 = smem + g.oT1;
    double* T2 = smem + g.oT2;
    double* As = smem + g.oAs;
    double* Ys = smem + g.oYs;
    int* offs = (int*)(smem + g.oOffs);
    const int r = g.r, R = g.R, nm = g.nm, l = g.l, s = t.s, S = t.S;
    const int ld1 = g.ld1, ld2 = g.ld2, ldA = g.ldA, ldY = g.ldY;
    if (mv_term_is_small(t, g, Ltc)) {
        mv_accumulate_term_small(t, x_blk, x_rs, x_ns, g, L0, Ltc, smem);
        return;
    }

    // operator core -> As[(nu, sig'), (sig, mu)]
    if (g.stageA) {
        for (int i = threadIdx.x; i < nm * S * s * nm; i += blockDim.x) {
            const int col = i % (s * nm), row = i / (s * nm);
            const int sg = col / nm, mu = col % nm, nu = row / S, sp = row % S;
            As[row * ldA + col] = t.A[sg * t.as_[0] + mu * t.as_[1] + nu * t.as_[2] + sp * t.as_[3]];
        }
    }
    // stage 1: T1[(rho, Lt), (nu, sig')] = sum_Rho x[(rho, nu), Rho] * P2[(L0 + Lt, sig'), Rho]
    tgemm(r * nm, Ltc * S, R, x_blk, ax2(nm, x_rs, x_ns), ax1(1),
          t.P2 + (long)L0 * t.p2s[0], ax1(t.p2s[2]), ax2(S, t.p2s[0], t.p2s[1]),
          [&](int m, int n, double v) {
              const int rho = m / nm, nu = m % nm, lt = n / S, sp = n % S;
              T1[(rho * Ltc + lt) * ld1 + nu * S + sp] = v;
          },
          offs);
    // stages 2 and 3 over chunks [sg0, sg0 + sn) of the operator rank (one chunk = everything unless the plan had to
    // shrink T2):
    //   T2[(sig, rho), (mu, Lt)] = sum_(nu, sig') T1[(rho, Lt), (nu, sig')] * A[sig, mu, nu, sig']
    //   Ys[lam, (mu, Lt)]       += alpha * sum_(sig, rho) P1[lam, sig, rho] * T2[(sig, rho), (mu, Lt)]
    const double alpha = t.alpha;
    for (int sg0 = 0; sg0 < s; sg0 += g.sc) {
        const int sn = imin(g.sc, s - sg0);
        auto store2 = [&](int m, int n, double v) {
            const int rho = m / Ltc, lt = m % Ltc, sg = n / nm, mu = n % nm;
            T2[(sg * r + rho) * ld2 + mu * Ltc + lt] = v;
        };
        if (g.stageA)
            tgemm(r * Ltc, sn * nm, nm * S, T1, ax1(ld1), ax1(1), As + sg0 * nm, ax1(ldA), ax1(1), store2, offs);
        else        // B[(nu, sig'), (sig, mu)] = A[sig, mu, nu, sig'] straight from the core (composite strides)
            tgemm(r * Ltc, sn * nm, nm * S, T1, ax1(ld1), ax1(1), t.A + (long)sg0 * t.as_[0], ax2(S, t.as_[2], t.as_[3]),
                  ax2(nm, t.as_[0], t.as_[1]), store2, offs);
        tgemm(l, nm * Ltc, sn * r, t.P1 + (long)sg0 * t.p1s[1], ax1(t.p1s[0]), ax2(r, t.p1s[1], t.p1s[2]), T2, ax1(ld2), ax1(1),
              [&](int m, int n, double v) { Ys[m * ldY + n] += alpha * v; }, offs);
    }
}

TT_DEV void mv_zero_tile(const MvGeom& g, double* smem) {
    double* Ys = smem + g.oYs;
    for (int i = threadIdx.x; i < g.l * g.ldY; i += blockDim.x) Ys[i] = 0.0;
    __syncthreads();
}
#endif

}  // namespace ttipm
