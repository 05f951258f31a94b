// Local eigenvalue problems of the step-size eigen sweeps (reference src/tt_als.py:931-1128,
// :1286-1389: scipy eigsh / lobpcg on the one- and two-site projections of a TT matrix).
//
//  k_eig_assemble : dense symmetric projection  0.5 (M + M^T),  M = 'lsr,smnk,kptS,LSR->lmpLrntR'
//                   (two-site, reference :952-959, :1305) or 'lsr,smnS,LSR->lmLrnR' (one-site, :1037-1041, :1346)
//  k_eig_lanczos  : extreme eigenpair of  cA * A + cD * D  (dense symmetric, m x m) in ONE persistent launch:
//                   Lanczos with full re-orthogonalisation (two classical Gram-Schmidt passes), explicit restarts,
//                   Ritz value by multisection of the Sturm sequence, Ritz vector by inverse iteration on the
//                   tridiagonal matrix.  The grid is one thread-block cluster (1..16 CTAs): every CTA owns a slice
//                   of rows of the operator -- resident in shared memory when the slices fit (the cluster's
//                   aggregate shared memory holds the whole matrix up to m ~ 600) -- and of every Lanczos vector;
//                   dot products are per-CTA partials summed in a fixed order after a hardware cluster barrier,
//                   the tiny tridiagonal eigenproblem is solved redundantly by every CTA (no broadcast).
//                   Also returns the Rayleigh quotient and residual of the start vector (the reference's
//                   `eig_val = p^T M p`, `old_res = ||M p - eig_val p||`).
//  ttipm_eig_gen_largest : largest eigenpair of the pencil (-D, A), A positive definite (reference :985, :1071:
//                   eigsh(-D, M=A, which="LA")): Cholesky reduction C = L^-1 (-D) L^-T through cuSOLVER / cuBLAS
//                   (plain library factorisations, like the dense Schur fallback), the same Lanczos kernel on C,
//                   back-substitution of the vector.
#include <float.h>
#include "api_util.h"
#ifndef TTIPM_EMU
#include <cublas_v2.h>
#include <cusolverDn.h>
#endif

namespace ttipm {

// ------------------------------------------------------------------------------------------------------------------
// dense symmetric assembly
//   M[(l,m,p,L), (r,n,t,R)] = sum_S W[l,r,m,n,p,t,S] P2[L,S,R],   W = (P1 . A1) . A2  built by two GEMMs on the host side
// (the contraction order opt_einsum picks for the reference's 'lsr,smnk,kptS,LSR->lmpLrntR': the operator ranks s, k
// only enter the two small GEMMs, never the m^2 outputs -- the step direction of a large Newton solve has TT ranks
// of ~100, where the naive s k S terms per output would be 10^6).
// ------------------------------------------------------------------------------------------------------------------
struct EigAsmParams {
    const double* W;    // (l, l, n1, n1, n2, n2, S) contiguous
    const double* P2;   // (L, S, L) through p2s
    int p2s[3];
    int l, S, L, n1, n2;
    int symmetrise;
    double* out;        // (l n1 n2 L)^2 row-major
};

TT_DEV double eig_asm_entry(const EigAsmParams& p, int lam, int m1, int m2, int Lam, int rho, int q1, int q2, int Rho) {
    const double* w = p.W + ((((((long)lam * p.l + rho) * p.n1 + m1) * p.n1 + q1) * p.n2 + m2) * p.n2 + q2) * p.S;
    const double* p2 = p.P2 + (long)Lam * p.p2s[0] + (long)Rho * p.p2s[2];
    double acc0 = 0.0, acc1 = 0.0;
    int sp = 0;
    for (; sp + 1 < p.S; sp += 2) {
        acc0 += w[sp] * p2[(long)sp * p.p2s[1]];
        acc1 += w[sp + 1] * p2[(long)(sp + 1) * p.p2s[1]];
    }
    if (sp < p.S) acc0 += w[sp] * p2[(long)sp * p.p2s[1]];
    return acc0 + acc1;
}

TT_GLOBAL void k_eig_assemble(const EigAsmParams p) {
    pdl_entry();
    const long m = (long)p.l * p.n1 * p.n2 * p.L;
    const long total = m * m, stride = (long)gridDim.x * blockDim.x;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long row = e / m, col = e % m;
        if (p.symmetrise && col < row) continue;            // the thread of (col, row) writes both
        int t = (int)row;
        const int Lam = t % p.L; t /= p.L;
        const int m2 = t % p.n2; t /= p.n2;
        const int m1 = t % p.n1; const int lam = t / p.n1;
        t = (int)col;
        const int Rho = t % p.L; t /= p.L;
        const int q2 = t % p.n2; t /= p.n2;
        const int q1 = t % p.n1; const int rho = t / p.n1;
        double v = eig_asm_entry(p, lam, m1, m2, Lam, rho, q1, q2, Rho);
        if (p.symmetrise) {
            if (col != row) {
                v = 0.5 * (v + eig_asm_entry(p, rho, q1, q2, Rho, lam, m1, m2, Lam));
                p.out[col * m + row] = v;
            }
        }
        p.out[row * m + col] = v;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Lanczos kernel
// ------------------------------------------------------------------------------------------------------------------
#define EIG_NT 512
#define EIG_PART 8          // partial-sum slots beyond the basis coefficients

struct EigParams {
    const double* A;
    const double* D;        // may be nullptr
    double cA, cD;
    int m;
    const double* v0;       // may be nullptr
    double* x;              // m
    double* out;            // 10
    double* V;              // (K + 1) x m
    double* wbuf;           // 2 x m  (unnormalised new vector, double buffered)
    double* part;           // 2 x G x (K + EIG_PART)
    int K, max_cycles, largest, G, ch;
    double tol;
    int cluster;
    int resM;               // the CTA's rows of the combined operator live in shared memory
    int oM, oV, oW, oH, oT, oRed;   // shared-memory offsets (doubles)
};

struct EigCtx {
    const EigParams& p;
    double* smem;
    int lane, wid, nw, e0, nr;
    int parity;
    TT_DEVM EigCtx(const EigParams& pp, double* s) : p(pp), smem(s), parity(0) {
        lane = threadIdx.x & 31;
        wid = threadIdx.x >> 5;
        nw = blockDim.x >> 5;
        e0 = blockIdx.x * pp.ch;
        nr = imin(pp.ch, pp.m - e0);
        if (nr < 0) nr = 0;
    }
    TT_DEVM void sync() {
        if (p.cluster) cluster_sync_all();
        else __syncthreads();
    }
};

// hs[0..n) <- sum over CTAs of the per-CTA values the caller left in hs[0..n) (fixed order; one cluster barrier)
TT_DEV void eig_allreduce(EigCtx& c, double* hs, int n) {
    const EigParams& p = c.p;
    if (p.G == 1) {
        __syncthreads();
        return;
    }
    const int stride = p.K + EIG_PART;
    double* mine = p.part + ((long)c.parity * p.G + blockIdx.x) * stride;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = hs[i];
    c.sync();
    const double* base = p.part + (long)c.parity * p.G * stride;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        double t = 0.0;
        for (int g = 0; g < p.G; ++g) t += ld_cg(base + (long)g * stride + i);
        hs[i] = t;
    }
    c.parity ^= 1;
    __syncthreads();
}

// wl[0..nr) = sigma * (cA A + cD D)[rows of this CTA] . v,  v = scale * src (src global, length m); ends with a block barrier
TT_DEV void eig_matvec(EigCtx& c, const double* src, double scale, double sigma) {
    const EigParams& p = c.p;
    double* vs = c.smem + p.oV;
    double* wl = c.smem + p.oW;
    const int m = p.m;
    __syncthreads();
    for (int e = threadIdx.x; e < m; e += blockDim.x) vs[e] = ld_cg(src + e) * scale;
    __syncthreads();
    for (int rr = c.wid; rr < c.nr; rr += c.nw) {
        double acc0 = 0.0, acc1 = 0.0;
        if (p.resM) {
            const double* row = c.smem + p.oM + (long)rr * m;
            int j = c.lane;
            for (; j + 32 < m; j += 64) {
                acc0 += row[j] * vs[j];
                acc1 += row[j + 32] * vs[j + 32];
            }
            if (j < m) acc0 += row[j] * vs[j];
        } else {
            const double* ra = p.A + (long)(c.e0 + rr) * m;
            const double* rd = p.D ? p.D + (long)(c.e0 + rr) * m : nullptr;
            for (int j = c.lane; j < m; j += 32) {
                double t = p.cA * ra[j];
                if (rd) t += p.cD * rd[j];
                if (j & 32) acc1 += t * vs[j];
                else acc0 += t * vs[j];
            }
        }
        const double s = warp_sum(acc0 + acc1);
        if (c.lane == 0) wl[rr] = sigma * s;
    }
    __syncthreads();
}

// number of eigenvalues of the symmetric tridiagonal (a[0..n), b[0..n-1)) that are < x
TT_DEV int eig_sturm(const double* a, const double* b, int n, double x, double tiny) {
    int cnt = 0;
    double d = a[0] - x;
    if (fabs(d) < tiny) d = -tiny;
    if (d < 0.0) ++cnt;
    for (int i = 1; i < n; ++i) {
        d = a[i] - x - b[i - 1] * b[i - 1] / d;
        if (fabs(d) < tiny) d = -tiny;
        if (d < 0.0) ++cnt;
    }
    return cnt;
}

// smallest eigenpair of the tridiagonal (ta, tb) of order n: theta by multisection (every thread one section point per
// round), y by inverse iteration (thread 0).  T[0] <- theta, y[0..n) unit vector.  All threads call; ends with a barrier.
TT_DEV void eig_tridiag_smallest(EigCtx& c, int n) {
    const EigParams& p = c.p;
    double* T = c.smem + p.oT;
    const int K1 = p.K + 1;
    double* ta = T;                // K1
    double* tb = ta + K1;          // K1
    double* ty = tb + K1;          // K1
    double* lu = ty + K1;          // 5 * K1: d, dl, du, du2, rhs
    double* sc = lu + 5 * K1;      // [0] theta, [1] lo, [2] hi, [3] tiny
    int* cnt = (int*)(sc + 8);     // blockDim.x + 1 ints
    const int nt = blockDim.x;
    if (threadIdx.x == 0) {
        double lo = ta[0], hi = ta[0], nrm = 0.0;
        for (int i = 0; i < n; ++i) {
            const double r = (i > 0 ? fabs(tb[i - 1]) : 0.0) + (i + 1 < n ? fabs(tb[i]) : 0.0);
            lo = fmin(lo, ta[i] - r);
            hi = fmax(hi, ta[i] + r);
            nrm = fmax(nrm, fabs(ta[i]) + r);
        }
        const double tiny = fmax(nrm, 1e-300) * DBL_EPSILON * 0.25;
        sc[1] = lo - 4.0 * tiny;
        sc[2] = hi + 4.0 * tiny;
        sc[3] = tiny;
    }
    __syncthreads();
    const double tiny = sc[3];
    for (int round = 0; round < 12; ++round) {
        const double lo = sc[1], hi = sc[2];
        __syncthreads();
        if (hi - lo <= 2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)) + 2.0 * tiny) break;
        const double x = lo + (hi - lo) * (double)(threadIdx.x + 1) / (double)(nt + 1);
        const int cn = eig_sturm(ta, tb, n, x, tiny);
        cnt[threadIdx.x + 1] = cn;
        if (threadIdx.x == 0) cnt[0] = 0;
        __syncthreads();
        // the first section point with an eigenvalue below it
        if (cn >= 1 && cnt[threadIdx.x] == 0) {
            sc[2] = x;
            if (threadIdx.x > 0) sc[1] = lo + (hi - lo) * (double)threadIdx.x / (double)(nt + 1);
        }
        if (threadIdx.x == nt - 1 && cn == 0) sc[1] = x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double theta = 0.5 * (sc[1] + sc[2]);
        sc[0] = theta;
        double* d = lu;
        double* dl = d + K1;
        double* du = dl + K1;
        double* du2 = du + K1;
        double* rhs = du2 + K1;
        // LU of T - theta I with partial pivoting (tridiagonal; one extra super-diagonal of fill)
        for (int i = 0; i < n; ++i) {
            d[i] = ta[i] - theta;
            dl[i] = i + 1 < n ? tb[i] : 0.0;
            du[i] = i + 1 < n ? tb[i] : 0.0;
            du2[i] = 0.0;
        }
        int* piv = cnt;
        for (int i = 0; i + 1 < n; ++i) {
            if (fabs(d[i]) >= fabs(dl[i])) {
                if (fabs(d[i]) < tiny) d[i] = tiny;
                const double f = dl[i] / d[i];
                dl[i] = f;
                d[i + 1] -= f * du[i];
                piv[i] = 0;
            } else {
                const double f = d[i] / dl[i];
                d[i] = dl[i];
                dl[i] = f;
                const double tmp = du[i];
                du[i] = d[i + 1];
                d[i + 1] = tmp - f * d[i + 1];
                if (i + 2 < n) {
                    du2[i] = du[i + 1];
                    du[i + 1] = -f * du[i + 1];
                }
                piv[i] = 1;
            }
        }
        if (fabs(d[n - 1]) < tiny) d[n - 1] = tiny;
        for (int i = 0; i < n; ++i) rhs[i] = 1.0 / sqrt((double)n) * ((i & 1) ? 0.9 : 1.0);
        for (int itn = 0; itn < 3; ++itn) {
            for (int i = 0; i + 1 < n; ++i) {
                if (piv[i]) {
                    const double tmp = rhs[i];
                    rhs[i] = rhs[i + 1];
                    rhs[i + 1] = tmp - dl[i] * rhs[i];
                } else {
                    rhs[i + 1] -= dl[i] * rhs[i];
                }
            }
            rhs[n - 1] /= d[n - 1];
            if (n > 1) rhs[n - 2] = (rhs[n - 2] - du[n - 2] * rhs[n - 1]) / d[n - 2];
            for (int i = n - 3; i >= 0; --i) rhs[i] = (rhs[i] - du[i] * rhs[i + 1] - du2[i] * rhs[i + 2]) / d[i];
            double big = 0.0;
            for (int i = 0; i < n; ++i) big = fmax(big, fabs(rhs[i]));
            if (!(big > 0.0) || !(big - big == 0.0)) {
                for (int i = 0; i < n; ++i) rhs[i] = i == 0 ? 1.0 : 0.0;
                big = 1.0;
            }
            double nn = 0.0;
            for (int i = 0; i < n; ++i) {
                rhs[i] /= big;
                nn += rhs[i] * rhs[i];
            }
            nn = 1.0 / sqrt(nn);
            for (int i = 0; i < n; ++i) rhs[i] *= nn;
        }
        for (int i = 0; i < n; ++i) ty[i] = rhs[i];
    }
    __syncthreads();
}

TT_DEV double eig_pseudo(int e, int salt) {
    unsigned h = (unsigned)e * 2654435761u + (unsigned)salt * 40503u + 12345u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    return ((double)(h & 0xFFFFFF) / (double)0x1000000) - 0.5;
}

TT_GLOBAL void __launch_bounds__(EIG_NT) k_eig_lanczos(const EigParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    EigCtx c(p, smem);
    const int m = p.m, K = p.K, K1 = K + 1;
    const double sigma = p.largest ? -1.0 : 1.0;
    double* wl = smem + p.oW;
    double* hs = smem + p.oH;                // K + EIG_PART coefficients / partials
    double* ta = smem + p.oT;
    double* tb = ta + K1;
    double* ty = tb + K1;
    const double* sc = ty + K1 + 5 * K1;
    const int lane = c.lane, wid = c.wid, nw = c.nw, e0 = c.e0, nr = c.nr;

    if (p.resM) {
        for (long i = threadIdx.x; i < (long)nr * m; i += blockDim.x) {
            const long g = (long)e0 * m + i;
            double t = p.cA * p.A[g];
            if (p.D) t += p.cD * p.D[g];
            smem[p.oM + i] = t;
        }
    }
    // ---- start vector: Rayleigh quotient and residual of the raw vector, then normalise -----------------------
    bool have_v0 = p.v0 != nullptr;
    double nrm0 = 0.0;
    if (have_v0) {
        double s = 0.0;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) s += p.v0[e0 + e] * p.v0[e0 + e];
        s = block_sum(s, smem + p.oRed);
        if (threadIdx.x == 0) hs[0] = s;
        eig_allreduce(c, hs, 1);
        nrm0 = sqrt(hs[0]);
        __syncthreads();
        if (!(nrm0 > 0.0) || !(nrm0 - nrm0 == 0.0)) have_v0 = false;
    }
    double rq_raw = 0.0, res_raw = 0.0, early_theta = 0.0, early_res = 0.0;
    int early = 0;
    if (have_v0) {
        eig_matvec(c, p.v0, 1.0, 1.0);
        double s = 0.0;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) s += p.v0[e0 + e] * wl[e];
        s = block_sum(s, smem + p.oRed);
        if (threadIdx.x == 0) hs[0] = s;
        eig_allreduce(c, hs, 1);
        rq_raw = hs[0];
        __syncthreads();
        s = 0.0;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) {
            const double t = wl[e] - rq_raw * p.v0[e0 + e];
            s += t * t;
        }
        s = block_sum(s, smem + p.oRed);
        if (threadIdx.x == 0) hs[0] = s;
        eig_allreduce(c, hs, 1);
        res_raw = sqrt(hs[0]);
        __syncthreads();
        for (int e = threadIdx.x; e < nr; e += blockDim.x) p.V[e0 + e] = p.v0[e0 + e] / nrm0;
        // Residual of the normalised start vector: if it already is an eigenvector to the requested tolerance AND the
        // iteration below does not find a lower eigenvalue (beyond that tolerance), the start vector is returned
        // unchanged -- the sweeps call with the previous local solution, and replacing a converged vector by another
        // member of a cluster of nearly equal eigenvalues only inflates the TT ranks of the eigenvector train.
        // (The iteration itself always runs: a start vector can be an exact eigenvector of a larger eigenvalue.)
        if (p.max_cycles > 0) {
            const double th0 = rq_raw / (nrm0 * nrm0);
            s = 0.0;
            for (int e = threadIdx.x; e < nr; e += blockDim.x) {
                const double t = wl[e] - th0 * p.v0[e0 + e];
                s += t * t;
            }
            s = block_sum(s, smem + p.oRed);
            if (threadIdx.x == 0) hs[0] = s;
            eig_allreduce(c, hs, 1);
            const double r0 = sqrt(hs[0]) / nrm0;
            __syncthreads();
            if (r0 <= p.tol) {
                early = 1;
                early_theta = th0;
                early_res = r0;
            }
        }
    } else {
        double s = 0.0;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) {
            const double t = eig_pseudo(e0 + e, 1) + 1.0;
            p.V[e0 + e] = t;
            s += t * t;
        }
        s = block_sum(s, smem + p.oRed);
        if (threadIdx.x == 0) hs[0] = s;
        eig_allreduce(c, hs, 1);
        const double inv = 1.0 / sqrt(hs[0]);
        __syncthreads();
        for (int e = threadIdx.x; e < nr; e += blockDim.x) p.V[e0 + e] *= inv;
    }
    c.sync();                                           // V[0] visible to every CTA

    double theta = 0.0, res_true = 0.0;
    int matvecs = have_v0 ? 1 : 0, cycles = 0, converged = 0;
    const int Keff = imin(K, m);
    for (int cyc = 0; cyc < p.max_cycles && !converged; ++cyc) {
        ++cycles;
        int nb = 0;                                     // order of the tridiagonal built in this cycle
        double beta_last = 0.0, tnorm = 0.0;
        const double* vsrc = p.V;                       // vector to multiply: V[0], then wbuf / beta
        double vscale = 1.0;
        bool space_done = false;
        for (int j = 0; j < Keff; ++j) {
            eig_matvec(c, vsrc, vscale, sigma);
            ++matvecs;
            double alpha = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int i = wid; i <= j; i += nw) {
                    const double* vi = p.V + (long)i * m + e0;
                    double s = 0.0;
                    for (int e = lane; e < nr; e += 32) s += vi[e] * wl[e];
                    s = warp_sum(s);
                    if (lane == 0) hs[i] = s;
                }
                eig_allreduce(c, hs, j + 1);
                for (int e = threadIdx.x; e < nr; e += blockDim.x) {
                    double t = wl[e];
                    for (int i = 0; i <= j; ++i) t -= hs[i] * p.V[(long)i * m + e0 + e];
                    wl[e] = t;
                }
                alpha += hs[j];
                __syncthreads();
            }
            double s = 0.0;
            for (int e = threadIdx.x; e < nr; e += blockDim.x) s += wl[e] * wl[e];
            s = block_sum(s, smem + p.oRed);
            double* wb = p.wbuf + (long)(j & 1) * m;
            for (int e = threadIdx.x; e < nr; e += blockDim.x) wb[e0 + e] = wl[e];
            if (threadIdx.x == 0) hs[0] = s;
            eig_allreduce(c, hs, 1);                    // also publishes wb
            const double beta = sqrt(hs[0]);
            __syncthreads();
            if (threadIdx.x == 0) {
                ta[j] = alpha;
                tb[j] = beta;
            }
            tnorm = fmax(tnorm, fabs(alpha) + beta);
            nb = j + 1;
            beta_last = beta;
            const bool breakdown = !(beta > 64.0 * DBL_EPSILON * fmax(tnorm, 1e-300));
            if (breakdown) {
                if (threadIdx.x == 0) tb[j] = 0.0;
                beta_last = 0.0;
                if (j + 1 >= m) space_done = true;
            }
            // convergence check on a geometric schedule, at the end of the cycle and at a breakdown
            const bool check = breakdown || j + 1 == Keff || (j + 1 >= 6 && ((j + 1) % 6 == 0));
            __syncthreads();
            if (check) {
                eig_tridiag_smallest(c, nb);
                const double est = fabs(beta_last * ty[nb - 1]);
                if (est <= p.tol || j + 1 == Keff || breakdown) {
                    // for a breakdown inside a larger space the Ritz pair is exact for the invariant subspace reached;
                    // the true-residual test below decides, a restart from the Ritz vector plus a perturbation follows
                    break;
                }
            }
            if (j + 1 < Keff) {
                for (int e = threadIdx.x; e < nr; e += blockDim.x) p.V[(long)(j + 1) * m + e0 + e] = wl[e] / beta;
                vsrc = wb;
                vscale = 1.0 / beta;
            }
            __syncthreads();
        }
        // ---- Ritz vector, true residual ----------------------------------------------------------------------
        theta = sc[0];
        for (int e = threadIdx.x; e < nr; e += blockDim.x) {
            double t = 0.0;
            for (int i = 0; i < nb; ++i) t += ty[i] * p.V[(long)i * m + e0 + e];
            p.x[e0 + e] = t;
        }
        c.sync();
        eig_matvec(c, p.x, 1.0, sigma);
        ++matvecs;
        double s = 0.0, s2 = 0.0;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) {
            s += p.x[e0 + e] * wl[e];
            s2 += p.x[e0 + e] * p.x[e0 + e];
        }
        s = block_sum(s, smem + p.oRed);
        s2 = block_sum(s2, smem + p.oRed);
        if (threadIdx.x == 0) {
            hs[0] = s;
            hs[1] = s2;
        }
        eig_allreduce(c, hs, 2);
        const double xx = hs[1];
        theta = hs[0] / xx;
        __syncthreads();
        s = 0.0;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) {
            const double t = wl[e] - theta * p.x[e0 + e];
            s += t * t;
        }
        s = block_sum(s, smem + p.oRed);
        if (threadIdx.x == 0) hs[0] = s;
        eig_allreduce(c, hs, 1);
        const double xn = sqrt(xx);
        res_true = sqrt(hs[0]) / xn;
        __syncthreads();
        // unit Ritz vector; it is also the start vector of the next cycle (perturbed after a breakdown so that a start
        // inside an invariant subspace cannot hide the rest of the spectrum)
        const bool stuck = beta_last == 0.0 && !space_done && nb < m;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) {
            const double t = p.x[e0 + e] / xn;
            p.x[e0 + e] = t;
            p.V[e0 + e] = stuck ? t + 1e-3 * eig_pseudo(e0 + e, 7 + cyc) : t;
        }
        if (res_true <= p.tol && !(stuck && cyc == 0)) converged = 1;
        if (nb >= m && !stuck) converged = 1;           // the whole space was spanned: nothing more to gain
        if (stuck && !converged) {
            // renormalise the perturbed start vector
            double q = 0.0;
            for (int e = threadIdx.x; e < nr; e += blockDim.x) q += p.V[e0 + e] * p.V[e0 + e];
            q = block_sum(q, smem + p.oRed);
            if (threadIdx.x == 0) hs[0] = q;
            eig_allreduce(c, hs, 1);
            const double inv = 1.0 / sqrt(hs[0]);
            __syncthreads();
            for (int e = threadIdx.x; e < nr; e += blockDim.x) p.V[e0 + e] *= inv;
        }
        c.sync();
    }
    if (early && converged && sigma * early_theta - theta <= p.tol) {       // nothing lower was found: keep the start vector
        for (int e = threadIdx.x; e < nr; e += blockDim.x) p.x[e0 + e] = p.v0[e0 + e] / nrm0;
        theta = sigma * early_theta;
        res_true = early_res;
        __syncthreads();
    }
    // ||M v0 - lambda v0|| with the eigenvalue just computed (reference src/tt_als.py:1315, :1327: tt_min_eig's old_res)
    double res_shift = 0.0;
    if (have_v0 && p.max_cycles > 0) {
        eig_matvec(c, p.v0, 1.0, 1.0);
        const double lam = sigma * theta;
        double s = 0.0;
        for (int e = threadIdx.x; e < nr; e += blockDim.x) {
            const double t = wl[e] - lam * p.v0[e0 + e];
            s += t * t;
        }
        s = block_sum(s, smem + p.oRed);
        if (threadIdx.x == 0) hs[0] = s;
        eig_allreduce(c, hs, 1);
        res_shift = sqrt(hs[0]);
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        p.out[8] = res_shift;
        p.out[0] = sigma * theta;
        p.out[1] = res_true;
        p.out[2] = (double)matvecs;
        p.out[3] = (double)converged;
        p.out[4] = rq_raw;
        p.out[5] = res_raw;
        p.out[6] = (double)cycles;
        p.out[7] = nrm0;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------------
struct EigPlan {
    EigParams p;
    int threads;
    long smem_bytes;
    long ws_doubles;
};

static int g_eig_force_G = 0;

static int eig_plan(EigPlan& pl, int m, int K) {
    if (m < 1) return fail(1, "eig: m=%d", m);
    if (K < 2) K = 2;
    if (K > m) K = m;
    if (K < 1) K = 1;
    EigParams& p = pl.p;
    p.m = m;
    p.K = K;
    const DevInfo di = dev_info();
    const long K1 = K + 1;
    const long small = (long)m /* vs */ + (K + EIG_PART) + 8 * K1 + 8 + (EIG_NT + 2) / 2 + 2 + 64;
    const long avail = (long)di.smem_optin / 8 - small - 64;
    // rows per CTA: resident operator slices when the cluster (<= 16 CTAs) can hold the matrix, else streamed rows
    int G = 1, resM = 0;
    for (int g = 1; g <= 16; g *= 2) {
        const long ch = (m + g - 1) / g;
        if (ch * m + ch <= avail) {
            G = g;
            resM = 1;
            break;
        }
    }
    if (!resM) G = m >= 2048 ? 16 : (m >= 512 ? 8 : 4);
    if (g_eig_force_G > 0) {
        G = g_eig_force_G;
        const long ch = (m + G - 1) / G;
        resM = ch * m + ch <= avail;
    }
    while (G > 1 && (m + G - 1) / G * (G - 1) >= m) G /= 2;     // no empty CTA
    p.G = G;
    p.ch = (m + G - 1) / G;
    p.resM = resM;
    p.cluster = G > 1;
    long o = 0;
    p.oM = (int)o; o += resM ? (long)p.ch * m : 0;
    o = (o + 1) & ~1L;
    p.oV = (int)o; o += m;
    p.oW = (int)o; o += p.ch;
    p.oH = (int)o; o += K + EIG_PART;
    p.oT = (int)o; o += 8 * K1 + 8 + (EIG_NT + 2) / 2 + 2;
    p.oRed = (int)o; o += 64;
    pl.smem_bytes = o * 8;
    pl.threads = EIG_NT;
#ifdef TTIPM_EMU
    pl.threads = block_threads();
#endif
    pl.ws_doubles = K1 * m + 2L * m + 2L * G * (K + EIG_PART) + 16;
    if (pl.smem_bytes > di.smem_optin) return fail(4, "eig: m=%d needs %ld B of shared memory", m, pl.smem_bytes);
    return 0;
}

static int eig_launch(const double* A, double cA, const double* D, double cD, int m, const double* v0, int largest, int K,
                      int max_cycles, double tol, double* x, double* out, double* ws, tt_stream_t st) {
    EigPlan pl;
    int rc = eig_plan(pl, m, K);
    if (rc) return rc;
    EigParams& p = pl.p;
    p.A = A; p.D = D; p.cA = cA; p.cD = cD; p.v0 = v0; p.x = x; p.out = out;
    p.largest = largest; p.max_cycles = max_cycles; p.tol = tol;
    p.V = ws;
    p.wbuf = p.V + (long)(p.K + 1) * m;
    p.part = p.wbuf + 2L * m;
    if (p.G > 1) {
        if (!cluster_launch_possible(k_eig_lanczos, p.G, 1, dim3(pl.threads), (size_t)pl.smem_bytes))
            return fail(4, "eig: a cluster of %d CTAs with %ld B of shared memory cannot be launched", p.G, pl.smem_bytes);
        return launch_kernel_cluster("k_eig_lanczos", k_eig_lanczos, dim3(p.G), dim3(pl.threads), (size_t)pl.smem_bytes, st, p);
    }
    return launch_kernel("k_eig_lanczos", k_eig_lanczos, dim3(1), dim3(pl.threads), (size_t)pl.smem_bytes, st, false, p);
}

}  // namespace ttipm

using namespace ttipm;

extern "C" int ttipm_eig_force_cluster(int ctas) {
    g_eig_force_G = ctas;
    return 0;
}

static bool eig_contig(const int64_t* st, const int* dims, int nd) {
    long acc = 1;
    for (int i = nd - 1; i >= 0; --i) {
        if (dims[i] != 1 && st[i] != acc) return false;
        acc *= dims[i];
    }
    return true;
}

extern "C" int ttipm_eig_assemble(const ttipm_eig_op* op, int symmetrise, double* out, void* stream) {
    tt_stream_t st = (tt_stream_t)stream;
    const int l = op->l, s = op->s, L = op->L, n1 = op->n1;
    const bool two = op->A2 != nullptr;
    const int k = op->k, S = two ? op->S : op->k, n2 = two ? op->n2 : 1;
    if (l < 1 || L < 1 || n1 < 1 || n2 < 1 || s < 1 || S < 1 || k < 1) return fail(1, "eig_assemble: bad dims");
    const long m = (long)l * n1 * n2 * L;
    if (m * m > (1L << 31)) return fail(1, "eig_assemble: m=%ld too large for a dense projection", m);
    const int d1[3] = {l, s, l}, dA1[4] = {s, n1, n1, k}, dA2[4] = {k, n2, n2, S};
    if (!eig_contig(op->p1_strides, d1, 3) || !eig_contig(op->a1_strides, dA1, 4) ||
        (two && !eig_contig(op->a2_strides, dA2, 4)))
        return fail(1, "eig_assemble: P1, A1, A2 must be contiguous");
    for (int i = 0; i < 3; ++i)
        if (!fits_int(op->p2_strides[i])) return fail(1, "eig_assemble: stride overflow");
    // W = ((l r) x s) . (s x n1 n1 k) -> ((l r n1 n1) x k) . (k x n2 n2 S)
    const long nP = (long)l * l * s, nT1 = (long)l * l * n1 * n1 * k, nT2 = two ? (long)l * l * n1 * n1 * n2 * n2 * S : 0;
    double* scratch = nullptr;
#ifdef TTIPM_EMU
    scratch = (double*)malloc(sizeof(double) * (size_t)(nP + nT1 + nT2 + 8));
#else
    if (cudaMallocAsync((void**)&scratch, sizeof(double) * (size_t)(nP + nT1 + nT2 + 8), st) != cudaSuccess)
        return fail(91, "eig_assemble: device allocation failed");
#endif
    double* P1p = scratch;
    double* T1 = P1p + nP;
    double* T2 = T1 + nT1;
    int rc = 0;
    {
        int32_t dims[4] = {1, l, s, l}, perm[4] = {0, 1, 3, 2};
        rc = ttipm_permute4(op->P1, dims, perm, P1p, nullptr, 0, 1, stream);                  // (l, s, r) -> (l, r, s)
    }
    const long c1 = (long)n1 * n1 * k;
    if (!rc) rc = ttipm_gemm((int)((long)l * l), (int)c1, s, 1.0, P1p, s, 1, 0, op->A1, c1, 1, 0, 0.0, T1, c1, 1, 0, 1, stream);
    const double* W = T1;
    if (!rc && two) {
        const long c2 = (long)n2 * n2 * S;
        rc = ttipm_gemm((int)((long)l * l * n1 * n1), (int)c2, k, 1.0, T1, k, 1, 0, op->A2, c2, 1, 0, 0.0, T2, c2, 1, 0, 1, stream);
        W = T2;
    }
    if (!rc) {
        EigAsmParams p;
        p.W = W; p.P2 = op->P2;
        for (int i = 0; i < 3; ++i) p.p2s[i] = (int)op->p2_strides[i];
        p.l = l; p.S = S; p.L = L; p.n1 = n1; p.n2 = n2; p.symmetrise = symmetrise; p.out = out;
        const int bt = block_threads();
        long blocks = (m * m + bt - 1) / bt;
        const long cap = (long)dev_info().sms * 8;
        if (blocks > cap) blocks = cap;
        rc = launch_kernel("k_eig_assemble", k_eig_assemble, dim3((unsigned)blocks), dim3(bt), 0, st, false, p);
    }
#ifdef TTIPM_EMU
    free(scratch);
#else
    cudaFreeAsync(scratch, st);
#endif
    return rc;
}

extern "C" int64_t ttipm_eig_workspace(int m, int K) {
    EigPlan pl;
    if (eig_plan(pl, m, K)) return -1;
    return pl.ws_doubles + 2L * m * m + 2L * m;      // + Cholesky factor, reduced matrix and two vectors of the pencil form
}

extern "C" int ttipm_eig_lanczos(const double* A, double cA, const double* D, double cD, int m, const double* v0,
                                 int largest, int K, int max_cycles, double tol, double* x, double* out, double* ws,
                                 void* stream) {
    if (check_bound_device()) return 6;
    return eig_launch(A, cA, D, cD, m, v0, largest, K, max_cycles, tol, x, out, ws, (tt_stream_t)stream);
}

// largest eigenpair of (-D) x = lambda A x; out[0] = lambda, x normalised to unit Euclidean length; out[3] = 1 on
// success, 0 when A is not positive definite (the reference's eigsh(-D, M=A) raises there and the caller shrinks the step)
extern "C" int ttipm_eig_gen_largest(const double* A, const double* D, int m, const double* v0, int K, int max_cycles,
                                     double tol, double* x, double* out, double* ws, void* stream) {
    if (check_bound_device()) return 6;
    tt_stream_t st = (tt_stream_t)stream;
    EigPlan pl;
    int rc = eig_plan(pl, m, K);
    if (rc) return rc;
    double* Lc = ws + pl.ws_doubles;
    double* Cm = Lc + (long)m * m;
    double* y0 = Cm + (long)m * m;
    double* y = y0 + m;
    const size_t mm = (size_t)m * m * sizeof(double);
    if (dev_copy(Lc, A, mm, st) || dev_copy(Cm, D, mm, st)) return fail(5, "eig_gen: copy failed");
#ifdef TTIPM_EMU
    // in-place lower Cholesky factor of the row-major matrix
    for (int j = 0; j < m; ++j) {
        double d = Lc[(long)j * m + j];
        for (int k = 0; k < j; ++k) d -= Lc[(long)j * m + k] * Lc[(long)j * m + k];
        if (!(d > 0.0)) {
            out[3] = 0.0;
            return 0;
        }
        d = sqrt(d);
        Lc[(long)j * m + j] = d;
        for (int i = j + 1; i < m; ++i) {
            double v = Lc[(long)i * m + j];
            for (int k = 0; k < j; ++k) v -= Lc[(long)i * m + k] * Lc[(long)j * m + k];
            Lc[(long)i * m + j] = v / d;
        }
    }
    // C = L^-1 (-D) L^-T
    for (long i = 0; i < (long)m * m; ++i) Cm[i] = -Cm[i];
    for (int col = 0; col < m; ++col)                 // L^-1 from the left
        for (int i = 0; i < m; ++i) {
            double v = Cm[(long)i * m + col];
            for (int k = 0; k < i; ++k) v -= Lc[(long)i * m + k] * Cm[(long)k * m + col];
            Cm[(long)i * m + col] = v / Lc[(long)i * m + i];
        }
    for (int row = 0; row < m; ++row)                 // L^-T from the right: X L^T = B
        for (int j = 0; j < m; ++j) {
            double v = Cm[(long)row * m + j];
            for (int k = 0; k < j; ++k) v -= Cm[(long)row * m + k] * Lc[(long)j * m + k];
            Cm[(long)row * m + j] = v / Lc[(long)j * m + j];
        }
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < i; ++j) {
            const double v = 0.5 * (Cm[(long)i * m + j] + Cm[(long)j * m + i]);
            Cm[(long)i * m + j] = Cm[(long)j * m + i] = v;
        }
    const double* start = nullptr;
    if (v0) {                                          // y0 = L^T v0
        for (int i = 0; i < m; ++i) {
            double v = 0.0;
            for (int k = i; k < m; ++k) v += Lc[(long)k * m + i] * v0[k];
            y0[i] = v;
        }
        start = y0;
    }
    rc = eig_launch(Cm, 1.0, nullptr, 0.0, m, start, 1, K, max_cycles, tol, y, out, ws, st);
    if (rc) return rc;
    for (int i = m; i-- > 0;) {                        // x = L^-T y
        double v = y[i];
        for (int k = i + 1; k < m; ++k) v -= Lc[(long)k * m + i] * x[k];
        x[i] = v / Lc[(long)i * m + i];
    }
    double nn = 0.0;
    for (int i = 0; i < m; ++i) nn += x[i] * x[i];
    nn = 1.0 / sqrt(nn);
    for (int i = 0; i < m; ++i) x[i] *= nn;
    out[3] = 1.0;
    return 0;
#else
    cublasHandle_t blas = (cublasHandle_t)blas_handle(st);
    cusolverDnHandle_t sol = (cusolverDnHandle_t)solver_handle(st);
    if (!blas || !sol) return fail(93, "eig_gen: cuBLAS / cuSOLVER handle creation failed");
    int lwork = 0;
    cusolverDnDpotrf_bufferSize(sol, CUBLAS_FILL_MODE_UPPER, m, Lc, m, &lwork);
    double* work = nullptr;
    if (cudaMallocAsync((void**)&work, sizeof(double) * ((size_t)lwork + 2), st) != cudaSuccess)
        return fail(91, "eig_gen: device allocation failed");
    int* info = (int*)(work + lwork);
    int hinfo = 0;
    bool ok = cusolverDnDpotrf(sol, CUBLAS_FILL_MODE_UPPER, m, Lc, m, work, lwork, info) == CUSOLVER_STATUS_SUCCESS;
    if (ok) ok = cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
                 cudaStreamSynchronize(st) == cudaSuccess;
    cudaFreeAsync(work, st);
    if (!ok) return fail(93, "eig_gen: potrf failed");
    double flag[1] = {hinfo == 0 ? 1.0 : 0.0};
    if (hinfo != 0) {
        cudaMemcpyAsync(out + 3, flag, sizeof(double), cudaMemcpyHostToDevice, st);
        cudaStreamSynchronize(st);
        return 0;
    }
    // column-major view: the buffer holds U (upper) with A = U^T U, U = L^T.  C = U^-T (-D) U^-1
    const double minus_one = -1.0, one = 1.0;
    if (cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, m, &minus_one, Lc, m,
                    Cm, m) != CUBLAS_STATUS_SUCCESS ||
        cublasDtrsm(blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, m, m, &one, Lc, m, Cm,
                    m) != CUBLAS_STATUS_SUCCESS)
        return fail(93, "eig_gen: trsm failed");
    const double* start = nullptr;
    if (v0) {                                          // y0 = U v0
        if (dev_copy(y0, v0, sizeof(double) * m, st)) return fail(5, "eig_gen: copy failed");
        if (cublasDtrmv(blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, Lc, m, y0, 1) != CUBLAS_STATUS_SUCCESS)
            return fail(93, "eig_gen: trmv failed");
        start = y0;
    }
    rc = eig_launch(Cm, 1.0, nullptr, 0.0, m, start, 1, K, max_cycles, tol, y, out, ws, st);
    if (rc) return rc;
    if (dev_copy(x, y, sizeof(double) * m, st)) return fail(5, "eig_gen: copy failed");
    if (cublasDtrsv(blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, m, Lc, m, x, 1) != CUBLAS_STATUS_SUCCESS)
        return fail(93, "eig_gen: trsv failed");
    double nrm = 0.0;
    cublasSetPointerMode(blas, CUBLAS_POINTER_MODE_HOST);
    if (cublasDnrm2(blas, m, x, 1, &nrm) != CUBLAS_STATUS_SUCCESS || !(nrm > 0.0)) return fail(93, "eig_gen: nrm2 failed");
    const double inv = 1.0 / nrm;
    if (cublasDscal(blas, m, &inv, x, 1) != CUBLAS_STATUS_SUCCESS) return fail(93, "eig_gen: scal failed");
    cudaMemcpyAsync(out + 3, flag, sizeof(double), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    return 0;
#endif
}
