// Dense factorisations of core unfoldings: Householder QR and the "left" SVD (U, s, S*V^T).
//
//   * QR replaces scipy.linalg.qr / LAPACK geqrf+orgqr at reference cy_src/tt_ops_cy.pyx:147,
//     src/tt_als.py:358, :482.
//   * The SVD replaces scipy.linalg.svd / LAPACK gesvd, gesdd at cy_src/tt_ops_cy.pyx:205, :404, :418 and
//     src/tt_als.py:270, :331, :457.  Every caller on the hot path only needs U, s and S*V^T, so the kernel
//     returns exactly those: U is a product of Householder reflectors and plane rotations (orthonormal even
//     for zero singular values) and W = S V^T is never divided by s.
//
// ONE kernel (k_linalg) serves every shape.  A single large matrix runs as a persistent cooperative launch
// over several CTAs (grid.x > 1, grid barriers between phases); small or batched matrices use one CTA per
// matrix (grid.y = batch) and the same code with block barriers only.  Phases:
//   QR      panel Householder: the panel of TT_QR_PB columns is factored in shared memory (redundantly by
//           every CTA, so no broadcast is needed), trailing columns are owned by warps across the grid and
//           held in registers while the panel's reflectors are applied; one grid barrier per panel.
//   Q       reflectors are staged panel by panel in shared memory and applied to register-resident vectors
//           (unit vectors -> rows of Q^T, or [g; 0] -> rows of G Q^T), one vector per warp, no grid barrier.
//   Jacobi  QR-preconditioned BLOCK one-sided Jacobi on the rows of the triangular factor R (K x K):
//           tall A = Q1 R1, R1^T = Q2 R2, R2^T = Q3 R3 -> rotate rows of R3, accumulator starts from Q3^T (K x K),
//                                  U = Q1 (accumulator)^T, W = (rotated rows) Q2^T at the end
//                                  (each QR of the transpose grades the factor further: on the sweep's real
//                                  unfoldings rows of R1 need 27-31 sweeps, rows of R3 ~10; rows are 2K long
//                                  instead of K + M).  ttipm_linalg_tall_triple_qr(0) selects the single-QR form
//                                  A = Q R -> rotate rows of R, accumulator starts from Q^T.
//           wide A^T = Q1 R1, R1^T = Q2 R2 -> rotate rows of R2, accumulator starts from Q2^T,
//                                  W = (rotated rows) Q1^T at the end
//           (rows of an unpreconditioned wide matrix need 40-50 sweeps, rows of R need ~10).
//           A CTA owns a pair of row blocks per round, keeps their rows of R and of the accumulator in shared
//           memory and orthogonalises every cross pair (intra-block pairs once per sweep); block pairs follow
//           a round-robin tournament with one grid barrier per round (K / nb rounds per sweep instead of K).
//
// Blackwell specifics: a multi-CTA launch is ONE thread-block cluster (<= 16 CTAs), so every phase boundary is the
// hardware cluster barrier instead of an atomic counter in global memory; the rows [R | accumulator] of a Jacobi
// block are contiguous in the workspace and arrive in shared memory as one bulk asynchronous copy per block
// (cp.async.bulk + mbarrier), as do the columns of a reflector panel.
#include "api_util.h"

namespace ttipm {

#define TT_QR_PB 8
#define TT_LIN_HDR 72          // doubles of per-batch flags ahead of the batch workspaces
#define TT_LIN_REG 16          // register-resident vector length = 32 * TT_LIN_REG

struct LinParams {
    const double* A;
    long a_rs, a_cs, a_bs;
    int M, N;
    int mode;            // 0 = SVD: U (M x K), S (K), Wt (K x N)     1 = QR: U = Q (M x K), Wt = R (K x N)
    double* U;
    double* S;
    double* Wt;
    int* info;           // 16 ints per batch entry or NULL
    double* ws;          // [nbatch x TT_LIN_HDR doubles: sweep flags (ints 0..59), barrier (64..), big-rotation flags (80..139)][nbatch x ws_per]
    long ws_per;
    long oW1, oTau1, oW2, oTau2, oG, oJt, oSv;      // offsets inside one batch workspace
    long oW3, oTau3;     // third factor of the tall SVD
    int triple;          // tall SVD preconditioned by three QR factorisations (accumulator K x K)
    long oPg;            // >= 0: the reflector panel lives in the batch workspace (columns too long for shared memory)
    int M1, N1;          // first QR: M x N (tall, square, QR mode) or N x M (wide SVD, factors A^T)
    int Mj;              // accumulator row length
    int nb;              // Jacobi block rows
    int resident;        // single CTA with every row [R | accumulator] in shared memory
    double floor_factor; // rows below floor_factor * eps * ||R||_F are left alone by the Jacobi iteration (0 = off)
    int ldp;             // panel leading dimension in shared memory
    int oOrd, oSvS, oTaus, oP;   // shared-memory offsets (doubles)
    int oRows;           // shared-memory offset of the Jacobi rows (even)
    int ld1;             // leading dimension of the column-major working copy W1 (even)
    int ldk;             // leading dimension of the K x K factors W2, W3 (even)
    int ldg;             // row stride of the Jacobi rows [R (K) | accumulator (Mj) | pad], global and shared (even)
    int cluster;         // the grid is one thread-block cluster: hardware barrier between phases
    int early_exit;      // stop after a sweep whose rotations were all small (lin_rot_level)
    int nbatch;
};

TT_DEV long long lin_now_raw();

struct LinCtx {
    const LinParams& p;
    double* smem;
    unsigned* barrier;
    unsigned epoch;
    int lane, wid, nw, gw, GW;
    double* panel;       // TT_QR_PB reflector columns: shared memory, or global scratch for very tall matrices
    unsigned long long* mbar;    // byte-counting barrier of the bulk copies (shared memory)
    unsigned mphase;
    bool bulk;           // panel in shared memory: bulk copies apply
    long long tq[4];     // QR phase timers of this CTA (ns): panel load, panel factorisation, trailing update, barrier
    TT_DEVM LinCtx(const LinParams& pp, double* s, unsigned* bar) : p(pp), smem(s), barrier(bar), epoch(0) {
        panel = s + pp.oP;
        mbar = (unsigned long long*)(s + 38);
        mphase = 0;
        bulk = pp.oPg < 0;
        tq[0] = tq[1] = tq[2] = tq[3] = 0;
        lane = threadIdx.x & 31;
        wid = threadIdx.x >> 5;
        nw = blockDim.x >> 5;
        gw = blockIdx.x * nw + wid;
        GW = gridDim.x * nw;
    }
    TT_DEVM void sync() {
        if (p.cluster) cluster_sync_all();
        else grid_sync(barrier, epoch);
    }
    // phase timers only when the caller asked for the info record: a %globaltimer read is not free and the Jacobi loop
    // would take seven of them per round
    TT_DEVM long long now() const { return p.info ? lin_now_raw() : 0; }
};

// the kernel's dynamic shared memory, re-derived where it is used heavily so that the compiler keeps the
// shared address space (LDS / STS instead of generic loads)
TT_DEV double* lin_smem() {
    TT_SMEM_DECL(lin_smem_raw);
    return (double*)lin_smem_raw;
}

#ifndef TTIPM_EMU
TT_DEV long long lin_now_raw() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
TT_DEV double lin_rsqrt(double x) { return rsqrt(x); }
#else
TT_DEV long long lin_now_raw() { return 0; }
TT_DEV double lin_rsqrt(double x) { return 1.0 / sqrt(x); }
#endif

// Jacobi rotation that orthogonalises two rows with squared norms saa, sbb and inner product sab != 0:
//   a' = cs a - sn b,  b' = sn a + cs b,  |theta| <= pi / 4.
// With d = sbb - saa and h = sqrt(d^2 + 4 sab^2):  cos 2theta = |d| / h,  sin 2theta = sign(d) 2 sab / h, so
//   cs^2 = (1 + |d| / h) / 2,  sn = sign(d) sab / (h cs)
// -- two reciprocal square roots in sequence instead of sqrt, division and reciprocal square root (the pair step of the
// block Jacobi iteration is a dependent chain; this shortens its scalar part).
TT_DEV void lin_rotation(double saa, double sbb, double sab, double& cs, double& sn) {
    const double d = sbb - saa;
    const double rh = lin_rsqrt(d * d + 4.0 * sab * sab);
    const double c2 = 0.5 + 0.5 * fabs(d) * rh;
    const double rc = lin_rsqrt(c2);
    cs = c2 * rc;
    sn = copysign(sab * rh * rc, d * sab);
}

// x <- (I - tj v v^T) x for a vector held in registers with the fixed mapping element i <-> (lane, q = i / 32);
// v[j] = 1 implicit, v[i] given for j < i < len (v may point before its first valid element), zero above j.
// NQ = register chunks actually in use (32 * NQ >= len).
template <int NQ>
TT_DEV void lin_reflect_reg(double (&reg)[NQ], const double* v, int j, int len, double tj, int lane) {
    double d2[2] = {0.0, 0.0};
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int i = lane + 32 * q;
        const double vv = (i > j && i < len) ? v[i] : (i == j ? 1.0 : 0.0);
        d2[q & 1] += vv * reg[q];
    }
    const double d = warp_sum(d2[0] + d2[1]) * tj;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int i = lane + 32 * q;
        const double vv = (i > j && i < len) ? v[i] : (i == j ? 1.0 : 0.0);
        reg[q] -= d * vv;
    }
}
// same for a vector in memory (len > 32 * TT_LIN_REG); element i at x[i * xs]; a thread only re-reads its own writes
TT_DEV void lin_reflect_mem(double* x, long xs, const double* v, int j, int len, double tj, int lane) {
    double d = 0.0;
    for (int i = lane; i < len; i += 32) {
        if (i > j) d += v[i] * ld_cg(x + (long)i * xs);
        else if (i == j) d += ld_cg(x + (long)i * xs);
    }
    d = warp_sum(d) * tj;
    for (int i = lane; i < len; i += 32) {
        if (i > j) x[(long)i * xs] = ld_cg(x + (long)i * xs) - d * v[i];
        else if (i == j) x[(long)i * xs] = ld_cg(x + (long)i * xs) - d;
    }
}

// stage reflectors [j0, j0 + pw) of W (column-major, leading dimension ld, rows j0..Mq-1 only) and their tau into the
// panel P.  Shared-memory panels are filled by one bulk asynchronous copy per column; ends with the data visible to
// every thread of the CTA.  Callers put a block barrier before the call (previous readers of P are done).
TT_DEV void lin_load_panel(LinCtx& c, const double* W, const double* tau, int ld, int Mq, int j0, int pw, bool with_tau,
                           double* P, double* taus, int ldp) {
    const int rows = Mq - j0;
    if (with_tau && (int)threadIdx.x < pw) taus[threadIdx.x] = ld_cg(tau + j0 + threadIdx.x);
    if (c.bulk) {
        if (threadIdx.x == 0) {
            const unsigned bytes = (unsigned)((rows + 1) & ~1) * 8u;       // ld and ldp are even: the pad element exists
            fence_proxy_async();
            mbar_expect_tx(c.mbar, bytes * (unsigned)pw);
            for (int j = 0; j < pw; ++j) bulk_g2s(P + j * ldp, W + (long)(j0 + j) * ld + j0, bytes, c.mbar);
        }
        mbar_wait(c.mbar, c.mphase);
        c.mphase ^= 1u;
        __syncthreads();                                                    // taus
        return;
    }
    for (int i = threadIdx.x; i < pw * rows; i += blockDim.x) {
        const int j = i / rows, q = i % rows;
        P[j * ldp + q] = ld_cg(W + (long)(j0 + j) * ld + j0 + q);
    }
    __syncthreads();
}

// trailing column x (global, rows elements from the panel's first row) <- H_{pw-1} ... H_0 x, held in registers
template <int NQ>
TT_DEV void lin_trailing_col(double* x, int rows, const double* P, int ldp, const double* taus, int pw, int lane) {
    double reg[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int i = lane + 32 * q;
        reg[q] = i < rows ? ld_cg(x + i) : 0.0;
    }
    for (int j = 0; j < pw; ++j)
        if (taus[j] != 0.0) lin_reflect_reg<NQ>(reg, P + j * ldp, j, rows, taus[j], lane);
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int i = lane + 32 * q;
        if (i < rows) x[i] = reg[q];
    }
}

// W (Mq x Nq column-major, leading dimension ld, global) -> R in the upper triangle, reflector tails below,
// tau[min(Mq, Nq)].  Ends with a grid barrier.
TT_DEV void lin_qr_factor(LinCtx& c, double* W, double* tau, int Mq, int Nq, int ld) {
    const int K = imin(Mq, Nq), ldp = c.p.ldp, lane = c.lane, wid = c.wid, nw = c.nw;
    const bool psh = c.bulk;                       // panel in shared memory
    double* taus = lin_smem() + c.p.oTaus;
    double* scl = taus + TT_QR_PB;                 // 1 / (alpha - beta) of every panel column
    double* bet = scl + TT_QR_PB;                  // its new diagonal entry
    double* P = psh ? lin_smem() + c.p.oP : c.panel;
    for (int p0 = 0; p0 < K; p0 += TT_QR_PB) {
        const int pw = imin(TT_QR_PB, K - p0), rows = Mq - p0;
        __syncthreads();
        const long long tq0 = c.now();
        lin_load_panel(c, W, nullptr, ld, Mq, p0, pw, false, P, taus, ldp);
        const long long tq1 = c.now();
        // Panel factorisation: every warp derives reflector j from column j redundantly (no broadcast, no serial
        // section), warp w then updates panel columns j + 1 + w, ...; the tails stay unscaled until the panel is done,
        // so a column step is one block barrier.
        for (int j = 0; j < pw; ++j) {
            const double* col = P + j * ldp;
            double s4[4] = {0.0, 0.0, 0.0, 0.0};
            int i = j + 1 + lane;
            for (; i + 96 < rows; i += 128) {
#pragma unroll
                for (int u = 0; u < 4; ++u) s4[u] += col[i + 32 * u] * col[i + 32 * u];
            }
            for (int u = 0; i < rows; i += 32, ++u) s4[u] += col[i] * col[i];
            const double s = warp_sum((s4[0] + s4[1]) + (s4[2] + s4[3]));
            const double alpha = col[j];
            double tj = 0.0, scale = 0.0, beta = alpha;
            if (s != 0.0) {
                // beta = -sign(alpha) * nrm, tau = (beta - alpha) / beta, scale = 1 / (alpha - beta) with one reciprocal
                // square root and one division:  1 / beta = -sign(alpha) * rn,  alpha - beta = sign(alpha) (|alpha| + nrm)
                const double q2 = alpha * alpha + s, rn = lin_rsqrt(q2), nrm = q2 * rn, an = fabs(alpha) + nrm;
                beta = -copysign(nrm, alpha);
                tj = an * rn;
                scale = copysign(1.0 / an, alpha);
            }
            if (tj != 0.0) {
                for (int cc = j + 1 + wid; cc < pw; cc += nw) {
                    double* x = P + cc * ldp;
                    double d4[4] = {0.0, 0.0, 0.0, 0.0};
                    int i2 = j + 1 + lane;
                    for (; i2 + 96 < rows; i2 += 128) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) d4[u] += col[i2 + 32 * u] * x[i2 + 32 * u];
                    }
                    for (int u = 0; i2 < rows; i2 += 32, ++u) d4[u] += col[i2] * x[i2];
                    double d = (d4[0] + d4[1]) + (d4[2] + d4[3]);
                    d = (warp_sum(d) * scale + x[j]) * tj;
                    __syncwarp();                              // every lane has read x[j]
                    if (lane == 0) x[j] -= d;
                    const double ds = d * scale;
                    for (int i = j + 1 + lane; i < rows; i += 32) x[i] -= ds * col[i];
                }
            }
            if (wid == nw - 1 && lane == 0) {
                taus[j] = tj;
                scl[j] = scale;
                bet[j] = beta;
            }
            __syncthreads();
        }
        for (int i = threadIdx.x; i < pw * rows; i += blockDim.x) {
            const int j = i / rows, q = i % rows;
            if (q > j) P[j * ldp + q] *= scl[j];
            else if (q == j) P[j * ldp + j] = bet[j];
        }
        __syncthreads();
        if (blockIdx.x == 0) {
            for (int i = threadIdx.x; i < pw * rows; i += blockDim.x) {
                const int j = i / rows, q = i % rows;
                W[(long)(p0 + j) * ld + p0 + q] = P[j * ldp + q];
            }
            if ((int)threadIdx.x < pw) tau[p0 + threadIdx.x] = taus[threadIdx.x];
        }
        const long long tq2 = c.now();
        for (int cc = p0 + pw + c.gw; cc < Nq; cc += c.GW) {
            double* x = W + (long)cc * ld + p0;
            if (rows <= 128) lin_trailing_col<4>(x, rows, P, ldp, taus, pw, lane);
            else if (rows <= 256) lin_trailing_col<8>(x, rows, P, ldp, taus, pw, lane);
            else if (rows <= 32 * TT_LIN_REG) lin_trailing_col<TT_LIN_REG>(x, rows, P, ldp, taus, pw, lane);
            else {
                for (int j = 0; j < pw; ++j)
                    if (taus[j] != 0.0) lin_reflect_mem(x, 1, P + j * ldp, j, rows, taus[j], lane);
            }
        }
        const long long tq3 = c.now();
        fence_proxy_async();                   // columns written here are read by the next panel's bulk copies
        c.sync();
        c.tq[0] += tq1 - tq0;
        c.tq[1] += tq2 - tq1;
        c.tq[2] += tq3 - tq2;
        c.tq[3] += c.now() - tq3;
    }
}

// One vector per warp task t < ntask:  x_t <- H_0 H_1 ... H_{Kq-1} x_t   (= Q x_t, reflectors of a factored W).
//   init 0: x_t = e_t (only H_j with j <= t act)            -> column t of Q
//   init 1: x_t = [src[ord[t] * src_rs + 0:Kq] ; 0]          -> Q [g; 0]
// Result element i of task t goes to dst[t * row_stride + i * elem_stride].  W, tau, src must be globally visible.
template <int NQ>
TT_DEV void lin_apply_q_t(LinCtx& c, const double* W, const double* tau, int ld, int Mq, int Kq, int ntask, int init,
                          const double* src, long src_rs, const int* ord, double* dst, long row_stride, long elem_stride) {
    const int ldp = c.p.ldp, lane = c.lane;
    double* taus = lin_smem() + c.p.oTaus;
    double* P = c.bulk ? lin_smem() + c.p.oP : c.panel;
    const bool in_regs = NQ > 0;
    constexpr int NR = NQ > 0 ? NQ : 1;
    for (int base = 0; base < ntask; base += c.GW) {
        const int first = base + blockIdx.x * c.nw;
        if (first >= ntask) break;                              // uniform over the CTA
        const int t = first + c.wid;
        const bool active = t < ntask;
        const int tmax = imin(ntask - 1, first + c.nw - 1);
        const int jtop = init == 0 ? imin(tmax, Kq - 1) : Kq - 1;
        double reg[NR];
        double* x = dst + (long)t * row_stride;
        if (active) {
            if (in_regs) {
#pragma unroll
                for (int q = 0; q < NR; ++q) {
                    const int i = lane + 32 * q;
                    if (init == 0) reg[q] = i == t ? 1.0 : 0.0;
                    else reg[q] = i < Kq ? ld_cg(src + (long)ord[t] * src_rs + i) : 0.0;
                }
            } else {
                for (int i = lane; i < Mq; i += 32) {
                    double v;
                    if (init == 0) v = i == t ? 1.0 : 0.0;
                    else v = i < Kq ? ld_cg(src + (long)ord[t] * src_rs + i) : 0.0;
                    x[(long)i * elem_stride] = v;
                }
            }
        }
        for (int j0 = (jtop / TT_QR_PB) * TT_QR_PB; j0 >= 0; j0 -= TT_QR_PB) {
            const int pw = imin(TT_QR_PB, Kq - j0);
            __syncthreads();
            lin_load_panel(c, W, tau, ld, Mq, j0, pw, true, P, taus, ldp);
            if (!active) continue;
            for (int jj = pw - 1; jj >= 0; --jj) {
                const int j = j0 + jj;
                if (init == 0 && j > t) continue;
                const double tj = taus[jj];
                if (tj == 0.0) continue;
                const double* v = P + jj * ldp - j0;            // v[i] valid for j < i < Mq
                if (in_regs) lin_reflect_reg<NR>(reg, v, j, Mq, tj, lane);
                else lin_reflect_mem(x, elem_stride, v, j, Mq, tj, lane);
            }
        }
        if (active && in_regs) {
#pragma unroll
            for (int q = 0; q < NR; ++q) {
                const int i = lane + 32 * q;
                if (i < Mq) x[(long)i * elem_stride] = reg[q];
            }
        }
    }
}
TT_DEV void lin_apply_q(LinCtx& c, const double* W, const double* tau, int ld, int Mq, int Kq, int ntask, int init,
                        const double* src, long src_rs, const int* ord, double* dst, long row_stride, long elem_stride) {
    if (Mq <= 128) lin_apply_q_t<4>(c, W, tau, ld, Mq, Kq, ntask, init, src, src_rs, ord, dst, row_stride, elem_stride);
    else if (Mq <= 256) lin_apply_q_t<8>(c, W, tau, ld, Mq, Kq, ntask, init, src, src_rs, ord, dst, row_stride, elem_stride);
    else if (Mq <= 32 * TT_LIN_REG)
        lin_apply_q_t<TT_LIN_REG>(c, W, tau, ld, Mq, Kq, ntask, init, src, src_rs, ord, dst, row_stride, elem_stride);
    else lin_apply_q_t<0>(c, W, tau, ld, Mq, Kq, ntask, init, src, src_rs, ord, dst, row_stride, elem_stride);
}

// orthogonalise two rows held in shared memory: [row of R (K) | accumulator row (Mj)], total length Ls
//
// `floor2` = (eps * ||R||_F)^2: a row whose norm has dropped below the rounding level of the largest singular value is
// numerically zero -- it is left alone (the unfoldings of the AMEn sweep are strongly graded: at maxcut_13 half of
// the singular values sit 1e-14 below the largest, and orthogonalising those noise rows against each other to full
// RELATIVE accuracy cost 27-31 sweeps instead of the ~10 the significant part needs).  U stays orthonormal and
// U * W = A holds regardless (only rotations are applied); singular values above the floor are unaffected, the ones
// below it carry an absolute error of eps * sigma_max like LAPACK's.
// Rotation level of a pair step: 0 = none, 1 = a "small" rotation, 3 = a rotation that can still move other pairs.
// A sweep in which every rotation was small -- inner product below 1e-11 |a| |b| AND sine below 1e-6 -- leaves every
// pair it touched orthogonal to <= K * 1e-17 < tol (each later rotation perturbs an inner product by its sine times
// another inner product of the same size), so the iteration stops without the extra all-skip sweep that would only
// confirm it (one sweep of ~9 on the sweep's unfoldings).
#define TT_JAC_SMALL_RATIO2 1e-22
#define TT_JAC_SMALL_SIN2 1e-12
TT_DEV int lin_rot_level(double saa, double sbb, double sab, double sn) {
    return (sab * sab > TT_JAC_SMALL_RATIO2 * saa * sbb || sn * sn > TT_JAC_SMALL_SIN2) ? 3 : 1;
}

TT_DEV int lin_jacobi_pair(double* ra, double* rb, int K, int Ls, double tol2, double floor2, int lane) {
    double saa = 0.0, sbb = 0.0, sab = 0.0, taa = 0.0, tbb = 0.0, tab = 0.0;
    int i0 = lane;
    for (; i0 + 32 < K; i0 += 64) {
        const double x = ra[i0], y = rb[i0], x2 = ra[i0 + 32], y2 = rb[i0 + 32];
        saa += x * x;
        sbb += y * y;
        sab += x * y;
        taa += x2 * x2;
        tbb += y2 * y2;
        tab += x2 * y2;
    }
    if (i0 < K) {
        const double x = ra[i0], y = rb[i0];
        saa += x * x;
        sbb += y * y;
        sab += x * y;
    }
    saa += taa;
    sbb += tbb;
    sab += tab;
#ifndef TTIPM_EMU
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {           // three interleaved butterflies
        saa += __shfl_xor_sync(0xffffffffu, saa, o);
        sbb += __shfl_xor_sync(0xffffffffu, sbb, o);
        sab += __shfl_xor_sync(0xffffffffu, sab, o);
    }
#else
    saa = warp_sum(saa);
    sbb = warp_sum(sbb);
    sab = warp_sum(sab);
#endif
    if (!(sab * sab > tol2 * saa * sbb) || fmin(saa, sbb) <= floor2) return 0;
    // tan of the rotation angle: zeta = (sbb - saa) / (2 sab), tg = sign(zeta) / (|zeta| + sqrt(1 + zeta^2))
    double cs, sn;
    lin_rotation(saa, sbb, sab, cs, sn);
    const double c1 = cs, s1 = -sn, c2 = sn, s2 = cs;
#pragma unroll 4
    for (int i = lane; i < Ls; i += 32) {
        const double x = ra[i], y = rb[i];
        ra[i] = c1 * x + s1 * y;
        rb[i] = c2 * x + s2 * y;
    }
    return lin_rot_level(saa, sbb, sab, sn);
}

// Cross pairs of two row blocks held in shared memory (block a: slots 0..nb-1, block b: slots nb..2nb-1).  Warp i
// keeps row i of block a in REGISTERS for all nb steps of the round and meets row (i + u) % nb of block b at step u, so
// the shared-memory traffic of a step is one read and one write of the b row (the smem-only form moves 2.5x as much
// and was bound by shared-memory bandwidth).  NQ = register chunks per row (32 * NQ >= Ls).
template <int NQ, bool RB_REGS>
TT_DEV int lin_cross_pairs_reg(double* rowsS, int nb, int na, int nbb, int K, int Ls, int ldg, double tol2, double floor2,
                                int lane, int wid, int nw) {
    int rot = 0;
    for (int i0 = 0; i0 < nb; i0 += nw) {                 // one pass when the CTA has at least nb warps
        const int i = i0 + wid;
        const bool have = i < na;
        double ra[NQ];
        if (have) {
            const double* rap = rowsS + (long)i * ldg;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int e = lane + 32 * q;
                ra[q] = e < Ls ? rap[e] : 0.0;
            }
        }
        for (int u = 0; u < nb; ++u) {
            const int jb = (i + u) % nb;
            if (have && jb < nbb) {
                double* rbp = rowsS + (long)(nb + jb) * ldg;
                double rb[RB_REGS ? NQ : 1];                  // long rows: the b row is re-read for the rotation
                // fp64 FMA latency on B200 is ~23 cycles (tools/lat_bench.cu): split every sum over independent
                // accumulators so the dot products are not one dependent chain
                double paa[2] = {0.0, 0.0}, pbb[2] = {0.0, 0.0}, pab[2] = {0.0, 0.0};
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int e = lane + 32 * q;
                    if (RB_REGS) {
                        rb[q] = e < Ls ? rbp[e] : 0.0;
                        if (e < K) {
                            paa[q & 1] += ra[q] * ra[q];
                            pbb[q & 1] += rb[q] * rb[q];
                            pab[q & 1] += ra[q] * rb[q];
                        }
                    } else if (e < K) {
                        const double y = rbp[e];
                        paa[q & 1] += ra[q] * ra[q];
                        pbb[q & 1] += y * y;
                        pab[q & 1] += ra[q] * y;
                    }
                }
                double saa = paa[0] + paa[1], sbb = pbb[0] + pbb[1], sab = pab[0] + pab[1];
#ifndef TTIPM_EMU
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    saa += __shfl_xor_sync(0xffffffffu, saa, o);
                    sbb += __shfl_xor_sync(0xffffffffu, sbb, o);
                    sab += __shfl_xor_sync(0xffffffffu, sab, o);
                }
#else
                saa = warp_sum(saa);
                sbb = warp_sum(sbb);
                sab = warp_sum(sab);
#endif
                if (sab * sab > tol2 * saa * sbb && fmin(saa, sbb) > floor2) {
                    double cs, sn;
                    lin_rotation(saa, sbb, sab, cs, sn);
#pragma unroll
                    for (int q = 0; q < NQ; ++q) {
                        const int e = lane + 32 * q;
                        const double x = ra[q], y = RB_REGS ? rb[q] : (e < Ls ? rbp[e] : 0.0);
                        ra[q] = cs * x - sn * y;
                        if (e < Ls) rbp[e] = sn * x + cs * y;
                    }
                    rot |= lin_rot_level(saa, sbb, sab, sn);
                }
            }
            __syncthreads();
        }
        if (have) {
            double* rap = rowsS + (long)i * ldg;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int e = lane + 32 * q;
                if (e < Ls) rap[e] = ra[q];
            }
        }
    }
    return rot;
}

// (eps * ||G||_F)^2 of the K x K matrix G (row stride ldg, global memory), computed redundantly by every CTA
TT_DEV double lin_noise_floor2(LinCtx& c, const double* G, int K, int ldg) {
    if (c.p.floor_factor == 0.0) return 0.0;
    double s = 0.0;
    for (long i = threadIdx.x; i < (long)K * K; i += blockDim.x) {
        const double v = ld_cg(G + (i / K) * ldg + i % K);
        s += v * v;
    }
    __syncthreads();
    s = block_sum(s, c.smem);
    const double e = 2.220446049250313e-16 * c.p.floor_factor;
    return e * e * s;
}

// block one-sided Jacobi on the K rows [G | Jt | pad] of GJ (row stride ldg, row-major, global memory).
// Ends with a grid barrier (all rows globally visible).  Returns the number of sweeps.
template <int NT>
TT_DEV int lin_jacobi(LinCtx& c, double* GJ, int K, int Mj, int* flags, long long* tm) {
    const int nb = c.p.nb, ldg = c.p.ldg, Ls = K + Mj, lane = c.lane, wid = c.wid, nw = c.nw;
    const int nblk = (K + nb - 1) / nb, nbe = nblk + (nblk & 1), npairs = nbe / 2, rounds = nbe - 1;
    const double tol = 2.220446049250313e-16 * sqrt((double)K), tol2 = tol * tol;
    int* rotated = (int*)(c.smem + 36);
    double* rowsS = lin_smem() + c.p.oRows;
    int sweeps = 0;
    if (K < 2) return 0;
    const double floor2 = lin_noise_floor2(c, GJ, K, ldg);
    for (; sweeps < 60; ++sweeps) {
        for (int t = 0; t < rounds; ++t) {
            for (int pi = blockIdx.x; pi < npairs; pi += gridDim.x) {
                int a, b;
                if (pi == 0) {
                    a = nbe - 1;
                    b = t;
                } else {
                    a = (t + pi) % (nbe - 1);
                    b = (t - pi + nbe - 1) % (nbe - 1);
                }
                if (a > b) { const int q = a; a = b; b = q; }       // a < b, only b can be the phantom block
                const bool bvalid = b < nblk;
                if (!bvalid && t != 0) continue;
                const int na = imin(nb, K - a * nb), nbb = bvalid ? imin(nb, K - b * nb) : 0;
                const long long t0 = c.now();
                __syncthreads();
                // the rows of a block are contiguous: one bulk copy per block (slot q < nb -> block a, nb + q -> block b)
                if (threadIdx.x == 0) {
                    rotated[0] = 0;
                    rotated[1] = 0;
                    const unsigned ba = (unsigned)(na * ldg) * 8u, bb = (unsigned)(nbb * ldg) * 8u;
                    fence_proxy_async();
                    mbar_expect_tx(c.mbar, ba + bb);
                    bulk_g2s(rowsS, GJ + (long)a * nb * ldg, ba, c.mbar);
                    if (bb) bulk_g2s(rowsS + (long)nb * ldg, GJ + (long)b * nb * ldg, bb, c.mbar);
                }
                mbar_wait(c.mbar, c.mphase);
                c.mphase ^= 1u;
                const long long t1 = c.now();
                int rot = 0;
                if (t == 0 && nb > 1) {
                    // intra-block pairs of both blocks: round-robin over the nb (even) slots of each block
                    const int half = nb / 2;
                    for (int u = 0; u < nb - 1; ++u) {
                        for (int task = wid; task < 2 * half; task += nw) {
                            const int sel = task / half, q = task % half;
                            const int cnt = sel == 0 ? na : nbb;
                            int x, y;
                            if (q == 0) {
                                x = nb - 1;
                                y = u;
                            } else {
                                x = (u + q) % (nb - 1);
                                y = (u - q + nb - 1) % (nb - 1);
                            }
                            if (x >= cnt || y >= cnt) continue;
                            if (x > y) { const int z = x; x = y; y = z; }
                            rot |= lin_jacobi_pair(rowsS + (long)(sel * nb + x) * ldg, rowsS + (long)(sel * nb + y) * ldg, K, Ls,
                                                   tol2, floor2, lane);
                        }
                        __syncthreads();
                    }
                }
                if (bvalid) {
                    if (Ls <= 128) rot |= lin_cross_pairs_reg<4, true>(rowsS, nb, na, nbb, K, Ls, ldg, tol2, floor2, lane, wid, nw);
                    else if (Ls <= 256) rot |= lin_cross_pairs_reg<8, true>(rowsS, nb, na, nbb, K, Ls, ldg, tol2, floor2, lane, wid, nw);
                    else if (Ls <= 384) rot |= lin_cross_pairs_reg<12, true>(rowsS, nb, na, nbb, K, Ls, ldg, tol2, floor2, lane, wid, nw);
                    else if (NT <= 384 && Ls <= 512) rot |= lin_cross_pairs_reg<NT <= 384 ? 16 : 1, true>(rowsS, nb, na, nbb, K, Ls, ldg, tol2, floor2, lane, wid, nw);
                    else if (NT <= 384 && Ls <= 704) rot |= lin_cross_pairs_reg<NT <= 384 ? 22 : 1, false>(rowsS, nb, na, nbb, K, Ls, ldg, tol2, floor2, lane, wid, nw);
                    else {
                        for (int u = 0; u < nb; ++u) {
                            for (int i = wid; i < nb; i += nw) {
                                const int jb = (i + u) % nb;
                                if (i >= na || jb >= nbb) continue;
                                rot |= lin_jacobi_pair(rowsS + (long)i * ldg, rowsS + (long)(nb + jb) * ldg, K, Ls, tol2, floor2, lane);
                            }
                            __syncthreads();
                        }
                    }
                }
                if (rot && lane == 0) rotated[0] = 1;
                if ((rot & 2) && lane == 0) rotated[1] = 1;
                __syncthreads();
                const long long t2 = c.now();
                if (rotated[0]) {
                    double* ga = GJ + (long)a * nb * ldg;
                    double* gb = GJ + (long)b * nb * ldg;
                    for (int i = threadIdx.x; i < na * ldg; i += blockDim.x) ga[i] = rowsS[i];
                    for (int i = threadIdx.x; i < nbb * ldg; i += blockDim.x) gb[i] = rowsS[(long)nb * ldg + i];
                    if (threadIdx.x == 0) {
                        flags[sweeps] = 1;
                        if (rotated[1]) flags[80 + sweeps] = 1;
                    }
                }
                tm[0] += t1 - t0;
                tm[1] += t2 - t1;
                tm[2] += c.now() - t2;
            }
            const long long t3 = c.now();
            fence_proxy_async();               // rows written through the generic proxy are read by bulk copies next round
            const long long t4 = c.now();
            c.sync();
            tm[2] += t4 - t3;
            tm[3] += c.now() - t4;
        }
        // done: nothing rotated, or only small rotations (see lin_rot_level)
        if (ld_cg_i(&flags[sweeps]) == 0 || (c.p.early_exit && ld_cg_i(&flags[80 + sweeps]) == 0)) {
            ++sweeps;
            break;
        }
    }
    return sweeps;
}

// Single-CTA variant for unfoldings whose rows [R | accumulator] fit in shared memory (K * ldg doubles): every
// row stays resident for the whole iteration, each warp owns one row pair of a round-robin step, and a step costs one
// block barrier.  The rows are read once (one bulk copy) and written once.
TT_DEV int lin_jacobi_resident(LinCtx& c, double* GJ, int K, int Mj) {
    const int ldg = c.p.ldg, Ls = K + Mj, lane = c.lane, wid = c.wid, nw = c.nw;
    const int Ke = K + (K & 1), half = Ke / 2;
    const double tol = 2.220446049250313e-16 * sqrt((double)K), tol2 = tol * tol;
    int* rotated = (int*)(c.smem + 36);
    double* rowsS = lin_smem() + c.p.oRows;
    if (K < 2) return 0;
    const double floor2 = lin_noise_floor2(c, GJ, K, ldg);
    __syncthreads();
    if (threadIdx.x == 0) {
        fence_proxy_async();
        mbar_expect_tx(c.mbar, (unsigned)(K * ldg) * 8u);
        bulk_g2s(rowsS, GJ, (unsigned)(K * ldg) * 8u, c.mbar);
    }
    mbar_wait(c.mbar, c.mphase);
    c.mphase ^= 1u;
    int sweeps = 0;
    for (; sweeps < 60;) {
        __syncthreads();
        if (threadIdx.x == 0) {
            rotated[0] = 0;
            rotated[1] = 0;
        }
        __syncthreads();
        int rot = 0;
        for (int u = 0; u < Ke - 1; ++u) {
            for (int q = wid; q < half; q += nw) {
                int x, y;
                if (q == 0) {
                    x = Ke - 1;
                    y = u;
                } else {
                    x = (u + q) % (Ke - 1);
                    y = (u - q + Ke - 1) % (Ke - 1);
                }
                if (x >= K || y >= K) continue;
                if (x > y) { const int z = x; x = y; y = z; }
                rot |= lin_jacobi_pair(rowsS + (long)x * ldg, rowsS + (long)y * ldg, K, Ls, tol2, floor2, lane);
            }
            __syncthreads();
        }
        if (rot && lane == 0) rotated[0] = 1;
        if ((rot & 2) && lane == 0) rotated[1] = 1;
        __syncthreads();
        ++sweeps;
        if (rotated[0] == 0 || (c.p.early_exit && rotated[1] == 0)) break;
    }
    for (int i = threadIdx.x; i < K * ldg; i += blockDim.x) GJ[i] = rowsS[i];
    c.sync();
    return sweeps;
}

// grid = (CTAs per matrix, batch)
template <int NT>
TT_GLOBAL void __launch_bounds__(NT) k_linalg(const LinParams p) {
    pdl_entry();
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    const int batch = blockIdx.y;
    int* flags = (int*)(p.ws + (long)batch * TT_LIN_HDR);
    LinCtx c(p, smem, (unsigned*)(flags + 64));
    double* ws = p.ws + (long)p.nbatch * TT_LIN_HDR + (long)batch * p.ws_per;
    if (p.oPg >= 0) c.panel = ws + p.oPg;
    const double* A = p.A + (long)batch * p.a_bs;
    const int M = p.M, N = p.N, K = imin(M, N), M1 = p.M1, Mj = p.Mj, ld1 = p.ld1, ldk = p.ldk, ldg = p.ldg;
    const bool wide = p.mode == 0 && M < N;
    double* W1 = ws + p.oW1;
    double* tau1 = ws + p.oTau1;
    double* W2 = ws + p.oW2;
    double* tau2 = ws + p.oTau2;
    double* G = ws + p.oG;            // rows [G (K) | Jt (Mj) | pad], stride ldg
    double* Jt = G + K;
    double* sv = ws + p.oSv;
    double* U = p.U + (long)batch * M * K;
    double* Wt = p.Wt + (long)batch * K * N;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x, gth = (long)gridDim.x * blockDim.x;
    const long long t_start = c.now();
    long long tm[4] = {0, 0, 0, 0};
    if (threadIdx.x == 0) mbar_init(c.mbar, 1);
    // the sweep flags are zeroed here instead of by a memset node ahead of the launch (a memset between two kernels breaks
    // the programmatic dependent launch chain); they are first read after several phase barriers.  The atomic grid barrier's
    // counter (flags + 64 .. 79, cooperative non-cluster launches only) is still cleared by the host.
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < 2 * TT_LIN_HDR; i += blockDim.x)
            if (i < 64 || i >= 80) flags[i] = 0;
    __syncthreads();

    // working copy, column-major: A (M x N), or A^T (N x M) for a wide SVD
    for (long i = gtid; i < (long)M * N; i += gth) {
        const long row = i / N, col = i % N;
        const double v = A[row * p.a_rs + col * p.a_cs];
        if (!wide) W1[row + col * ld1] = v;
        else W1[col + row * ld1] = v;
    }
    fence_proxy_async();
    c.sync();
    lin_qr_factor(c, W1, tau1, M1, p.N1, ld1);
    if (p.mode == 1) {
        for (long i = gtid; i < (long)K * N; i += gth) {
            const long row = i / N, col = i % N;
            Wt[i] = row <= col ? ld_cg(W1 + row + col * ld1) : 0.0;
        }
        fence_proxy_async();
        c.sync();                                   // the factored panels written by CTA 0 are read by bulk copies
        lin_apply_q(c, W1, tau1, ld1, M, K, K, 0, nullptr, 0, nullptr, U, 1, K);   // column t of Q -> Q (M x K) row-major
        return;
    }
    const double* Wq = W1;      // factor whose R is rotated and whose Q^T seeds the accumulator
    const double* tq = tau1;
    int Mq = M1, ldq = ld1;
    if (wide || p.triple) {
        // L = R1^T (K x K lower triangular), column-major working copy, second QR
        for (long i = gtid; i < (long)K * K; i += gth) {
            const long row = i % K, col = i / K;
            W2[row + col * ldk] = col <= row ? ld_cg(W1 + col + row * ld1) : 0.0;
        }
        fence_proxy_async();
        c.sync();
        lin_qr_factor(c, W2, tau2, K, K, ldk);
        Wq = W2;
        tq = tau2;
        Mq = K;
        ldq = ldk;
    }
    if (p.triple) {
        // tall: a third factorisation R2^T = Q3 R3 puts the accumulator back on the U side:
        //   A = Q1 R1 = Q1 R2^T Q2^T = (Q1 Q3) R3 Q2^T,  rows of R3 rotated, accumulator starts from Q3^T (K x K)
        double* W3 = ws + p.oW3;
        double* tau3 = ws + p.oTau3;
        for (long i = gtid; i < (long)K * K; i += gth) {
            const long row = i % K, col = i / K;
            W3[row + col * ldk] = col <= row ? ld_cg(W2 + col + row * ldk) : 0.0;
        }
        fence_proxy_async();
        c.sync();
        lin_qr_factor(c, W3, tau3, K, K, ldk);
        Wq = W3;
        tq = tau3;
        Mq = K;
        ldq = ldk;
    }
    const long long t_qr = c.now();
    for (long i = gtid; i < (long)K * K; i += gth) {
        const long row = i / K, col = i % K;
        G[row * ldg + col] = row <= col ? ld_cg(Wq + row + col * ldq) : 0.0;
    }
    for (long i = gtid; i < (long)K * (ldg - K - Mj); i += gth)      // pad columns travel with the bulk copies
        G[(i / (ldg - K - Mj)) * ldg + K + Mj + i % (ldg - K - Mj)] = 0.0;
    fence_proxy_async();
    c.sync();
    lin_apply_q(c, Wq, tq, ldq, Mq, K, K, 0, nullptr, 0, nullptr, Jt, ldg, 1);      // rows of Q^T (K x Mq), Mj == Mq
    fence_proxy_async();
    c.sync();
    const long long t_q = c.now();
    const int sweeps = p.resident ? lin_jacobi_resident(c, G, K, Mj) : lin_jacobi<NT>(c, G, K, Mj, flags, tm);
    const long long t_jac = c.now();
    // singular values = row norms
    for (int i = c.gw; i < K; i += c.GW) {
        double s = 0.0;
        for (int q = c.lane; q < K; q += 32) {
            const double g = ld_cg(G + (long)i * ldg + q);
            s += g * g;
        }
        s = warp_sum(s);
        if (c.lane == 0) sv[i] = sqrt(s);
    }
    c.sync();
    // descending order (stable counting rank), computed redundantly by every CTA in shared memory
    int* ord = (int*)(smem + p.oOrd);
    double* svS = smem + p.oSvS;
    for (int i = threadIdx.x; i < K; i += blockDim.x) svS[i] = ld_cg(sv + i);
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        int rank = 0;
        const double si = svS[i];
        for (int j = 0; j < K; ++j) rank += (svS[j] > si || (svS[j] == si && j < i)) ? 1 : 0;
        ord[rank] = i;
    }
    __syncthreads();
    double* S = p.S + (long)batch * K;
    for (long i = gtid; i < K; i += gth) S[i] = svS[ord[i]];
    if (p.triple) {
        lin_apply_q(c, W1, tau1, ld1, M1, K, K, 1, Jt, ldg, ord, U, 1, K);          // U columns = Q1 [(Q3 J) column; 0]
        lin_apply_q(c, W2, tau2, ldk, K, K, K, 1, G, ldg, ord, Wt, N, 1);           // W rows = Q2 g  (N == K)
    } else if (!wide) {
        for (long i = gtid; i < (long)M * K; i += gth) U[i] = ld_cg(Jt + (long)ord[i % K] * ldg + i / K);
        for (long i = gtid; i < (long)K * N; i += gth) Wt[i] = ld_cg(G + (long)ord[i / N] * ldg + i % N);
    } else {
        for (long i = gtid; i < (long)M * K; i += gth) U[i] = ld_cg(Jt + (long)ord[i % K] * ldg + i / K);
        lin_apply_q(c, W1, tau1, ld1, M1, K, K, 1, G, ldg, ord, Wt, N, 1);          // W rows = Q1 [g; 0]
    }
    if (gtid == 0 && p.info) {
        // [0] sweeps; ns: [1] QR(s), [2] Q^T set-up, [3] Jacobi, [4] grid, [5] block rows,
        // Jacobi breakdown on CTA 0 (ns): [6] row loads, [7] rotations, [8] row stores, [9] grid barriers; [10] total
        int* info = p.info + 16 * batch;
        info[0] = sweeps;
        info[1] = (int)(t_qr - t_start);
        info[2] = (int)(t_q - t_qr);
        info[3] = (int)(t_jac - t_q);
        info[4] = (int)gridDim.x;
        info[5] = p.nb;
        info[6] = (int)tm[0];
        info[7] = (int)tm[1];
        info[8] = (int)tm[2];
        info[9] = (int)tm[3];
        info[10] = (int)(c.now() - t_start);
        info[11] = p.cluster;
        info[12] = (int)c.tq[0];
        info[13] = (int)c.tq[1];
        info[14] = (int)c.tq[2];
        info[15] = (int)c.tq[3];
    }
}

static int g_coop_min_dim = 17;
static int g_resident_max_dim = 32;
static double g_floor_factor = 0.0;
static int g_tall_triple_qr = 1;
static int g_coop_threads = 256;
static int g_use_cluster = 1;
static int g_wide_cta_min_dim = 96;
static int g_min_block_rows = 8;
static int g_early_exit = 1;

struct LinPlan {
    LinParams p;
    int grid;
    int threads;
    long smem_bytes;
    long ws_total;
    double* ws;
};

static long even_up(long v) { return v + (v & 1); }

// shapes, workspace carve-up, shared-memory layout and grid of one call; returns 0 or an error code
static int lin_plan(LinPlan& pl, int mode, int M, int N, int nbatch) {
    LinParams& p = pl.p;
    const long K = imin(M, N);
    const bool wide = mode == 0 && M < N;
    p.M = M; p.N = N; p.mode = mode; p.nbatch = nbatch;
    p.floor_factor = g_floor_factor;
    p.early_exit = g_early_exit;
    p.M1 = wide ? N : M;
    p.N1 = wide ? M : N;
    // three QR factorisations pay off through the sweep count of graded unfoldings; below ~16 columns the sweeps are few
    // either way and the two extra factorisations are a third of the kernel (measured on B200: 8 x 6 .. 216 x 6 unfoldings
    // 58 -> 40 us, 32 x 24 and 68 x 24 0.25-0.29 -> 0.21-0.22 ms with the single-QR form; equal at 52 x 39)
    p.triple = (mode == 0 && !wide && g_tall_triple_qr && K > 32) ? 1 : 0;
    p.Mj = (wide || p.triple) ? (int)K : M;
    p.ld1 = (int)even_up(p.M1);
    p.ldk = (int)even_up(K);
    p.ldg = (int)even_up(K + p.Mj);
    p.cluster = 0;
    // every array starts at an even offset (16-byte aligned: bulk copies)
    long o = 0;
    p.oW1 = o; o += (long)p.ld1 * p.N1;
    p.oTau1 = o; o += even_up(K);
    p.oW2 = o; o += (wide || p.triple) ? (long)p.ldk * K : 0;
    p.oTau2 = o; o += (wide || p.triple) ? even_up(K) : 0;
    p.oW3 = o; o += p.triple ? (long)p.ldk * K : 0;
    p.oTau3 = o; o += p.triple ? even_up(K) : 0;
    p.oG = o; o += mode == 0 ? K * p.ldg : 0;
    p.oJt = p.oG + K;
    p.oSv = o; o += even_up(K);
    DevInfo di = dev_info();
    // 256-thread CTAs (8 warps = the 8 concurrent row pairs of a block-pair step; measured on B200: the pairwise
    // rotations are bound by shared-memory bandwidth, 16 warps per CTA do not finish a step sooner); the resident
    // single-CTA variant runs 512 threads above 16 rows
#ifdef TTIPM_EMU
    pl.threads = block_threads();
#else
    pl.threads = g_coop_threads;
#endif
    p.ldp = p.M1 + (p.M1 & 1);
    p.oOrd = 40;
    p.oSvS = p.oOrd + (int)(K + 1) / 2 + 1;
    p.oTaus = (int)even_up(p.oSvS + (int)K);
    p.oP = p.oTaus + 3 * TT_QR_PB;                  // tau, 1 / (alpha - beta), beta of the panel columns
    p.oRows = p.oTaus;
    // a panel of TT_QR_PB columns of M1 rows normally sits in shared memory; columns too long for that keep the
    // panel in the workspace and the matrix is factored by one CTA (rare: unfoldings with > ~3000 rows)
    const bool panel_global = ((long)p.oP + (long)TT_QR_PB * p.ldp) * 8 > di.smem_optin - 1024;
    p.oPg = -1;
    if (panel_global) { p.oPg = o; o += (long)TT_QR_PB * p.ldp; }
    p.ws_per = even_up(o);
    pl.ws_total = (long)nbatch * TT_LIN_HDR + (long)nbatch * p.ws_per;
    const long ldg = p.ldg;
    const long qr_doubles = p.oP + (panel_global ? 0 : (long)TT_QR_PB * p.ldp);
    // Jacobi block rows: 8, raised (even) until the block pairs of a round fit one thread-block cluster of 16 CTAs,
    // lowered until two blocks fit the shared memory of a CTA
    int nb = g_min_block_rows;
    while ((K + nb - 1) / nb > 32 && (p.oRows + 2L * (nb + 2) * ldg) * 8 <= di.smem_optin - 1024) nb += 2;
    while (nb > 1 && (p.oRows + 2L * nb * ldg) * 8 > di.smem_optin - 1024) nb /= 2;
    p.nb = nb;
    long jac_doubles = mode == 0 ? p.oRows + 2L * nb * ldg : 0;
    // every row resident in one CTA's shared memory when it fits (forced multi-CTA runs of the tests excepted)
    // (measured on B200: one SM's issue slots bound the resident variant, so beyond K ~ 32 a single matrix is faster
    // spread over several CTAs; batched calls keep one CTA per matrix)
    p.resident = (mode == 0 && K >= 2 && (p.oRows + K * ldg) * 8 <= di.smem_optin - 1024 &&
                  (nbatch > 1 || K <= g_resident_max_dim) && !(nbatch == 1 && g_coop_min_dim <= 1)) ? 1 : 0;
    if (p.resident) jac_doubles = p.oRows + K * ldg;
    pl.smem_bytes = 8 * (qr_doubles > jac_doubles ? qr_doubles : jac_doubles);
    if (pl.smem_bytes > di.smem_optin)
        return fail(4, "linalg: %d x %d does not fit the kernel's shared memory (%ld B)", M, N, pl.smem_bytes);
    int G = 1;
    if (nbatch == 1 && K >= g_coop_min_dim && K >= 2 && !panel_global && !p.resident) {
        const int nblk = (int)((K + nb - 1) / nb), npairs = (nblk + 1) / 2;
        if (mode == 0) G = npairs;                      // Jacobi keeps npairs CTAs busy
        else G = (int)((K + 7) / 8);                    // the QR phases: one warp per trailing column
        // one thread-block cluster (hardware barrier) up to 16 CTAs; unfoldings with >= 32 block pairs per round (K >= ~500:
        // the rank-exploded intermediates of the zip-up products, src/tt_ipm.py:1074 at maxcut_13) spread over up to 128 CTAs
        // of a cooperative grid instead
        G = G >= 32 ? imin(G, 128) : imin(G, 16);
        if (g_coop_min_dim <= 1) G = imax(G, 2);        // forced (tests): always exercise the multi-CTA path
        G = imax(1, imin(G, imin(128, di.sms)));
    }
    pl.grid = G;
#ifdef TTIPM_EMU
    if (p.resident && K > 16) pl.threads = 2 * block_threads();
#else
    if (p.resident && K > 16) pl.threads = 512;
    if (G > 1 && (nb > 8 || K >= g_wide_cta_min_dim)) pl.threads = 512;     // a warp per block row / trailing column
    // rows longer than 384 doubles: 16 / 22 register chunks per row need the 384-thread build of the kernel (170 registers)
    if (G > 1 && !p.resident && mode == 0 && ldg > 384 && ldg <= 704 && nb <= 12) pl.threads = 384;
#endif
    return 0;
}

template <int NT>
static int lin_launch_t(LinPlan& pl, int nbatch, tt_stream_t st) {
    LinParams& p = pl.p;
    int G = pl.grid;
    if (G > 1 && g_use_cluster &&
        cluster_launch_possible(k_linalg<NT>, G, nbatch, dim3(pl.threads), (size_t)pl.smem_bytes)) {
        p.cluster = 1;
        return launch_kernel_cluster("k_linalg", k_linalg<NT>, dim3(G, nbatch), dim3(pl.threads), (size_t)pl.smem_bytes, st, p);
    }
#ifndef TTIPM_EMU
    if (G > 1) {
        DevInfo di = dev_info();
        cudaFuncSetAttribute((const void*)k_linalg<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, di.smem_optin);
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_linalg<NT>, pl.threads, (size_t)pl.smem_bytes);
        if (per_sm < 1) return fail(4, "linalg: kernel does not fit on an SM with %ld B shared memory", pl.smem_bytes);
        if (G > per_sm * di.sms) G = per_sm * di.sms;
    }
#endif
    if (G > 1) {
        if (dev_memset(pl.ws, 0, (size_t)nbatch * TT_LIN_HDR * 8, st)) return fail(5, "linalg: memset failed");
        return launch_kernel("k_linalg", k_linalg<NT>, dim3(G, nbatch), dim3(pl.threads), (size_t)pl.smem_bytes, st, true, p);
    }
    return launch_kernel("k_linalg", k_linalg<NT>, dim3(1, nbatch), dim3(pl.threads), (size_t)pl.smem_bytes, st, false, p);
}

static int lin_launch(int mode, const double* A, long a_rs, long a_cs, long a_bs, int M, int N, double* U, double* S,
                      double* Wt, double* ws, int* info, int nbatch, tt_stream_t st) {
    LinPlan pl;
    int rc = lin_plan(pl, mode, M, N, nbatch);
    if (rc) return rc;
    LinParams& p = pl.p;
    p.A = A; p.a_rs = a_rs; p.a_cs = a_cs; p.a_bs = a_bs;
    p.U = U; p.S = S; p.Wt = Wt; p.info = info; p.ws = ws;
    if (((uintptr_t)ws & 15) != 0) return fail(1, "linalg: workspace must be 16-byte aligned");
    pl.ws = ws;
    if (pl.threads == 384) return lin_launch_t<384>(pl, nbatch, st);
    return lin_launch_t<512>(pl, nbatch, st);
}

}  // namespace ttipm

using namespace ttipm;

extern "C" int ttipm_linalg_coop_min_dim(int min_dim) {
    const int old = g_coop_min_dim;
    if (min_dim > 0) g_coop_min_dim = min_dim;
    return old;
}

extern "C" int ttipm_linalg_use_cluster(int on) {
    const int old = g_use_cluster;
    if (on >= 0) g_use_cluster = on ? 1 : 0;
    return old;
}

extern "C" int ttipm_linalg_threads(int threads) {
    const int old = g_coop_threads;
    if (threads == 256 || threads == 512) g_coop_threads = threads;
    g_wide_cta_min_dim = threads == 256 ? (1 << 30) : 96;           // 256: never widen (tuning runs)
    return old;
}

extern "C" int ttipm_linalg_block_rows(int nb) {
    const int old = g_min_block_rows;
    if (nb >= 2) g_min_block_rows = nb & ~1;
    return old;
}

extern "C" int ttipm_linalg_tall_triple_qr(int on) {
    const int old = g_tall_triple_qr;
    if (on >= 0) g_tall_triple_qr = on ? 1 : 0;
    return old;
}

extern "C" int ttipm_linalg_early_exit(int on) {
    const int old = g_early_exit;
    if (on >= 0) g_early_exit = on ? 1 : 0;
    return old;
}

extern "C" double ttipm_linalg_noise_floor(double factor) {
    const double old = g_floor_factor;
    if (factor >= 0.0) g_floor_factor = factor;
    return old;
}

extern "C" int64_t ttipm_qr_workspace(int M, int N, int nbatch) {
    LinPlan pl;
    if (M < 1 || N < 1 || nbatch < 1 || lin_plan(pl, 1, M, N, nbatch)) return 0;
    return pl.ws_total;
}

extern "C" int ttipm_qr(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int M, int N, double* Q, double* R,
                        double* workspace, int nbatch, void* stream) {
    if (M < 1 || N < 1 || nbatch < 1) return fail(1, "qr: bad dims %d x %d", M, N);
    if (!workspace) return fail(1, "qr: workspace required (ttipm_qr_workspace doubles)");
    return lin_launch(1, A, (long)a_rs, (long)a_cs, (long)a_bs, M, N, Q, nullptr, R, workspace, nullptr, nbatch,
                      (tt_stream_t)stream);
}

extern "C" int64_t ttipm_svd_workspace(int M, int N, int nbatch) {
    LinPlan pl;
    if (M < 1 || N < 1 || nbatch < 1 || lin_plan(pl, 0, M, N, nbatch)) return 0;
    return pl.ws_total;
}

extern "C" int ttipm_svd_left(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int M, int N, double* U,
                              double* S, double* Wt, double* workspace, int32_t* info, int nbatch, void* stream) {
    if (M < 1 || N < 1 || nbatch < 1) return fail(1, "svd_left: bad dims %d x %d", M, N);
    if (!workspace) return fail(1, "svd_left: workspace required (ttipm_svd_workspace doubles)");
    return lin_launch(0, A, (long)a_rs, (long)a_cs, (long)a_bs, M, N, U, S, Wt, workspace, info, nbatch,
                      (tt_stream_t)stream);
}
