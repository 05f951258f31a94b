// Small dense factorisations of core unfoldings, one CTA per matrix (batched over blockIdx.x):
//   * Householder QR  (replaces scipy.linalg.qr / LAPACK geqrf+orgqr at reference
//     cy_src/tt_ops_cy.pyx:147, src/tt_als.py:358, :482)
//   * "left" SVD by one-sided Jacobi on the rows with accumulated rotations
//     (replaces scipy.linalg.svd / LAPACK gesvd, gesdd at cy_src/tt_ops_cy.pyx:205, :404, :418 and
//     src/tt_als.py:270, :331, :457).  Every caller on the hot path only needs U, s and S*V^T, so
//     the kernel returns exactly those: U is a product of plane rotations (orthonormal even for zero
//     singular values) and W = S V^T are the rotated rows themselves (no division by s anywhere).
// Working arrays live in shared memory when they fit, otherwise in a caller-provided workspace.
#include "api_util.h"

namespace ttipm {

// ---- Householder QR on a column-major working copy W (M x N, leading dim M) ------------------------
// On exit W holds R in its upper triangle and the reflector tails below the diagonal, tau[K].
TT_DEV void qr_factor(double* W, double* tau, int M, int N, double* scr) {
    const int K = imin(M, N);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j = 0; j < K; ++j) {
        double* col = W + (long)j * M;
        double s = 0.0;
        for (int i = j + 1 + threadIdx.x; i < M; i += blockDim.x) s += col[i] * col[i];
        s = block_sum(s, scr);
        const double alpha = col[j];
        double tj = 0.0, scale = 0.0, beta = alpha;
        if (s != 0.0) {
            beta = -copysign(sqrt(alpha * alpha + s), alpha);
            tj = (beta - alpha) / beta;
            scale = 1.0 / (alpha - beta);
        }
        __syncthreads();
        for (int i = j + 1 + threadIdx.x; i < M; i += blockDim.x) col[i] *= scale;
        if (threadIdx.x == 0) {
            tau[j] = tj;
            col[j] = beta;
        }
        __syncthreads();
        if (tj != 0.0) {
            for (int c = j + 1 + wid; c < N; c += nw) {
                double* cc = W + (long)c * M;
                double d = lane == 0 ? cc[j] : 0.0;
                for (int i = j + 1 + lane; i < M; i += 32) d += col[i] * cc[i];
                d = warp_sum(d) * tj;
                if (lane == 0) cc[j] -= d;
                for (int i = j + 1 + lane; i < M; i += 32) cc[i] -= d * col[i];
            }
        }
        __syncthreads();
    }
}

// Q (M x K, element (i, c) at Q[i * q_rs + c * q_cs]) = H_0 ... H_{K-1} applied to the first K unit vectors
TT_DEV void qr_form_q(const double* W, const double* tau, int M, int N, double* Q, long q_rs, long q_cs) {
    const int K = imin(M, N);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
        const int row = i / K, c = i % K;
        Q[row * q_rs + c * q_cs] = row == c ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int j = K - 1; j >= 0; --j) {
        const double tj = tau[j];
        const double* v = W + (long)j * M;
        if (tj != 0.0) {
            for (int c = j + wid; c < K; c += nw) {
                double d = lane == 0 ? Q[j * q_rs + c * q_cs] : 0.0;
                for (int i = j + 1 + lane; i < M; i += 32) d += v[i] * Q[i * q_rs + c * q_cs];
                d = warp_sum(d) * tj;
                if (lane == 0) Q[j * q_rs + c * q_cs] -= d;
                for (int i = j + 1 + lane; i < M; i += 32) Q[i * q_rs + c * q_cs] -= d * v[i];
            }
        }
        __syncthreads();
    }
}

struct QrParams {
    const double* A;
    long a_rs, a_cs, a_bs;
    int M, N;
    double* Q;   // batch x M x K row-major
    double* R;   // batch x K x N row-major
    double* ws;  // batch x (M*N + K) when the working copy does not fit in shared memory
    int use_smem;
};

TT_GLOBAL void __launch_bounds__(TT_MAX_THREADS) k_qr(const QrParams p) {
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    const int M = p.M, N = p.N, K = imin(M, N);
    double* scr = smem;
    double* W = p.use_smem ? smem + 40 : p.ws + (long)blockIdx.x * ((long)M * N + K);
    double* tau = W + (long)M * N;
    const double* A = p.A + blockIdx.x * p.a_bs;
    for (int i = threadIdx.x; i < M * N; i += blockDim.x) {
        const int row = i / N, c = i % N;
        W[row + (long)c * M] = A[row * p.a_rs + c * p.a_cs];
    }
    __syncthreads();
    qr_factor(W, tau, M, N, scr);
    double* R = p.R + (long)blockIdx.x * K * N;
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) {
        const int row = i / N, c = i % N;
        R[i] = row <= c ? W[row + (long)c * M] : 0.0;
    }
    qr_form_q(W, tau, M, N, p.Q + (long)blockIdx.x * M * K, K, 1);
}

// ---- one-sided Jacobi on the rows of G (K x N row-major), Jt accumulates the rotations -------------
// On exit the rows of G are mutually orthogonal: G[i,:] = s_i v_i^T and column i of J = Jt[i,:] is u_i.
TT_DEV int jacobi_rows(double* G, double* Jt, int K, int N, int* flag) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int Ke = K + (K & 1);
    const double tol = 2.220446049250313e-16 * sqrt((double)N);
    int sweeps = 0;
    if (K < 2) return 0;
    for (; sweeps < 60; ++sweeps) {
        if (threadIdx.x == 0) *flag = 0;
        __syncthreads();
        for (int t = 0; t < Ke - 1; ++t) {
            for (int pi = wid; pi < Ke / 2; pi += nw) {
                int a, b;
                if (pi == 0) {
                    a = Ke - 1;
                    b = t;
                } else {
                    a = (t + pi) % (Ke - 1);
                    b = (t - pi + Ke - 1) % (Ke - 1);
                }
                if (a >= K || b >= K) continue;
                if (a > b) { const int q = a; a = b; b = q; }
                double* ga = G + (long)a * N;
                double* gb = G + (long)b * N;
                double saa = 0.0, sbb = 0.0, sab = 0.0;
                for (int i = lane; i < N; i += 32) {
                    const double x = ga[i], y = gb[i];
                    saa += x * x;
                    sbb += y * y;
                    sab += x * y;
                }
                saa = warp_sum(saa);
                sbb = warp_sum(sbb);
                sab = warp_sum(sab);
                if (fabs(sab) > tol * sqrt(saa * sbb) && sab != 0.0) {
                    const double zeta = (sbb - saa) / (2.0 * sab);
                    const double tg = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double cs = 1.0 / sqrt(1.0 + tg * tg), sn = cs * tg;
                    for (int i = lane; i < N; i += 32) {
                        const double x = ga[i], y = gb[i];
                        ga[i] = cs * x - sn * y;
                        gb[i] = sn * x + cs * y;
                    }
                    double* ja = Jt + (long)a * K;
                    double* jb = Jt + (long)b * K;
                    for (int i = lane; i < K; i += 32) {
                        const double x = ja[i], y = jb[i];
                        ja[i] = cs * x - sn * y;
                        jb[i] = sn * x + cs * y;
                    }
                    if (lane == 0) *flag = 1;
                }
            }
            __syncthreads();
        }
        const int any = *flag;
        __syncthreads();
        if (!any) break;
    }
    return sweeps;
}

struct SvdParams {
    const double* A;
    long a_rs, a_cs, a_bs;
    int M, N;
    double* U;      // batch x M x K row-major
    double* S;      // batch x K
    double* Wt;     // batch x K x N row-major  (S * V^T)
    double* ws;     // batch x ws_per
    long ws_per;
    int use_smem;
    int* info;      // batch: sweeps used
};

TT_GLOBAL void __launch_bounds__(TT_MAX_THREADS) k_svd_left(const SvdParams p) {
    TT_SMEM_DECL(smem_raw);
    double* smem = (double*)smem_raw;
    const int M = p.M, N = p.N, K = imin(M, N);
    const bool tall = M > N;
    double* scr = smem;
    int* flag = (int*)(smem + 36);
    double* base = p.use_smem ? smem + 40 : p.ws + blockIdx.x * p.ws_per;
    // workspace layout: G (K x N) | Jt (K x K) | sv (K) | ord (K ints, in doubles) | [tall: W (M x N) | tau (K) | Q (M x K)]
    double* G = base;
    double* Jt = G + (long)K * N;
    double* sv = Jt + (long)K * K;
    int* ord = (int*)(sv + K);
    double* W = sv + K + (K + 1) / 2 + 1;
    double* tau = W + (long)M * N;
    double* Q = tau + K;
    const double* A = p.A + blockIdx.x * p.a_bs;
    if (tall) {
        for (int i = threadIdx.x; i < M * N; i += blockDim.x) {
            const int row = i / N, c = i % N;
            W[row + (long)c * M] = A[row * p.a_rs + c * p.a_cs];
        }
        __syncthreads();
        qr_factor(W, tau, M, N, scr);
        for (int i = threadIdx.x; i < N * N; i += blockDim.x) {
            const int row = i / N, c = i % N;
            G[i] = row <= c ? W[row + (long)c * M] : 0.0;
        }
        qr_form_q(W, tau, M, N, Q, K, 1);
    } else {
        for (int i = threadIdx.x; i < M * N; i += blockDim.x) G[i] = A[(i / N) * p.a_rs + (i % N) * p.a_cs];
    }
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) Jt[i] = (i / K == i % K) ? 1.0 : 0.0;
    __syncthreads();
    const int sweeps = jacobi_rows(G, Jt, K, N, flag);
    // singular values = row norms; sort descending (stable rank by counting)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = wid; i < K; i += nw) {
        double s = 0.0;
        for (int c = lane; c < N; c += 32) s += G[(long)i * N + c] * G[(long)i * N + c];
        s = warp_sum(s);
        if (lane == 0) sv[i] = sqrt(s);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K; i += blockDim.x) {
        int rank = 0;
        const double si = sv[i];
        for (int j = 0; j < K; ++j) rank += (sv[j] > si || (sv[j] == si && j < i)) ? 1 : 0;
        ord[rank] = i;
    }
    __syncthreads();
    double* S = p.S + (long)blockIdx.x * K;
    double* Wt = p.Wt + (long)blockIdx.x * K * N;
    double* U = p.U + (long)blockIdx.x * M * K;
    for (int i = threadIdx.x; i < K; i += blockDim.x) S[i] = sv[ord[i]];
    for (int i = threadIdx.x; i < K * N; i += blockDim.x) Wt[i] = G[(long)ord[i / N] * N + i % N];
    if (!tall) {
        for (int i = threadIdx.x; i < M * K; i += blockDim.x) U[i] = Jt[(long)ord[i % K] * K + i / K];
    } else {
        // U[:, p] = Q * J[:, ord[p]]
        for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
            const int row = i / K, pcol = i % K;
            const double* jt = Jt + (long)ord[pcol] * K;
            const double* q = Q + (long)row * K;
            double acc = 0.0;
            for (int k = 0; k < K; ++k) acc += q[k] * jt[k];
            U[i] = acc;
        }
    }
    if (threadIdx.x == 0 && p.info) p.info[blockIdx.x] = sweeps;
}

static long svd_ws_doubles(int M, int N) {
    const long K = M < N ? M : N;
    long n = K * N + K * K + K + (K + 1) / 2 + 1;
    if (M > N) n += (long)M * N + K + (long)M * K;
    return n + 8;
}

}  // namespace ttipm

using namespace ttipm;

extern "C" int64_t ttipm_qr_workspace(int M, int N, int nbatch) {
    return (int64_t)nbatch * ((int64_t)M * N + (M < N ? M : N));
}

extern "C" int ttipm_qr(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int M, int N, double* Q, double* R,
                        double* workspace, int nbatch, void* stream) {
    if (M < 1 || N < 1 || nbatch < 1) return fail(1, "qr: bad dims %d x %d", M, N);
    QrParams p{A, (long)a_rs, (long)a_cs, (long)a_bs, M, N, Q, R, workspace, 0};
    const long need = ((long)M * N + imin(M, N) + 40) * 8;
    DevInfo di = dev_info();
    p.use_smem = need <= di.smem_optin;
    if (!p.use_smem && !workspace) return fail(1, "qr: %d x %d needs a workspace", M, N);
    return launch_kernel("k_qr", k_qr, dim3(nbatch), dim3(block_threads()), p.use_smem ? need : 40 * 8,
                         (tt_stream_t)stream, false, p);
}

extern "C" int64_t ttipm_svd_workspace(int M, int N, int nbatch) { return (int64_t)nbatch * svd_ws_doubles(M, N); }

extern "C" int ttipm_svd_left(const double* A, int64_t a_rs, int64_t a_cs, int64_t a_bs, int M, int N, double* U,
                              double* S, double* Wt, double* workspace, int32_t* info, int nbatch, void* stream) {
    if (M < 1 || N < 1 || nbatch < 1) return fail(1, "svd_left: bad dims %d x %d", M, N);
    SvdParams p{A, (long)a_rs, (long)a_cs, (long)a_bs, M, N, U, S, Wt, workspace, svd_ws_doubles(M, N), 0, info};
    const long need = (p.ws_per + 40) * 8;
    DevInfo di = dev_info();
    p.use_smem = need <= di.smem_optin;
    if (!p.use_smem && !workspace) return fail(1, "svd_left: %d x %d needs a workspace", M, N);
    return launch_kernel("k_svd_left", k_svd_left, dim3(nbatch), dim3(block_threads()), p.use_smem ? need : 40 * 8,
                         (tt_stream_t)stream, false, p);
}
