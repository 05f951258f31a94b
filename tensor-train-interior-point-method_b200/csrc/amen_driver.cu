// Native sweep driver: the whole block AMEn solve of one KKT system behind ONE C-ABI object.
//
// Mirrors tt_block_amen / _bck_sweep / _fwd_sweep (reference src/tt_als.py:277-670) and the two local
// solvers (reference src/tt_ipm.py:183-401).  The control flow is the reference's; every numerical step is
// a kernel of this library launched back-to-back on one stream from C++, device buffers come from the
// stream-ordered allocator, and the host only waits for the handful of scalars the reference branches on
// (residual norms, singular values, truncation norms) through a pinned staging buffer.  The Python layer
// (ttipm_b200.amen) only uploads the operands, makes the reference's NumPy RNG draws and fetches the result.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <algorithm>
#include "tensor.h"
#include "drv_ops.h"

#ifndef TTIPM_EMU
#include <cublas_v2.h>
#include <cusolverDn.h>
#endif

namespace ttipm {
namespace drv {

// ---------------------------------------------------------------------------------------------------
// per-launch profile (profiling passes only)
// ---------------------------------------------------------------------------------------------------
ProfScope::ProfScope(Ctx& ctx, int cat, double work) : c(ctx), on(ctx.prof) {
#ifndef TTIPM_EMU
    if (!on) return;
    ProfRec r;
    r.cat = cat;
    r.work = work;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    r.e0 = a;
    r.e1 = b;
    cudaEventRecord(a, c.st);
    idx = c.recs.size();
    c.recs.push_back(r);
#else
    (void)cat; (void)work;
    on = false;
#endif
}
ProfScope::~ProfScope() {
#ifndef TTIPM_EMU
    if (on) cudaEventRecord((cudaEvent_t)c.recs[idx].e1, c.st);
#endif
}

// ---------------------------------------------------------------------------------------------------
// memory and transfers
// ---------------------------------------------------------------------------------------------------
void* dev_alloc(Ctx& c, size_t bytes) {
    void* p = nullptr;
#ifdef TTIPM_EMU
    p = malloc(bytes ? bytes : 8);
#else
    if (cudaMallocAsync(&p, bytes ? bytes : 8, c.st) != cudaSuccess) throw DriverError(91, "device allocation failed");
#endif
    c.bytes_live += bytes;
    if (c.bytes_live > c.bytes_peak) c.bytes_peak = c.bytes_live;
    return p;
}
void dev_free(Ctx& c, void* p, size_t bytes) {
#ifdef TTIPM_EMU
    free(p);
#else
    cudaFreeAsync(p, c.st);
#endif
    c.bytes_live -= bytes;
}
// The pinned staging buffer for scalar read-backs is process wide (like the library handles below): cudaMallocHost /
// cudaFreeHost cost hundreds of microseconds and serialise the device, and a drop-in run creates one solver object per
// Newton system.  One caller thread and one stream per process, so the buffer is never shared concurrently.
static double* g_pinned = nullptr;
static size_t g_pinned_cap = 0;
static void ensure_pinned(Ctx& c, size_t n) {
    if (g_pinned_cap < n) {
        const size_t want = n < 4096 ? 4096 : n + n / 2;
#ifdef TTIPM_EMU
        free(g_pinned);
        g_pinned = (double*)malloc(want * sizeof(double));
#else
        if (g_pinned) cudaFreeHost(g_pinned);
        g_pinned = nullptr;
        if (cudaMallocHost((void**)&g_pinned, want * sizeof(double)) != cudaSuccess) throw DriverError(91, "pinned allocation failed");
#endif
        g_pinned_cap = want;
    }
    c.pinned = g_pinned;
    c.pinned_cap = g_pinned_cap;
}
void to_host(Ctx& c, const double* dev, size_t n, double* host) {
#ifdef TTIPM_EMU
    memcpy(host, dev, n * sizeof(double));
#else
    ensure_pinned(c, n);
    cudaMemcpyAsync(c.pinned, dev, n * sizeof(double), cudaMemcpyDeviceToHost, c.st);
    if (cudaStreamSynchronize(c.st) != cudaSuccess) throw DriverError(92, std::string("stream sync failed: ") +
                                                                           cudaGetErrorString(cudaGetLastError()));
    memcpy(host, c.pinned, n * sizeof(double));
#endif
    c.syncs++;
}
void from_host(Ctx& c, const double* host, size_t n, double* dev) {
#ifdef TTIPM_EMU
    memcpy(dev, host, n * sizeof(double));
#else
    cudaMemcpyAsync(dev, host, n * sizeof(double), cudaMemcpyHostToDevice, c.st);
    cudaStreamSynchronize(c.st);       // the host buffer is pageable and may be released by the caller
#endif
}
void dev_to_dev(Ctx& c, const double* src, size_t n, double* dst) {
    if (dev_copy(dst, src, n * sizeof(double), c.st)) throw DriverError(92, "device copy failed");
}
// ---------------------------------------------------------------------------------------------------
// dense Schur fallback: cuSOLVER / cuBLAS on row-major device matrices (reference src/tt_ipm.py:196-223, :298-334)
// ---------------------------------------------------------------------------------------------------
#ifndef TTIPM_EMU
#endif
struct Dense {
    Ctx& c;
#ifndef TTIPM_EMU
    cublasHandle_t blas = nullptr;
    cusolverDnHandle_t sol = nullptr;
#endif
    explicit Dense(Ctx& ctx) : c(ctx) {
#ifndef TTIPM_EMU
        blas = (cublasHandle_t)blas_handle(c.st);          // one pair per process (api.cu), re-bound to this stream
        sol = (cusolverDnHandle_t)solver_handle(c.st);
        if (!blas) throw DriverError(93, "cublasCreate failed");
        if (!sol) throw DriverError(93, "cusolverDnCreate failed");
#endif
    }
    ~Dense() {}
    // C (M x N) = alpha * op(A) op(B) + beta * C, all row-major contiguous
    void gemm(bool ta, bool tb, long M, long N, long K, double alpha, const double* A, const double* B, double beta, double* C) {
#ifdef TTIPM_EMU
        std::vector<double> out((size_t)(M * N));
        for (long i = 0; i < M; ++i)
            for (long j = 0; j < N; ++j) {
                double acc = 0.0;
                for (long k = 0; k < K; ++k) acc += (ta ? A[k * M + i] : A[i * K + k]) * (tb ? B[j * K + k] : B[k * N + j]);
                out[i * N + j] = alpha * acc + (beta != 0.0 ? beta * C[i * N + j] : 0.0);
            }
        memcpy(C, out.data(), out.size() * sizeof(double));
#else
        // row-major C = A B  <=>  column-major C^T = B^T A^T
        const long lda = ta ? M : K, ldb = tb ? K : N;
        ProfScope ps(c, CAT_DENSE, 2.0 * M * N * K);
        if (cublasDgemm(blas, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N, (int)N, (int)M, (int)K, &alpha,
                        B, (int)ldb, A, (int)lda, &beta, C, (int)N) != CUBLAS_STATUS_SUCCESS)
            throw DriverError(93, "cublasDgemm failed");
#endif
        c.launches++;
    }
    // in-place Cholesky of a symmetric row-major matrix: lower factor in the lower triangle; false if not SPD
    bool cholesky(double* A, long m) {
#ifdef TTIPM_EMU
        for (long j = 0; j < m; ++j) {
            double d = A[j * m + j];
            for (long k = 0; k < j; ++k) d -= A[j * m + k] * A[j * m + k];
            if (!(d > 0.0)) return false;
            d = sqrt(d);
            A[j * m + j] = d;
            for (long i = j + 1; i < m; ++i) {
                double v = A[i * m + j];
                for (long k = 0; k < j; ++k) v -= A[i * m + k] * A[j * m + k];
                A[i * m + j] = v / d;
            }
        }
        return true;
#else
        ProfScope ps(c, CAT_DENSE, (double)m * m * m / 3.0);
        int lwork = 0;
        cusolverDnDpotrf_bufferSize(sol, CUBLAS_FILL_MODE_UPPER, (int)m, A, (int)m, &lwork);
        Tensor work = Tensor::empty(c, {(long)lwork + 2});
        int* info = (int*)(work.p + lwork);
        if (cusolverDnDpotrf(sol, CUBLAS_FILL_MODE_UPPER, (int)m, A, (int)m, work.p, lwork, info) != CUSOLVER_STATUS_SUCCESS)
            throw DriverError(93, "potrf failed");
        double h;
        to_host(c, work.p + lwork, 1, &h);
        int hi;
        memcpy(&hi, &h, sizeof(int));
        c.launches++;
        return hi == 0;
#endif
    }
    // X = (L L^T)^{-1} B in place, B row-major (m x k)
    void chol_solve(const double* Lc, long m, double* B, long k) {
#ifdef TTIPM_EMU
        for (long col = 0; col < k; ++col) {
            for (long i = 0; i < m; ++i) {
                double v = B[i * k + col];
                for (long j = 0; j < i; ++j) v -= Lc[i * m + j] * B[j * k + col];
                B[i * k + col] = v / Lc[i * m + i];
            }
            for (long i = m; i-- > 0;) {
                double v = B[i * k + col];
                for (long j = i + 1; j < m; ++j) v -= Lc[j * m + i] * B[j * k + col];
                B[i * k + col] = v / Lc[i * m + i];
            }
        }
#else
        // column-major view: B^T (k x m), factor U = L^T upper with A = U^T U; solve X^T U^T U = B^T from the right
        const double one = 1.0;
        ProfScope ps(c, CAT_DENSE, 2.0 * m * m * k);
        if (cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, (int)k, (int)m,
                        &one, Lc, (int)m, B, (int)k) != CUBLAS_STATUS_SUCCESS ||
            cublasDtrsm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)k, (int)m,
                        &one, Lc, (int)m, B, (int)k) != CUBLAS_STATUS_SUCCESS)
            throw DriverError(93, "trsm failed");
        c.launches += 2;
#endif
    }
    // LU factorisation (partial pivoting) of a row-major matrix in place; piv has m ints (stored in doubles)
    struct LU { Tensor a; Tensor piv; long m; };
    std::vector<Tensor> lu_infos;      // device info words of the getrf calls of this solve (checked by lu_check)
    // scipy's lu_factor / solve raise on an exactly singular factor (reference src/tt_ipm.py:204,215,320-326 -> the
    // except branch sets direct_solve_failure): one read-back for all factorizations of the solve
    void lu_check() {
#ifndef TTIPM_EMU
        for (const Tensor& t : lu_infos) {
            double h;
            to_host(c, t.p, 1, &h);
            int hi;
            memcpy(&hi, &h, sizeof(int));
            if (hi != 0) {
                lu_infos.clear();
                throw DriverError(94, "local Schur complement is singular (getrf info > 0)");
            }
        }
#endif
        lu_infos.clear();
    }
    LU lu_factor(const Tensor& A, long m) {
        LU f;
        f.m = m;
        f.a = Tensor::empty(c, {m, m});
        dev_to_dev(c, A.p, (size_t)(m * m), f.a.p);
        f.piv = Tensor::empty(c, {m + 2});
#ifdef TTIPM_EMU
        int* piv = (int*)f.piv.p;
        double* a = f.a.p;           // factor the row-major matrix directly (row pivoting)
        for (long k = 0; k < m; ++k) {
            long best = k;
            for (long i = k + 1; i < m; ++i)
                if (fabs(a[i * m + k]) > fabs(a[best * m + k])) best = i;
            piv[k] = (int)best;
            if (best != k)
                for (long j = 0; j < m; ++j) std::swap(a[k * m + j], a[best * m + j]);
            if (a[k * m + k] == 0.0) throw DriverError(94, "local Schur complement is singular");
            for (long i = k + 1; i < m; ++i) {
                a[i * m + k] /= a[k * m + k];
                for (long j = k + 1; j < m; ++j) a[i * m + j] -= a[i * m + k] * a[k * m + j];
            }
        }
#else
        ProfScope ps(c, CAT_DENSE, 2.0 * m * m * m / 3.0);
        int lwork = 0;
        cusolverDnDgetrf_bufferSize(sol, (int)m, (int)m, f.a.p, (int)m, &lwork);
        Tensor work = Tensor::empty(c, {(long)lwork + 2});
        Tensor info_t = Tensor::empty(c, {2});
        int* info = (int*)info_t.p;
        if (cusolverDnDgetrf(sol, (int)m, (int)m, f.a.p, (int)m, work.p, (int*)f.piv.p, info) != CUSOLVER_STATUS_SUCCESS)
            throw DriverError(93, "getrf failed");
        lu_infos.push_back(info_t);
        c.launches++;
#endif
        return f;
    }
    // X = A^{-1} B for row-major B (m x k); returns a new row-major tensor
    Tensor lu_solve(const LU& f, const Tensor& B, long k) {
        const long m = f.m;
#ifdef TTIPM_EMU
        Tensor X = Tensor::empty(c, {m, k});
        memcpy(X.p, B.p, sizeof(double) * (size_t)(m * k));
        const int* piv = (const int*)f.piv.p;
        const double* a = f.a.p;
        for (long col = 0; col < k; ++col) {
            for (long i = 0; i < m; ++i)
                if (piv[i] != i) std::swap(X.p[i * k + col], X.p[piv[i] * k + col]);
            for (long i = 0; i < m; ++i)
                for (long j = 0; j < i; ++j) X.p[i * k + col] -= a[i * m + j] * X.p[j * k + col];
            for (long i = m; i-- > 0;) {
                for (long j = i + 1; j < m; ++j) X.p[i * k + col] -= a[i * m + j] * X.p[j * k + col];
                X.p[i * k + col] /= a[i * m + i];
            }
        }
        return X;
#else
        // memory of row-major A is column-major A^T =: C (factored); A X = B <=> C^T X = B; getrs wants column-major
        // right-hand sides, i.e. B^T in row-major memory, and returns X^T the same way
        Tensor Bt = k == 1 ? Tensor::empty(c, {m, 1}) : copy2d(c, B.t2());
        if (k == 1) dev_to_dev(c, B.p, (size_t)m, Bt.p);
        Tensor info = Tensor::empty(c, {2});
        ProfScope ps(c, CAT_DENSE, 2.0 * m * m * k);
        if (cusolverDnDgetrs(sol, CUBLAS_OP_T, (int)m, (int)k, f.a.p, (int)m, (const int*)f.piv.p, Bt.p, (int)m,
                             (int*)info.p) != CUSOLVER_STATUS_SUCCESS)
            throw DriverError(93, "getrs failed");
        c.launches++;
        if (k == 1) return Bt.reshape({m, 1});
        Tensor Xt = Bt.reshape({k, m});
        return copy2d(c, Xt.t2());
#endif
    }
};

// ---------------------------------------------------------------------------------------------------
// the solver object
// ---------------------------------------------------------------------------------------------------
// TTIPM_AMEN_LOG=1: one stderr line per core step of a sweep (wall clock, shapes, residuals, which local solver ran) --
// for runs that do not come back (e2e graphm_3), off by default and free when off
static bool amen_log_on() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("TTIPM_AMEN_LOG");
        on = (e && e[0] && e[0] != '0') ? 1 : 0;
    }
    return on == 1;
}
static double amen_wall_ms() {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int g_host_krylov_mode = 1;

struct Amen {
    Ctx c;
    int d = 0, bs = 0;
    bool ineq = false;
    std::map<Key, std::vector<Tensor>> A;      // operator cores (s, n, n, S)
    std::map<Key, Key> aliases, transposes;
    std::map<int, std::vector<Tensor>> b;      // rhs cores (rb, n, rb')
    std::vector<Tensor> x, z;
    std::vector<KeyMap> XAX, ZAX;
    std::vector<RowMap> Xb, Zb;
    std::vector<long> rx, rz, N;
    bool amen = true;
    int kick_rank = 2, r_max = 100;
    double eps = 1e-12, trunc_tol = 0.0;
    bool direct_solve_failure = false;
    int sweeps = 0;
    long local_solves = 0, lgmres_its = 0, lgmres_matvecs = 0, lgmres_calls = 0, dense_solves = 0, krylov_failures = 0;
    long host_krylov_solves = 0;
    int host_krylov_mode = g_host_krylov_mode; // 0 never, 1 automatic (host_krylov_wanted), 2 always (tests)
    std::string last_krylov_error;
    std::vector<double> trace;                 // (swp, k, res_old, res_new, r*R) per local solve
    std::unique_ptr<Dense> dense;
    std::vector<Tensor> lg_infos;              // device info blocks of the Krylov solves (read lazily)
    struct LgProf {
#ifndef TTIPM_EMU
        cudaEvent_t e0, e1;
#endif
        double mv_flops, nv;
        int restart;
    };
    std::vector<LgProf> lg_prof;               // one record per Krylov solve (same order as lg_infos)
    bool profile = false, stats_done = false;
    double lg_time = 0.0, lg_flops = 0.0;
    double cat_seconds[CAT_COUNT] = {0}, cat_work[CAT_COUNT] = {0}, cat_launches[CAT_COUNT] = {0};
    std::string error;

    Tensor ones(std::initializer_list<long> dims) {
        Tensor t = Tensor::empty(c, dims);
        std::vector<double> h((size_t)t.numel(), 1.0);
        from_host(c, h.data(), h.size(), t.p);
        return t;
    }

    // ---- term lists ------------------------------------------------------------------------------
    Terms full_terms(int k) {
        Terms tl;
        for (auto& kv : A) {
            const Key key = kv.first;
            const Tensor& Ak = kv.second[k];
            const Tensor &P1 = XAX[k].at(key), &P2 = XAX[k + 1].at(key);
            tl.add(P1, Ak, P2, key.second, key.first);
            auto tr = transposes.find(key);
            if (tr != transposes.end())
                tl.add(P1.permute({2, 1, 0}), Ak.permute({0, 2, 1, 3}), P2.permute({2, 1, 0}), tr->second.second, tr->second.first);
            auto al = aliases.find(key);
            if (al != aliases.end()) tl.add(P1, Ak, P2, al->second.second, al->second.first);
        }
        return tl;
    }
    // compressed_/lcompressed_/rcompressed_block_local_product (reference src/tt_als.py:202-238)
    Terms mixed_terms(int k, const KeyMap& left, const KeyMap& right, bool left_is_z, bool right_is_z) {
        Terms tl;
        for (auto& kv : A) {
            const Key key = kv.first;
            const Tensor& Ak = kv.second[k];
            tl.add(left.at(key), Ak, right.at(key), key.second, key.first);
            auto tr = transposes.find(key);
            if (tr != transposes.end()) {
                const Key pt = tr->second;
                Tensor Pl = left_is_z ? left.at(pt) : left.at(key).permute({2, 1, 0});
                Tensor Pr = right_is_z ? right.at(pt) : right.at(key).permute({2, 1, 0});
                tl.add(Pl, Ak.permute({0, 2, 1, 3}), Pr, pt.second, pt.first);
            }
            auto al = aliases.find(key);
            if (al != aliases.end()) tl.add(left.at(key), Ak, right.at(key), al->second.second, al->second.first);
        }
        return tl;
    }
    Tensor rhs_block(int k, const RowMap& Xl, const RowMap& Xr, long r, long n, long R) {
        std::vector<Tensor> X1, Bc, X2;
        std::vector<int> rows;
        for (auto& kv : b) {
            rows.push_back(kv.first);
            X1.push_back(Xl.at(kv.first));
            Bc.push_back(kv.second[k]);
            X2.push_back(Xr.at(kv.first));
        }
        return rhs_project(c, X1, Bc, X2, rows, r, bs, n, R);
    }
    Tensor apply_block(const Tensor& P1, const Tensor& Ak, const Tensor& P2, const Tensor& v3) {   // v3: (r, n, R)
        Terms tl;
        tl.add(P1, Ak, P2, 0, 0);
        Tensor x4 = v3.reshape({v3.d[0], 1, v3.d[1], v3.d[2]});
        Tensor y = block_matvec(c, tl, x4, false, 1, P1.d[0], P2.d[0], nullptr, 1.0, 0.0, nullptr);
        return y.reshape({P1.d[0], v3.d[1], P2.d[0]});
    }

    // ---- local solver (reference src/tt_ipm.py:183-401) ---------------------------------------------------
    struct LocalOut { Tensor sol, rhs; double res_old, res_new, norm_rhs; };

    Tensor contiguous3(const Tensor& v) {          // x[:, j] slice -> contiguous (r, n, R)
        Tensor out = Tensor::empty(c, {v.d[0], v.d[1], v.d[2]});
        ewise(c, v, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &out, nullptr);
        return out;
    }

    Tensor dense_solve(int k, const Tensor& rhs, const Tensor& inv_I, long r, long n, long R) {
        if (!dense) dense.reset(new Dense(c));
        Dense& D = *dense;
        const long m = r * n * R;
        auto blk = [&](int i, int j) { return local_dense(c, XAX[k].at({i, j}), A.at({i, j})[k], XAX[k + 1].at({i, j})); };
        auto col = [&](int i) { return contiguous3(rhs.select(1, i)).reshape({m, 1}); };
        auto scale_cols = [&](const Tensor& M) {          // M * inv_I[None, :]
            Tensor out = Tensor::empty(c, {m, m});
            check_rc(ttipm_scale2d(M.p, m, 1, (int)m, (int)m, inv_I.p, 1, 0, out.p, c.st), "scale2d");
            c.launches++;
            return out;
        };
        auto axpy = [&](const Tensor& a, double beta, const Tensor& bb) {   // a + beta * bb (new tensor)
            Tensor out = Tensor::empty(c, {a.d[0], a.d[1]});
            ewise(c, a, 1.0, &bb, beta, nullptr, 0.0, nullptr, &out, nullptr);
            return out;
        };
        auto mm = [&](const Tensor& X, const Tensor& Y, bool ty = false) {
            const long M = X.d[0], K = X.d[1], N = ty ? Y.d[0] : Y.d[1];
            Tensor out = Tensor::empty(c, {M, N});
            D.gemm(false, ty, M, N, K, 1.0, X.p, Y.p, 0.0, out.p);
            return out;
        };
        auto add_diag = [&](Tensor& M) {
            Tensor e = Tensor::empty(c, {m});
            std::vector<double> h((size_t)m, 1e-11);
            from_host(c, h.data(), h.size(), e.p);
            // diag += 1e-11 through the strided panel form (rows = m, inner = 1)
            check_rc(ttipm_ewise((int)m, 1, 1.0, M.p, m + 1, 1.0, e.p, 1, 0.0, nullptr, 0, nullptr, 0, M.p, m + 1, nullptr, c.st),
                     "diag shift");
            c.launches++;
        };
        Tensor Lc = blk(2, 1);
        if (!D.cholesky(Lc.p, m)) throw DriverError(94, "local L_Z block is not positive definite");
        Tensor sol = Tensor::empty(c, {r, (long)bs, n, R});
        auto put = [&](int j, const Tensor& v) {
            Tensor dst = sol.select(1, j);
            Tensor src = v.reshape({r, n, R});
            ewise(c, src, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &dst, nullptr);
        };
        const Tensor &P1_01 = XAX[k].at({0, 1}), &A01 = A.at({0, 1})[k], &P2_01 = XAX[k + 1].at({0, 1});
        auto applyT01 = [&](const Tensor& y3) {
            return apply_block(P1_01.permute({2, 1, 0}), A01.permute({0, 2, 1, 3}), P2_01.permute({2, 1, 0}), y3);
        };
        if (!ineq) {
            Tensor Rp = col(0), Rd = col(1), Rc = col(2);
            Tensor LXI = scale_cols(blk(2, 2));
            Tensor Leq = blk(0, 1);
            Tensor t1 = axpy(Rc, -1.0, mm(LXI, Rd));
            D.chol_solve(Lc.p, m, t1.p, 1);
            Tensor bb = axpy(Rp, -1.0, mm(Leq, t1));
            Tensor Zm = Tensor::empty(c, {m, m});
            dev_to_dev(c, LXI.p, (size_t)(m * m), Zm.p);
            D.chol_solve(Lc.p, m, Zm.p, m);
            Tensor S = mm(Leq, mm(Zm, Leq, true));
            Tensor K00 = blk(0, 0);
            S = axpy(S, 1.0, K00);
            add_diag(S);
            Dense::LU f = D.lu_factor(S, m);
            Tensor y = D.lu_solve(f, bb, 1);
            put(0, y);
            Tensor y3 = contiguous3(sol.select(1, 0));
            Tensor kty = applyT01(y3).reshape({m, 1});
            Tensor zz = Tensor::empty(c, {m, 1});
            Tensor invc = inv_I.reshape({m, 1});
            ewise(c, Rd, 1.0, &kty, -1.0, nullptr, 0.0, &invc, &zz, nullptr);
            put(2, zz);
            Tensor z3 = zz.reshape({r, n, R});
            Tensor k22z = apply_block(XAX[k].at({2, 2}), A.at({2, 2})[k], XAX[k + 1].at({2, 2}), z3).reshape({m, 1});
            Tensor xx = axpy(Rc, -1.0, k22z);
            D.chol_solve(Lc.p, m, xx.p, 1);
            put(1, xx);
        } else {
            Tensor Rp = col(0), Rd = col(1), Rc = col(2), Rt = col(3);
            Tensor LZc = Tensor::empty(c, {m, 1});
            dev_to_dev(c, Rc.p, (size_t)m, LZc.p);
            D.chol_solve(Lc.p, m, LZc.p, 1);
            Tensor LZX = blk(2, 2);
            D.chol_solve(Lc.p, m, LZX.p, m);
            Tensor LZXI = scale_cols(LZX);
            Tensor Leq = blk(0, 1), Top = blk(3, 1);
            Tensor w = axpy(LZc, -1.0, mm(LZXI, Rd));
            Tensor u = axpy(Rp, -1.0, mm(Leq, w));
            Tensor v = axpy(Rt, -1.0, mm(Top, w));
            Tensor Am = axpy(blk(0, 0), 1.0, mm(mm(Leq, LZXI), Leq, true));
            Tensor Dm = axpy(blk(3, 3), 1.0, mm(Top, LZX));
            add_diag(Dm);
            Tensor TopS = mm(mm(Top, LZXI), Leq, true);
            Tensor LeqS = mm(Leq, LZX);
            Dense::LU fD = D.lu_factor(Dm, m);
            Tensor rhs_l = axpy(u, -1.0, mm(LeqS, D.lu_solve(fD, v, 1)));
            Tensor lhs_l = axpy(Am, -1.0, mm(LeqS, D.lu_solve(fD, TopS, m)));
            Dense::LU fl = D.lu_factor(lhs_l, m);
            Tensor y = D.lu_solve(fl, rhs_l, 1);
            put(0, y);
            Tensor tt = D.lu_solve(fD, axpy(v, -1.0, mm(TopS, y)), 1);
            put(3, tt);
            Tensor y3 = contiguous3(sol.select(1, 0));
            Tensor kty = applyT01(y3).reshape({m, 1});
            Tensor zz = Tensor::empty(c, {m, 1});
            Tensor invc = inv_I.reshape({m, 1});
            ewise(c, Rd, 1.0, &kty, -1.0, &tt, -1.0, &invc, &zz, nullptr);
            put(2, zz);
            Tensor z3 = zz.reshape({r, n, R});
            Tensor k22z = apply_block(XAX[k].at({2, 2}), A.at({2, 2})[k], XAX[k + 1].at({2, 2}), z3).reshape({m, 1});
            Tensor xx = axpy(Rc, -1.0, k22z);
            D.chol_solve(Lc.p, m, xx.p, 1);
            put(1, xx);
        }
        D.lu_check();
        dense_solves++;
        return sol;
    }

    // ---- host-driven LGMRES (same algorithm as k_lgmres; reference cy_src/lgmres_cy.pyx:203-510 over PETSc LGMRES) ---------
    // For local blocks whose reduced-operator matvec tiles badly inside the persistent kernel (see host_krylov_wanted):
    // the matvec goes through ttipm_block_matvec (grouped contraction GEMMs over the whole machine above its flop
    // threshold), the vector algebra through ttipm_cgs_project / ttipm_lincomb / k_ewise, the Givens-rotated Hessenberg
    // matrix lives on the host; one synchronising read-back per inner step.
    struct ReducedOp {
        Terms tA, tB;
        Tensor inv_I, xB;
        long r, n, R, m;
        int nred;
    };
    ReducedOp reduced_op(int k, const Tensor& inv_I, long r, long n, long R) {
        ReducedOp op;
        op.r = r; op.n = n; op.R = R; op.m = r * n * R; op.nred = ineq ? 3 : 2;
        op.inv_I = inv_I;
        auto P1 = [&](Key q) -> const Tensor& { return XAX[k].at(q); };
        auto P2 = [&](Key q) -> const Tensor& { return XAX[k + 1].at(q); };
        auto Ak = [&](Key q) -> const Tensor& { return A.at(q)[k]; };
        // phase A: out0 = K00 v0 + K01 v1, out1 (raw) = K01^T v0, [out2 = K31 v1 + K33 v2]
        op.tA.add(P1({0, 0}), Ak({0, 0}), P2({0, 0}), 0, 0);
        op.tA.add(P1({0, 1}), Ak({0, 1}), P2({0, 1}), 1, 0);
        op.tA.add(P1({0, 1}).permute({2, 1, 0}), Ak({0, 1}).permute({0, 2, 1, 3}), P2({0, 1}).permute({2, 1, 0}), 0, 1);
        if (ineq) {
            op.tA.add(P1({3, 1}), Ak({3, 1}), P2({3, 1}), 1, 2);
            op.tA.add(P1({3, 3}), Ak({3, 3}), P2({3, 3}), 2, 2);
        }
        // phase B on xB = [v1 ; inv_I .* out1 (+ v2)]: out1 = K21 v1 - K22 xB1
        op.tB.add(P1({2, 1}), Ak({2, 1}), P2({2, 1}), 0, 0);
        op.tB.add(P1({2, 2}), Ak({2, 2}), P2({2, 2}), 1, 0, -1.0);
        op.xB = Tensor::empty(c, {2, r, n, R});
        return op;
    }
    void reduced_apply(ReducedOp& op, const Tensor& src, Tensor& dst) {       // src, dst: (nred, r, n, R) contiguous
        const long m = op.m, nR = op.n * op.R;
        check_rc(ttipm_block_matvec(op.tA.v.data(), (int)op.tA.v.size(), (int)op.r, (int)op.R, (int)op.r, (int)op.R, (int)op.n,
                                    op.nred, src.p, m, nR, op.R, 0, dst.p, m, nR, op.R, 0, 1.0, nullptr, 0.0, nullptr, 1, c.st),
                 "host krylov: phase A");
        Tensor d1 = dst.select(0, 1), s1 = src.select(0, 1), x0 = op.xB.select(0, 0), x1 = op.xB.select(0, 1);
        ewise(c, s1, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &x0, nullptr);
        if (ineq) {
            Tensor s2 = src.select(0, 2);
            ewise(c, d1, 1.0, nullptr, 0.0, &s2, 1.0, &op.inv_I, &x1, nullptr);
        } else {
            ewise(c, d1, 1.0, nullptr, 0.0, nullptr, 0.0, &op.inv_I, &x1, nullptr);
        }
        check_rc(ttipm_block_matvec(op.tB.v.data(), (int)op.tB.v.size(), (int)op.r, (int)op.R, (int)op.r, (int)op.R, (int)op.n, 1,
                                    op.xB.p, m, nR, op.R, 0, dst.p + m, m, nR, op.R, 0, 1.0, nullptr, 0.0, nullptr, 1, c.st),
                 "host krylov: phase B");
        c.launches += 2;
    }

    Tensor host_lgmres(int k, const Tensor& inv_I, const Tensor& bvec, long r, long n, long R, int max_k, int aug_dim, int max_it,
                       double rtol, bool apply_only, Tensor* info_out) {
        enum { R_NONE = 0, R_RTOL = 1, R_ATOL = 2, R_ITS = 3, R_DTOL = -1, R_BREAKDOWN = -2, R_NULL = -3, R_NAN = -4 };
        ReducedOp op = reduced_op(k, inv_I, r, n, R);
        const long nv = (long)op.nred * op.m;
        Tensor x = Tensor::empty(c, {(long)op.nred, r, n, R});
        if (apply_only) {
            reduced_apply(op, bvec, x);
            return x;
        }
        const double abstol = 1e-50, dtol = 1e5, haptol = 1e-30;
        const int ldh = max_k + 2, it_arnoldi = max_k - aug_dim, nparts = ttipm_cgs_parts(nv);
        Tensor V = Tensor::empty(c, {(long)max_k + 2, nv});
        Tensor AUG = Tensor::empty(c, {(long)std::max(aug_dim, 1), nv}), AAUG = Tensor::empty(c, {(long)std::max(aug_dim, 1), nv});
        Tensor upd = Tensor::empty(c, {(long)op.nred, r, n, R});
        Tensor partials = Tensor::empty(c, {(long)nparts, 112}), hb = Tensor::empty(c, {112 + 128});
        auto vec = [&](const Tensor& M, long i) { return M.select(0, i).reshape({(long)op.nred, r, n, R}); };
        auto norm_of = [&](const Tensor& v) {
            Tensor ss;
            ewise(c, v, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, nullptr, &ss);
            return sqrt(host_sums(c, {&ss})[0]);
        };
        auto scale_inplace = [&](Tensor v, double f) { ewise(c, v, f, nullptr, 0.0, nullptr, 0.0, nullptr, &v, nullptr); };
        auto lincomb = [&](const std::vector<const double*>& vs, const std::vector<double>& cs, double scale, Tensor& out) {
            check_rc(ttipm_lincomb((int)vs.size(), vs.data(), cs.data(), scale, nullptr, 0.0, out.p, nv, c.st), "lincomb");
            c.launches++;
        };
        std::vector<double> hh((size_t)ldh * (max_k + 1), 0.0), hes((size_t)ldh * (max_k + 1), 0.0);
        std::vector<double> cc(ldh, 0.0), ss(ldh, 0.0), grs(ldh, 0.0), nrs(ldh, 0.0), col(ldh, 0.0), tS(ldh, 0.0);
        if (dev_memset(x.p, 0, sizeof(double) * (size_t)nv, c.st)) throw DriverError(92, "memset failed");
        int its = 0, matvecs = 0, aug_ct = 0, reason = R_NONE, cycles = 0;
        int aug_order[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        double ttol = 0.0, rnorm0 = 0.0, res = 0.0;
        bool first = true;
        auto converged = [&](int it, double rn) {
            if (it == 0) {
                ttol = std::max(rtol * rn, abstol);
                rnorm0 = rn;
            }
            if (!(rn == rn) || std::isinf(rn)) return (int)R_NAN;
            if (rn <= ttol) return rn < abstol ? (int)R_ATOL : (int)R_RTOL;
            if (rn >= dtol * rnorm0) return (int)R_DTOL;
            return (int)R_NONE;
        };
        auto aug_spot = [&](int order) {
            for (int ii = 0; ii < aug_dim; ++ii)
                if (aug_order[ii] == order) return ii;
            return 0;
        };
        for (;;) {
            ++cycles;
            Tensor V0 = vec(V, 0);
            double res_norm;
            if (first) {
                Tensor sq;
                ewise(c, bvec, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &V0, &sq);
                res_norm = sqrt(host_sums(c, {&sq})[0]);
            } else {
                reduced_apply(op, x, V0);
                ++matvecs;
                Tensor sq;
                ewise(c, bvec, 1.0, &V0, -1.0, nullptr, 0.0, nullptr, &V0, &sq);
                res_norm = sqrt(host_sums(c, {&sq})[0]);
            }
            first = false;
            res = res_norm;
            grs[0] = res_norm;
            if (res == 0.0) {
                reason = R_ATOL;
                break;
            }
            scale_inplace(V0, 1.0 / res_norm);
            const int it_total = it_arnoldi + aug_ct;
            reason = converged(its, res);
            int loc_it = 0;
            bool hapend = false;
            double grs_loc = res_norm;
            while (reason == R_NONE && loc_it < it_total && its < max_it) {
                Tensor Vn = vec(V, loc_it + 1);
                if (loc_it < it_arnoldi) {
                    reduced_apply(op, vec(V, loc_it), Vn);
                    ++matvecs;
                } else {
                    Tensor src = vec(AAUG, aug_spot(loc_it - it_arnoldi + 1));
                    ewise(c, src, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &Vn, nullptr);
                }
                // classical Gram-Schmidt, one pass: h = V^T w, w -= V h, ||w||
                check_rc(ttipm_cgs_project(V.p, nv, loc_it + 1, Vn.p, nv, partials.p, hb.p, hb.p + 112, c.st), "cgs_project");
                c.launches += 2;
                std::vector<double> hv = read_vec(c, hb);
                double s2 = 0.0;
                for (int g = 0; g < nparts; ++g) s2 += hv[112 + g];
                const double tt = sqrt(s2);
                double hapbnd = fabs(tt / grs_loc);
                if (hapbnd > haptol) hapbnd = haptol;
                if (tt > hapbnd) scale_inplace(Vn, 1.0 / tt);
                else hapend = true;
                double* hcol = hh.data() + (size_t)loc_it * ldh;
                double* ucol = hes.data() + (size_t)loc_it * ldh;
                for (int i = 0; i <= loc_it + 1; ++i) {
                    const double v = i <= loc_it ? hv[i] : tt;
                    col[i] = v;
                    ucol[i] = v;
                }
                for (int j = 0; j < loc_it; ++j) {
                    const double t0 = col[j];
                    col[j] = cc[j] * t0 + ss[j] * col[j + 1];
                    col[j + 1] = cc[j] * col[j + 1] - ss[j] * t0;
                }
                double newres = 0.0;
                bool null_pivot = false;
                if (!hapend) {
                    const double t0 = sqrt(col[loc_it] * col[loc_it] + col[loc_it + 1] * col[loc_it + 1]);
                    if (t0 == 0.0) {
                        null_pivot = true;
                    } else {
                        cc[loc_it] = col[loc_it] / t0;
                        ss[loc_it] = col[loc_it + 1] / t0;
                        grs[loc_it + 1] = -(ss[loc_it] * grs[loc_it]);
                        grs[loc_it] = cc[loc_it] * grs[loc_it];
                        col[loc_it] = cc[loc_it] * col[loc_it] + ss[loc_it] * col[loc_it + 1];
                        newres = fabs(grs[loc_it + 1]);
                    }
                }
                for (int i = 0; i <= loc_it + 1; ++i) hcol[i] = col[i];
                if (null_pivot) {
                    reason = R_NULL;
                    break;
                }
                res = newres;
                grs_loc = grs[loc_it + 1];
                ++loc_it;
                ++its;
                reason = converged(its, res);
                if (hapend && reason == R_NONE) {
                    reason = R_BREAKDOWN;
                    break;
                }
            }
            // solution of this cycle
            const int it = loc_it - 1;
            int n_arn = 0, n_aug = 0;
            if (it >= 0) {
                if (it_arnoldi >= it + 1) {
                    n_arn = it + 1;
                } else {
                    n_arn = it_arnoldi;
                    n_aug = it + 1 - it_arnoldi;
                }
                for (int i = 0; i <= it; ++i) tS[i] = grs[i];
                for (int q = it; q >= 0; --q) {
                    const double dgl = hh[(size_t)q * ldh + q];
                    const double yk = (q == it && dgl == 0.0) ? 0.0 : tS[q] / dgl;
                    nrs[q] = yk;
                    for (int i = 0; i < q; ++i) tS[i] -= hh[(size_t)q * ldh + i] * yk;
                }
                std::vector<const double*> vs;
                std::vector<double> cs;
                for (int i = 0; i < n_arn; ++i) {
                    vs.push_back(V.p + (long)i * nv);
                    cs.push_back(nrs[i]);
                }
                for (int ii = 0; ii < n_aug; ++ii) {
                    vs.push_back(AUG.p + (long)aug_spot(ii + 1) * nv);
                    cs.push_back(nrs[n_arn + ii]);
                }
                lincomb(vs, cs, 1.0, upd);
                ewise(c, x, 1.0, &upd, 1.0, nullptr, 0.0, nullptr, &x, nullptr);
            }
            // harvest the error approximation for the next cycle
            if (reason == R_NONE && its < max_it && aug_dim > 0) {
                int spot = 0;
                if (aug_ct == 0) {
                    spot = 0;
                    ++aug_ct;
                } else if (aug_ct < aug_dim) {
                    spot = aug_ct;
                    ++aug_ct;
                } else {
                    for (int ii = 0; ii < aug_dim; ++ii)
                        if (aug_order[ii] == aug_dim) spot = ii;
                }
                const double inv = 1.0 / norm_of(upd);
                std::vector<const double*> vs;
                std::vector<double> cs;
                for (int jj = 0; jj <= it_total; ++jj) {
                    double a = 0.0;
                    for (int ii = std::max(0, jj - 1); ii < it_total; ++ii) a += hes[(size_t)ii * ldh + jj] * nrs[ii];
                    vs.push_back(V.p + (long)jj * nv);
                    cs.push_back(a);
                }
                Tensor aug = vec(AUG, spot), aaug = vec(AAUG, spot);
                ewise(c, upd, inv, nullptr, 0.0, nullptr, 0.0, nullptr, &aug, nullptr);
                lincomb(vs, cs, inv, aaug);
                for (int ii = 0; ii < aug_dim; ++ii) aug_order[ii] += 1;
                aug_order[spot] = 1;
            }
            if (reason != R_NONE) break;
            if (its >= max_it) {
                reason = R_ITS;
                break;
            }
        }
        if (info_out) {
            const double iv[6] = {(double)its, (double)matvecs, (double)reason, (double)cycles, res, -1.0};   // grid -1: host-driven
            from_host(c, iv, 6, info_out->p);
        }
        host_krylov_solves++;
        return x;
    }

    // The persistent kernel tiles the matvec over output columns L of the right interface and keeps the first intermediate
    // T1 (r x Lt x n S) of a tile in shared memory: when ONE column of T1 takes most of an SM's shared memory (left rank x
    // operator rank large: graphm_3 from IPM iteration 3 on, r ~ 60-150 with S ~ 30) it runs on R <= 16-64 CTAs that each
    // re-read the whole left interface per column (measured: 0.6 ms per inner step at r = 130, R = 16, ~0.6 TFLOP/s).
    bool host_krylov_wanted(long r, long n, long R, const Terms& ops, double mv_flops) const {
        if (host_krylov_mode == 0) return false;
        if (host_krylov_mode == 2) return true;
        long Smax = 1;
        for (const ttipm_term& t : ops.v) Smax = std::max<long>(Smax, std::max(t.s, t.S));
        const double t1_col_bytes = 8.0 * (double)r * (double)(n * Smax + 4);
        return mv_flops >= 1e8 && t1_col_bytes * 4.0 > 200.0 * 1024.0;      // fewer than 4 columns per tile
    }

    void log_phase(const char* what) {             // TTIPM_AMEN_LOG: drain the stream, say which phase just finished
        if (!amen_log_on()) return;
        double probe = 0.0;
        Tensor t = Tensor::empty(c, {1});
        dev_memset(t.p, 0, sizeof(double), c.st);
        to_host(c, t.p, 1, &probe);
        fprintf(stderr, "[amen]   %.1f ms: %s\n", amen_wall_ms(), what);
    }

    LocalOut solve_local(int k, const Tensor& prev, int size_limit, bool dense_ok, const Terms& full) {
        const long r = prev.d[0], n = prev.d[2], R = prev.d[3], m = r * n * R;
        const double rtol = 1e-5;
        LocalOut o;
        o.rhs = rhs_block(k, Xb[k], Xb[k + 1], r, n, R);
        Tensor rhs_ss, res_ss;
        ewise(c, o.rhs, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, nullptr, &rhs_ss);
        Tensor inv_I = local_diag_inv(c, XAX[k].at({1, 2}), A.at({1, 2})[k], XAX[k + 1].at({1, 2}));
        log_phase("rhs block, diagonal");
        block_matvec(c, full, prev, false, bs, r, R, &o.rhs, 1.0, -1.0, &res_ss);
        std::vector<double> sums = host_sums(c, {&rhs_ss, &res_ss});
        log_phase("residual of the previous solution");
        o.norm_rhs = std::max(sqrt(sums[0]), 1e-10);
        o.res_old = sqrt(sums[1]) / o.norm_rhs;
        const double limit = ineq ? 0.95 * size_limit : (double)size_limit;
        bool dense_now = (sqrt((double)(r * R)) <= limit) && dense_ok && (o.res_old >= rtol);
        bool direct_fail = !dense_now;
        if (dense_now) {
            try {
                o.sol = dense_solve(k, o.rhs, inv_I, r, n, R);
            } catch (const DriverError& e) {
                if (e.code != 94 && e.code != 93) throw;
                direct_fail = true;
            }
        }
        if (!dense_now || direct_fail) {
            const int nred = ineq ? 3 : 2;
            const int src[3] = {0, 1, 3};
            Terms ops;      // K00, K01, K21, K22, K31, K33
            const Key keys[6] = {{0, 0}, {0, 1}, {2, 1}, {2, 2}, {3, 1}, {3, 3}};
            for (int q = 0; q < (ineq ? 6 : 4); ++q) ops.add(XAX[k].at(keys[q]), A.at(keys[q])[k], XAX[k + 1].at(keys[q]), 0, 0);
            const ttipm_term* T = ops.v.data();
            const int restart = (int)std::min<long>(m, 100), aug = std::max(restart / 10, 3);
            Tensor ws = Tensor::empty(c, {(long)ttipm_lgmres_workspace(ineq, (int)r, (int)R, (int)n, restart, aug)});
            double mvf = 0.0;
            for (int q = 0; q < (ineq ? 6 : 4); ++q) {
                const double s_ = (double)ops.v[q].s, S_ = (double)ops.v[q].S;
                mvf += (q == 1 ? 2.0 : 1.0) * (2.0 * r * n * R * R * S_ + 2.0 * r * R * s_ * n * n * S_ + 2.0 * r * n * R * r * s_);
            }
            const bool on_host = host_krylov_wanted(r, n, R, ops, mvf);
            auto lg = [&](const Tensor& in, bool apply_only, Tensor* info) {
                if (on_host) {
                    ProfScope ps(c, CAT_KRYLOV, 0.0);
                    return host_lgmres(k, inv_I, in, r, n, R, restart, aug, 300, rtol, apply_only, info);
                }
                Tensor out = Tensor::empty(c, {(long)nred, r, n, R});
                ProfScope ps(c, CAT_KRYLOV, 0.0);         // work is known after the solve (inner steps): see lg_flops
                check_rc(ttipm_local_lgmres(ineq, T + 0, T + 1, T + 2, T + 3, ineq ? T + 4 : nullptr, ineq ? T + 5 : nullptr,
                                            inv_I.p, (int)r, (int)R, (int)n, in.p, out.p, ws.p, ws.numel(), restart, aug, 300,
                                            rtol, apply_only ? 1 : 0, 0, info ? info->p : nullptr, c.st), "local_lgmres");
                c.launches++;
                return out;
            };
            Tensor lrhs = Tensor::empty(c, {(long)nred, r, n, R});
            Tensor l0 = lrhs.select(0, 0), l1 = lrhs.select(0, 1);
            ewise(c, o.rhs.select(1, 0), 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &l0, nullptr);
            Tensor t = Tensor::empty(c, {r, n, R});
            ewise(c, o.rhs.select(1, 1), 1.0, nullptr, 0.0, nullptr, 0.0, &inv_I, &t, nullptr);
            Tensor k22t = apply_block(XAX[k].at({2, 2}), A.at({2, 2})[k], XAX[k + 1].at({2, 2}), t);
            ewise(c, o.rhs.select(1, 2), 1.0, &k22t, -1.0, nullptr, 0.0, nullptr, &l1, nullptr);
            if (ineq) {
                Tensor l2 = lrhs.select(0, 2);
                ewise(c, o.rhs.select(1, 3), 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &l2, nullptr);
            }
            Tensor prev_red = Tensor::empty(c, {(long)nred, r, n, R});
            for (int q = 0; q < nred; ++q) {
                Tensor dst = prev_red.select(0, q);
                ewise(c, prev.select(1, src[q]), 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, &dst, nullptr);
            }
            log_phase("reduced right-hand side");
            Tensor lvec;
            try {
                lvec = lg(prev_red, true, nullptr);
                log_phase("reduced operator applied to the previous solution");
            } catch (const DriverError& e) {
                // a local block the Krylov kernel cannot take (e.g. shared-memory plan): like any exception inside the
                // reference's local solver -- keep previous_solution, flag direct_solve_failure, let the sweep go on
                krylov_failures++;
                last_krylov_error = e.what();
                o.sol = prev;
                o.res_new = o.res_old;
                direct_solve_failure = true;
                local_solves++;
                return o;
            }
            Tensor n0, n1, diff = Tensor::empty(c, {(long)nred, r, n, R});
            ewise(c, lrhs, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, nullptr, &n0);
            ewise(c, lrhs, 1.0, &lvec, -1.0, nullptr, 0.0, nullptr, &diff, &n1);
            std::vector<double> nn = host_sums(c, {&n0, &n1});
            const bool use_prev = sqrt(nn[1]) < sqrt(nn[0]);
            Tensor info = Tensor::empty(c, {6});
            bool krylov_failed = false;
            LgProf pr;
            pr.nv = (double)nred * m;
            pr.restart = restart;
            pr.mv_flops = 0.0;
            for (int q = 0; q < (ineq ? 6 : 4); ++q) {
                const double s_ = (double)ops.v[q].s, S_ = (double)ops.v[q].S;
                const double f = 2.0 * r * n * R * R * S_ + 2.0 * r * R * s_ * n * n * S_ + 2.0 * r * n * R * r * s_;
                pr.mv_flops += (q == 1 ? 2.0 : 1.0) * f;       // K01 is applied both ways
            }
#ifndef TTIPM_EMU
            if (profile) {
                cudaEventCreate(&pr.e0);
                cudaEventCreate(&pr.e1);
                cudaEventRecord(pr.e0, c.st);
            }
#endif
            Tensor xs;
            try {
                xs = lg(use_prev ? diff : lrhs, false, &info);
            } catch (const DriverError& e) {
                // the reference catches every exception of the local solve, keeps previous_solution and flags
                // direct_solve_failure (src/tt_ipm.py:262-280, :379-399); the sweep goes on and the restart ladder of
                // tt_restarted_block_amen decides
                krylov_failures++;
                last_krylov_error = e.what();
                krylov_failed = true;
            }
#ifndef TTIPM_EMU
            if (profile) cudaEventRecord(pr.e1, c.st);
#endif
            log_phase("Krylov solve");
            if (krylov_failed) {
#ifndef TTIPM_EMU
                if (profile) {
                    cudaEventDestroy(pr.e0);
                    cudaEventDestroy(pr.e1);
                }
#endif
                o.sol = prev;
                o.res_new = o.res_old;
                direct_solve_failure = true;
                local_solves++;
                return o;
            }
            lg_prof.push_back(pr);
            lg_infos.push_back(info);
            lgmres_calls++;
            o.sol = Tensor::empty(c, {r, (long)bs, n, R});
            for (int q = 0; q < nred; ++q) {
                Tensor dst = o.sol.select(1, src[q]);
                Tensor pv = prev.select(1, src[q]);
                ewise(c, xs.select(0, q), 1.0, use_prev ? &pv : nullptr, 1.0, nullptr, 0.0, nullptr, &dst, nullptr);
            }
            Tensor y3 = contiguous3(o.sol.select(1, 0));
            Tensor kty = apply_block(XAX[k].at({0, 1}).permute({2, 1, 0}), A.at({0, 1})[k].permute({0, 2, 1, 3}),
                                     XAX[k + 1].at({0, 1}).permute({2, 1, 0}), y3);
            Tensor dst = o.sol.select(1, 2);
            Tensor tcol = ineq ? o.sol.select(1, 3) : Tensor();
            ewise(c, o.rhs.select(1, 1), 1.0, &kty, -1.0, ineq ? &tcol : nullptr, -1.0, &inv_I, &dst, nullptr);
        }
        Tensor new_ss;
        block_matvec(c, full, o.sol, false, bs, r, R, &o.rhs, 1.0, -1.0, &new_ss);
        std::vector<double> s2 = host_sums(c, {&new_ss});
        log_phase("residual of the new solution");
        const double res_new = sqrt(s2[0]) / o.norm_rhs;
        if (!(res_new <= o.res_old)) o.sol = prev;           // also keeps prev when res_new is not finite
        o.res_new = std::isfinite(res_new) ? std::min(o.res_old, res_new) : o.res_old;
        direct_solve_failure = direct_fail;
        local_solves++;
        return o;
    }

    // ---- interface updates (reference src/tt_als.py:372-387, :499-514) ---------------------------------------
    void update_interfaces(int k, bool bck, bool zside) {
        const int src = bck ? k + 1 : k, dst = bck ? k : k + 1;
        const Tensor& xk = x[k];
        std::vector<Tensor> Xs, Bc;
        std::vector<int> rows;
        std::vector<Tensor> phis, cores;
        std::vector<Key> keys;
        const std::vector<KeyMap>& PH = zside ? ZAX : XAX;
        for (auto& kv : A) {
            keys.push_back(kv.first);
            phis.push_back(PH[src].at(kv.first));
            cores.push_back(kv.second[k]);
        }
        if (zside)
            for (auto& tr : transposes) {
                keys.push_back(tr.second);
                phis.push_back(ZAX[src].at(tr.second));
                cores.push_back(A.at(tr.first)[k].permute({0, 2, 1, 3}));
            }
        const Tensor& left = zside ? z[k] : xk;
        std::vector<Tensor> outs = phi_update(c, phis, cores, left, xk, !bck);
        KeyMap nm;
        for (size_t q = 0; q < keys.size(); ++q) nm[keys[q]] = outs[q];
        (zside ? ZAX : XAX)[dst] = nm;
        const std::vector<RowMap>& XB = zside ? Zb : Xb;
        for (auto& kv : b) {
            rows.push_back(kv.first);
            Xs.push_back(XB[src].at(kv.first));
            Bc.push_back(kv.second[k]);
        }
        std::vector<Tensor> ro = phi_rhs_update(c, Xs, Bc, left, !bck);
        RowMap rm;
        for (size_t q = 0; q < rows.size(); ++q) rm[rows[q]] = ro[q];
        (zside ? Zb : Xb)[dst] = rm;
    }

    // ---- one half sweep --------------------------------------------------------------------------------
    void sweep(int direction, int swp, bool last, double& local_res, double& local_dx) {
        const bool bck = direction > 0;
        local_res = swp == 0 ? INFINITY : 0.0;
        local_dx = swp == 0 ? INFINITY : 0.0;
        const bool solving = swp > 0 && !last;
        for (int step = 0; step < d; ++step) {
            const int k = bck ? d - 1 - step : step;
            const bool inner = bck ? k > 0 : k < d - 1;
            const long n = N[k], r_k = rx[k], R_k = rx[k + 1];
            Tensor sol, resz, rhs;
            double r_new = 0.0, norm_rhs = 1.0;
            Terms full;
            const double t_step = amen_log_on() ? amen_wall_ms() : 0.0;
            if (solving) {
                Tensor prev = x[k];
                full = full_terms(k);
                const long lg0 = lgmres_calls, dn0 = dense_solves, kf0 = krylov_failures;
                LocalOut lo = solve_local(k, prev, 3 * d, !direct_solve_failure, full);
                if (amen_log_on()) {
                    std::vector<double> inf(6, 0.0);
                    if (lgmres_calls > lg0) inf = read_vec(c, lg_infos.back());
                    fprintf(stderr, "[amen] swp %d k %d r %ld R %ld n %ld: local solve %.1f ms (%s%s) res %.3e -> %.3e"
                                    " krylov its %.0f matvecs %.0f reason %.0f grid %.0f\n",
                            swp, k, r_k, R_k, n, amen_wall_ms() - t_step, dense_solves > dn0 ? "dense " : "",
                            lgmres_calls > lg0 ? "krylov" : (krylov_failures > kf0 ? "krylov FAILED" : ""), lo.res_old,
                            lo.res_new, inf[0], inf[1], inf[2], inf[5]);
                }
                sol = lo.sol; rhs = lo.rhs; r_new = lo.res_new; norm_rhs = lo.norm_rhs;
                trace.insert(trace.end(), {(double)swp, (double)k, lo.res_old, lo.res_new, (double)(r_k * R_k)});
                local_res = std::max(local_res, lo.res_old);
                Tensor dnum, dden;
                ewise(c, sol, 1.0, &prev, -1.0, nullptr, 0.0, nullptr, nullptr, &dnum);
                ewise(c, sol, 1.0, nullptr, 0.0, nullptr, 0.0, nullptr, nullptr, &dden);
                std::vector<double> dd = host_sums(c, {&dnum, &dden});
                local_dx = std::max(local_dx, sqrt(dd[0]) / sqrt(dd[1]));
                if (amen) {
                    Tensor rhsz = rhs_block(k, Zb[k], Zb[k + 1], rz[k], n, rz[k + 1]);
                    resz = block_matvec(c, mixed_terms(k, ZAX[k], ZAX[k + 1], true, true), sol, false, bs, rz[k], rz[k + 1],
                                        &rhsz, -1.0, 1.0, nullptr);
                }
            } else {
                sol = x[k];
                if (amen && !last) resz = z[k];
            }
            Tensor scales = block_norms(c, sol);
            Tensor S, mat, rzm;
            if (bck) {
                S = permute4(c, sol, 0, 1, 2, 3, &scales, 1, false);                 // (r, b, n, R)
                mat = S.reshape({r_k * bs, n * R_k}).t2();                            // (n R, r b) view
                if (resz.defined()) rzm = resz.reshape({rz[k] * bs, n * rz[k + 1]}).t2();
            } else {
                S = permute4(c, sol, 0, 2, 1, 3, &scales, 2, false);                 // (r, n, b, R)
                mat = S.reshape({r_k * n, bs * R_k});
                if (resz.defined()) rzm = permute4(c, resz, 0, 2, 1, 3, nullptr, 0, false).reshape({rz[k] * n, bs * rz[k + 1]});
            }
            if (!inner) {
                x[k] = bck ? permute4(c, S, 0, 1, 2, 3, &scales, 1, true) : permute4(c, S, 0, 2, 1, 3, &scales, 1, true);
                if (amen && !last) z[k] = permute4(c, resz, 0, 1, 2, 3, &scales, 1, true);
                continue;
            }
            Tensor U, Sv, W;
            svd_left(c, mat, U, Sv, W);
            std::vector<double> s_host = read_vec(c, Sv);
            const long Kk = U.d[1];
            long r;
            Tensor uk, vk;
            if (solving) {
                const double trunc_lim = std::max(2 * trunc_tol, r_new);
                const long r0 = std::min<long>(prune_singular_vals(s_host, eps), r_max);
                Tensor sol_r0;
                if (bck) sol_r0 = gemm(c, W.slice(0, 0, r0).t2(), U.slice(1, 0, r0).t2()).reshape({r_k, (long)bs, n, R_k});
                else sol_r0 = gemm(c, U.slice(1, 0, r0), W.slice(0, 0, r0)).reshape({r_k, n, (long)bs, R_k});
                Tensor res = block_matvec(c, full, sol_r0, !bck, bs, r_k, R_k, &rhs, 1.0, -1.0, nullptr);
                r = r0;
                if (r0 > 1) {
                    Tensor terms;
                    if (bck) terms = gemm(c, W.slice(0, 1, r0).unsqueeze(2), U.slice(1, 1, r0).t2().unsqueeze(1))
                                         .reshape({r0 - 1, r_k, (long)bs, n, R_k});
                    else terms = gemm(c, U.slice(1, 1, r0).t2().unsqueeze(2), W.slice(0, 1, r0).unsqueeze(1))
                                     .reshape({r0 - 1, r_k, n, (long)bs, R_k});
                    Tensor Y = block_matvec(c, full, terms, !bck, bs, r_k, R_k, nullptr, 1.0, 0.0, nullptr);
                    Tensor parts = Tensor::empty(c, {r0 - 1, 256});
                    ProfScope ps(c, CAT_EWISE, 8.0 * (double)(res.numel() + Y.numel()));
                    check_rc(ttipm_trunc_resnorms(res.p, Y.p, (int)(r0 - 1), res.numel(), parts.p, c.st), "trunc_resnorms");
                    c.launches++;
                    std::vector<double> ph = read_vec(c, parts);
                    r = 1;
                    for (long q = r0 - 1; q >= 1; --q) {
                        if (sqrt(sum_host(ph, (size_t)(q - 1) * 256, 256)) / norm_rhs > trunc_lim) {
                            r = q;
                            break;
                        }
                    }
                }
                r += 1;
                r = std::min(r, Kk);
                uk = U.slice(1, 0, r);
                vk = W.slice(0, 0, r);
                if (amen) {
                    Tensor Uz, Sz, Wz;
                    long kr;
                    if (bck) {
                        Tensor rhsxz = rhs_block(k, Zb[k], Xb[k + 1], rz[k], n, R_k);
                        Tensor resxz = block_matvec(c, mixed_terms(k, ZAX[k], XAX[k + 1], true, false), sol_r0, false, bs,
                                                    rz[k], R_k, &rhsxz, -1.0, 1.0, nullptr);
                        kr = std::min<long>(kick_rank, std::min(rz[k] * bs, n * R_k));
                        svd_left(c, resxz.reshape({rz[k] * bs, n * R_k}).t2(), Uz, Sz, Wz);
                    } else {
                        Tensor sol_r = gemm(c, uk, vk).reshape({r_k, n, (long)bs, R_k});
                        Tensor rhsxz = rhs_block(k, Xb[k], Zb[k + 1], r_k, n, rz[k + 1]);
                        Tensor resxz = block_matvec(c, mixed_terms(k, XAX[k], ZAX[k + 1], false, true), sol_r, true, bs, r_k,
                                                    rz[k + 1], &rhsxz, -1.0, 1.0, nullptr);
                        kr = std::min<long>(kick_rank, std::min(r_k * n, bs * rz[k + 1]));
                        svd_left(c, permute4(c, resxz, 0, 2, 1, 3, nullptr, 0, false).reshape({r_k * n, bs * rz[k + 1]}), Uz, Sz, Wz);
                    }
                    Tensor cat = Tensor::empty(c, {U.d[0], r + kr});
                    copy_rows(uk, cat.slice(1, 0, r));
                    copy_rows(Uz.slice(1, 0, kr), cat.slice(1, r, r + kr));
                    Tensor Q, Rf;
                    qr(c, cat, Q, Rf);
                    vk = gemm(c, Rf.slice(1, 0, r), vk);
                    uk = Q;
                    r = uk.d[1];
                }
            } else {
                r = std::min<long>(prune_singular_vals(s_host, eps), r_max);
                uk = U.slice(1, 0, r);
                vk = W.slice(0, 0, r);
            }
            if (bck) {
                x[k] = copy2d(c, uk.t2()).reshape({r, n, R_k});
                Tensor vT = copy2d(c, vk.t2());                                         // (r_k * bs, r)
                const long a = x[k - 1].d[0], dd = x[k - 1].d[1], cc = x[k - 1].d[2];
                Tensor G = gemm(c, x[k - 1].reshape({a * dd, cc}), vT.reshape({cc, bs * r})).reshape({a, dd, (long)bs, r});
                x[k - 1] = permute4(c, G, 0, 2, 1, 3, &scales, 1, true);
                rx[k] = r;
            } else {
                x[k] = copy2d(c, uk).reshape({r_k, n, r});
                const long Rn = x[k + 1].d[0], dd = x[k + 1].d[1], k2 = x[k + 1].d[2];
                Tensor vc = vk.contiguous() ? vk : copy2d(c, vk);
                Tensor G = gemm(c, vc.reshape({r * bs, Rn}), x[k + 1].reshape({Rn, dd * k2}));
                x[k + 1] = permute4(c, G.reshape({r, (long)bs, dd, k2}), 0, 1, 2, 3, &scales, 1, true);
                rx[k + 1] = r;
            }
            update_interfaces(k, bck, false);
            if (amen_log_on()) {
                double probe = 0.0;
                to_host(c, x[k].p, 1, &probe);             // drain the stream: the time below is the step's device time
                fprintf(stderr, "[amen] swp %d k %d: step %.1f ms, unfolding %ld x %ld, new rank %ld\n", swp, k,
                        amen_wall_ms() - t_step, mat.d[0], mat.d[1], r);
            }
            if (amen && !last) {
                const long kr = std::min<long>(kick_rank, std::min(rzm.d[0], rzm.d[1]));
                Tensor Uz, Sz, Wz;
                svd_left(c, rzm, Uz, Sz, Wz);
                if (bck) {
                    z[k] = copy2d(c, Uz.slice(1, 0, kr).t2()).reshape({kr, n, rz[k + 1]});
                    Tensor vT = copy2d(c, Wz.slice(0, 0, kr).t2());
                    const long a = z[k - 1].d[0], dd = z[k - 1].d[1], cc = z[k - 1].d[2];
                    Tensor G = gemm(c, z[k - 1].reshape({a * dd, cc}), vT.reshape({cc, bs * kr})).reshape({a, dd, (long)bs, kr});
                    z[k - 1] = permute4(c, G, 0, 2, 1, 3, &scales, 1, true);
                    rz[k] = kr;
                } else {
                    z[k] = copy2d(c, Uz.slice(1, 0, kr)).reshape({rz[k], n, kr});
                    const long Rn = z[k + 1].d[0], dd = z[k + 1].d[1], k2 = z[k + 1].d[2];
                    Tensor wc = Wz.slice(0, 0, kr);
                    Tensor G = gemm(c, wc.reshape({kr * bs, Rn}), z[k + 1].reshape({Rn, dd * k2}));
                    z[k + 1] = permute4(c, G.reshape({kr, (long)bs, dd, k2}), 0, 1, 2, 3, &scales, 1, true);
                    rz[k + 1] = kr;
                }
                update_interfaces(k, bck, true);
            }
        }
    }

    void copy_rows(const Tensor& src, Tensor dst) {      // 2-D views with unit column stride
        check_rc(ttipm_ewise((int)src.d[0], (int)src.d[1], 1.0, src.p, src.s[0], 0.0, nullptr, 0, 0.0, nullptr, 0, nullptr, 0,
                             dst.p, dst.s[0], nullptr, c.st), "copy_rows");
        c.launches++;
    }

    // ---- whole solve (reference src/tt_als.py:586-670) -------------------------------------------------------
    double run(double term_tol, int rmax, double eps_, int nswp, int direction) {
        r_max = rmax;
        eps = eps_;
        trunc_tol = term_tol / sqrt((double)d);
        direct_solve_failure = false;
        bool last = false;
        double final_res = INFINITY;
        sweeps = 0;
        for (int swp = 0; swp <= nswp; ++swp) {
            double local_res, local_dx;
            sweep(direction, swp, last, local_res, local_dx);
            sweeps = swp;
            if (last) break;
            if (local_res < term_tol || local_dx < eps || swp == nswp - 2) {
                last = true;
                final_res = local_res;
            }
            direction = -direction;
        }
        return final_res;
    }
};

}  // namespace drv
}  // namespace ttipm

using namespace ttipm;
using namespace ttipm::drv;

struct ttipm_amen {
    Amen a;
};

#define TT_TRY(h, ...)                                   \
    try {                                                \
        __VA_ARGS__;                                     \
        return 0;                                        \
    } catch (const DriverError& e) {                     \
        (h)->a.error = e.what();                         \
        return fail(e.code ? e.code : 1, "%s", e.what()); \
    } catch (const std::exception& e) {                  \
        (h)->a.error = e.what();                         \
        return fail(99, "%s", e.what());                 \
    }

extern "C" ttipm_amen* ttipm_amen_create(int d, int block_size, int ineq, void* stream) {
    if (check_bound_device()) return nullptr;          // process-wide handles / staging belong to one device
    pool_keep_freed_blocks();
    ttipm_amen* h = new ttipm_amen();
    h->a.c.st = (tt_stream_t)stream;
    h->a.d = d;
    h->a.bs = block_size;
    h->a.ineq = ineq != 0;
    h->a.x.resize(d);
    h->a.z.resize(d);
    h->a.rx.assign(d + 1, 1);
    h->a.rz.assign(d + 1, 1);
    h->a.N.assign(d, 4);
    return h;
}

extern "C" void ttipm_amen_destroy(ttipm_amen* h) {
    if (!h) return;
    h->a.c.pinned = nullptr;          // the staging buffer is process wide (ensure_pinned)
    delete h;
}

static Tensor upload(Ctx& c, const double* host, std::initializer_list<long> dims) {
    Tensor t = Tensor::empty(c, dims);
    from_host(c, host, (size_t)t.numel(), t.p);
    return t;
}

extern "C" int ttipm_amen_set_block(ttipm_amen* h, int i, int j, int k, const double* core_host, int s, int n, int S) {
    TT_TRY(h, {
        auto& v = h->a.A[{i, j}];
        if ((int)v.size() != h->a.d) v.resize(h->a.d);
        v[k] = upload(h->a.c, core_host, {(long)s, (long)n, (long)n, (long)S});
    });
}

extern "C" int ttipm_amen_add_alias(ttipm_amen* h, int i, int j, int p, int t, int is_transpose) {
    TT_TRY(h, { (is_transpose ? h->a.transposes : h->a.aliases)[{i, j}] = {p, t}; });
}

extern "C" int ttipm_amen_set_rhs(ttipm_amen* h, int i, int k, const double* core_host, int rb, int n, int rb2) {
    TT_TRY(h, {
        auto& v = h->a.b[i];
        if ((int)v.size() != h->a.d) v.resize(h->a.d);
        v[k] = upload(h->a.c, core_host, {(long)rb, (long)n, (long)rb2});
    });
}

// which: 0 = solution train x, 1 = residual train z; nb = 0 for an ordinary core (r, n, R), block size for the block core
extern "C" int ttipm_amen_set_core(ttipm_amen* h, int which, int k, const double* core_host, int r, int nb, int n, int R) {
    TT_TRY(h, {
        Amen& a = h->a;
        Tensor t = nb > 0 ? upload(a.c, core_host, {(long)r, (long)nb, (long)n, (long)R})
                          : upload(a.c, core_host, {(long)r, (long)n, (long)R});
        (which == 0 ? a.x : a.z)[k] = t;
        (which == 0 ? a.rx : a.rz)[k] = r;
        (which == 0 ? a.rx : a.rz)[k + 1] = R;
        a.N[k] = n;
    });
}

extern "C" int ttipm_amen_run(ttipm_amen* h, double term_tol, int r_max, double eps, int nswp, int kick_rank, int use_amen,
                              int direction, double* final_res, int* sweeps) {
    TT_TRY(h, {
        Amen& a = h->a;
        a.amen = use_amen != 0;
        a.kick_rank = kick_rank;
        const int d = a.d;
        a.XAX.assign(d + 1, KeyMap());
        a.Xb.assign(d + 1, RowMap());
        a.ZAX.assign(d + 1, KeyMap());
        a.Zb.assign(d + 1, RowMap());
        for (auto& kv : a.A) {
            a.XAX[0][kv.first] = a.ones({1, 1, 1});
            a.XAX[d][kv.first] = a.ones({1, 1, 1});
            if (a.amen) {
                a.ZAX[0][kv.first] = a.ones({1, 1, 1});
                a.ZAX[d][kv.first] = a.ones({1, 1, 1});
            }
        }
        if (a.amen)
            for (auto& tr : a.transposes) {
                a.ZAX[0][tr.second] = a.ones({1, 1, 1});
                a.ZAX[d][tr.second] = a.ones({1, 1, 1});
            }
        for (auto& kv : a.b) {
            a.Xb[0][kv.first] = a.ones({1, 1});
            a.Xb[d][kv.first] = a.ones({1, 1});
            if (a.amen) {
                a.Zb[0][kv.first] = a.ones({1, 1});
                a.Zb[d][kv.first] = a.ones({1, 1});
            }
        }
        a.trace.clear();
        a.stats_done = false;
        a.c.prof = a.profile;
        const double res = a.run(term_tol, r_max, eps, nswp, direction);
        if (final_res) *final_res = res;
        if (sweeps) *sweeps = a.sweeps;
    });
}

// shape of solution core k: dims[0..3] = (r, nb or 0, n, R)
extern "C" int ttipm_amen_host_krylov(int mode) {
    const int old = g_host_krylov_mode;
    if (mode >= 0 && mode <= 2) g_host_krylov_mode = mode;
    return old;
}

extern "C" int ttipm_amen_set_profile(ttipm_amen* h, int on) {
    h->a.profile = on != 0;
    return 0;
}

extern "C" int ttipm_amen_core_shape(ttipm_amen* h, int k, int32_t* dims) {
    TT_TRY(h, {
        const Tensor& t = h->a.x[k];
        if (t.nd == 4) { dims[0] = (int)t.d[0]; dims[1] = (int)t.d[1]; dims[2] = (int)t.d[2]; dims[3] = (int)t.d[3]; }
        else { dims[0] = (int)t.d[0]; dims[1] = 0; dims[2] = (int)t.d[1]; dims[3] = (int)t.d[2]; }
    });
}

extern "C" int ttipm_amen_get_core(ttipm_amen* h, int k, double* dst_host) {
    TT_TRY(h, {
        const Tensor& t = h->a.x[k];
        to_host(h->a.c, t.p, (size_t)t.numel(), dst_host);
    });
}

// Per-category profile of the last run() made with set_profile(1): out[3 * cat + {0, 1, 2}] = seconds between the
// CUDA events bracketing the category's launches, algorithmic work (flops; bytes for cat 6), launches.  Categories:
// 0 block matvec, 1 interface update, 2 rhs contraction, 3 bond GEMM, 4 QR, 5 SVD, 6 memory-bound helpers,
// 7 dense Schur fallback (cuSOLVER/cuBLAS), 8 Krylov kernel.  Call ttipm_amen_stats first (it resolves the events).
extern "C" int ttipm_amen_profile(ttipm_amen* h, double* out) {
    TT_TRY(h, {
        Amen& a = h->a;
        for (int q = 0; q < CAT_COUNT; ++q) {
            out[3 * q] = a.cat_seconds[q];
            out[3 * q + 1] = a.cat_work[q];
            out[3 * q + 2] = a.cat_launches[q];
        }
    });
}

// stats[0..11]: [10] = seconds inside the Krylov kernel (only with profiling on), [11] = its algorithmic flops
// stats[0..9] = sweeps, local solves, dense solves, Krylov solves, Krylov inner steps, Krylov matvecs, launches,
//               host syncs, peak device bytes, trace rows; trace (may be NULL) receives 5 doubles per local solve
extern "C" int ttipm_amen_stats(ttipm_amen* h, double* stats, double* trace, int max_trace_rows) {
    TT_TRY(h, {
        Amen& a = h->a;
        if (!a.stats_done) {
            for (size_t q = 0; q < a.lg_infos.size(); ++q) {
                std::vector<double> v = read_vec(a.c, a.lg_infos[q]);
                a.lgmres_its += (long)v[0];
                a.lgmres_matvecs += (long)v[1];
                const Amen::LgProf& pr = a.lg_prof[q];
                const double kk = v[0] <= pr.restart ? v[0] * (v[0] + 1) / 2.0 : v[0] * (pr.restart + 1) / 2.0;
                a.lg_flops += v[1] * pr.mv_flops + 4.0 * pr.nv * kk;
#ifndef TTIPM_EMU
                if (a.profile) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, pr.e0, pr.e1);
                    a.lg_time += ms * 1e-3;
                    cudaEventDestroy(pr.e0);
                    cudaEventDestroy(pr.e1);
                }
#endif
            }
#ifndef TTIPM_EMU
            for (const ProfRec& r : a.c.recs) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, (cudaEvent_t)r.e0, (cudaEvent_t)r.e1);
                a.cat_seconds[r.cat] += ms * 1e-3;
                a.cat_work[r.cat] += r.work;
                a.cat_launches[r.cat] += 1.0;
                cudaEventDestroy((cudaEvent_t)r.e0);
                cudaEventDestroy((cudaEvent_t)r.e1);
            }
#endif
            a.c.recs.clear();
            a.cat_work[CAT_KRYLOV] = a.lg_flops;
            a.profile = false;
            a.c.prof = false;
            a.stats_done = true;
        }
        const double its = (double)a.lgmres_its, mv = (double)a.lgmres_matvecs;
        stats[10] = a.lg_time;
        stats[11] = a.lg_flops;
        stats[0] = a.sweeps; stats[1] = (double)a.local_solves; stats[2] = (double)a.dense_solves;
        stats[3] = (double)a.lgmres_calls; stats[4] = its; stats[5] = mv; stats[6] = (double)a.c.launches;
        stats[7] = (double)a.c.syncs; stats[8] = (double)a.c.bytes_peak; stats[9] = (double)(a.trace.size() / 5);
        if (trace) {
            const size_t rows = std::min<size_t>(a.trace.size() / 5, (size_t)max_trace_rows);
            memcpy(trace, a.trace.data(), rows * 5 * sizeof(double));
        }
    });
}
