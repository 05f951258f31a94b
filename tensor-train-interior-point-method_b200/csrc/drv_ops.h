// Kernel wrappers shared by the native drivers (amen_driver.cu: block AMEn sweep; tt_driver.cu: TT algebra):
// the same roles as ttipm_b200/kernels.py, on drv::Tensor views, launching through the C ABI on the driver's stream.
#pragma once
#include <math.h>
#include <string.h>
#include <algorithm>
#include "tensor.h"

namespace ttipm {
namespace drv {

static inline std::vector<double> read_vec(Ctx& c, const Tensor& t) {
    std::vector<double> h((size_t)t.numel());
    to_host(c, t.p, h.size(), h.data());
    return h;
}

// ---------------------------------------------------------------------------------------------------
// kernel wrappers (same roles as ttipm_b200/kernels.py)
// ---------------------------------------------------------------------------------------------------
struct Terms {
    std::vector<ttipm_term> v;
    void add(const Tensor& P1, const Tensor& A, const Tensor& P2, int in_blk, int out_blk, double alpha = 1.0) {
        ttipm_term t;
        t.P1 = P1.p; t.A = A.p; t.P2 = P2.p;
        for (int i = 0; i < 3; ++i) { t.p1_strides[i] = P1.s[i]; t.p2_strides[i] = P2.s[i]; }
        for (int i = 0; i < 4; ++i) t.a_strides[i] = A.s[i];
        t.s = (int)A.d[0]; t.S = (int)A.d[3];
        t.in_block = in_blk; t.out_block = out_blk; t.alpha = alpha;
        v.push_back(t);
    }
};

// x: (r, b, n, R) ["rbnR"] or (r, n, b, R) ["rnbR"], optionally with a leading batch axis
static inline Tensor block_matvec(Ctx& c, const Terms& tl, const Tensor& x, bool rnbR, int nb_out, long l, long L,
                           const Tensor* sub, double y_scale, double sub_scale, Tensor* sumsq) {
    const bool batched = x.nd == 5;
    const int o = batched ? 1 : 0;
    const long B = batched ? x.d[0] : 1;
    const long r = x.d[o], n = rnbR ? x.d[o + 1] : x.d[o + 2], R = x.d[o + 3];
    const long x_rs = x.s[o], x_ns = rnbR ? x.s[o + 1] : x.s[o + 2], x_bs = rnbR ? x.s[o + 2] : x.s[o + 1];
    Tensor y = batched ? Tensor::empty(c, {B, l, (long)nb_out, n, L}) : Tensor::empty(c, {l, (long)nb_out, n, L});
    if (sumsq) *sumsq = Tensor::empty(c, {B, nb_out * L});
    double work = 0.0;
    for (const ttipm_term& t : tl.v)
        work += (double)B * (2.0 * r * n * R * L * t.S + 2.0 * r * L * t.s * n * n * t.S + 2.0 * l * n * L * r * t.s);
    ProfScope ps(c, CAT_MATVEC, work);
    check_rc(ttipm_block_matvec(tl.v.data(), (int)tl.v.size(), (int)l, (int)L, (int)r, (int)R, (int)n, nb_out, x.p, x_bs,
                                x_rs, x_ns, batched ? x.s[0] : 0, y.p, n * L, nb_out * n * L, L, l * nb_out * n * L,
                                y_scale, sub ? sub->p : nullptr, sub_scale, sumsq ? sumsq->p : nullptr, (int)B, c.st),
             "block_matvec");
    c.launches++;
    return y;
}

static inline Tensor local_diag_inv(Ctx& c, const Tensor& P1, const Tensor& A, const Tensor& P2) {
    Terms t;
    t.add(P1, A, P2, 0, 0);
    Tensor out = Tensor::empty(c, {P1.d[0], A.d[1], P2.d[0]});
    ProfScope ps(c, CAT_EWISE, 8.0 * (double)(P1.numel() + A.numel() + P2.numel() + out.numel()));
    check_rc(ttipm_local_diag(t.v.data(), (int)P1.d[0], (int)P2.d[0], (int)A.d[1], 1, out.p, c.st), "local_diag");
    c.launches++;
    return out;
}

static inline Tensor local_dense(Ctx& c, const Tensor& P1, const Tensor& A, const Tensor& P2) {
    Terms t;
    t.add(P1, A, P2, 0, 0);
    const long l = P1.d[0], r = P1.d[2], L = P2.d[0], R = P2.d[2], n = A.d[1];
    Tensor out = Tensor::empty(c, {l * n * L, r * n * R});
    ProfScope ps(c, CAT_DENSE, 2.0 * (double)out.numel() * (double)(A.d[0] + A.d[3]));
    check_rc(ttipm_local_dense(t.v.data(), (int)l, (int)L, (int)r, (int)R, (int)n, out.p, c.st), "local_dense");
    c.launches++;
    return out;
}

static inline std::vector<Tensor> phi_update(Ctx& c, const std::vector<Tensor>& phis, const std::vector<Tensor>& cores,
                                      const Tensor& U, const Tensor& V, bool forward) {
    const size_t n = phis.size();
    std::vector<ttipm_phi_term> t(n);
    std::vector<Tensor> outs(n);
    const long ul = U.d[0], nm = U.d[1], uL = U.d[2], vr = V.d[0], vR = V.d[2];
    for (size_t q = 0; q < n; ++q) {
        const long s = cores[q].d[0], S = cores[q].d[3];
        outs[q] = forward ? Tensor::empty(c, {uL, S, vR}) : Tensor::empty(c, {ul, s, vr});
        t[q].Phi = phis[q].p; t[q].A = cores[q].p; t[q].out = outs[q].p;
        for (int i = 0; i < 4; ++i) t[q].a_strides[i] = cores[q].s[i];
        t[q].s = (int)s; t[q].S = (int)S;
    }
    double work = 0.0;
    for (size_t q = 0; q < n; ++q) {
        const double s_ = (double)cores[q].d[0], S_ = (double)cores[q].d[3];
        // 2 l s r N R + 2 l R s N M S + 2 l M L S R (forward), mirrored for the backward update
        work += forward ? 2.0 * ul * s_ * vr * nm * vR + 2.0 * ul * vR * s_ * nm * nm * S_ + 2.0 * ul * nm * uL * S_ * vR
                        : 2.0 * uL * S_ * vR * nm * vr + 2.0 * uL * vr * S_ * nm * nm * s_ + 2.0 * uL * nm * ul * s_ * vr;
    }
    ProfScope ps(c, CAT_PHI, work);
    check_rc(ttipm_phi_update(t.data(), (int)n, forward ? 1 : 0, U.p, (int)ul, (int)uL, V.p, (int)vr, (int)vR, (int)nm,
                              c.st), "phi_update");
    c.launches++;
    return outs;
}

// out (r, nb, n, R) zero-initialised; rows[q] gives the block of term q
static inline Tensor rhs_project(Ctx& c, const std::vector<Tensor>& X1, const std::vector<Tensor>& Bc,
                          const std::vector<Tensor>& X2, const std::vector<int>& rows, long r, int nb, long n, long R) {
    Tensor out = Tensor::empty(c, {r, (long)nb, n, R});
    if (rows.empty()) {
        if (dev_memset(out.p, 0, sizeof(double) * (size_t)out.numel(), c.st)) throw DriverError(92, "memset failed");
        return out;
    }
    std::vector<ttipm_rhs_term> t(rows.size());
    for (size_t q = 0; q < rows.size(); ++q) {
        t[q].Xb1 = X1[q].p; t[q].B = Bc[q].p; t[q].Xb2 = X2[q].p;
        t[q].out = out.p + rows[q] * out.s[1];
        t[q].b = (int)Bc[q].d[0]; t[q].Bp = (int)Bc[q].d[2];
    }
    double work = 0.0;
    for (size_t q = 0; q < rows.size(); ++q)
        work += 2.0 * t[q].b * r * n * t[q].Bp + 2.0 * r * n * t[q].Bp * R;
    // blocks without a right-hand-side core: "zero terms" (B == NULL), cleared by the kernel itself
    for (int j = 0; j < nb; ++j) {
        bool has = false;
        for (int row : rows) has = has || row == j;
        if (has) continue;
        ttipm_rhs_term z;
        z.Xb1 = nullptr; z.B = nullptr; z.Xb2 = nullptr; z.out = out.p + j * out.s[1]; z.b = 0; z.Bp = 0;
        t.push_back(z);
    }
    ProfScope ps(c, CAT_RHS, work);
    check_rc(ttipm_rhs_contract(t.data(), (int)t.size(), 0, nullptr, (int)r, (int)R, (int)n, out.s[0], c.st), "rhs_project");
    c.launches++;
    return out;
}

static inline std::vector<Tensor> phi_rhs_update(Ctx& c, const std::vector<Tensor>& Xb, const std::vector<Tensor>& Bc,
                                          const Tensor& core, bool forward) {
    const size_t n = Bc.size();
    std::vector<ttipm_rhs_term> t(n);
    std::vector<Tensor> outs(n);
    const long r = core.d[0], nm = core.d[1], R = core.d[2];
    for (size_t q = 0; q < n; ++q) {
        t[q].B = Bc[q].p; t[q].b = (int)Bc[q].d[0]; t[q].Bp = (int)Bc[q].d[2];
        if (forward) { outs[q] = Tensor::empty(c, {Bc[q].d[2], R}); t[q].Xb1 = Xb[q].p; t[q].Xb2 = nullptr; }
        else { outs[q] = Tensor::empty(c, {Bc[q].d[0], r}); t[q].Xb1 = nullptr; t[q].Xb2 = Xb[q].p; }
        t[q].out = outs[q].p;
    }
    if (n) {
        double work = 0.0;
        for (size_t q = 0; q < n; ++q) work += 2.0 * r * nm * R * (forward ? t[q].b : t[q].Bp) + 2.0 * t[q].b * nm * t[q].Bp * (forward ? R : r);
        ProfScope ps(c, CAT_RHS, work);
        check_rc(ttipm_rhs_contract(t.data(), (int)n, forward ? 1 : 2, core.p, (int)r, (int)R, (int)nm, 0, c.st),
                 "phi_rhs_update");
        c.launches++;
    }
    return outs;
}

// C = A @ B for 2-D strided views or batched 3-D views
static inline Tensor gemm(Ctx& c, const Tensor& A, const Tensor& B) {
    const bool batched = A.nd == 3;
    const int o = batched ? 1 : 0;
    const long nb = batched ? A.d[0] : 1, M = A.d[o], K = A.d[o + 1], N = B.d[o + 1];
    Tensor C = batched ? Tensor::empty(c, {nb, M, N}) : Tensor::empty(c, {M, N});
    ProfScope ps(c, CAT_GEMM, 2.0 * nb * M * N * K);
    check_rc(ttipm_gemm((int)M, (int)N, (int)K, 1.0, A.p, A.s[o], A.s[o + 1], batched && nb > 1 ? A.s[0] : 0, B.p, B.s[o],
                        B.s[o + 1], batched && nb > 1 ? B.s[0] : 0, 0.0, C.p, N, 1, M * N, (int)nb, c.st), "gemm");
    c.launches++;
    return C;
}

static inline void qr(Ctx& c, const Tensor& A, Tensor& Q, Tensor& R) {
    const long M = A.d[0], N = A.d[1], K = std::min(M, N);
    Q = Tensor::empty(c, {M, K});
    R = Tensor::empty(c, {K, N});
    Tensor ws = Tensor::empty(c, {(long)ttipm_qr_workspace((int)M, (int)N, 1)});
    // Householder QR with explicit Q (geqrf + orgqr): 4 M N K - (4/3) K^3
    ProfScope ps(c, CAT_QR, 4.0 * M * N * K - 4.0 / 3.0 * K * K * K);
    check_rc(ttipm_qr(A.p, A.s[0], A.s[1], 0, (int)M, (int)N, Q.p, R.p, ws.p, 1, c.st), "qr");
    c.launches++;
}

static inline void svd_left(Ctx& c, const Tensor& A, Tensor& U, Tensor& S, Tensor& W) {
    const long M = A.d[0], N = A.d[1], K = std::min(M, N);
    U = Tensor::empty(c, {M, K});
    S = Tensor::empty(c, {K});
    W = Tensor::empty(c, {K, N});
    Tensor ws = Tensor::empty(c, {(long)ttipm_svd_workspace((int)M, (int)N, 1)});
    // economy SVD with both factors, Golub & Van Loan R-SVD count: 6 m n^2 + 20 n^3 (m >= n)
    ProfScope ps(c, CAT_SVD, 6.0 * std::max(M, N) * K * K + 20.0 * K * K * K);
    check_rc(ttipm_svd_left(A.p, A.s[0], A.s[1], 0, (int)M, (int)N, U.p, S.p, W.p, ws.p, nullptr, 1, c.st), "svd_left");
    c.launches++;
}

// materialised permutation of a contiguous 4-D tensor, optional scaling along an OUTPUT axis
static inline Tensor permute4(Ctx& c, const Tensor& x, int p0, int p1, int p2, int p3, const Tensor* scale, int axis, bool divide) {
    if (!x.contiguous() || x.nd != 4) throw DriverError(90, "permute4 needs a contiguous 4-D tensor");
    int32_t dims[4] = {(int32_t)x.d[0], (int32_t)x.d[1], (int32_t)x.d[2], (int32_t)x.d[3]};
    int32_t perm[4] = {p0, p1, p2, p3};
    Tensor out = Tensor::empty(c, {x.d[p0], x.d[p1], x.d[p2], x.d[p3]});
    ProfScope ps(c, CAT_EWISE, 16.0 * (double)x.numel());
    check_rc(ttipm_permute4(x.p, dims, perm, out.p, scale ? scale->p : nullptr, axis, divide ? 2 : 1, c.st), "permute4");
    c.launches++;
    return out;
}

static inline Tensor block_norms(Ctx& c, const Tensor& x) {
    Tensor out = Tensor::empty(c, {x.d[1]});
    ProfScope ps(c, CAT_EWISE, 8.0 * (double)x.numel());
    check_rc(ttipm_block_norms(x.p, (int)x.d[0], (int)x.d[1], (int)(x.numel() / (x.d[0] * x.d[1])), 1e-10, out.p, c.st),
             "block_norms");
    c.launches++;
    return out;
}

// (rows, inner, stride) view of a contiguous tensor or an x[:, j]-style slice
struct Panel { const double* p; long rs; };
static inline void panels(const Tensor* ts[], int n, long& rows, long& inner, long rs[]) {
    const Tensor* first = nullptr;
    bool all_contig = true;
    for (int i = 0; i < n; ++i)
        if (ts[i]) {
            if (!first) first = ts[i];
            all_contig = all_contig && ts[i]->contiguous();
        }
    if (all_contig) {
        rows = 1;
        inner = first->numel();
        for (int i = 0; i < n; ++i) rs[i] = ts[i] ? inner : 0;
        return;
    }
    rows = first->d[0];
    inner = first->numel() / rows;
    for (int i = 0; i < n; ++i) rs[i] = ts[i] ? ts[i]->s[0] : 0;
}

// out = w .* (alpha a + beta b) + gamma c ; any of b, c, w, out may be null; optional sum-of-squares partials
static inline void ewise(Ctx& c, const Tensor& a, double alpha, const Tensor* b, double beta, const Tensor* cc, double gamma,
                  const Tensor* w, Tensor* out, Tensor* sumsq) {
    const Tensor* ts[5] = {&a, b, cc, w, out};
    long rows, inner, rs[5];
    panels(ts, 5, rows, inner, rs);
    if (sumsq) *sumsq = Tensor::empty(c, {256});
    ProfScope ps(c, CAT_EWISE, 8.0 * (double)rows * inner * (1 + (b ? 1 : 0) + (cc ? 1 : 0) + (w ? 1 : 0) + (out ? 1 : 0)));
    check_rc(ttipm_ewise((int)rows, (int)inner, alpha, a.p, rs[0], beta, b ? b->p : nullptr, rs[1], gamma,
                         cc ? cc->p : nullptr, rs[2], w ? w->p : nullptr, rs[3], out ? out->p : nullptr, rs[4],
                         sumsq ? sumsq->p : nullptr, c.st), "ewise");
    c.launches++;
}

static inline Tensor copy2d(Ctx& c, const Tensor& A) {     // contiguous copy of a strided 2-D view
    Tensor out = Tensor::empty(c, {A.d[0], A.d[1]});
    ProfScope ps(c, CAT_EWISE, 16.0 * (double)out.numel());
    check_rc(ttipm_scale2d(A.p, A.s[0], A.s[1], (int)A.d[0], (int)A.d[1], nullptr, 0, 0, out.p, c.st), "copy2d");
    c.launches++;
    return out;
}

static inline double sum_host(const std::vector<double>& h, size_t a, size_t n) {
    double t = 0.0;
    for (size_t i = 0; i < n; ++i) t += h[a + i];
    return t;
}

// several partial-sum buffers -> totals with ONE synchronising transfer
static inline std::vector<double> host_sums(Ctx& c, const std::vector<const Tensor*>& ts) {
    size_t tot = 0;
    for (auto t : ts) tot += (size_t)t->numel();
    Tensor flat = Tensor::empty(c, {(long)tot});
    size_t o = 0;
    for (auto t : ts) {
        dev_to_dev(c, t->p, (size_t)t->numel(), flat.p + o);
        o += (size_t)t->numel();
    }
    std::vector<double> h = read_vec(c, flat);
    std::vector<double> out;
    o = 0;
    for (auto t : ts) {
        out.push_back(sum_host(h, o, (size_t)t->numel()));
        o += (size_t)t->numel();
    }
    return out;
}

static inline int prune_singular_vals(const std::vector<double>& s, double eps) {   // reference cy_src/tt_ops_cy.pyx:162-177
    const size_t n = s.size();
    bool allzero = true;
    for (double v : s) allzero = allzero && v == 0.0;
    if (allzero) return 1;
    std::vector<double> sc(n);
    double acc = 0.0;
    for (size_t i = n; i-- > 0;) {
        acc += fabs(s[i]) * fabs(s[i]);
        sc[i] = acc;
    }
    int R = 0;
    for (size_t i = 0; i < n; ++i)
        if (sc[i] < eps * eps) { R = (int)i; break; }
    R = std::max(R, 1);
    if (sc[n - 1] > eps * eps) R = (int)n;
    return R;
}


}  // namespace drv
}  // namespace ttipm
