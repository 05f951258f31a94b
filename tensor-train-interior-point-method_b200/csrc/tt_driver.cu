// Native TT-algebra driver: device-resident tensor trains behind C-ABI handles.
//
// A `ttipm_tt` is a list of device cores (r, n, R) / (r, m, n, R).  The operations below are the TT algebra the IPM
// driver chains to form X / Y / Z updates, residuals and the KKT blocks (reference src/tt_ipm.py:404-475, :510-568):
//   rl-orthogonalisation and rounding     cy_src/tt_ops_cy.pyx:132-226  (+ the energy-collecting variant of :262-388)
//   addition                              cy_src/tt_ops_cy.pyx:229-258
//   inner product                         cy_src/tt_ops_cy.pyx:506-520
//   zip-up products (mat-vec, mat-mat,    cy_src/tt_ops_cy.pyx:393-502
//   Hadamard) with swap_cores
// Each call runs its whole loop from C++ on one stream: per core a QR / SVD launch (k_linalg), the bond GEMMs (k_gemm),
// the layout kernels (k_permute4, k_scale2d, k_block_diag), and ONE synchronising read-back of the singular values for
// the reference's rank rule.  Trains stay in HBM between calls; the Python layer (ttipm_b200.tt) materialises NumPy
// cores only when the unchanged driver code touches them.
#include "drv_ops.h"

namespace ttipm {
namespace drv {

struct Core {
    Tensor t;          // contiguous (r, n1[, n2], R)
    int nd = 3;
    long r = 1, n1 = 1, n2 = 0, R = 1;
    long nm() const { return n1 * (n2 > 0 ? n2 : 1); }
    long numel() const { return r * nm() * R; }
};

static Core make_core(const Tensor& t, long r, long n1, long n2, long R) {
    Core c;
    c.r = r; c.n1 = n1; c.n2 = n2; c.R = R;
    c.nd = n2 > 0 ? 4 : 3;
    c.t = n2 > 0 ? t.reshape({r, n1, n2, R}) : t.reshape({r, n1, R});
    return c;
}

// One context for every train of the process (one caller thread, one stream -- SURVEY 8b): device buffers are shared
// between trains (clones, untouched cores), so their bookkeeping must outlive any single handle.
static Ctx& tt_ctx() {
    static Ctx ctx;
    return ctx;
}

struct TT {
    Ctx& c;
    std::vector<Core> cores;
    TT() : c(tt_ctx()) {}
};

// (r, nm, R) -> (R, nm, r) materialised
static Tensor reverse3(Ctx& c, const Core& k) {
    return permute4(c, k.t.reshape({1, k.r, k.nm(), k.R}), 0, 3, 2, 1, nullptr, 0, false);
}

// cy_src/tt_ops_cy.pyx:132-159 (in place)
static void rl_orthogonalise(TT& tt) {
    Ctx& c = tt.c;
    const int d = (int)tt.cores.size();
    for (int i = d - 1; i > 0; --i) {
        Core& ci = tt.cores[i];
        Core& cp = tt.cores[i - 1];
        Tensor Q, Rm;
        qr(c, ci.t.reshape({ci.r, ci.nm() * ci.R}).t2(), Q, Rm);                // (rest, K), (K, r)
        const long nr = Rm.d[0];
        Tensor qt = copy2d(c, Q.t2());
        Tensor left = gemm(c, cp.t.reshape({cp.r * cp.nm(), cp.R}), Rm.t2());    // (.., nr)
        ci = make_core(qt, nr, ci.n1, ci.n2, ci.R);
        cp = make_core(left, cp.r, cp.n1, cp.n2, nr);
    }
}

// left-to-right truncation sweep (cy_src/tt_ops_cy.pyx:197-224; collect: the tail-energy rule of :283-318, :349-384)
static double round_sweep(TT& tt, double eps, bool collect) {
    Ctx& c = tt.c;
    const int d = (int)tt.cores.size();
    double dropped = 0.0;
    for (int idx = 0; idx + 1 < d; ++idx) {
        Core& ck = tt.cores[idx];
        Core& cn = tt.cores[idx + 1];
        Tensor U, S, W;
        svd_left(c, ck.t.reshape({ck.r * ck.nm(), ck.R}), U, S, W);
        std::vector<double> s = read_vec(c, S);
        const long K = (long)s.size();
        long nr;
        if (collect) {
            std::vector<double> sc((size_t)K);
            double acc = 0.0;
            for (long i = K; i-- > 0;) {
                acc += s[i] * s[i];
                sc[i] = acc;
            }
            nr = 0;
            for (long i = 0; i < K; ++i)
                if (sc[i] < eps * eps) { nr = i; break; }
            nr = std::max<long>(nr, 1);
            if (sc[K - 1] > eps * eps) nr = K;
            if (nr < K) dropped += sc[nr];
        } else {
            nr = prune_singular_vals(s, eps);
        }
        Tensor u = nr == K ? U : copy2d(c, U.slice(1, 0, nr));
        Tensor right = gemm(c, W.slice(0, 0, nr), cn.t.reshape({cn.r, cn.nm() * cn.R}));
        ck = make_core(u, ck.r, ck.n1, ck.n2, nr);
        cn = make_core(right, nr, cn.n1, cn.n2, cn.R);
    }
    return dropped;
}

static bool all_rank_one(const TT& tt) {
    if (tt.cores.size() == 1) return true;
    for (size_t i = 0; i + 1 < tt.cores.size(); ++i)
        if (tt.cores[i].R != 1) return false;
    return true;
}

// cy_src/tt_ops_cy.pyx:393-426: (ca, cb) -> cores with the two modes exchanged, truncated at eps
static void swap_cores(Ctx& c, Core& a, Core& b, double eps) {
    const long a0 = a.r, am = a.nm(), a3 = a.R, bm = b.nm(), b3 = b.R;
    Tensor t = gemm(c, a.t.reshape({a0 * am, a3}), b.t.reshape({b.r, bm * b3}));          // (a0 am) x (bm b3)
    t = permute4(c, t.reshape({a0, am, bm, b3}), 0, 2, 1, 3, nullptr, 0, false);          // (a0, bm, am, b3)
    Tensor U, S, W;
    svd_left(c, t.reshape({a0 * bm, am * b3}), U, S, W);
    std::vector<double> s = read_vec(c, S);
    const long rp = prune_singular_vals(s, eps);
    // the reference keeps u * s on the left and v on the right; W = s * v, so left = U * s, right = W / s
    Tensor left = Tensor::empty(c, {a0 * bm, rp});
    Tensor right = Tensor::empty(c, {rp, am * b3});
    Tensor Us = U.slice(1, 0, rp), Ws = W.slice(0, 0, rp);
    check_rc(ttipm_scale2d(Us.p, Us.s[0], Us.s[1], (int)Us.d[0], (int)rp, S.p, 1, 0, left.p, c.st), "scale_cols");
    check_rc(ttipm_scale2d(Ws.p, Ws.s[0], Ws.s[1], (int)rp, (int)Ws.d[1], S.p, 0, 1, right.p, c.st), "scale_rows");
    c.launches += 2;
    const Core a_old = a, b_old = b;
    a = make_core(left, a0, b_old.n1, b_old.n2, rp);
    b = make_core(right, rp, a_old.n1, a_old.n2, b3);
}

// kind 0: matrix x vector, 1: matrix x matrix, 2: Hadamard (cy_src/tt_ops_cy.pyx:430-502)
static void zipup(TT& out, const TT& A, const TT& B, int kind, double eps) {
    Ctx& c = out.c;
    const int d = (int)A.cores.size();
    if ((int)B.cores.size() != d) throw DriverError(1, "zipup: trains of different length");
    std::vector<Core> cores(d);
    for (int q = 0; q < d; ++q) {                        // reversed and transposed second operand
        const Core& b = B.cores[d - 1 - q];
        cores[q] = make_core(reverse3(c, b), b.R, b.n1, b.n2, b.r);
    }
    const double loop_eps = d > 1 ? eps / sqrt((double)(d - 1)) : eps;
    for (int i = 0; i < d; ++i) {
        const Core& a = A.cores[d - 1 - i];
        Core& c0 = cores[0];
        if (kind == 2) {
            // out[r, i.., K] = sum_R a[r, i.., R] c0[R, i.., K], batched over the mode
            if (a.nm() != c0.nm() || a.R != c0.r) throw DriverError(1, "hadamard: shapes do not chain");
            const long nn = a.nm(), Kk = c0.R;
            Tensor A3 = a.t.reshape({a.r, nn, a.R}).permute({1, 0, 2});
            Tensor B3 = c0.t.reshape({c0.r, nn, Kk}).permute({1, 0, 2});
            Tensor o = gemm(c, A3, B3);                                              // (nn, r, K)
            o = permute4(c, o.reshape({1, nn, a.r, Kk}), 0, 2, 1, 3, nullptr, 0, false);
            c0 = make_core(o, a.r, a.n1, a.n2, Kk);
        } else {
            // tensordot(A[p] (s, m, n, S), c0 (S, n[, j], K), axes = ([3, 2], [0, 1])) -> (s, m[, j], K)
            if (a.nd != 4) throw DriverError(1, "zipup: the first operand must be a TT matrix");
            const long s = a.r, m = a.n1, n = a.n2, S = a.R;
            const long j = kind == 1 ? c0.n2 : 1, Kk = c0.R;
            if (c0.r != S || c0.n1 != n) throw DriverError(1, "zipup: shapes do not chain");
            Tensor Ap = permute4(c, a.t, 0, 1, 3, 2, nullptr, 0, false).reshape({s * m, S * n});
            Tensor o = gemm(c, Ap, c0.t.reshape({S * n, j * Kk}));
            c0 = make_core(o, s, m, kind == 1 ? j : 0, Kk);
        }
        if (i != d - 1)
            for (int q = i; q >= 0; --q) swap_cores(c, cores[q], cores[q + 1], loop_eps);
    }
    out.cores = cores;
}

// cy_src/tt_ops_cy.pyx:244-258
static void add(TT& out, const TT& a, const TT& b) {
    Ctx& c = out.c;
    const int d = (int)a.cores.size();
    if ((int)b.cores.size() != d) throw DriverError(1, "tt_add: trains of different length");
    out.cores.resize(d);
    for (int k = 0; k < d; ++k) {
        const Core& x = a.cores[k];
        const Core& y = b.cores[k];
        if (x.nm() != y.nm()) throw DriverError(1, "tt_add: mode sizes differ");
        if (d == 1) {
            Tensor o = Tensor::empty(c, {x.numel()});
            Tensor xf = x.t.reshape({x.numel()}), yf = y.t.reshape({y.numel()});
            ewise(c, xf, 1.0, &yf, 1.0, nullptr, 0.0, nullptr, &o, nullptr);
            out.cores[k] = make_core(o, x.r, x.n1, x.n2, x.R);
            continue;
        }
        const int mode = k == 0 ? 0 : (k == d - 1 ? 2 : 1);
        const long ro = mode == 0 ? x.r : x.r + y.r, Ro = mode == 2 ? x.R : x.R + y.R;
        Tensor o = Tensor::empty(c, {ro * x.nm() * Ro});
        check_rc(ttipm_block_diag(x.t.p, y.t.p, o.p, (int)x.r, (int)x.R, (int)y.r, (int)y.R, (int)x.nm(), mode, c.st), "block_diag");
        c.launches++;
        out.cores[k] = make_core(o, ro, x.n1, x.n2, Ro);
    }
}

// cy_src/tt_ops_cy.pyx:506-520
static double inner(Ctx& c, const TT& a, const TT& b) {
    {   // one launch for the whole chain while the running matrix and the intermediate fit one CTA's shared memory
        const int d = (int)a.cores.size();
        if (d >= 1 && d == (int)b.cores.size() && d <= 40) {
            std::vector<const double*> pa(d), pb(d);
            std::vector<int32_t> ra(d + 1), rb(d + 1), nm(d);
            bool ok = true;
            for (int k = 0; k < d; ++k) {
                const Core& c1 = a.cores[k];
                const Core& c2 = b.cores[k];
                if (c2.nm() != c1.nm()) throw DriverError(1, "tt_inner_prod: mode sizes differ");
                ok = ok && c1.t.contiguous() && c2.t.contiguous();
                pa[k] = c1.t.p; pb[k] = c2.t.p;
                ra[k] = (int32_t)c1.r; rb[k] = (int32_t)c2.r; nm[k] = (int32_t)c1.nm();
                ra[k + 1] = (int32_t)c1.R; rb[k + 1] = (int32_t)c2.R;
            }
            if (ok && ra[0] == 1 && rb[0] == 1 && ra[d] == 1 && rb[d] == 1) {
                Tensor out = Tensor::empty(c, {1});
                const int rc = ttipm_tt_inner_chain(d, pa.data(), pb.data(), ra.data(), rb.data(), nm.data(), out.p, c.st);
                if (rc == 0) {
                    c.launches++;
                    double v = 0.0;
                    to_host(c, out.p, 1, &v);
                    return v;
                }
                if (rc > 0) check_rc(rc, "tt_inner_chain");
            }
        }
    }
    Tensor res = Tensor::empty(c, {1, 1});
    const double one = 1.0;
    from_host(c, &one, 1, res.p);
    for (size_t k = 0; k < a.cores.size(); ++k) {
        const Core& c1 = a.cores[k];
        const Core& c2 = b.cores[k];
        const long nn = c1.nm();
        if (c2.nm() != nn) throw DriverError(1, "tt_inner_prod: mode sizes differ");
        Tensor T = gemm(c, res.t2(), c1.t.reshape({c1.r, nn * c1.R}));                 // (r2, nn R1)
        res = gemm(c, T.reshape({c2.r * nn, c1.R}).t2(), c2.t.reshape({c2.r * nn, c2.R}));   // (R1, R2)
    }
    double v = 0.0;
    to_host(c, res.p, 1, &v);
    return v;
}

}  // namespace drv
}  // namespace ttipm

using namespace ttipm;
using namespace ttipm::drv;

struct ttipm_tt {
    TT tt;
};

#define TT_GUARD(...)                                      \
    try {                                                  \
        __VA_ARGS__;                                       \
        return 0;                                          \
    } catch (const DriverError& e) {                       \
        return fail(e.code ? e.code : 1, "%s", e.what());  \
    } catch (const std::exception& e) {                    \
        return fail(99, "%s", e.what());                   \
    }

extern "C" ttipm_tt* ttipm_tt_create(int d, void* stream) {
    if (check_bound_device()) return nullptr;
    pool_keep_freed_blocks();
    ttipm_tt* h = new ttipm_tt();
    h->tt.c.st = (tt_stream_t)stream;
    h->tt.cores.resize(d > 0 ? d : 0);
    return h;
}

extern "C" void ttipm_tt_destroy(ttipm_tt* h) {
    delete h;
}

extern "C" int ttipm_tt_length(const ttipm_tt* h) { return (int)h->tt.cores.size(); }

// All cores from one host buffer (concatenated, row-major); dims = 4 ints per core: r, n1, n2 (0 for a 3-D core), R.
// One host-to-device copy; the cores are views into one device allocation.
extern "C" int ttipm_tt_set_cores(ttipm_tt* h, const double* host, const int32_t* dims) {
    TT_GUARD({
        TT& tt = h->tt;
        const int d = (int)tt.cores.size();
        long total = 0;
        for (int k = 0; k < d; ++k) total += (long)dims[4 * k] * dims[4 * k + 1] * (dims[4 * k + 2] > 0 ? dims[4 * k + 2] : 1) * dims[4 * k + 3];
        Tensor all = Tensor::empty(tt.c, {total});
        from_host(tt.c, host, (size_t)total, all.p);
        long o = 0;
        for (int k = 0; k < d; ++k) {
            const long r = dims[4 * k], n1 = dims[4 * k + 1], n2 = dims[4 * k + 2], R = dims[4 * k + 3];
            const long n = r * n1 * (n2 > 0 ? n2 : 1) * R;
            tt.cores[k] = make_core(all.slice(0, o, o + n), r, n1, n2, R);
            o += n;
        }
    });
}

extern "C" int ttipm_tt_shapes(const ttipm_tt* h, int32_t* dims) {
    const TT& tt = h->tt;
    for (size_t k = 0; k < tt.cores.size(); ++k) {
        dims[4 * k] = (int32_t)tt.cores[k].r;
        dims[4 * k + 1] = (int32_t)tt.cores[k].n1;
        dims[4 * k + 2] = (int32_t)tt.cores[k].n2;
        dims[4 * k + 3] = (int32_t)tt.cores[k].R;
    }
    return 0;
}

// All cores into one host buffer (concatenated like ttipm_tt_set_cores): packed on the device, ONE synchronising copy.
extern "C" int ttipm_tt_get_cores(ttipm_tt* h, double* host) {
    TT_GUARD({
        TT& tt = h->tt;
        long total = 0;
        for (const Core& k : tt.cores) total += k.numel();
        Tensor all = Tensor::empty(tt.c, {total});
        long o = 0;
        for (const Core& k : tt.cores) {
            dev_to_dev(tt.c, k.t.p, (size_t)k.numel(), all.p + o);
            o += k.numel();
        }
        to_host(tt.c, all.p, (size_t)total, host);
    });
}

// a new handle sharing the device buffers of `src` (cores are immutable once built; operations replace them)
extern "C" ttipm_tt* ttipm_tt_clone(const ttipm_tt* src) {
    ttipm_tt* h = new ttipm_tt();
    h->tt.c.st = src->tt.c.st;
    h->tt.cores = src->tt.cores;
    return h;
}

// core k <- alpha * core k (tt_scale multiplies ONE core, cy_src/tt_ops_cy.pyx:96-114)
extern "C" int ttipm_tt_scale_core(ttipm_tt* h, int k, double alpha) {
    TT_GUARD({
        Core& ck = h->tt.cores.at(k);
        Tensor o = Tensor::empty(h->tt.c, {ck.numel()});
        Tensor in = ck.t.reshape({ck.numel()});
        ewise(h->tt.c, in, alpha, nullptr, 0.0, nullptr, 0.0, nullptr, &o, nullptr);
        ck = make_core(o, ck.r, ck.n1, ck.n2, ck.R);
    });
}

extern "C" int ttipm_tt_rl_orthogonalise(ttipm_tt* h) {
    TT_GUARD({
        if (h->tt.cores.size() > 1) rl_orthogonalise(h->tt);
    });
}

// rl-orthogonalisation + truncation sweep with the PER-BOND tolerance eps (the caller applies the reference's
// eps / sqrt(d - 1) and eps / 2 scalings); collect != 0: tail-energy rule, *dropped receives the discarded energy.
// Trains whose ranks are all 1 are left untouched, as in the reference.
extern "C" int ttipm_tt_round(ttipm_tt* h, double eps, int collect, double* dropped) {
    TT_GUARD({
        if (dropped) *dropped = 0.0;
        if (!all_rank_one(h->tt)) {
            rl_orthogonalise(h->tt);
            const double dr = round_sweep(h->tt, eps, collect != 0);
            if (dropped) *dropped = dr;
        }
    });
}

extern "C" ttipm_tt* ttipm_tt_add(const ttipm_tt* a, const ttipm_tt* b) {
    ttipm_tt* h = new ttipm_tt();
    h->tt.c.st = a->tt.c.st;
    try {
        add(h->tt, a->tt, b->tt);
        return h;
    } catch (const std::exception& e) {
        fail(1, "%s", e.what());
        delete h;
        return nullptr;
    }
}

extern "C" int ttipm_tt_inner(ttipm_tt* a, const ttipm_tt* b, double* out) {
    TT_GUARD({ *out = inner(a->tt.c, a->tt, b->tt); });
}

extern "C" ttipm_tt* ttipm_tt_zipup(int kind, const ttipm_tt* A, const ttipm_tt* B, double eps) {
    ttipm_tt* h = new ttipm_tt();
    h->tt.c.st = A->tt.c.st;
    try {
        zipup(h->tt, A->tt, B->tt, kind, eps);
        return h;
    } catch (const std::exception& e) {
        fail(1, "%s", e.what());
        delete h;
        return nullptr;
    }
}

// same data, other mode split: every core (r, n1 * n2', R) viewed as (r, n1, n2, R) (n2 = 0: one mode) -- tt_reshape
// (reference src/tt_ops.py:330-333, without the core-merging branch); shares the device buffers
extern "C" ttipm_tt* ttipm_tt_reshape(const ttipm_tt* src, int n1, int n2) {
    ttipm_tt* h = new ttipm_tt();
    try {
        for (const Core& k : src->tt.cores) {
            if ((long)n1 * (n2 > 0 ? n2 : 1) != k.nm()) throw DriverError(1, "tt_reshape: mode size changes");
            h->tt.cores.push_back(make_core(k.t.reshape({k.numel()}), k.r, n1, n2, k.R));
        }
        return h;
    } catch (const std::exception& e) {
        fail(1, "%s", e.what());
        delete h;
        return nullptr;
    }
}

// every 4-D core (r, m, n, R) -> (r, n, m, R) materialised (tt_transpose, cy_src/tt_ops_cy.pyx:57-78, for trains whose
// cores are all 4-D); NULL if a core is 3-D
extern "C" ttipm_tt* ttipm_tt_transpose(const ttipm_tt* src) {
    ttipm_tt* h = new ttipm_tt();
    try {
        for (const Core& k : src->tt.cores) {
            if (k.nd != 4) throw DriverError(1, "tt_transpose: 3-D core");
            Tensor o = permute4(h->tt.c, k.t, 0, 2, 1, 3, nullptr, 0, false);
            h->tt.cores.push_back(make_core(o, k.r, k.n2, k.n1, k.R));
        }
        return h;
    } catch (const std::exception& e) {
        fail(1, "%s", e.what());
        delete h;
        return nullptr;
    }
}

// Kronecker / diagonal embeddings per core (reference src/tt_ops.py:360-375, :312-316):
// kind 0: I (x) M, 1: M (x) I  ((r, 2, 2, R) -> (r, 4, 4, R));  2: diag ((r, q, R) or (r, m, n, R) with q = m n -> (r, q, q, R))
extern "C" ttipm_tt* ttipm_tt_embed(const ttipm_tt* src, int kind) {
    ttipm_tt* h = new ttipm_tt();
    try {
        Ctx& c = h->tt.c;
        for (const Core& k : src->tt.cores) {
            if (kind != 2 && !(k.nd == 4 && k.n1 == 2 && k.n2 == 2)) throw DriverError(1, "tt_embed: expected (r, 2, 2, R) cores");
            const long q = kind == 2 ? k.nm() : 4;
            Tensor o = Tensor::empty(c, {k.r * q * q * k.R});
            check_rc(ttipm_embed(k.t.p, o.p, (int)k.r, (int)k.R, (int)q, kind, c.st), "embed");
            c.launches++;
            h->tt.cores.push_back(make_core(o, k.r, q, q, k.R));
        }
        return h;
    } catch (const std::exception& e) {
        fail(1, "%s", e.what());
        delete h;
        return nullptr;
    }
}

// launches and host synchronisations issued through this handle so far
extern "C" int ttipm_tt_counters(const ttipm_tt* h, int64_t* launches, int64_t* syncs) {
    if (launches) *launches = h->tt.c.launches;
    if (syncs) *syncs = h->tt.c.syncs;
    return 0;
}
