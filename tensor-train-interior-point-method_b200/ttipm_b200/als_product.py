"""ALS-fitted TT products on the device: tt_approx_mat_mat_mul (reference src/tt_als.py:1502-1628)
and tt_approx_mat_vec_mul (src/tt_als.py:1637-1762), the branch the reference's product dispatchers take when a rank
product exceeds 40 / 80 (src/tt_als.py:1631-1634, :1765-1768).

One engine serves both: a vector train (b, k, B) is a matrix train with a column mode of 1.  Cores, interfaces and
intermediates stay on the device for the whole fit; the host sees the scalars the reference branches on (local
residual norms, core / interface norms, singular values for the rank rule) and draws the Gaussian enrichment columns
with the global NumPy RNG in the reference's order (tt_random_gaussian, add_kick_rank).

Contractions (SURVEY 8a' family K7), each a chain of three strided GEMMs (`k_gemm`) with `k_permute4` in between:
  local solution  'rab,amkA,bknB,RAB->rmnR'
  backward update 'RAB,amkA,bknB,rmnR->rab'
  forward update  'rab,amkA,bknB,rmnR->RAB'
"""
import math

import numpy as np

from . import kernels as K
from .runtime import get_runtime

STATS = {"fits": 0, "half_sweeps": 0}       # process-wide counters (end-to-end harness, tools)


def symmetric_powers_of_two(length):
    """cy_src/tt_ops_cy.pyx:538-555."""
    half = (length + 1) // 2
    up = [1 << (i + 1) for i in range(half)]
    return np.array((up + up[:length // 2][::-1])[:length] if length > 0 else [], dtype=np.int64)


def add_kick_rank(u, v, r_add=2):
    """cy_src/tt_ops_cy.pyx:559-579 on NumPy operands: r_add Gaussian columns (global RNG) appended to the orthonormal
    factor u, re-orthogonalised by the device QR, the R factor carried into v."""
    rt = get_runtime()
    Q, W, r = _enrich(rt.to_device(u), rt.to_device(v), u.shape[1], int(r_add), rt)
    return rt.to_host(Q), rt.to_host(W), r


def _enrich(U, W, r, add, rt):
    M = U.shape[0]
    cat = rt.empty(M, r + add)
    cat[:, :r].copy_(U[:, :r])
    cat[:, r:].copy_(rt.to_device(np.random.randn(M, add)))
    Q, Rm = K.qr(cat, rt=rt)
    return Q, K.gemm(Rm[:, :r], W[:r], rt=rt), Q.shape[1]


def _sumsq(t, rt, other=None):
    if other is None:
        _, ss = K.ewise(t, want_sumsq=True, store=False, rt=rt)
    else:
        _, ss = K.ewise(t, 1.0, b=other, beta=-1.0, want_sumsq=True, store=False, rt=rt)
    return float(rt.to_host(ss).sum())


class _Operands:
    """Per-core operand layouts of A (a, m, k, A') and D (b, k, n, B), permuted once per fit."""

    def __init__(self, A, D, rt):
        self.A, self.D = A, D
        self.A_ak = [K.permute4(c, (0, 2, 1, 3), rt=rt) for c in A]      # (a, k, m, A')
        self.A_mAk = [K.permute4(c, (0, 1, 3, 2), rt=rt) for c in A]     # (a, m, A', k)
        self.D_bnk = [K.permute4(c, (0, 2, 1, 3), rt=rt) for c in D]     # (b, n, k, B)
        self.D_nB_bk = [_d_nB_bk(c, rt) for c in D]                      # (n B) x (b k)


def _local(G0, G1, ops, k, scale, rt):
    """'rab,amkA,bknB,RAB->rmnR' times scale."""
    A, D = ops.A[k], ops.D[k]
    r, a, b = G0.shape
    _, m, kk, Ap = A.shape
    _, _, n, B = D.shape
    R = G1.shape[0]
    T1 = K.gemm(G0.reshape(r * a, b), D.reshape(b, kk * n * B), rt=rt)                        # (r, a, k, n, B)
    T2 = K.gemm(T1.reshape(r, a * kk, n * B).permute(0, 2, 1), ops.A_ak[k].reshape(1, a * kk, m * Ap), rt=rt)
    T2 = K.permute4(T2.reshape(r * n, B, m, Ap), (0, 2, 1, 3), rt=rt)                         # (r n, m, B, A')
    G1p = K.permute4(G1.reshape(R, Ap, B, 1), (2, 1, 0, 3), rt=rt)                            # (B, A', R)
    sol = K.gemm(T2.reshape(r * n * m, B * Ap), G1p.reshape(B * Ap, R), alpha=scale, rt=rt)   # (r, n, m, R)
    if n == 1:
        return sol.reshape(r, m, n, R)
    return K.permute4(sol.reshape(r, n, m, R), (0, 2, 1, 3), rt=rt)


def _phi_bck(G1, ops, k, x, rt):
    """'RAB,amkA,bknB,rmnR->rab'."""
    A, D = ops.A[k], ops.D[k]
    a, m, kk, Ap = A.shape
    b, _, n, B = D.shape
    r, R = x.shape[0], x.shape[3]
    U1 = K.gemm(x.reshape(r * m * n, R), G1.reshape(R, Ap * B), rt=rt)                        # (r, m, n, A', B)
    U1 = K.permute4(U1.reshape(r * m, n, Ap, B), (0, 2, 1, 3), rt=rt)                         # (r m, A', n, B)
    U2 = K.gemm(U1.reshape(r * m * Ap, n * B), ops.D_nB_bk[k], rt=rt)                         # (r, m, A', b, k)
    U2 = K.permute4(U2.reshape(r, m * Ap, b, kk), (0, 2, 1, 3), rt=rt)                        # (r, b, m A', k)
    Gt = K.gemm(U2.reshape(r * b, m * Ap * kk), ops.A_mAk[k].reshape(a, m * Ap * kk).t(), rt=rt)   # (r, b, a)
    return K.permute4(Gt.reshape(r, b, a, 1), (0, 2, 1, 3), rt=rt).reshape(r, a, b)


def _d_nB_bk(D, rt):
    """D (b, k, n, B) as the (n B) x (b k) matrix of the backward update: a materialised permutation, because the
    composite row index (n, B) is not expressible with one stride."""
    b, kk, n, B = D.shape
    return K.permute4(D.reshape(b * kk, n, B, 1), (1, 2, 0, 3), rt=rt).reshape(n * B, b * kk)


def _phi_fwd(G0, ops, k, x, rt):
    """'rab,amkA,bknB,rmnR->RAB'."""
    A, D = ops.A[k], ops.D[k]
    a, m, kk, Ap = A.shape
    b, _, n, B = D.shape
    r, R = x.shape[0], x.shape[3]
    V1 = K.gemm(G0.reshape(r, a * b).t(), x.reshape(r, m * n * R), rt=rt)                     # (a, b, m, n R)
    V1 = K.permute4(V1.reshape(a, b, m, n * R), (1, 0, 2, 3), rt=rt)                          # (b, a, m, n R)
    V2 = K.gemm(V1.reshape(b, a * m, n * R).permute(0, 2, 1), A.reshape(1, a * m, kk * Ap), rt=rt)   # (b, n, R, k, A')
    V2 = K.permute4(V2.reshape(b * n, R, kk, Ap), (0, 2, 1, 3), rt=rt)                        # (b n, k, R, A')
    return K.gemm(V2.reshape(b * n * kk, R * Ap).t(), ops.D_bnk[k].reshape(b * n * kk, B), rt=rt).reshape(R, Ap, B)


def _split(mat, bond, last, tol, kick_rank, rt):
    """SVD of the local unfolding, tail-energy rank rule (cy_src/tt_ops_cy.pyx:162-177), Gaussian enrichment unless this
    is the closing sweep (add_kick_rank, cy_src/tt_ops_cy.pyx:559-579).  Returns U (M, r'), W (r', N), r'."""
    from .tt import prune_singular_vals
    U, S, W = K.svd_left(mat, rt=rt)
    r = prune_singular_vals(rt.to_host(S), tol)
    if last:
        return U[:, :r], W[:r], r
    return _enrich(U, W, r, int(kick_rank[bond]), rt)


def als_fit_product(A, D, x0=None, kick_rank=None, nswp=50, tol=1e-6, verbose=False, trace=None):
    """NumPy cores in, NumPy cores out; same arguments and RNG draws as the reference's two functions."""
    from .tt import tt_random_gaussian, tt_ranks
    rt = get_runtime()
    STATS["fits"] += 1
    vec = D[0].ndim == 3
    if x0 is None:
        max_ranks = np.maximum((np.array(tt_ranks(A)) + np.array(tt_ranks(D))) / 2, 2).astype(int)
        x_host = tt_random_gaussian(list(max_ranks), (A[0].shape[2],) if vec else A[0].shape[1:-1])
    else:
        x_host, max_ranks = x0, np.array(tt_ranks(x0))
    if kick_rank is None:
        kick_rank = np.maximum((symmetric_powers_of_two(len(A) - 1) - max_ranks) / (nswp / 2), 2).astype(int)
    d = len(x_host)
    up4 = lambda c: rt.to_device(c).reshape(c.shape[0], c.shape[1], 1, c.shape[2]) if c.ndim == 3 else rt.to_device(c)
    x = [up4(c) for c in x_host]
    ops = _Operands([rt.to_device(c) for c in A], [up4(c) for c in D], rt)
    rx = [1] + [int(q) for q in tt_ranks(x_host)] + [1]
    one = rt.to_device(np.ones((1, 1, 1)))
    G = [one] + [None] * (d - 1) + [one]
    normAD, normx, nrmsc = np.ones(d - 1), np.ones(d - 1), 1.0
    tol = tol / math.sqrt(d)
    max_res = 0.0

    def local(k):
        nonlocal max_res
        sol = _local(G[k], G[k + 1], ops, k, nrmsc, rt)
        res = math.sqrt(_sumsq(sol, rt, other=x[k])) / max(math.sqrt(_sumsq(sol, rt)), 1e-8)
        max_res = max(max_res, res)
        return sol

    def unit(t):
        nrm = math.sqrt(_sumsq(t, rt))
        return nrm, K.ewise(t, 1.0 / nrm, rt=rt)

    last = False
    for swp in range(nswp):
        max_res = math.inf if swp == 0 else 0.0
        for k in range(d - 1, -1, -1):
            sol = local(k) if swp > 0 else x[k]
            m, n = x[k].shape[1], x[k].shape[2]
            if k == 0:
                x[k] = sol
                continue
            U, W, r = _split(sol.reshape(rx[k], m * n * rx[k + 1]).t(), k - 1, last, tol, kick_rank, rt)
            nrmsc *= normx[k - 1] / normAD[k - 1]
            x[k] = U.t().contiguous().reshape(r, m, n, rx[k + 1])
            sh = x[k - 1].shape
            nrm, x[k - 1] = unit(K.gemm(x[k - 1].reshape(-1, sh[3]), W.t(), rt=rt).reshape(sh[0], sh[1], sh[2], r))
            normx[k - 1] *= nrm
            rx[k] = r
            Gk = _phi_bck(G[k + 1], ops, k, x[k], rt)
            nrm = math.sqrt(_sumsq(Gk, rt))
            nrm = nrm if nrm > 0 else 1.0
            G[k] = K.ewise(Gk, 1.0 / nrm, rt=rt)
            normAD[k - 1] = nrm
            nrmsc *= normAD[k - 1] / normx[k - 1]
        STATS["half_sweeps"] += 1
        if trace is not None:
            trace.append(("bck", swp, float(max_res), list(rx)))
        if last:
            break
        if max_res < tol or swp == nswp - 1:
            last = True
        max_res = 0.0
        for k in range(d):
            sol = local(k)
            m, n = x[k].shape[1], x[k].shape[2]
            if k == d - 1:
                x[k] = sol
                continue
            nrmsc *= normx[k] / normAD[k]
            U, W, r = _split(sol.reshape(rx[k] * m * n, rx[k + 1]), k, last, tol, kick_rank, rt)
            x[k] = U.contiguous().reshape(rx[k], m, n, r)
            sh = x[k + 1].shape
            nrm, x[k + 1] = unit(K.gemm(W, x[k + 1].reshape(sh[0], -1), rt=rt).reshape(r, sh[1], sh[2], sh[3]))
            normx[k] *= nrm
            rx[k + 1] = r
            Gk = _phi_fwd(G[k], ops, k, x[k], rt)
            nrm = math.sqrt(_sumsq(Gk, rt))
            nrm = nrm if nrm > 0 else 1.0
            G[k + 1] = K.ewise(Gk, 1.0 / nrm, rt=rt)
            normAD[k] = nrm
            nrmsc *= normAD[k] / normx[k]
        STATS["half_sweeps"] += 1
        if trace is not None:
            trace.append(("fwd", swp, float(max_res), list(rx)))
        if last:
            break
        if max_res < tol:
            last = True
        if verbose:
            print(f"\tStarting Sweep: {swp}\n\tResidual {max_res}\n\tTT-sol rank: {rx[1:-1]}", flush=True)

    scale = float(np.exp(np.sum(np.log(normx)) / d))
    out = [rt.to_host(K.ewise(c, scale, rt=rt)) for c in x]
    return [c[:, :, 0] if vec else c for c in out]


def tt_approx_mat_mat_mul(A, D, x0=None, kick_rank=None, nswp=50, tol=1e-6, verbose=False):
    """src/tt_als.py:1502-1628."""
    return als_fit_product(A, D, x0, kick_rank, nswp, tol, verbose)


def tt_approx_mat_vec_mul(A, d_vec, x0=None, kick_rank=None, nswp=50, tol=1e-6, verbose=False):
    """src/tt_als.py:1637-1762."""
    return als_fit_product(A, d_vec, x0, kick_rank, nswp, tol, verbose)
