"""ORACLE (test infrastructure, not product code).

A plain NumPy/SciPy restatement of the reference's Newton-system hot path: the
block AMEn/ALS sweep (reference src/tt_als.py:12-825), the two local KKT-block
solvers and their matrix-free operators (src/tt_ipm.py:183-401,
cy_src/lgmres_cy.pyx:126-153,291-331,490-510) and the TT algebra they call
(cy_src/tt_ops_cy.pyx, src/tt_ops.py).  Every function cites the reference
lines it follows.  The restatement is pinned against the real reference run in
this container (oracle/ref_harness, fixtures under tests/golden/, see
tests/test_oracle_vs_golden.py); the Krylov solver is the one part whose
iterate-level parity is UNPINNED (PETSc is absent, see oracle/lgmres_ref.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.  The product path never does.
"""
import numpy as np
import scipy.linalg as sla

from lgmres_ref import lgmres as _lgmres

F64 = np.float64


# ----------------------------------------------------------------------------
# K1..K4 contractions (SURVEY 8a', reference src/tt_als.py:190-265)
# ----------------------------------------------------------------------------
def local_matvec(P1, A, P2, x):
    """'lsr,smnS,LSR,rnR->lmL' in the reference's 3-GEMM order
    (cy_src/lgmres_cy.pyx:146-153)."""
    l, s, r = P1.shape
    _, m, n, S = A.shape
    L, _, R = P2.shape
    t1 = x.reshape(r * n, R) @ P2.reshape(L * S, R).T                      # (r n, L S)
    t1 = t1.reshape(r, n, L, S).transpose(0, 2, 1, 3).reshape(r * L, n * S)
    t2 = t1 @ A.transpose(2, 3, 0, 1).reshape(n * S, s * m)                # (r L, s m)
    t2 = t2.reshape(r, L, s, m).transpose(2, 0, 3, 1).reshape(s * r, m * L)
    return (P1.reshape(l, s * r) @ t2).reshape(l, m, L)


def local_matvec_T(P1, A, P2, y):
    """'lsr,smnS,LSR,lmL->rnR' (src/tt_als.py:196): apply the transposed block."""
    return local_matvec(P1.transpose(2, 1, 0), A.transpose(0, 2, 1, 3), P2.transpose(2, 1, 0), y)


def local_diag(P1, A, P2):
    """'lsr,smnS,LSR->lmL' diagonal of the projected block (src/tt_ipm.py:191)."""
    d1 = np.einsum("lsl->ls", P1)
    dA = np.einsum("smmS->smS", A)
    d2 = np.einsum("LSL->LS", P2)
    return np.einsum("ls,smS,LS->lmL", d1, dA, d2)


def local_dense(P1, A, P2):
    """'lsr,smnS,LSR->lmLrnR' reshaped to (m, m) (src/tt_ipm.py:201-212)."""
    l, s, r = P1.shape
    _, m, n, S = A.shape
    L, _, R = P2.shape
    t = np.einsum("lsr,smnS->lrmnS", P1, A)
    t = np.einsum("lrmnS,LSR->lmLrnR", t, P2)
    return t.reshape(l * m * L, r * n * R)


def phi_bck(Phi, left, A, right):
    """'LSR,lML,sMNS,rNR->lsr' (src/tt_als.py:252-253)."""
    t = np.einsum("LSR,rNR->LSrN", Phi, right)
    t = np.einsum("LSrN,sMNS->LrsM", t, A)
    return np.einsum("LrsM,lML->lsr", t, left)


def phi_fwd(Phi, left, A, right):
    """'lsr,lML,sMNS,rNR->LSR' (src/tt_als.py:256-257)."""
    t = np.einsum("lsr,rNR->lsNR", Phi, right)
    t = np.einsum("lsNR,sMNS->lMSR", t, A)
    return np.einsum("lMSR,lML->LSR", t, left)


def phi_bck_rhs(Phi, core_b, core):
    """'BR,bnB,rnR->br' (src/tt_als.py:260-261)."""
    return np.einsum("bnR,rnR->br", np.einsum("BR,bnB->bnR", Phi, core_b), core)


def phi_fwd_rhs(Phi, core_b, core):
    """'br,bnB,rnR->BR' (src/tt_als.py:264-265)."""
    return np.einsum("rnB,rnR->BR", np.einsum("br,bnB->rnB", Phi, core_b), core)


def rhs_project(Xb_k, core_b, Xb_k1):
    """'br,bnB,BR->rnR' (src/tt_als.py:82, src/tt_ipm.py:187-189)."""
    return np.einsum("rnB,BR->rnR", np.einsum("br,bnB->rnB", Xb_k, core_b), Xb_k1)


# ----------------------------------------------------------------------------
# Block containers (duck-typed like reference src/tt_als.py:16-162)
# ----------------------------------------------------------------------------
class BlockMatrix:
    def __init__(self, data=None, aliases=None, transposes=None):
        self._data = dict(data or {})
        self._aliases = dict(aliases or {})
        self._transposes = dict(transposes or {})

    def tkeys(self):
        return self._data.keys() | set(self._transposes.values())


class BlockVector:
    def __init__(self, data=None):
        self._data = dict(data or {})


def block_size_of(block_A):
    return max(k[0] for k in block_A._data.keys()) + 1


def block_local_product(block_A, k, XAX_k, XAX_k1, x):
    """src/tt_als.py:190-200.  x: (r, b, n, R)."""
    y = np.zeros_like(x, dtype=F64)
    for (i, j), cores in block_A._data.items():
        A = cores[k]
        y[:, i] += local_matvec(XAX_k[i, j], A, XAX_k1[i, j], x[:, j])
        if (i, j) in block_A._transposes:
            p, t = block_A._transposes[i, j]
            y[:, p] += local_matvec_T(XAX_k[i, j], A, XAX_k1[i, j], x[:, t])
        if (i, j) in block_A._aliases:
            p, t = block_A._aliases[i, j]
            y[:, p] += local_matvec(XAX_k[i, j], A, XAX_k1[i, j], x[:, t])
    return y


def mixed_block_local_product(block_A, k, left, right, x, shape, left_is_z, right_is_z):
    """The three 'compressed' products of src/tt_als.py:202-238 in one routine.

    left/right are the interface dicts used on each side; a side flagged *_is_z
    holds residual (Z) interfaces which, for a transposed key (p,t), were built
    with the m<->n swapped operator core, while an X interface of the stored
    key (i,j) has to be read reversed ('rsl' / 'RSL') instead."""
    y = np.zeros(shape, dtype=F64)
    for (i, j), cores in block_A._data.items():
        A = cores[k]
        y[:, i] += local_matvec(left[i, j], A, right[i, j], x[:, j])
        if (i, j) in block_A._transposes:
            p, t = block_A._transposes[i, j]
            Pl = left[p, t] if left_is_z else left[i, j].transpose(2, 1, 0)
            Pr = right[p, t] if right_is_z else right[i, j].transpose(2, 1, 0)
            y[:, p] += local_matvec(Pl, A.transpose(0, 2, 1, 3), Pr, x[:, t])
        if (i, j) in block_A._aliases:
            p, t = block_A._aliases[i, j]
            y[:, p] += local_matvec(left[i, j], A, right[i, j], x[:, t])
    return y


def block_rhs_product(block_b, k, Xb_k, Xb_k1, shape):
    """src/tt_als.py:79-83 with nrmsc = 1."""
    y = np.zeros(shape, dtype=F64)
    for i, cores in block_b._data.items():
        y[:, i] += rhs_project(Xb_k[i], cores[k], Xb_k1[i])
    return y


# ----------------------------------------------------------------------------
# Reduced (Schur) local operators (cy_src/lgmres_cy.pyx:291-331, 490-510)
# ----------------------------------------------------------------------------
class ReducedOperatorEq:
    """[y; x] -> [K00 y + K01 x ; K21 x - K22 (inv_I o K01^T y)], block-major flat
    vectors of length 2 r n R."""

    def __init__(self, P1, A, P2, inv_I):
        self.P1, self.A, self.P2, self.inv_I = P1, A, P2, inv_I
        self.shape = inv_I.shape

    def _mv(self, key, x):
        return local_matvec(self.P1[key], self.A[key], self.P2[key], x)

    def matvec(self, v):
        r, n, R = self.shape
        x = v.reshape(2, r, n, R)
        o0 = self._mv((0, 0), x[0]) + self._mv((0, 1), x[1])
        t = self.inv_I * local_matvec_T(self.P1[0, 1], self.A[0, 1], self.P2[0, 1], x[0])
        o1 = self._mv((2, 1), x[1]) - self._mv((2, 2), t)
        return np.stack([o0, o1]).reshape(-1)


class ReducedOperatorIneq(ReducedOperatorEq):
    """3-block version: unknowns [dY; dX; dT]."""

    def matvec(self, v):
        r, n, R = self.shape
        x = v.reshape(3, r, n, R)
        o0 = self._mv((0, 0), x[0]) + self._mv((0, 1), x[1])
        t = self.inv_I * local_matvec_T(self.P1[0, 1], self.A[0, 1], self.P2[0, 1], x[0]) + x[2]
        o1 = self._mv((2, 1), x[1]) - self._mv((2, 2), t)
        o2 = self._mv((3, 1), x[1]) + self._mv((3, 3), x[2])
        return np.stack([o0, o1, o2]).reshape(-1)


def _fb_sub(Lc, b):
    y = sla.solve_triangular(Lc, b, lower=True, check_finite=False)
    return sla.solve_triangular(Lc.T, y, lower=False, check_finite=False)


def _gather(block_A, k, XAX_k, XAX_k1, keys):
    P1 = {key: XAX_k[key] for key in keys}
    A = {key: block_A._data[key][k] for key in keys}
    P2 = {key: XAX_k1[key] for key in keys}
    return P1, A, P2


def local_solver_eq(XAX_k, block_A, k, XAX_k1, Xb_k, block_b, Xb_k1, prev, size_limit,
                    dense_solve=True, rtol=1e-5, stats=None):
    """src/tt_ipm.py:183-282 (3 blocks: dY, dX, dZ)."""
    r, b, n, R = prev.shape
    m = r * n * R
    rhs = np.zeros_like(prev)
    for i in range(3):
        if i in block_b._data:
            rhs[:, i] = rhs_project(Xb_k[i], block_b._data[i][k], Xb_k1[i])
    norm_rhs = max(np.linalg.norm(rhs), 1e-10)
    keys = [(0, 0), (0, 1), (2, 1), (2, 2), (1, 2)]
    P1, A, P2 = _gather(block_A, k, XAX_k, XAX_k1, keys)
    inv_I = 1.0 / local_diag(P1[1, 2], A[1, 2], P2[1, 2])
    res_old = np.linalg.norm(block_local_product(block_A, k, XAX_k, XAX_k1, prev) - rhs) / norm_rhs
    dense = (np.sqrt(r * R) <= size_limit) and dense_solve and (res_old >= rtol)
    direct_fail = not dense
    sol = None
    if dense:
        try:
            Rp, Rd, Rc = (rhs[:, i].reshape(m, 1) for i in range(3))
            LXI = local_dense(P1[2, 2], A[2, 2], P2[2, 2]) * inv_I.reshape(1, -1)
            Leq = local_dense(P1[0, 1], A[0, 1], P2[0, 1])
            Lc = sla.cholesky(local_dense(P1[2, 1], A[2, 1], P2[2, 1]), lower=True, check_finite=False)
            bb = Rp - Leq @ _fb_sub(Lc, Rc - LXI @ Rd)
            S = Leq @ (_fb_sub(Lc, LXI) @ Leq.T)
            S += local_dense(P1[0, 0], A[0, 0], P2[0, 0])
            S.flat[:: m + 1] += 1e-11
            sol = np.empty_like(prev)
            sol[:, 0] = sla.solve(S, bb, check_finite=False).reshape(r, n, R)
            sol[:, 2] = ((Rd - local_matvec_T(P1[0, 1], A[0, 1], P2[0, 1], sol[:, 0]).reshape(-1, 1))
                         * inv_I.reshape(-1, 1)).reshape(r, n, R)
            sol[:, 1] = _fb_sub(Lc, Rc - local_matvec(P1[2, 2], A[2, 2], P2[2, 2], sol[:, 2]).reshape(-1, 1)
                                ).reshape(r, n, R)
        except Exception:
            direct_fail = True
    if not dense or direct_fail:
        op = ReducedOperatorEq(P1, A, P2, inv_I)
        lrhs = np.empty((2, r, n, R))
        lrhs[0] = rhs[:, 0]
        lrhs[1] = rhs[:, 2] - local_matvec(P1[2, 2], A[2, 2], P2[2, 2], inv_I * rhs[:, 1])
        nrm = np.linalg.norm(lrhs)
        lvec = op.matvec(np.transpose(prev[:, :2], (1, 0, 2, 3)).reshape(-1)).reshape(2, r, n, R)
        use_prev = np.linalg.norm(lrhs - lvec) < nrm
        if use_prev:
            lrhs = lrhs - lvec
        restart = min(m, 100)
        aug = max(restart // 10, 3)
        out = _lgmres(op.matvec, lrhs.reshape(-1), rtol=rtol, max_it=300, restart=restart, augment=aug)
        if stats is not None:
            stats.append(dict(n=2 * m, its=out.its, matvecs=out.matvecs, reason=out.reason))
        part = np.transpose(out.x.reshape(2, r, n, R), (1, 0, 2, 3)).copy()
        if use_prev:
            part += prev[:, :2]
        z = inv_I * (rhs[:, 1] - local_matvec_T(P1[0, 1], A[0, 1], P2[0, 1], part[:, 0]))
        sol = np.concatenate((part, z.reshape(r, 1, n, R)), axis=1)
    res_new = np.linalg.norm(block_local_product(block_A, k, XAX_k, XAX_k1, sol) - rhs) / norm_rhs
    if res_old < res_new:
        sol = prev
    return sol, res_old, min(res_old, res_new), rhs, norm_rhs, direct_fail


def local_solver_ineq(XAX_k, block_A, k, XAX_k1, Xb_k, block_b, Xb_k1, prev, size_limit,
                      dense_solve=True, rtol=1e-5, stats=None):
    """src/tt_ipm.py:284-401 (4 blocks: dY, dX, dZ, dT)."""
    r, b, n, R = prev.shape
    m = r * n * R
    rhs = np.zeros_like(prev)
    for i in range(4):
        if i in block_b._data:
            rhs[:, i] = rhs_project(Xb_k[i], block_b._data[i][k], Xb_k1[i])
    keys = [(0, 0), (0, 1), (2, 1), (2, 2), (3, 1), (3, 3), (1, 2)]
    P1, A, P2 = _gather(block_A, k, XAX_k, XAX_k1, keys)
    inv_I = 1.0 / local_diag(P1[1, 2], A[1, 2], P2[1, 2])
    norm_rhs = max(np.linalg.norm(rhs), 1e-10)
    res_old = np.linalg.norm(block_local_product(block_A, k, XAX_k, XAX_k1, prev) - rhs) / norm_rhs
    dense = (np.sqrt(r * R) <= 0.95 * size_limit) and dense_solve and (res_old >= rtol)
    direct_fail = not dense
    sol = None
    if dense:
        try:
            Lc = sla.cholesky(local_dense(P1[2, 1], A[2, 1], P2[2, 1]), lower=True, check_finite=False)
            Rp, Rd, Rc, Rt = (rhs[:, i].reshape(m, 1) for i in range(4))
            LZc = _fb_sub(Lc, Rc)
            LZX = _fb_sub(Lc, local_dense(P1[2, 2], A[2, 2], P2[2, 2]))
            LZXI = LZX * inv_I.reshape(1, -1)
            Leq = local_dense(P1[0, 1], A[0, 1], P2[0, 1])
            Top = local_dense(P1[3, 1], A[3, 1], P2[3, 1])
            w = LZc - LZXI @ Rd
            u = Rp - Leq @ w
            v = Rt - Top @ w
            Am = local_dense(P1[0, 0], A[0, 0], P2[0, 0]) + Leq @ LZXI @ Leq.T
            D = local_dense(P1[3, 3], A[3, 3], P2[3, 3]) + Top @ LZX
            D.flat[:: m + 1] += 1e-11
            TopS = (Top @ LZXI) @ Leq.T
            LeqS = Leq @ LZX
            Dlu = sla.lu_factor(D, check_finite=False)
            rhs_l = u - LeqS @ sla.lu_solve(Dlu, v, check_finite=False)
            lhs_l = Am - LeqS @ sla.lu_solve(Dlu, TopS, check_finite=False)
            y = sla.lu_solve(sla.lu_factor(lhs_l, check_finite=False), rhs_l, check_finite=False)
            sol = np.empty_like(prev)
            sol[:, 0] = y.reshape(r, n, R)
            sol[:, 3] = sla.lu_solve(Dlu, v - TopS @ y, check_finite=False).reshape(r, n, R)
            sol[:, 2] = ((Rd - local_matvec_T(P1[0, 1], A[0, 1], P2[0, 1], sol[:, 0]).reshape(-1, 1))
                         * inv_I.reshape(-1, 1)).reshape(r, n, R) - sol[:, 3]
            sol[:, 1] = _fb_sub(Lc, Rc - local_matvec(P1[2, 2], A[2, 2], P2[2, 2], sol[:, 2]).reshape(-1, 1)
                                ).reshape(r, n, R)
        except Exception:
            direct_fail = True
    if not dense or direct_fail:
        op = ReducedOperatorIneq(P1, A, P2, inv_I)
        lrhs = np.empty((3, r, n, R))
        lrhs[0] = rhs[:, 0]
        lrhs[1] = rhs[:, 2] - local_matvec(P1[2, 2], A[2, 2], P2[2, 2], inv_I * rhs[:, 1])
        lrhs[2] = rhs[:, 3]
        nrm = np.linalg.norm(lrhs)
        lvec = op.matvec(np.transpose(prev[:, [0, 1, 3]], (1, 0, 2, 3)).reshape(-1)).reshape(3, r, n, R)
        use_prev = np.linalg.norm(lrhs - lvec) < nrm
        if use_prev:
            lrhs = lrhs - lvec
        restart = min(m, 100)
        aug = max(restart // 10, 3)
        out = _lgmres(op.matvec, lrhs.reshape(-1), rtol=rtol, max_it=300, restart=restart, augment=aug)
        if stats is not None:
            stats.append(dict(n=3 * m, its=out.its, matvecs=out.matvecs, reason=out.reason))
        part = np.transpose(out.x.reshape(3, r, n, R), (1, 0, 2, 3)).copy()
        if use_prev:
            part[:, 0] += prev[:, 0]
            part[:, 1] += prev[:, 1]
            part[:, 2] += prev[:, 3]
        z = inv_I * (rhs[:, 1] - local_matvec_T(P1[0, 1], A[0, 1], P2[0, 1], part[:, 0])) - part[:, 2]
        sol = np.concatenate((part[:, :2], z.reshape(r, 1, n, R), part[:, None, 2]), axis=1)
    res_new = np.linalg.norm(block_local_product(block_A, k, XAX_k, XAX_k1, sol) - rhs) / norm_rhs
    if res_old < res_new:
        sol = prev
    return sol, res_old, min(res_old, res_new), rhs, norm_rhs, direct_fail


# ----------------------------------------------------------------------------
# TT algebra (cy_src/tt_ops_cy.pyx, src/tt_ops.py)
# ----------------------------------------------------------------------------
def tt_ranks(tt):
    """cy_src/tt_ops_cy.pyx:82-92."""
    return [c.shape[0] for c in tt[1:]]


def tt_scale(alpha, tt):
    """cy_src/tt_ops_cy.pyx:96-114: ONE randomly drawn core is multiplied by
    float32(alpha); the draw advances the global NumPy RNG."""
    idx = np.random.randint(0, len(tt))
    out = list(tt)
    out[idx] = float(np.float32(alpha)) * tt[idx]
    return out


def _block_diag(c1, c2):
    out = np.zeros((c1.shape[0] + c2.shape[0], *c1.shape[1:-1], c1.shape[-1] + c2.shape[-1]))
    out[: c1.shape[0], ..., : c1.shape[-1]] = c1
    out[c1.shape[0]:, ..., c1.shape[-1]:] = c2
    return out


def tt_add(a, b):
    """cy_src/tt_ops_cy.pyx:244-258."""
    if len(a) == 1:
        return [a[0] + b[0]]
    mid = [_block_diag(x, y) for x, y in zip(a[1:-1], b[1:-1])]
    return [np.concatenate((a[0], b[0]), axis=-1)] + mid + [np.concatenate((a[-1], b[-1]), axis=0)]


def tt_sub(a, b):
    """src/tt_ops.py:189-190."""
    return tt_add(a, tt_scale(-1, b))


def tt_inner_prod(a, b):
    """cy_src/tt_ops_cy.pyx:506-520."""
    res = np.ones((1, 1))
    for c1, c2 in zip(a, b):
        t = np.tensordot(res, c1, axes=([0], [0]))
        ax = list(range(c1.ndim - 1))
        res = np.tensordot(t, c2, axes=(ax, ax))
    return float(res[0, 0])


def tt_norm(a):
    """src/tt_ops.py:306-310."""
    v = tt_inner_prod(a, a)
    return float(np.sqrt(v)) if v > 0 else 0.0


def tt_normalise(tt, radius=1):
    """cy_src/tt_ops_cy.pyx:524-526 (radius is a C int there)."""
    return tt_scale(np.divide(int(radius), np.sqrt(tt_inner_prod(tt, tt))), tt)


def tt_transpose(tt):
    """cy_src/tt_ops_cy.pyx:57-78."""
    split = int(np.argmax([c.ndim for c in tt]))
    return list(tt[:split]) + [np.swapaxes(c, 1, 2) for c in tt[split:]]


def tt_reshape(tt, shape):
    """src/tt_ops.py:330-333 (no core merging needed on the IPM path)."""
    return [c.reshape(c.shape[0], *shape, c.shape[-1]) for c in tt]


def tt_rl_orthogonalise(tt):
    """cy_src/tt_ops_cy.pyx:132-159; in place, norm ends in core 0."""
    d = len(tt)
    for i in range(d - 1, 0, -1):
        sh = tt[i].shape
        shm = tt[i - 1].shape
        q, rr = sla.qr(tt[i].reshape(sh[0], -1).T, mode="economic", check_finite=False)
        nr = rr.shape[0]
        tt[i] = q.T.reshape(nr, *sh[1:])
        tt[i - 1] = (tt[i - 1].reshape(-1, sh[0]) @ rr.T).reshape(*shm[:-1], nr)
    return tt


def prune_singular_vals(s, eps):
    """cy_src/tt_ops_cy.pyx:162-177."""
    if np.linalg.norm(s) == 0.0:
        return 1
    sc = np.cumsum(np.abs(s[::-1]) ** 2)[::-1]
    R = max(int(np.argmax(sc < eps ** 2)), 1)
    if sc[-1] > eps ** 2:
        R = s.size
    return R


def _round_sweep(tt, eps, collect=False):
    """Left-to-right gesvd truncation sweep shared by cy_src/tt_ops_cy.pyx:197-224,
    :283-318 and :349-384.  Returns the discarded-energy sum of the psd/mask variants."""
    d = len(tt)
    rank = 1
    dropped = 0.0
    for idx in range(d - 1):
        sh = tt[idx].shape
        shn = tt[idx + 1].shape
        u, s, vt = sla.svd(tt[idx].reshape(rank * int(np.prod(sh[1:-1])), -1), full_matrices=False,
                           check_finite=False, lapack_driver="gesvd")
        if collect:
            sc = np.cumsum(np.abs(s[::-1]) ** 2)[::-1]
            nr = max(int(np.argmax(sc < eps ** 2)), 1)
            if sc[-1] > eps ** 2:
                nr = s.size
            if nr < s.size:
                dropped += sc[nr]
        else:
            nr = prune_singular_vals(s, eps)
        tt[idx] = u[:, :nr].reshape(rank, *sh[1:-1], nr)
        tt[idx + 1] = ((s[:nr].reshape(-1, 1) * vt[:nr]) @ tt[idx + 1].reshape(shn[0], -1)
                       ).reshape(nr, *shn[1:-1], -1)
        rank = nr
    return dropped


def _all_rank_one(tt):
    return len(tt) == 1 or all(r == 1 for r in tt_ranks(tt))


def tt_rank_reduce(tt, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:180-226; mutates and returns the input list."""
    if _all_rank_one(tt):
        return tt
    eps = eps / np.sqrt(len(tt) - 1)
    tt = tt_rl_orthogonalise(tt)
    _round_sweep(tt, eps)
    return tt


def tt_psd_rank_reduce(tt, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:262-325."""
    d = len(tt)
    eps = eps / 2.0
    if _all_rank_one(tt):
        return tt
    eps = eps / np.sqrt(d - 1)
    tt = tt_rl_orthogonalise(tt)
    dropped = _round_sweep(tt, eps, collect=True)
    factor = pow(dropped, 1.0 / (2 * d))
    I = factor * np.eye(tt[0].shape[1]).reshape(1, *tt[0].shape[1:-1], 1)
    return tt_add(tt, [I] * d)


def tt_mask_rank_reduce(tt, mask_tt, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:329-388."""
    d = len(tt)
    eps = eps / 2.0
    if _all_rank_one(tt):
        return tt
    eps = eps / np.sqrt(d - 1)
    tt = tt_rl_orthogonalise(tt)
    dropped = _round_sweep(tt, eps, collect=True)
    factor = pow(dropped, 1.0 / (2 * d))
    return tt_add(tt, [factor * c for c in mask_tt])


def _swap_cores(ca, cb, eps):
    """cy_src/tt_ops_cy.pyx:393-426."""
    if ca.ndim == 3:
        t = np.tensordot(ca, cb, axes=([2], [0])).transpose(0, 2, 1, 3)
        u, s, v = sla.svd(t.reshape(ca.shape[0] * cb.shape[1], -1), full_matrices=False,
                          check_finite=False, lapack_driver="gesvd")
        rp = prune_singular_vals(s, eps)
        return ((u[:, :rp] * s[:rp]).reshape(ca.shape[0], cb.shape[1], -1),
                v[:rp].reshape(-1, ca.shape[1], cb.shape[2]))
    t = np.tensordot(ca, cb, axes=([3], [0])).transpose(0, 3, 4, 1, 2, 5)
    u, s, v = sla.svd(t.reshape(ca.shape[0] * cb.shape[1] * cb.shape[2], -1), full_matrices=False,
                      check_finite=False, lapack_driver="gesvd")
    rp = prune_singular_vals(s, eps)
    return ((u[:, :rp] * s[:rp]).reshape(ca.shape[0], cb.shape[1], cb.shape[2], -1),
            v[:rp].reshape(-1, ca.shape[1], ca.shape[2], cb.shape[3]))


def _zipup(first_fn, d, cores, eps):
    loop_eps = eps / np.sqrt(d - 1) if d > 1 else eps
    for i in range(d):
        cores[0] = first_fn(d - 1 - i, cores[0])
        if i != d - 1:
            for j in range(i, -1, -1):
                cores[j], cores[j + 1] = _swap_cores(cores[j], cores[j + 1], loop_eps)
    return cores


def tt_fast_matrix_vec_mul(M, v, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:430-447."""
    cores = [np.transpose(c, (2, 1, 0)) for c in reversed(v)]
    return _zipup(lambda p, c0: np.tensordot(M[p], c0, axes=([3, 2], [0, 1])), len(M), cores, eps)


def tt_fast_mat_mat_mul(A, B, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:451-464."""
    cores = [np.transpose(c, (3, 1, 2, 0)) for c in reversed(B)]
    return _zipup(lambda p, c0: np.tensordot(A[p], c0, axes=([3, 2], [0, 1])), len(A), cores, eps)


def tt_fast_hadamard(a, b, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:468-502."""
    if a[0].ndim == 4 and b[0].ndim == 4:
        cores = [np.transpose(c, (3, 1, 2, 0)) for c in reversed(b)]

        def first4(p, c0):          # 'rijR,RijK->rijK', same operation order as the reference
            t = np.tensordot(a[p], c0, axes=([3], [0]))
            t = np.diagonal(np.diagonal(t, axis1=1, axis2=3), axis1=1, axis2=2)
            return t.transpose(0, 2, 3, 1)
        return _zipup(first4, len(a), cores, eps)
    cores = [np.transpose(c, (2, 1, 0)) for c in reversed(b)]

    def first3(p, c0):              # 'riR,RiK->riK'
        t = np.tensordot(a[p], c0, axes=([2], [0]))
        return np.diagonal(t, axis1=1, axis2=2).transpose(0, 2, 1)
    return _zipup(first3, len(a), cores, eps)


def symmetric_powers_of_two(length):
    """cy_src/tt_ops_cy.pyx:538-555: 2, 4, 8, ... rising to the middle bond and mirrored."""
    half = (length + 1) // 2
    up = [1 << (i + 1) for i in range(half)]
    return np.array((up + up[:length // 2][::-1])[:length] if length > 0 else [], dtype=np.int64)


def add_kick_rank(u, v, r_add=2):
    """cy_src/tt_ops_cy.pyx:559-579: enrich the orthonormal factor with r_add Gaussian columns
    (global RNG), re-orthogonalise, carry the R factor into v."""
    old_r = u.shape[1]
    uk = np.random.randn(u.shape[0], r_add)
    q, rm = sla.qr(np.ascontiguousarray(np.concatenate((u, uk), axis=1)), mode="economic", check_finite=False)
    return q, rm[:, :old_r] @ v, q.shape[1]


def _als_fit_product(A, D, x0, kick_rank, nswp, tol, trace=None):
    """ALS fit of the TT product A*D, one engine for
    tt_approx_mat_mat_mul (src/tt_als.py:1502-1628, D cores (b,k,n,B)) and
    tt_approx_mat_vec_mul (src/tt_als.py:1637-1762, D cores (b,k,B) handled as n = 1).

    State as in the reference: unit-norm interfaces G[k] (r, a, b) with their norms in
    normAD, unit-norm non-orthogonal neighbour cores with their norms in normx, and the
    running scale nrmsc that converts a raw local contraction into the local solution.
    The three 4-operand contractions are explicit pairwise tensordot chains (BLAS), which is what
    opt_einsum does for the reference; np.einsum's greedy path falls back to its non-BLAS loop here."""
    vec = D[0].ndim == 3
    D4 = [c.reshape(c.shape[0], c.shape[1], 1, c.shape[2]) for c in D] if vec else D
    if x0 is None:
        max_ranks = np.maximum((np.array(tt_ranks(A)) + np.array(tt_ranks(D))) / 2, 2).astype(int)
        x = tt_random_gaussian(list(max_ranks), (A[0].shape[2],) if vec else A[0].shape[1:-1])
    else:
        x, max_ranks = x0, np.array(tt_ranks(x0))
    if kick_rank is None:
        kick_rank = np.maximum((symmetric_powers_of_two(len(A) - 1) - max_ranks) / (nswp / 2), 2).astype(int)
    d = len(x)
    x = [c.reshape(c.shape[0], c.shape[1], 1, c.shape[2]) if vec else c for c in x]
    rx = np.array([1] + tt_ranks(x) + [1])
    G = [np.ones((1, 1, 1))] + [None] * (d - 1) + [np.ones((1, 1, 1))]
    normAD, normx, nrmsc = np.ones(d - 1), np.ones(d - 1), 1.0
    tol = tol / np.sqrt(d)
    st = {"max_res": 0.0}

    def local(k):
        t = np.tensordot(G[k], D4[k], axes=([2], [0]))                       # 'rab,amkA,bknB,RAB->rmnR'
        t = np.tensordot(t, A[k], axes=([1, 2], [0, 2]))                     # (r, n, B, m, A')
        sol = np.tensordot(t, G[k + 1], axes=([4, 2], [1, 2])).transpose(0, 2, 1, 3) * nrmsc
        res = np.linalg.norm(sol - x[k]) / max(np.linalg.norm(sol), 1e-8)
        st["max_res"] = max(st["max_res"], res)
        return sol

    def split(mat, bond, last):
        u, s, v = sla.svd(mat, full_matrices=False, check_finite=False, lapack_driver="gesvd")
        v = s.reshape(-1, 1) * v
        r = prune_singular_vals(s, tol)
        if not last:
            return add_kick_rank(u[:, :r], v[:r], kick_rank[bond])
        return u[:, :r], v[:r], r

    last = False
    for swp in range(nswp):
        st["max_res"] = np.inf if swp == 0 else 0.0
        for k in range(d - 1, -1, -1):                                   # :1531-1565 / :1666-1700
            sol = local(k) if swp > 0 else x[k]
            nm = x[k].shape[1:3]
            if k == 0:
                x[k] = sol.reshape(rx[k], *nm, rx[k + 1])
                continue
            u, v, r = split(sol.reshape(rx[k], -1).T, k - 1, last)
            nrmsc *= normx[k - 1] / normAD[k - 1]
            x[k] = u.T.reshape(r, *nm, rx[k + 1])
            x[k - 1] = np.tensordot(x[k - 1], v.T, axes=([3], [0]))
            nrm = np.linalg.norm(x[k - 1])
            normx[k - 1] *= nrm
            x[k - 1] = x[k - 1] / nrm
            rx[k] = r
            t = np.tensordot(x[k], G[k + 1], axes=([3], [0]))                # 'RAB,amkA,bknB,rmnR->rab'
            t = np.tensordot(t, D4[k], axes=([2, 4], [2, 3]))                # (r, m, A', b, k)
            G[k] = np.tensordot(t, A[k], axes=([1, 2, 4], [1, 3, 2])).transpose(0, 2, 1)
            nrm = np.linalg.norm(G[k])
            nrm = nrm if nrm > 0 else 1.0
            G[k] = G[k] / nrm
            normAD[k - 1] = nrm
            nrmsc *= normAD[k - 1] / normx[k - 1]
        if trace is not None:
            trace.append(("bck", swp, float(st["max_res"]), [int(q) for q in rx]))
        if last:
            break
        if st["max_res"] < tol or swp == nswp - 1:
            last = True
        st["max_res"] = 0.0
        for k in range(d):                                               # :1571-1605 / :1706-1740
            sol = local(k)
            nm = x[k].shape[1:3]
            if k == d - 1:
                x[k] = sol.reshape(rx[k], *nm, rx[k + 1])
                continue
            nrmsc *= normx[k] / normAD[k]
            u, v, r = split(sol.reshape(-1, rx[k + 1]), k, last)
            x[k] = u.reshape(rx[k], *nm, r)
            x[k + 1] = np.tensordot(v, x[k + 1], axes=([1], [0]))
            nrm = np.linalg.norm(x[k + 1])
            normx[k] *= nrm
            x[k + 1] = x[k + 1] / nrm
            rx[k + 1] = r
            t = np.tensordot(G[k], x[k], axes=([0], [0]))                    # 'rab,amkA,bknB,rmnR->RAB'
            t = np.tensordot(t, A[k], axes=([0, 2], [0, 1]))                 # (b, n, R, k, A')
            G[k + 1] = np.tensordot(t, D4[k], axes=([0, 1, 3], [0, 2, 1]))
            nrm = np.linalg.norm(G[k + 1])
            nrm = nrm if nrm > 0 else 1.0
            G[k + 1] = G[k + 1] / nrm
            normAD[k] = nrm
            nrmsc *= normAD[k] / normx[k]
        if trace is not None:
            trace.append(("fwd", swp, float(st["max_res"]), [int(q) for q in rx]))
        if last:
            break
        if st["max_res"] < tol:
            last = True
    scale = np.exp(np.sum(np.log(normx)) / d)
    return [scale * (c[:, :, 0] if vec else c) for c in x]


def tt_approx_mat_mat_mul(A, D, x0=None, kick_rank=None, nswp=50, tol=1e-6, trace=None):
    """src/tt_als.py:1502-1628."""
    return _als_fit_product(A, D, x0, kick_rank, nswp, tol, trace)


def tt_approx_mat_vec_mul(A, d_vec, x0=None, kick_rank=None, nswp=50, tol=1e-6, trace=None):
    """src/tt_als.py:1637-1762."""
    return _als_fit_product(A, d_vec, x0, kick_rank, nswp, tol, trace)


def tt_random_gaussian(target_ranks, shape=(2,)):
    """cy_src/tt_ops_cy.pyx:529-533."""
    rr = [1] + list(target_ranks) + [1]
    return tt_normalise([np.divide(1, a * np.prod(shape) * b) * np.random.randn(a, *shape, b)
                         for a, b in zip(rr[:-1], rr[1:])])


def tt_mat_vec_mul(mat, vec, op_tol, eps):
    """src/tt_als.py:1765-1768."""
    if np.max(np.array(tt_ranks(mat)) * np.array(tt_ranks(vec))) <= 80:
        return tt_rank_reduce(tt_fast_matrix_vec_mul(mat, vec, eps), op_tol)
    return tt_approx_mat_vec_mul(mat, vec, tol=op_tol)


def tt_mat_mat_mul(a, b, op_tol, eps):
    """src/tt_als.py:1631-1634."""
    if np.max(np.array(tt_ranks(a)) * np.array(tt_ranks(b))) <= 40:
        return tt_rank_reduce(tt_fast_mat_mat_mul(a, b, eps), eps=op_tol)
    return tt_approx_mat_mat_mul(a, b, tol=op_tol)


def tt_IkronM(M):
    """src/tt_ops.py:360-363."""
    I = np.eye(2)
    return [np.einsum("mn,rijR->rminjR", I, c).reshape(c.shape[0], 4, 4, c.shape[-1]) for c in M]


def tt_MkronI(M):
    """src/tt_ops.py:365-368."""
    I = np.eye(2)
    return [np.einsum("rmnR,ij->rminjR", c, I).reshape(c.shape[0], 4, 4, c.shape[-1]) for c in M]


def _diag_embed(c3):
    out = np.zeros((c3.shape[0], c3.shape[1], c3.shape[1], c3.shape[2]))
    idx = np.arange(c3.shape[1])
    out[:, idx, idx, :] = c3
    return out


def tt_diag(vec_tt, eps=1e-18):
    """src/tt_ops.py:312-316."""
    return tt_rank_reduce([_diag_embed(c) for c in vec_tt], eps)


def tt_diag_op(matrix_tt, eps=1e-18):
    """src/tt_ops.py:371-375."""
    return tt_rank_reduce([_diag_embed(c.reshape(c.shape[0], -1, c.shape[-1])) for c in matrix_tt], eps)


def tt_entrywise_sum(tt):
    """src/tt_ops.py:342-352."""
    res = np.ones((1,))
    for c in tt:
        res = res @ c.reshape(c.shape[0], -1, c.shape[-1]).sum(axis=1)
    return float(np.sum(res))


def tt_rl_orthogonalise_py(tt):
    """src/tt_ops.py:30-42: loops down to i = 0, so core 0 is normalised and its
    1x1 R factor wraps around into the LAST core (index -1)."""
    d = len(tt)
    if d == 1:
        return tt
    for i in range(d - 1, -1, -1):
        sh = tt[i].shape
        shm = tt[i - 1].shape
        q, rr = sla.qr(tt[i].reshape(sh[0], -1).T, mode="economic", check_finite=False)
        tt[i] = q.T.reshape(-1, *sh[1:-1], sh[-1])
        tt[i - 1] = (tt[i - 1].reshape(-1, rr.shape[-1]) @ rr.T).reshape(-1, *shm[1:-1], tt[i].shape[0])
    return tt


def tt_rank_retraction(tt, upper_ranks):
    """src/tt_ops.py:132-152: cap ranks by top-k SVD (argpartition order kept)."""
    tt = tt_rl_orthogonalise_py(tt)
    rank = 1
    for idx, ur in enumerate(upper_ranks):
        sh = tt[idx].shape
        shn = tt[idx + 1].shape
        U, S, Vt = sla.svd(tt[idx].reshape(rank * int(np.prod(sh[1:-1])), -1), full_matrices=False,
                           check_finite=False, lapack_driver="gesvd")
        nr = min(ur, len(S))
        sel = np.argpartition(np.abs(S), -nr)[-nr:]
        tt[idx] = U[:, sel].reshape(rank, *sh[1:-1], nr)
        tt[idx + 1] = (np.diag(S[sel]) @ Vt[sel] @ tt[idx + 1].reshape(Vt.shape[-1], -1)
                       ).reshape(nr, *shn[1:-1], -1)
        rank = nr
    return tt


def tt_get_block(i, tt):
    """src/tt_als.py:12-14."""
    b = int(np.argmax([c.ndim for c in tt]))
    return list(tt[:b]) + [tt[b][:, i]] + list(tt[b + 1:])


# ----------------------------------------------------------------------------
# Block AMEn (src/tt_als.py:277-825)
# ----------------------------------------------------------------------------
def _truncated_svd(mat, k):
    """src/tt_als.py:269-274 (gesdd)."""
    u, s, v = sla.svd(mat, full_matrices=False, check_finite=False)
    return u[:, :k], s[:k].reshape(-1, 1) * v[:k]


def _block_scales(sol):
    """src/tt_als.py:321 per-block norm equilibration."""
    nb = sol.shape[1]
    return np.maximum(np.array([np.linalg.norm(sol[:, j]) for j in range(nb)]), 1e-10).reshape(1, -1, 1, 1)


class AmenState:
    pass


def _sweep(st, direction, swp, last, local_solver, trace=None):
    """One half-sweep; direction > 0 is the reference's _bck_sweep (k = d-1..0,
    src/tt_als.py:277-394), direction < 0 its _fwd_sweep (:397-522)."""
    d, bs, N = st.d, st.block_size, st.N
    A, rhsb = st.block_A, st.block_b
    x, z, XAX, ZAX, Xb, Zb, rx, rz = st.x, st.z, st.XAX, st.ZAX, st.Xb, st.Zb, st.rx, st.rz
    bck = direction > 0
    local_res = np.inf if swp == 0 else 0
    local_dx = np.inf if swp == 0 else 0
    order = range(d - 1, -1, -1) if bck else range(d)
    solving = swp > 0 and not last
    for k in order:
        inner = (k > 0) if bck else (k < d - 1)
        if solving:
            prev = x[k]
            sol, r_old, r_new, rhs, norm_rhs, st.direct_solve_failure = local_solver(
                XAX[k], A, k, XAX[k + 1], Xb[k], rhsb, Xb[k + 1], prev, 3 * d, not st.direct_solve_failure)
            if trace is not None:
                trace.append(dict(swp=swp, k=k, res_old=float(r_old), res_new=float(r_new), shape=prev.shape))
            local_res = max(local_res, r_old)
            local_dx = max(local_dx, np.linalg.norm(sol - prev) / np.linalg.norm(sol))
            if st.amen:
                zshape = (rz[k], bs, N[k], rz[k + 1])
                Az = mixed_block_local_product(A, k, ZAX[k], ZAX[k + 1], sol, zshape, True, True)
                resz = block_rhs_product(rhsb, k, Zb[k], Zb[k + 1], zshape) - Az
        else:
            sol = x[k]
            resz = z[k] if (st.amen and not last) else None
        scales = _block_scales(sol)
        sol = scales * sol
        if bck:
            mat = sol.reshape(rx[k] * bs, N[k] * rx[k + 1]).T                 # (n R, r b)
            rzm = resz.reshape(rz[k] * bs, N[k] * rz[k + 1]).T if resz is not None else None
        else:
            mat = sol.transpose(0, 2, 1, 3).reshape(rx[k] * N[k], bs * rx[k + 1])   # (r n, b R)
            rzm = (resz.transpose(0, 2, 1, 3).reshape(rz[k] * N[k], bs * rz[k + 1])
                   if resz is not None else None)
        if not inner:
            if bck:
                x[k] = mat.T.reshape(rx[k], bs, N[k], rx[k + 1]) / scales
                if st.amen and not last:
                    z[k] = rzm.T.reshape(rz[k], bs, N[k], rz[k + 1]) / scales
            else:
                x[k] = mat.reshape(rx[k], N[k], bs, rx[k + 1]).transpose(0, 2, 1, 3) / scales
                if st.amen and not last:
                    z[k] = rzm.reshape(rz[k], N[k], bs, rz[k + 1]).transpose(0, 2, 1, 3) / scales
            continue

        u, s, v = sla.svd(mat, full_matrices=False, check_finite=False)
        v = s.reshape(-1, 1) * v

        def unfold_to_core(m2):
            # m2: product u_part @ v_part in the unfolding layout -> (r, b, n, R)
            if bck:
                return m2.T.reshape(rx[k], bs, N[k], rx[k + 1])
            return m2.reshape(rx[k], N[k], bs, rx[k + 1]).transpose(0, 2, 1, 3)

        if solving:
            trunc_lim = max(2 * st.trunc_tol, r_new)
            r0 = min(prune_singular_vals(s, st.eps), st.r_max)
            sol_r0 = unfold_to_core(u[:, :r0] @ v[:r0])
            res = block_local_product(A, k, XAX[k], XAX[k + 1], sol_r0) - rhs
            r = r0
            for r in range(r0 - 1, 0, -1):
                res -= block_local_product(A, k, XAX[k], XAX[k + 1], unfold_to_core(u[:, r:r + 1] @ v[r:r + 1]))
                if np.linalg.norm(res) / norm_rhs > trunc_lim:
                    break
            r += 1
            uk, vk = u[:, :r], v[:r]
            r = uk.shape[1]
            if st.amen:
                if bck:
                    eshape = (rz[k], bs, N[k], rx[k + 1])
                    Axz = mixed_block_local_product(A, k, ZAX[k], XAX[k + 1], sol_r0, eshape, True, False)
                    resxz = block_rhs_product(rhsb, k, Zb[k], Xb[k + 1], eshape) - Axz
                    kr = min(st.kick_rank, rz[k] * bs, N[k] * rx[k + 1])
                    uz, _ = _truncated_svd(resxz.reshape(rz[k] * bs, N[k] * rx[k + 1]).T, kr)
                else:
                    eshape = (rx[k], bs, N[k], rz[k + 1])
                    Axz = mixed_block_local_product(A, k, XAX[k], ZAX[k + 1], unfold_to_core(uk @ vk),
                                                    eshape, False, True)
                    resxz = block_rhs_product(rhsb, k, Xb[k], Zb[k + 1], eshape) - Axz
                    kr = min(st.kick_rank, rx[k] * N[k], bs * rz[k + 1])
                    uz, _ = _truncated_svd(resxz.transpose(0, 2, 1, 3).reshape(rx[k] * N[k], bs * rz[k + 1]), kr)
                q, Rf = sla.qr(np.concatenate((uk, uz), axis=1), mode="economic", check_finite=False)
                vk = Rf[:, :r] @ vk
                uk = q
                r = uk.shape[1]
        else:
            r = min(prune_singular_vals(s, st.eps), st.r_max)
            uk, vk = u[:, :r], v[:r]

        if bck:
            x[k] = uk.T.reshape(r, N[k], rx[k + 1])
            vv = vk.T.reshape(rx[k], bs, r)
            x[k - 1] = np.einsum("adc,cbR->abdR", x[k - 1], vv) / scales
            rx[k] = r
            XAX[k] = {key: phi_bck(XAX[k + 1][key], x[k], A._data[key][k], x[k]) for key in A._data}
            Xb[k] = {i: phi_bck_rhs(Xb[k + 1][i], rhsb._data[i][k], x[k]) for i in rhsb._data}
        else:
            x[k] = uk.reshape(rx[k], N[k], r)
            vv = vk.reshape(r, bs, rx[k + 1])
            x[k + 1] = np.einsum("rbR,Rdk->rbdk", vv, x[k + 1]) / scales
            rx[k + 1] = r
            XAX[k + 1] = {key: phi_fwd(XAX[k][key], x[k], A._data[key][k], x[k]) for key in A._data}
            Xb[k + 1] = {i: phi_fwd_rhs(Xb[k][i], rhsb._data[i][k], x[k]) for i in rhsb._data}

        if st.amen and not last:
            kr = min(st.kick_rank, *rzm.shape)
            uz, vz = _truncated_svd(rzm, kr)
            if bck:
                z[k] = uz.T.reshape(kr, N[k], rz[k + 1])
                vzz = vz.T.reshape(rz[k], bs, kr)
                z[k - 1] = np.einsum("adc,cbR->abdR", z[k - 1], vzz) / scales
                rz[k] = kr
                ZAX[k] = {key: phi_bck(ZAX[k + 1][key], z[k], A._data[key][k], x[k]) for key in A._data}
                for (i, j), (p, t) in A._transposes.items():
                    ZAX[k][p, t] = phi_bck(ZAX[k + 1][p, t], z[k], A._data[i, j][k].transpose(0, 2, 1, 3), x[k])
                Zb[k] = {i: phi_bck_rhs(Zb[k + 1][i], rhsb._data[i][k], z[k]) for i in rhsb._data}
            else:
                z[k] = uz.reshape(rz[k], N[k], kr)
                vzz = vz.reshape(kr, bs, rz[k + 1])
                z[k + 1] = np.einsum("rbR,Rdk->rbdk", vzz, z[k + 1]) / scales
                rz[k + 1] = kr
                ZAX[k + 1] = {key: phi_fwd(ZAX[k][key], z[k], A._data[key][k], x[k]) for key in A._data}
                for (i, j), (p, t) in A._transposes.items():
                    ZAX[k + 1][p, t] = phi_fwd(ZAX[k][p, t], z[k], A._data[i, j][k].transpose(0, 2, 1, 3), x[k])
                Zb[k + 1] = {i: phi_fwd_rhs(Zb[k][i], rhsb._data[i][k], z[k]) for i in rhsb._data}
    return local_res, local_dx


def tt_block_amen(block_A, block_b, term_tol, r_max=100, eps=1e-12, nswp=22, x0=None,
                  local_solver=None, kick_rank=2, amen=False, trace=None):
    """src/tt_als.py:525-670.  Same NumPy RNG draw order as the reference."""
    st = AmenState()
    st.block_A, st.block_b = block_A, block_b
    st.block_size = bs = block_size_of(block_A)
    model = next(iter(block_b._data.values()))
    x_shape = model[0].shape[1:-1]

    def fresh():
        return (tt_normalise([np.random.randn(1, *c.shape[1:-1], 1) for c in model[:-1]])
                + [np.random.randn(1, bs, *x_shape, 1)])

    direction = 1
    if x0 is None:
        x = fresh()
    else:
        x = x0
        where = [i for i, c in enumerate(x) if c.ndim == 4 and c.shape[1] == bs]
        if len(where) != 1 or where[0] not in (0, len(x) - 1):
            x = fresh()
        elif where[0] == 0:
            direction = -1
    st.x = x
    st.N = [c.shape[-2] for c in x]
    st.d = d = len(st.N)
    one3 = lambda: np.ones((1, 1, 1))
    one2 = lambda: np.ones((1, 1))
    st.XAX = [{key: one3() for key in block_A._data}] + [dict() for _ in range(d - 1)] + \
             [{key: one3() for key in block_A._data}]
    st.Xb = [{i: one2() for i in block_b._data}] + [dict() for _ in range(d - 1)] + \
            [{i: one2() for i in block_b._data}]
    st.rx = np.array([1] + tt_ranks(x) + [1])
    st.amen = amen
    st.ZAX = st.Zb = st.z = st.rz = None
    if amen:
        tk = block_A.tkeys()
        st.ZAX = [{key: one3() for key in tk}] + [dict() for _ in range(d - 1)] + [{key: one3() for key in tk}]
        st.Zb = [{i: one2() for i in block_b._data}] + [dict() for _ in range(d - 1)] + \
                [{i: one2() for i in block_b._data}]
        kr = kick_rank
        z = [np.divide(1, np.prod(x[0].shape[1:-1]) * kr ** 2) * np.random.randn(*x[0].shape[:-1], kr)]
        z += [np.divide(1, np.prod(c.shape[1:-1]) * kr ** 2) * np.random.randn(kr, *c.shape[1:-1], kr)
              for c in x[1:-1]]
        z += [np.divide(1, np.prod(x[-1].shape[1:-1]) * kr ** 2) * np.random.randn(kr, *x[-1].shape[1:])]
        st.z = z
        st.rz = np.array([1] + tt_ranks(z) + [1])
    st.eps, st.r_max, st.kick_rank = eps, r_max, kick_rank
    st.trunc_tol = term_tol / np.sqrt(d)
    st.direct_solve_failure = False
    last = False
    final_res = np.inf
    sweeps = 0
    for swp in range(nswp + 1):
        local_res, local_dx = _sweep(st, direction, swp, last, local_solver, trace)
        sweeps = swp
        if last:
            break
        if local_res < term_tol or local_dx < eps or swp == nswp - 2:
            last = True
            final_res = local_res
        direction *= -1
    st.sweeps = sweeps
    return st.x, final_res, st


def block_norm(block_b):
    """src/tt_als.py:45-47."""
    return float(np.sqrt(sum(tt_inner_prod(v, v) for v in block_b._data.values())))


def block_product(block_A, x_cores, op_tol, eps=1e-12):
    """src/tt_als.py:132-155."""
    out = {}

    def acc(row, term):
        out[row] = tt_rank_reduce(tt_add(out[row], term), eps) if row in out else term

    for (i, j), blk in block_A._data.items():
        acc(i, tt_mat_vec_mul(blk, tt_get_block(j, x_cores), op_tol, eps))
        if (i, j) in block_A._transposes:
            p, t = block_A._transposes[i, j]
            acc(p, tt_mat_vec_mul(tt_transpose(blk), tt_get_block(t, x_cores), op_tol, eps))
        if (i, j) in block_A._aliases:
            p, t = block_A._aliases[i, j]
            acc(p, tt_mat_vec_mul(blk, tt_get_block(t, x_cores), op_tol, eps))
    return BlockVector(out)


def block_sub(a, b):
    """src/tt_als.py:49-53."""
    return BlockVector({i: tt_rank_reduce(tt_sub(a._data[i], b._data[i]), 1e-12) for i in a._data})


def tt_restarted_block_amen(block_A, block_b, rank_restriction, op_tol, termination_tol=1e-3, eps=1e-11,
                            num_restarts=3, inner_m=10, x0=None, local_solver=None, trace=None):
    """src/tt_als.py:744-825."""
    if x0 is not None:
        dim = len(x0)
        x0 = tt_rank_retraction(x0, [dim] * (dim - 1))
    orig = block_norm(block_b)
    if orig < 0.5 * op_tol:
        raise RuntimeError("Absolute tolerance already reached")
    x, res, _ = tt_block_amen(block_A, block_b, termination_tol, r_max=rank_restriction, eps=eps, nswp=inner_m,
                              x0=x0, local_solver=local_solver, kick_rank=2, amen=True, trace=trace)
    if res < termination_tol:
        return x, res
    rn = block_norm(block_sub(block_b, block_product(block_A, x, 0.1 * op_tol)))
    if rn < termination_tol * orig or rn < orig:
        return x, res
    for _ in range(1, num_restarts):
        dim = len(x)
        x = tt_rank_retraction(x, [2 * dim] * (dim - 1))
        x, res, _ = tt_block_amen(block_A, block_b, termination_tol, r_max=rank_restriction + 4, eps=eps,
                                  nswp=inner_m, x0=x, local_solver=local_solver, kick_rank=4, amen=True,
                                  trace=trace)
        rn = block_norm(block_sub(block_b, block_product(block_A, x, 0.1 * op_tol)))
        if rn < termination_tol * orig or rn < orig:
            return x, res
    raise RuntimeError("Number of restarts exhausted")
