"""TEST INFRASTRUCTURE (oracle): CPU restatement of the reference's step-size eigen sweeps
`tt_max_generalised_eigen` (src/tt_als.py:1132-1283) and `tt_min_eig` (:1392-1499) with their local solvers
(:931-1128, :1286-1389) in NumPy / SciPy.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product path never does.

Pinned against the real reference: tests/golden/eigen_*.npz hold calls traced from the reference IPM (inputs, RNG
state, returned step size / eigenvector train); tests/test_oracle_vs_golden.py checks the step size to 1e-6 relative
(the same bar the device path is held to) and the Rayleigh quotient of the eigenvector train.

Restatement notes.  The reference finds the extreme local eigenpairs with ARPACK (`eigsh`, start vector = previous
solution) or LOBPCG; both converge to the eigenpair that `scipy.linalg.eigh` returns directly, which is what this
oracle uses (the local problems are at most a few hundred unknowns wide in every traced call).  Truncation rule,
enrichment (`_add_kick_rank`, `_add_kick_rank_rev`) with its NumPy RNG draws, sweep order, stopping rules and the
final rescaling of the step follow the reference line by line.
"""
import numpy as np
import scipy.linalg as sla

import tt_oracle as O


def _proj2(P1, A1, A2, P2):
    """'lsr,smnk,kptS,LSR->lmpLrntR' as a symmetric matrix (reference :952-959)."""
    M = np.einsum("lsr,smnk,kptS,LSR->lmpLrntR", P1, A1, A2, P2, optimize=True)
    m = int(np.prod(M.shape[:4]))
    M = M.reshape(m, m)
    return 0.5 * (M + M.T)


def _proj1(P1, A, P2):
    """'lsr,smnS,LSR->lmLrnR' (reference :1037-1041)."""
    M = np.einsum("lsr,smnS,LSR->lmLrnR", P1, A, P2, optimize=True)
    m = int(np.prod(M.shape[:3]))
    M = M.reshape(m, m)
    return 0.5 * (M + M.T)


def _smallest(M):
    w, v = sla.eigh(M, subset_by_index=[0, 0])
    return float(w[0]), v[:, 0]


def _pencil(A, D, prev, step, eps):
    """Local step-size problem (reference :960-993): returns (unit solution, step, old_res)."""
    M = A / step + D
    lam, sol = _smallest(M)
    if lam < 0:
        try:
            w, v = sla.eigh(-D, A, subset_by_index=[A.shape[0] - 1, A.shape[0] - 1])
            step = max(0.0, min(step, 1.0 / w[0]))
            sol = v[:, 0]
        except Exception:
            sol = prev.copy()
            step *= (1 - eps)
    if step > 0:
        M = A / step + D
        rq = prev @ (M @ prev)
        old_res = float(np.linalg.norm(M @ prev - rq * prev))
    else:
        old_res = np.inf
    return sol / np.linalg.norm(sol), step, old_res


def _kick_fwd(u, v, r_add=4):
    """reference _add_kick_rank (:1108-1113)."""
    old = u.shape[-1]
    uk = np.random.randn(u.shape[0], r_add)
    q, rm = sla.qr(np.concatenate((u, uk), 1), mode="economic")
    return q, rm[:, :old] @ v


def _kick_bwd(u, v, r_add=4):
    """reference _add_kick_rank_rev (:1115-1120)."""
    old = v.shape[0]
    uk = np.random.randn(r_add, v.shape[-1])
    rm, q = sla.rq(np.concatenate((v, uk), 0), mode="economic")
    return u @ rm[:old], q


def _split_two(sol, shape, trunc_tol, max_rank, bwd):
    r0, n, t, R = shape
    mat = sol.reshape(r0 * n, t * R)
    if bwd:
        u, s, vt = sla.svd(mat.T, full_matrices=False, lapack_driver="gesvd")
        vt = s[:, None] * vt
        r = min(O.prune_singular_vals(s, trunc_tol), max_rank)
        a, b = _kick_bwd(vt[:r].T, u[:, :r].T)
    else:
        u, s, vt = sla.svd(mat, full_matrices=False, lapack_driver="gesvd")
        r = min(O.prune_singular_vals(s, trunc_tol), max_rank)
        a, b = _kick_fwd(u[:, :r], s[:r, None] * vt[:r])
    k = a.shape[1]
    return a.reshape(r0, n, k), b.reshape(k, t, R)


class _State:
    def __init__(self, mats, x, trunc_tol):
        self.mats = mats
        self.x = x
        self.d = len(x)
        self.trunc_tol = trunc_tol
        self.max_rank = int(np.floor(2 ** (self.d / 2)))
        self.phi = [[np.ones((1, 1, 1))] + [None] * (self.d - 1) + [np.ones((1, 1, 1))] for _ in mats]

    def bck(self, k):
        for M, ph in zip(self.mats, self.phi):
            ph[k] = O.phi_bck(ph[k + 1], self.x[k], M[k], self.x[k])

    def fwd(self, k):
        for M, ph in zip(self.mats, self.phi):
            ph[k + 1] = O.phi_fwd(ph[k], self.x[k], M[k], self.x[k])

    def rank(self, s):
        return min(O.prune_singular_vals(s, self.trunc_tol), self.max_rank)

    def orth_bck(self, k, sol=None):
        """x[k] -> right-orthogonal, remainder into x[k-1] (reference :1168-1174, :1243-1250)."""
        r0, n, r1 = self.x[k].shape
        mat = (self.x[k] if sol is None else sol).reshape(r0, n * r1).T
        u, s, vt = sla.svd(mat, full_matrices=False, lapack_driver="gesvd")
        vt = s[:, None] * vt
        r = self.rank(s)
        self.x[k] = u[:, :r].T.reshape(r, n, r1)
        self.x[k - 1] = np.einsum("rdc,cR->rdR", self.x[k - 1], vt[:r].T)

    def orth_fwd(self, k, sol):
        """reference :1190-1196."""
        r0, n, r1 = self.x[k].shape
        u, s, vt = sla.svd(sol.reshape(r0 * n, r1), full_matrices=False, lapack_driver="gesvd")
        vt = s[:, None] * vt
        r = self.rank(s)
        self.x[k] = u[:, :r].reshape(r0, n, r)
        self.x[k + 1] = np.einsum("ij,jkl->ikl", vt[:r], self.x[k + 1])

    def two(self, k):
        prev = np.einsum("rny,ytR->rntR", self.x[k], self.x[k + 1])
        return prev.reshape(-1), prev.shape

    def proj(self, which, k, two):
        M, ph = self.mats[which], self.phi[which]
        if two:
            return _proj2(ph[k], M[k], M[k + 1], ph[k + 2])
        return _proj1(ph[k], M[k], ph[k + 1])


def _stalled(prev_res, res, tol):
    return bool(np.isfinite(prev_res) and np.isfinite(res) and res <= 50 * tol and res >= 0.8 * prev_res)


def tt_max_generalised_eigen(A, Delta, x0=None, nswp=10, tol=1e-8, stats=None):
    x = O.tt_random_gaussian([2] * (len(A) - 1), (A[0].shape[2],)) if x0 is None else [c.copy() for c in x0]
    d = len(x)
    st = _State([A, Delta], x, tol / np.sqrt(d))
    step = 1.0
    local_res = np.inf * np.ones((2, d - 1))
    prev_step, prev_res = None, np.inf

    def solve_two(k, bwd):
        nonlocal step
        if (not np.isfinite(step)) or step <= 0:
            step = 0.0
            return np.inf
        prev, shape = st.two(k)
        sol, step, res = _pencil(st.proj(0, k, True), st.proj(1, k, True), prev, step, tol)
        st.x[k], st.x[k + 1] = _split_two(sol, shape, st.trunc_tol, st.max_rank, bwd)
        return res

    def solve_one(k):
        nonlocal step
        prev = st.x[k].reshape(-1)
        if (not np.isfinite(step)) or step <= 0:
            step = 0.0
            return prev
        sol, step, _ = _pencil(st.proj(0, k, False), st.proj(1, k, False), prev, step, tol)
        return sol

    swp = 0
    for swp in range(nswp):
        zero = False
        for k in range(d - 1, 0, -1):
            if swp > 0:
                local_res[0, k - 1] = solve_two(k - 1, True)
                if step <= 0:
                    zero = True
                    break
            else:
                st.orth_bck(k)
            st.bck(k)
        if zero:
            break
        if np.max(local_res) < tol or swp == nswp - 1:
            for k in range(d):
                sol = solve_one(k)
                if k < d - 1:
                    st.orth_fwd(k, sol)
                    st.fwd(k)
                else:
                    st.x[k] = sol.reshape(st.x[k].shape)
            break
        for k in range(d - 1):
            local_res[1, k] = solve_two(k, False)
            if step <= 0:
                zero = True
                break
            st.fwd(k)
        if zero:
            break
        if np.max(local_res) < tol:
            for k in range(d - 1, -1, -1):
                sol = solve_one(k)
                if k > 0:
                    st.orth_bck(k, sol)
                    st.bck(k)
                else:
                    st.x[k] = sol.reshape(st.x[k].shape)
            break
        sres = np.max(local_res)
        if swp >= 2 and prev_step is not None:
            scale = max(abs(step), abs(prev_step), 1.0)
            if abs(step - prev_step) <= max(10 * tol, 1e-12) * scale and _stalled(prev_res, sres, tol):
                break
        prev_step, prev_res = step, sres
    max_res = np.max(local_res)
    out = O.tt_normalise(st.x)
    if max_res > tol:
        step *= tol / max_res
    if stats is not None:
        stats.update(sweeps=swp + 1, max_res=float(max_res))
    return step, out


def tt_min_eig(A, x0=None, nswp=10, tol=1e-8, return_eig_val=False, stats=None):
    x = O.tt_random_gaussian([2] * (len(A) - 1), (A[0].shape[2],)) if x0 is None else [c.copy() for c in x0]
    d = len(x)
    st = _State([A], x, 0.1 * tol / np.sqrt(d))

    def solve_two(k, bwd):
        prev, shape = st.two(k)
        M = st.proj(0, k, True)
        lam, sol = _smallest(M)
        st.x[k], st.x[k + 1] = _split_two(sol, shape, st.trunc_tol, st.max_rank, bwd)
        return float(np.linalg.norm(lam * prev - M @ prev))

    def solve_one(k):
        return _smallest(st.proj(0, k, False))[1]

    max_res, prev_res = 0.0, np.inf
    swp = 0
    for swp in range(nswp):
        max_res = np.inf if swp == 0 else 0.0
        for k in range(d - 1, 0, -1):
            if swp > 0:
                max_res = max(max_res, solve_two(k - 1, True))
            else:
                st.orth_bck(k)
            st.bck(k)
        if max_res < tol or swp == nswp - 1:
            for k in range(d):
                sol = solve_one(k)
                if k < d - 1:
                    st.orth_fwd(k, sol)
                    st.fwd(k)
                else:
                    st.x[k] = sol.reshape(st.x[k].shape)
            break
        max_res = 0.0
        for k in range(d - 1):
            max_res = max(max_res, solve_two(k, False))
            st.fwd(k)
        if max_res < tol:
            for k in range(d - 1, -1, -1):
                sol = solve_one(k)
                if k > 0:
                    st.orth_bck(k, sol)
                    st.bck(k)
                else:
                    st.x[k] = sol.reshape(st.x[k].shape)
            break
        if swp >= 2 and _stalled(prev_res, max_res, tol):
            break
        prev_res = max_res
    out = O.tt_normalise(st.x)
    val = None
    if return_eig_val:
        val = O.tt_inner_prod(out, O.tt_fast_matrix_vec_mul(A, [c.copy() for c in out], 1e-12))
    if stats is not None:
        stats.update(sweeps=swp + 1, max_res=float(max_res))
    return out, val
