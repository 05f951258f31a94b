"""Step-size eigen sweeps on the device: `tt_max_generalised_eigen`, `tt_min_eig`
(reference src/tt_als.py:931-1499, called at src/tt_ipm.py:714-715, :735, :791).

Both are two-site DMRG-type sweeps over a TT vector x with local eigenvalue problems on the projections of one or
two TT matrices.  Everything numeric runs on the GPU and stays there for the whole call:

  * cores of A / Delta and of x, the interfaces XAX / XDX (updated one bond per step by `k_phi_update`),
  * the local projections (`k_eig_assemble`: dense symmetric one- / two-site operator),
  * the local eigen solves (`k_eig_lanczos`, one persistent cluster launch per solve; the pencil (-D, A) through a
    Cholesky reduction + the same kernel) -- these replace scipy's ARPACK `eigsh` and `lobpcg` calls,
  * truncation SVDs and the enrichment QR (`k_linalg`), bond products (`k_gemm`).

The host keeps what the reference's control flow branches on: eigenvalue signs, step sizes, local residuals, the
rank rule (`prune_singular_vals`) and the NumPy RNG draws of the enrichment (`_add_kick_rank[_rev]`, drawn in the
reference's order so a seeded run consumes the global RNG identically while ranks agree).

Differences from the reference that are deliberate: local problems larger than its dense limits (r R > size_limit)
are still assembled densely and solved by the same Lanczos kernel (the reference switches to matrix-free LOBPCG with
a 100-iteration cap and silently keeps the previous iterate when that does not converge); one-site projections are
symmetrised like the two-site ones.  Both only change results at the rounding level of a converged solve.
"""
import numpy as np

from . import kernels as K
from . import tt as T
from .runtime import TTIPMError, get_runtime

_LANCZOS_BASIS = 64
_LANCZOS_CYCLES = 40
_DENSE_LIMIT = 16384          # largest local problem assembled densely (2 x m^2 doubles: 4.3 GB at the limit)


def _eigen_residual_stalled(prev_res, res, tol):
    """reference src/tt_als.py:905-911."""
    return bool(np.isfinite(prev_res) and np.isfinite(res) and res <= 50 * tol and res >= 0.8 * prev_res)


def _eigen_step_stalled(prev_step, step, prev_res, res, tol):
    """reference src/tt_als.py:914-921."""
    if prev_step is None:
        return False
    scale = max(abs(step), abs(prev_step), 1.0)
    return bool(abs(step - prev_step) <= max(10 * tol, 1e-12) * scale and _eigen_residual_stalled(prev_res, res, tol))


class _Sweeper:
    """Device state of one eigen sweep: operator cores, x cores, interfaces."""

    def __init__(self, mats, x_cores, tol, trunc_tol, rt):
        self.rt = rt
        self.mats = [[rt.to_device(c) for c in M] for M in mats]          # [A] or [A, Delta]
        self.x = [rt.to_device(c) for c in x_cores]
        self.d = len(self.x)
        self.N = [int(c.shape[1]) for c in self.x]
        self.rx = [1] + [int(c.shape[2]) for c in self.x[:-1]] + [1]
        one = rt.to_device(np.ones((1, 1, 1)))
        self.phi = [[one] + [None] * (self.d - 1) + [one] for _ in self.mats]
        self.tol = tol
        self.trunc_tol = trunc_tol
        self.max_rank = int(np.floor(2 ** (self.d / 2)))
        self.ltol = 0.1 * tol
        self.solves = 0
        self.matvecs = 0

    # ---- small pieces ------------------------------------------------------------------------------------------
    def _rank(self, s_dev):
        s = self.rt.to_host(s_dev)
        return int(min(T.prune_singular_vals(s, self.trunc_tol), self.max_rank))

    def update_bck(self, k):
        """XAX[k] = compute_phi_bck_A(XAX[k+1], x[k], A[k], x[k]) for every operator (reference :1176-1177)."""
        phis = [ph[k + 1] for ph in self.phi]
        outs = K.phi_update(phis, [M[k] for M in self.mats], self.x[k], self.x[k], forward=False, rt=self.rt)
        for ph, o in zip(self.phi, outs):
            ph[k] = o
        self.rx[k] = int(self.x[k].shape[0])

    def update_fwd(self, k):
        phis = [ph[k] for ph in self.phi]
        outs = K.phi_update(phis, [M[k] for M in self.mats], self.x[k], self.x[k], forward=True, rt=self.rt)
        for ph, o in zip(self.phi, outs):
            ph[k + 1] = o
        self.rx[k + 1] = int(self.x[k].shape[2])

    def orth_step_bck(self, k):
        """First backward pass: x[k] <- right-orthogonal factor, x[k-1] absorbs the rest (reference :1168-1174)."""
        r0, n, r1 = self.x[k].shape
        U, S, W = K.svd_left(self.x[k].reshape(r0, n * r1).t(), rt=self.rt)      # (n r1) x r0
        r = self._rank(S)
        self.x[k] = U[:, :r].t().contiguous().reshape(r, n, r1)
        left = self.x[k - 1]
        a, b, _ = left.shape
        self.x[k - 1] = K.gemm(left.reshape(a * b, r0), W[:r].t(), rt=self.rt).reshape(a, b, r)

    def split_fwd(self, k, sol):
        """x[k] <- left-orthogonal factor of the (r n) x R solution, x[k+1] absorbs S V^T (reference :1190-1198)."""
        r0, n, r1 = self.x[k].shape
        U, S, W = K.svd_left(sol.reshape(r0 * n, r1), rt=self.rt)
        r = self._rank(S)
        self.x[k] = U[:, :r].contiguous().reshape(r0, n, r)
        nxt = self.x[k + 1]
        _, n2, r2 = nxt.shape
        self.x[k + 1] = K.gemm(W[:r].contiguous(), nxt.reshape(r1, n2 * r2), rt=self.rt).reshape(r, n2, r2)

    def split_bck(self, k, sol):
        """reference :1243-1253."""
        r0, n, r1 = self.x[k].shape
        U, S, W = K.svd_left(sol.reshape(r0, n * r1).t(), rt=self.rt)
        r = self._rank(S)
        self.x[k] = U[:, :r].t().contiguous().reshape(r, n, r1)
        left = self.x[k - 1]
        a, b, _ = left.shape
        self.x[k - 1] = K.gemm(left.reshape(a * b, r0), W[:r].t(), rt=self.rt).reshape(a, b, r)

    def two_site(self, k):
        """previous_solution = 'rny,ytR->rntR' of x[k], x[k+1] as a flat device vector, and its shape."""
        a, b = self.x[k], self.x[k + 1]
        r0, n, y = a.shape
        _, t, R = b.shape
        return K.gemm(a.reshape(r0 * n, y), b.reshape(y, t * R), rt=self.rt).reshape(-1), (r0, n, t, R)

    def project(self, which, k, two):
        """Dense symmetric projection of operator `which` on sites k[, k+1]."""
        M = self.mats[which]
        ph = self.phi[which]
        right = ph[k + 2] if two else ph[k + 1]
        shape_m = int(ph[k].shape[0]) * self.N[k] * (self.N[k + 1] if two else 1) * int(right.shape[0])
        if shape_m > _DENSE_LIMIT:
            raise TTIPMError(f"eigen sweep: local problem of size {shape_m} exceeds the dense limit {_DENSE_LIMIT}")
        if two:
            return K.eig_assemble(ph[k], M[k], M[k + 1], ph[k + 2], rt=self.rt)
        return K.eig_assemble(ph[k], M[k], None, ph[k + 1], rt=self.rt)

    def lanczos(self, A, cA, D, cD, v0, largest=False, cycles=_LANCZOS_CYCLES):
        x, out = K.eig_lanczos(A, cA, D, cD, v0=v0, largest=largest, K=_LANCZOS_BASIS, max_cycles=cycles, tol=self.ltol,
                               rt=self.rt)
        o = self.rt.to_host(out)
        self.solves += 1
        self.matvecs += int(o[2])
        return x, o

    def split_two_site(self, k, sol, shape, bwd):
        """Truncate the two-site solution and enrich the bond with 4 random directions
        (reference :1090-1105 with _add_kick_rank / _add_kick_rank_rev, :1108-1120)."""
        r0, n, t, R = shape
        rt = self.rt
        if bwd:
            U, S, W = K.svd_left(sol.reshape(r0 * n, t * R).t(), rt=rt)          # (t R) x (r0 n)
            r = self._rank(S)
            u_arg = W[:r].t()                                                    # (r0 n) x r
            v_arg = U[:, :r].t()                                                 # r x (t R)
            uk = rt.to_device(np.random.randn(4, t * R))
            stacked = rt.empty(r + 4, t * R)
            stacked[:r] = v_arg
            stacked[r:] = uk
            Q, Rm = K.qr(stacked.t(), rt=rt)                                     # stacked = Rm^T Q^T  (LQ form)
            kk = Q.shape[1]
            new2 = Q.t().contiguous()                                            # kk x (t R), orthonormal rows
            new1 = K.gemm(u_arg, Rm[:, :r].t(), rt=rt)                           # (r0 n) x kk
        else:
            U, S, W = K.svd_left(sol.reshape(r0 * n, t * R), rt=rt)
            r = self._rank(S)
            uk = rt.to_device(np.random.randn(r0 * n, 4))
            stacked = rt.empty(r0 * n, r + 4)
            stacked[:, :r] = U[:, :r]
            stacked[:, r:] = uk
            Q, Rm = K.qr(stacked, rt=rt)
            kk = Q.shape[1]
            new1 = Q
            new2 = K.gemm(Rm[:, :r], W[:r], rt=rt)                               # kk x (t R)
        self.x[k] = new1.contiguous().reshape(r0, n, kk)
        self.x[k + 1] = new2.contiguous().reshape(kk, t, R)


# -------------------------------------------------------------------------------------------------------------------
# tt_max_generalised_eigen
# -------------------------------------------------------------------------------------------------------------------
def _pencil_solve(sw, A, D, prev, step, eps):
    """Local step-size problem on a dense projection pair (reference _step_size_local_solve :931-1089 and
    _step_size_local_solve_last :1123-1128): smallest eigenpair of A / step + D; if it is negative the step shrinks to
    1 / lambda_max(-D, A).  Returns (solution vector (device, unit norm), step, old_res)."""
    x, o = sw.lanczos(A, 1.0 / step, D, 1.0, prev)
    eig_val, rq_raw, res_raw = o[0], o[4], o[5]
    sol = x
    new_step = step
    if eig_val < 0:
        xg, og = K.eig_gen_largest(A, D, v0=x, K=_LANCZOS_BASIS, max_cycles=_LANCZOS_CYCLES, tol=sw.ltol, rt=sw.rt)
        og = sw.rt.to_host(og)
        sw.solves += 1
        sw.matvecs += int(og[2])
        if og[3] > 0 and np.isfinite(og[0]) and og[0] != 0:
            new_step = max(0.0, min(step, 1.0 / og[0]))
            sol = xg
        else:                       # the reference's exception branch (A not positive definite)
            sol = K.ewise(prev, alpha=1.0 / o[7], rt=sw.rt) if o[7] > 0 else prev
            new_step = step * (1 - eps)
    if new_step != step:
        if new_step > 0:
            _, o2 = sw.lanczos(A, 1.0 / new_step, D, 1.0, prev, cycles=0)
            res_raw = o2[5]
        else:
            res_raw = np.inf
    return sol, new_step, float(res_raw)


def _step_two_site(sw, k, step, eps, bwd):
    if (not np.isfinite(step)) or step <= 0:
        return 0.0, np.inf
    prev, shape = sw.two_site(k)
    A = sw.project(0, k, True)
    D = sw.project(1, k, True)
    sol, step, old_res = _pencil_solve(sw, A, D, prev, step, eps)
    sw.split_two_site(k, sol, shape, bwd)
    return step, old_res


def _step_one_site(sw, k, step, eps):
    prev = sw.x[k].reshape(-1)
    if (not np.isfinite(step)) or step <= 0:
        return prev, 0.0
    A = sw.project(0, k, False)
    D = sw.project(1, k, False)
    sol, step, _ = _pencil_solve(sw, A, D, prev, step, eps)
    return sol, step


def tt_max_generalised_eigen(A, Delta, x0=None, nswp=10, tol=1e-8, size_limit=256, verbose=False, _stats=None):
    """Largest step t such that A + t * Delta stays positive semidefinite, as the reference computes it
    (src/tt_als.py:1132-1283).  Returns (step_size, x_cores)."""
    rt = get_runtime()
    if x0 is None:
        x_cores = T.tt_random_gaussian([2] * (len(A) - 1), (A[0].shape[2],))
    else:
        x_cores = x0
    d = len(x_cores)
    sw = _Sweeper([A, Delta], x_cores, tol, tol / np.sqrt(d), rt)
    step = 1.0
    local_res = np.inf * np.ones((2, d - 1))
    prev_sweep_step, prev_sweep_res = None, np.inf
    swp = 0
    for swp in range(nswp):
        zero_step = False
        for k in range(d - 1, 0, -1):
            if swp > 0:
                step, res = _step_two_site(sw, k - 1, step, tol, bwd=True)
                local_res[0, k - 1] = res
                if step <= 0:
                    zero_step = True
                    break
            else:
                sw.orth_step_bck(k)
            sw.update_bck(k)
        if zero_step:
            break
        if np.max(local_res) < tol or swp == nswp - 1:
            for k in range(d):
                sol, step = _step_one_site(sw, k, step, tol)
                if k < d - 1:
                    sw.split_fwd(k, sol)
                    sw.update_fwd(k)
                else:
                    sw.x[k] = sol.reshape(sw.x[k].shape).contiguous()
            break
        for k in range(d - 1):
            step, res = _step_two_site(sw, k, step, tol, bwd=False)
            local_res[1, k] = res
            if step <= 0:
                zero_step = True
                break
            sw.update_fwd(k)
        if zero_step:
            break
        if np.max(local_res) < tol:
            for k in range(d - 1, -1, -1):
                sol, step = _step_one_site(sw, k, step, tol)
                if k > 0:
                    sw.split_bck(k, sol)
                    sw.update_bck(k)
                else:
                    sw.x[k] = sol.reshape(sw.x[k].shape).contiguous()
            break
        sweep_res = np.max(local_res)
        if swp >= 2 and _eigen_step_stalled(prev_sweep_step, step, prev_sweep_res, sweep_res, tol):
            break
        prev_sweep_step, prev_sweep_res = step, sweep_res

    max_res = np.max(local_res)
    out = T.tt_normalise([rt.to_host(c) for c in sw.x])
    if verbose:
        print(f"\t Solution rank is {sw.rx[1:-1]}\n\t Step size: {step:f}\n\t Residual {max_res}\n\t Number of sweeps {swp + 1}",
              flush=True)
    if max_res > tol:
        print('\t Target Residual not reached!', flush=True)
        step *= (tol / max_res)
    if _stats is not None:
        _stats.update(sweeps=swp + 1, local_solves=sw.solves, matvecs=sw.matvecs, max_res=float(max_res))
    return step, out


# -------------------------------------------------------------------------------------------------------------------
# tt_min_eig
# -------------------------------------------------------------------------------------------------------------------
def _min_two_site(sw, k, bwd):
    """reference _eigen_local_solve (src/tt_als.py:1286-1343); returns old_res = ||lambda p - A p||."""
    prev, shape = sw.two_site(k)
    A = sw.project(0, k, True)
    x, o = sw.lanczos(A, 1.0, None, 0.0, prev)
    sw.split_two_site(k, x, shape, bwd)
    return float(o[8])


def _min_one_site(sw, k):
    prev = sw.x[k].reshape(-1)
    A = sw.project(0, k, False)
    x, _ = sw.lanczos(A, 1.0, None, 0.0, prev)
    return x


def tt_min_eig(A, x0=None, nswp=10, tol=1e-8, size_limit=64, return_eig_val=False, verbose=False, _stats=None):
    """Eigenvector train of the smallest eigenvalue of the TT matrix A (reference src/tt_als.py:1392-1499).
    Returns (x_cores, eigenvalue or None)."""
    rt = get_runtime()
    if x0 is None:
        x_cores = T.tt_random_gaussian([2] * (len(A) - 1), (A[0].shape[2],))
    else:
        x_cores = x0
    d = len(x_cores)
    sw = _Sweeper([A], x_cores, tol, 0.1 * tol / np.sqrt(d), rt)
    max_res = 0.0
    prev_sweep_res = np.inf
    swp = 0
    for swp in range(nswp):
        max_res = np.inf if swp == 0 else 0.0
        for k in range(d - 1, 0, -1):
            if swp > 0:
                max_res = max(max_res, _min_two_site(sw, k - 1, bwd=True))
            else:
                sw.orth_step_bck(k)
            sw.update_bck(k)
        if max_res < tol or swp == nswp - 1:
            for k in range(d):
                sol = _min_one_site(sw, k)
                if k < d - 1:
                    sw.split_fwd(k, sol)
                    sw.update_fwd(k)
                else:
                    sw.x[k] = sol.reshape(sw.x[k].shape).contiguous()
            break
        max_res = 0.0
        for k in range(d - 1):
            max_res = max(max_res, _min_two_site(sw, k, bwd=False))
            sw.update_fwd(k)
        if max_res < tol:
            for k in range(d - 1, -1, -1):
                sol = _min_one_site(sw, k)
                if k > 0:
                    sw.split_bck(k, sol)
                    sw.update_bck(k)
                else:
                    sw.x[k] = sol.reshape(sw.x[k].shape).contiguous()
            break
        if swp >= 2 and _eigen_residual_stalled(prev_sweep_res, max_res, tol):
            break
        prev_sweep_res = max_res
    if verbose:
        print(f"\t Solution rank is {sw.rx[1:-1]}\n\t Residual {max_res}\n\t Number of sweeps {swp + 1}", flush=True)
    out = T.tt_normalise([rt.to_host(c) for c in sw.x])
    value = None
    if return_eig_val:
        value = T.tt_inner_prod(out, T.tt_fast_matrix_vec_mul(A, out, 1e-12))
    if _stats is not None:
        _stats.update(sweeps=swp + 1, local_solves=sw.solves, matvecs=sw.matvecs, max_res=float(max_res))
    return out, value
