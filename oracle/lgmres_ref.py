"""ORACLE (test infrastructure, not product code): PETSc-style LGMRES in NumPy.

PARITY UNPINNED: the reference solves its local KKT blocks with PETSc KSP
`lgmres` through petsc4py (reference src/tt_ipm.py:101-162, called at :251-254
and :364-367).  PETSc 3.25.1 / petsc4py 3.25.1 (reference env.yaml:13-14) are
not vendored under /root/reference and are not installed here, so this file
restates the published algorithm (Baker, Jessup, Manteuffel, "A technique for
accelerating the convergence of restarted GMRES", SIMAX 2005; PETSc's
KSPLGMRES implementation of it) from its documentation: zero initial guess, no
preconditioner, classical Gram-Schmidt without refinement, `restart` = total
size of the approximation space, `augment` error-approximation vectors, default
convergence test ||r|| <= max(rtol*||b||, abstol), `max_it` total inner steps.
The only reference test that pins anything at this boundary
(tests/test_tt_preprocessing.py:25-36, a 2x2 SPD system solved to 1e-10) is
reproduced in tests/test_oracle_lgmres.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.
"""
import numpy as np


class LgmresResult:
    __slots__ = ("x", "its", "matvecs", "reason", "rnorm", "history")

    def __init__(self, x, its, matvecs, reason, rnorm, history):
        self.x = x
        self.its = its
        self.matvecs = matvecs
        self.reason = reason
        self.rnorm = rnorm
        self.history = history


def lgmres(matvec, b, rtol=1e-5, max_it=300, restart=100, augment=10,
           abstol=1e-50, dtol=1e5, haptol=1e-30):
    """Solve A x = b from x0 = 0.

    restart  : size of the approximation space per cycle (PETSc -ksp_gmres_restart)
    augment  : number of error-approximation vectors kept (-ksp_lgmres_augment)
    Each cycle runs (restart - augment) Arnoldi steps followed by one step per
    stored augmentation vector; a step of either kind counts as one iteration.
    reason: 'rtol', 'atol', 'its', 'dtol', 'breakdown', 'null'.
    """
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    n = b.size
    x = np.zeros(n)
    max_k = int(restart)
    aug_dim = int(augment)
    it_arnoldi_nominal = max_k - aug_dim
    aug_vecs = np.zeros((max(aug_dim, 1), n))
    a_aug_vecs = np.zeros((max(aug_dim, 1), n))
    aug_order = [0] * max(aug_dim, 1)
    aug_ct = 0
    its = 0
    matvecs = 0
    reason = None
    ttol = None
    rnorm0 = None
    history = []
    first_cycle = True
    rnorm = 0.0

    def converged(it, rn):
        nonlocal ttol, rnorm0
        if it == 0:
            ttol = max(rtol * rn, abstol)
            rnorm0 = rn
        if not np.isfinite(rn):
            return "nan"
        if rn <= ttol:
            return "atol" if rn < abstol else "rtol"
        if rn >= dtol * rnorm0:
            return "dtol"
        return None

    V = np.zeros((max_k + 2, n))
    hh = np.zeros((max_k + 2, max_k + 1))    # rotated Hessenberg (becomes upper triangular)
    hes = np.zeros((max_k + 2, max_k + 1))   # un-rotated Hessenberg
    grs = np.zeros(max_k + 2)
    cc = np.zeros(max_k + 1)
    ss = np.zeros(max_k + 1)

    while True:
        # initial residual of this cycle
        if first_cycle:
            r = b.copy()
        else:
            r = b - matvec(x)
            matvecs += 1
            if its >= max_it:
                reason = "its"
                break
        first_cycle = False

        it_arnoldi = it_arnoldi_nominal
        it_total = it_arnoldi + aug_ct
        hh[:] = 0.0
        hes[:] = 0.0
        grs[:] = 0.0
        res_norm = float(np.linalg.norm(r))
        res = res_norm
        grs[0] = res_norm
        rnorm = res
        if res == 0.0:
            reason = "atol"
            break
        V[0] = r / res_norm
        loc_it = 0
        hapend = False
        reason = converged(its, res)
        while reason is None and loc_it < it_total and its < max_it:
            history.append(res)
            if loc_it < it_arnoldi:
                w = np.array(matvec(V[loc_it]), dtype=np.float64, copy=True).reshape(-1)
                matvecs += 1
            else:
                order = loc_it - it_arnoldi + 1
                spot = 0
                for ii in range(aug_dim):
                    if aug_order[ii] == order:
                        spot = ii
                        break
                w = a_aug_vecs[spot].copy()
            # classical Gram-Schmidt, one pass
            h = V[: loc_it + 1] @ w
            w -= h @ V[: loc_it + 1]
            hh[: loc_it + 1, loc_it] = h
            hes[: loc_it + 1, loc_it] = h
            tt = float(np.linalg.norm(w))
            hh[loc_it + 1, loc_it] = tt
            hes[loc_it + 1, loc_it] = tt
            hapbnd = abs(tt / grs[loc_it]) if grs[loc_it] != 0.0 else np.inf
            if hapbnd > haptol:
                hapbnd = haptol
            if tt > hapbnd:
                V[loc_it + 1] = w / tt
            else:
                V[loc_it + 1] = w
                hapend = True
            # apply previous rotations to the new column, then a new rotation
            col = hh[:, loc_it]
            for j in range(loc_it):
                t = col[j]
                col[j] = cc[j] * t + ss[j] * col[j + 1]
                col[j + 1] = cc[j] * col[j + 1] - ss[j] * t
            if not hapend:
                t = np.sqrt(col[loc_it] * col[loc_it] + col[loc_it + 1] * col[loc_it + 1])
                if t == 0.0:
                    reason = "null"
                    break
                cc[loc_it] = col[loc_it] / t
                ss[loc_it] = col[loc_it + 1] / t
                grs[loc_it + 1] = -(ss[loc_it] * grs[loc_it])
                grs[loc_it] = cc[loc_it] * grs[loc_it]
                col[loc_it] = cc[loc_it] * col[loc_it] + ss[loc_it] * col[loc_it + 1]
                res = abs(grs[loc_it + 1])
            else:
                res = 0.0
            loc_it += 1
            its += 1
            rnorm = res
            reason = converged(its, res)
            if hapend and reason is None:
                reason = "breakdown"
                break
        history.append(res)

        # ---- form the solution of this cycle -------------------------------
        it = loc_it - 1
        update = np.zeros(n)
        nrs = np.zeros(max(loc_it, 1))
        if it >= 0:
            if it_arnoldi >= it + 1:
                n_arn, n_aug = it + 1, 0
            else:
                n_arn, n_aug = it_arnoldi, (it + 1) - it_arnoldi
            nrs[it] = grs[it] / hh[it, it] if hh[it, it] != 0.0 else 0.0
            for k in range(it - 1, -1, -1):
                t = grs[k]
                for j in range(k + 1, it + 1):
                    t -= hh[k, j] * nrs[j]
                nrs[k] = t / hh[k, k]
            update = nrs[:n_arn] @ V[:n_arn]
            for ii in range(n_aug):
                spot = 0
                for jj in range(aug_dim):
                    if aug_order[jj] == ii + 1:
                        spot = jj
                        break
                update = update + nrs[n_arn + ii] * aug_vecs[spot]
            x = x + update

        # ---- harvest the augmentation vector for the next cycle ---------------
        if reason is None and its < max_it and aug_dim > 0:
            if aug_ct == 0:
                spot = 0
                aug_ct += 1
            elif aug_ct < aug_dim:
                spot = aug_ct
                aug_ct += 1
            else:
                spot = 0
                for ii in range(aug_dim):
                    if aug_order[ii] == aug_dim:
                        spot = ii
            tmp_norm = float(np.linalg.norm(update))
            inv = 1.0 / tmp_norm
            aug_vecs[spot] = update * inv
            for ii in range(aug_dim):
                aug_order[ii] += 1
            aug_order[spot] = 1
            # A * aug = V_{it_total+1} * (Hes * y)
            avec = np.zeros(it_total + 1)
            for ii in range(it_total):
                hi = min(ii + 2, it_total + 1)
                avec[:hi] += hes[:hi, ii] * nrs[ii]
            a_aug_vecs[spot] = (avec @ V[: it_total + 1]) * inv

        if reason is not None:
            break
        if its >= max_it:
            # PETSc recomputes the residual once more before flagging DIVERGED_ITS;
            # the solution is unchanged by that, so stop here.
            reason = "its"
            break

    return LgmresResult(x, its, matvecs, reason, rnorm, history)
