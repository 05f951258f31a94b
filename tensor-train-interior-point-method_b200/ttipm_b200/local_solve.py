"""Device local KKT-block solvers of the AMEn sweep.

Drop-in for the `local_solver` callbacks the reference passes into tt_restarted_block_amen:
`_ipm_local_solver` (3 blocks dY, dX, dZ; reference src/tt_ipm.py:183-282) and
`_ipm_local_solver_ineq` (4 blocks, + dT; :284-401).  The sweep hands over device tensors; the
matrix-free branch runs entirely on the GPU (persistent LGMRES kernel), the dense Schur branch
assembles the m x m blocks with the K4 kernel and factorises them with cuSOLVER/cuBLAS through
torch.linalg (plain library factorisations).  The host only sees the scalars that the
reference's control flow branches on.
"""
import math

import numpy as np
import torch

from . import kernels as K


class LocalSystem:
    """Operands of one local block system: interfaces, operator cores and rhs pieces on the device."""

    def __init__(self, rt, P1, A, P2, aliases, transposes, Xb1, B, Xb2, nblk):
        self.rt = rt
        self.P1, self.A, self.P2 = P1, A, P2
        self.aliases, self.transposes = aliases, transposes
        self.Xb1, self.B, self.Xb2 = Xb1, B, Xb2
        self.nblk = nblk
        self._full = None

    def full_terms(self):
        """Term list of TTBlockMatrixView.block_local_product (reference src/tt_als.py:190-200)."""
        if self._full is None:
            tl = K.TermList()
            for (i, j), A in self.A.items():
                tl.add(self.P1[i, j], A, self.P2[i, j], j, i)
                if (i, j) in self.transposes:
                    p, t = self.transposes[i, j]
                    tl.add(self.P1[i, j].permute(2, 1, 0), A.permute(0, 2, 1, 3), self.P2[i, j].permute(2, 1, 0), t, p)
                if (i, j) in self.aliases:
                    p, t = self.aliases[i, j]
                    tl.add(self.P1[i, j], A, self.P2[i, j], t, p)
            self._full = tl
        return self._full

    def project_rhs(self, shape):
        rt = self.rt
        rhs = rt.zeros(*shape)
        rows = sorted(self.B.keys())
        if rows:
            K.rhs_project([self.Xb1[i] for i in rows], [self.B[i] for i in rows], [self.Xb2[i] for i in rows], rhs,
                          rows, rt=rt)
        return rhs

    def residual(self, x, rhs, store=False):
        """(A_loc x - rhs, device partials of its squared norm)."""
        r, _, n, R = x.shape
        return K.block_matvec(self.full_terms(), x, self.nblk, (r, R), sub=rhs, want_norm=True, rt=self.rt)

    def apply_T01(self, y):
        """'lsr,smnS,LSR,lmL->rnR' with block (0,1): K01^T y (reference src/tt_ipm.py:217, :274)."""
        key = (0, 1)
        tl = K.TermList().add(self.P1[key].permute(2, 1, 0), self.A[key].permute(0, 2, 1, 3),
                              self.P2[key].permute(2, 1, 0), 0, 0)
        r, n, R = y.shape
        return K.block_matvec(tl, y.reshape(r, 1, n, R), 1, (r, R), rt=self.rt).reshape(r, n, R)

    def apply(self, key, x):
        tl = K.TermList().add(self.P1[key], self.A[key], self.P2[key], 0, 0)
        r, n, R = x.shape
        return K.block_matvec(tl, x.reshape(r, 1, n, R), 1, (r, R), rt=self.rt).reshape(r, n, R)

    def dense(self, key):
        return K.local_dense(self.P1[key], self.A[key], self.P2[key], rt=self.rt)


def _host_sums(rt, *tensors):
    """One device->host transfer for several partial-sum buffers; returns the totals."""
    flat = torch.cat([t.reshape(-1) for t in tensors])
    h = rt.to_host(flat)
    out, o = [], 0
    for t in tensors:
        n = t.numel()
        out.append(float(h[o:o + n].sum()))
        o += n
    return out


def _fb_sub(Lc, b):
    y = torch.linalg.solve_triangular(Lc, b, upper=False)
    return torch.linalg.solve_triangular(Lc.transpose(0, 1), y, upper=True)


def _dense_eq(sys_, rhs, inv_I, shape):
    """Dense Schur-complement solve, reference src/tt_ipm.py:196-223."""
    r, b, n, R = shape
    m = r * n * R
    Rp, Rd, Rc = (rhs[:, i].reshape(m, 1) for i in range(3))
    inv = inv_I.reshape(1, m)
    LXI = sys_.dense((2, 2)) * inv
    Leq = sys_.dense((0, 1))
    Lc, info = torch.linalg.cholesky_ex(sys_.dense((2, 1)))
    if int(info) != 0:
        raise np.linalg.LinAlgError("local L_Z block is not positive definite")
    bb = Rp - Leq @ _fb_sub(Lc, Rc - LXI @ Rd)
    S = Leq @ (_fb_sub(Lc, LXI) @ Leq.transpose(0, 1))
    S += sys_.dense((0, 0))
    S.diagonal().add_(1e-11)
    sol = sys_.rt.empty(r, b, n, R)
    y = torch.linalg.solve(S, bb)
    sol[:, 0] = y.reshape(r, n, R)
    z = (Rd - sys_.apply_T01(sol[:, 0].contiguous()).reshape(m, 1)) * inv_I.reshape(m, 1)
    sol[:, 2] = z.reshape(r, n, R)
    x = _fb_sub(Lc, Rc - sys_.apply((2, 2), sol[:, 2].contiguous()).reshape(m, 1))
    sol[:, 1] = x.reshape(r, n, R)
    return sol


def _dense_ineq(sys_, rhs, inv_I, shape):
    """Nested dense Schur solve of the 4-block system, reference src/tt_ipm.py:298-334."""
    r, b, n, R = shape
    m = r * n * R
    Lc, info = torch.linalg.cholesky_ex(sys_.dense((2, 1)))
    if int(info) != 0:
        raise np.linalg.LinAlgError("local L_Z block is not positive definite")
    Rp, Rd, Rc, Rt = (rhs[:, i].reshape(m, 1) for i in range(4))
    inv = inv_I.reshape(1, m)
    LZc = _fb_sub(Lc, Rc)
    LZX = _fb_sub(Lc, sys_.dense((2, 2)))
    LZXI = LZX * inv
    Leq = sys_.dense((0, 1))
    Top = sys_.dense((3, 1))
    w = LZc - LZXI @ Rd
    u = Rp - Leq @ w
    v = Rt - Top @ w
    Am = sys_.dense((0, 0)) + Leq @ LZXI @ Leq.transpose(0, 1)
    D = sys_.dense((3, 3)) + Top @ LZX
    D.diagonal().add_(1e-11)
    TopS = (Top @ LZXI) @ Leq.transpose(0, 1)
    LeqS = Leq @ LZX
    LU, piv = torch.linalg.lu_factor(D)
    rhs_l = u - LeqS @ torch.linalg.lu_solve(LU, piv, v)
    lhs_l = Am - LeqS @ torch.linalg.lu_solve(LU, piv, TopS)
    y = torch.linalg.solve(lhs_l, rhs_l)
    sol = sys_.rt.empty(r, b, n, R)
    sol[:, 0] = y.reshape(r, n, R)
    sol[:, 3] = torch.linalg.lu_solve(LU, piv, v - TopS @ y).reshape(r, n, R)
    z = (Rd - sys_.apply_T01(sol[:, 0].contiguous()).reshape(m, 1)) * inv_I.reshape(m, 1)
    sol[:, 2] = z.reshape(r, n, R) - sol[:, 3]
    x = _fb_sub(Lc, Rc - sys_.apply((2, 2), sol[:, 2].contiguous()).reshape(m, 1))
    sol[:, 1] = x.reshape(r, n, R)
    return sol


def solve_local(sys_, prev, size_limit, dense_solve, ineq, rtol=1e-5, stats=None):
    """Device version of _ipm_local_solver / _ipm_local_solver_ineq.

    prev: (r, b, n, R) device tensor.  Returns (solution, res_old, min(res_old, res_new), rhs,
    norm_rhs, direct_solve_failure) with the three scalars on the host, like the reference."""
    rt = sys_.rt
    r, b, n, R = prev.shape
    m = r * n * R
    rhs = sys_.project_rhs((r, b, n, R))
    _, rhs_ss = K.ewise(rhs, want_sumsq=True, store=False, rt=rt)
    inv_I = K.local_diag(sys_.P1[1, 2], sys_.A[1, 2], sys_.P2[1, 2], invert=True, rt=rt)
    _, res_ss = sys_.residual(prev, rhs)
    rhs2, res2 = _host_sums(rt, rhs_ss, res_ss)
    norm_rhs = max(math.sqrt(rhs2), 1e-10)
    res_old = math.sqrt(res2) / norm_rhs
    limit = 0.95 * size_limit if ineq else size_limit
    dense = (math.sqrt(r * R) <= limit) and dense_solve and (res_old >= rtol)
    direct_fail = not dense
    sol = None
    if dense:
        try:
            sol = (_dense_ineq if ineq else _dense_eq)(sys_, rhs, inv_I, (r, b, n, R))
        except Exception:
            direct_fail = True
    if not dense or direct_fail:
        nred = 3 if ineq else 2
        src = [0, 1, 3] if ineq else [0, 1]
        op = K.ReducedOperator(sys_.P1, sys_.A, sys_.P2, inv_I, ineq, rt=rt)
        lrhs = rt.empty(nred, r, n, R)
        K.ewise(rhs[:, 0], out=lrhs[0], rt=rt)
        t = K.ewise(rhs[:, 1], w=inv_I, rt=rt)                               # inv_I o r_d
        K.ewise(rhs[:, 2], 1.0, b=sys_.apply((2, 2), t), beta=-1.0, out=lrhs[1], rt=rt)
        if ineq:
            K.ewise(rhs[:, 3], out=lrhs[2], rt=rt)
        prev_red = rt.empty(nred, r, n, R)
        for q, j in enumerate(src):
            K.ewise(prev[:, j], out=prev_red[q], rt=rt)
        lvec = op.matvec(prev_red)
        _, n0 = K.ewise(lrhs, want_sumsq=True, store=False, rt=rt)
        diff, n1 = K.ewise(lrhs, 1.0, b=lvec, beta=-1.0, want_sumsq=True, rt=rt)
        n0, n1 = _host_sums(rt, n0, n1)
        use_prev = math.sqrt(n1) < math.sqrt(n0)
        if use_prev:
            lrhs = diff
        restart = min(m, 100)
        aug = max(restart // 10, 3)
        prof = stats.get("profile") if stats is not None else None
        if prof is not None and rt.is_cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        xs, info = op.solve(lrhs, restart, aug, max_it=300, rtol=rtol)
        if prof is not None and rt.is_cuda:
            ev1.record()
            keys = K.ReducedOperator.KEYS_INEQ if ineq else K.ReducedOperator.KEYS_EQ
            fl = 0.0
            for key in keys:
                s_, S_ = sys_.A[key].shape[0], sys_.A[key].shape[3]
                fl += (2.0 * r * n * R * R * S_ + 2.0 * r * R * s_ * n * n * S_ + 2.0 * r * n * R * r * s_) * \
                      (2 if key == (0, 1) else 1)
            prof.append((ev0, ev1, info, dict(nv=nred * m, mv_flops=fl, restart=restart, r=r, R=R)))
        if stats is not None:
            stats["lgmres"].append(info)
        sol = rt.empty(r, b, n, R)
        for q, j in enumerate(src):
            if use_prev:
                K.ewise(xs[q], 1.0, b=prev[:, j], beta=1.0, out=sol[:, j], rt=rt)
            else:
                K.ewise(xs[q], out=sol[:, j], rt=rt)
        kty = sys_.apply_T01(sol[:, 0].contiguous())
        # z = inv_I o (r_d - K01^T y) [- t]
        if ineq:
            K.ewise(rhs[:, 1], 1.0, b=kty, beta=-1.0, w=inv_I, c=sol[:, 3], gamma=-1.0, out=sol[:, 2], rt=rt)
        else:
            K.ewise(rhs[:, 1], 1.0, b=kty, beta=-1.0, w=inv_I, out=sol[:, 2], rt=rt)
    _, new_ss = sys_.residual(sol, rhs)
    (new2,) = _host_sums(rt, new_ss)
    res_new = math.sqrt(new2) / norm_rhs
    if res_old < res_new:
        sol = prev
    return sol, res_old, min(res_old, res_new), rhs, norm_rhs, direct_fail
