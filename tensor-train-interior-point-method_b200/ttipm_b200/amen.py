"""Device-resident block AMEn/ALS sweep for the TT-format KKT system.

Mirror of tt_block_amen / _bck_sweep / _fwd_sweep (reference src/tt_als.py:277-670): the operator
cores, the solution and residual trains and all interface tensors (XAX, Xb, ZAX, Zb) are uploaded
once and stay in HBM for the whole solve; every contraction, factorisation and local solve is a
CUDA kernel of libttipm_b200.  The host keeps only what the reference's control flow branches on
(ranks, residual scalars, the dense/iterative switch) and the NumPy RNG draws, which are made in
the reference's order so that seeded runs follow the same trajectory.
"""
import math

import numpy as np
import torch

from . import kernels as K
from .local_solve import LocalSystem, solve_local, _host_sums
from .runtime import get_runtime


def prune_singular_vals(s, eps):
    """reference cy_src/tt_ops_cy.pyx:162-177 (host decision on the device-computed singular values)."""
    if np.linalg.norm(s) == 0.0:
        return 1
    sc = np.cumsum(np.abs(s[::-1]) ** 2)[::-1]
    R = max(int(np.argmax(sc < eps ** 2)), 1)
    if sc[-1] > eps ** 2:
        R = s.size
    return R


class _State:
    pass


def _tkeys(A_keys, transposes):
    return list(A_keys) + [v for v in transposes.values() if v not in A_keys]


class DeviceBlockAmen:
    def __init__(self, block_A, aliases, transposes, block_b, ineq, rt=None, stats=None):
        """block_A: {(i,j): [numpy cores (s,4,4,S)]}, block_b: {i: [numpy cores (b,4,B)]}."""
        self.rt = rt or get_runtime()
        rt = self.rt
        self.A = {key: [rt.to_device(c) for c in cores] for key, cores in block_A.items()}
        self.b = {i: [rt.to_device(c) for c in cores] for i, cores in block_b.items()}
        self.aliases = dict(aliases)
        self.transposes = dict(transposes)
        self.block_size = max(k[0] for k in block_A.keys()) + 1
        self.ineq = bool(ineq)
        self.stats = stats if stats is not None else {}
        self.stats.setdefault("lgmres", [])
        self.stats.setdefault("local_solves", 0)
        self.trace = []

    # ------------------------------------------------------------------------------------------
    def _local_system(self, st, k):
        A_k = {key: cores[k] for key, cores in self.A.items()}
        B_k = {i: cores[k] for i, cores in self.b.items()}
        return LocalSystem(self.rt, st.XAX[k], A_k, st.XAX[k + 1], self.aliases, self.transposes, st.Xb[k], B_k,
                           st.Xb[k + 1], self.block_size)

    def _mixed_terms(self, k, left, right, left_is_z, right_is_z):
        """Term lists of compressed_/lcompressed_/rcompressed_block_local_product
        (reference src/tt_als.py:202-238)."""
        tl = K.TermList()
        for (i, j), cores in self.A.items():
            A = cores[k]
            tl.add(left[i, j], A, right[i, j], j, i)
            if (i, j) in self.transposes:
                p, t = self.transposes[i, j]
                Pl = left[p, t] if left_is_z else left[i, j].permute(2, 1, 0)
                Pr = right[p, t] if right_is_z else right[i, j].permute(2, 1, 0)
                tl.add(Pl, A.permute(0, 2, 1, 3), Pr, t, p)
            if (i, j) in self.aliases:
                p, t = self.aliases[i, j]
                tl.add(left[i, j], A, right[i, j], t, p)
        return tl

    def _rhs_block(self, k, Xl, Xr, shape):
        rt = self.rt
        out = rt.zeros(*shape)
        rows = sorted(self.b.keys())
        K.rhs_project([Xl[i] for i in rows], [self.b[i][k] for i in rows], [Xr[i] for i in rows], out, rows, rt=rt)
        return out

    def _update_interfaces(self, st, k, bck, which):
        """XAX/Xb (which='x') or ZAX/Zb (which='z') after core k changed (src/tt_als.py:372-387, :499-514)."""
        rt = self.rt
        xk = st.x[k]
        src, dst = (k + 1, k) if bck else (k, k + 1)
        rows = sorted(self.b.keys())
        if which == "x":
            keys = list(self.A.keys())
            outs = K.phi_update([st.XAX[src][key] for key in keys], [self.A[key][k] for key in keys], xk, xk,
                                not bck, rt=rt)
            st.XAX[dst] = dict(zip(keys, outs))
            outs = K.phi_rhs_update([st.Xb[src][i] for i in rows], [self.b[i][k] for i in rows], xk, not bck, rt=rt)
            st.Xb[dst] = dict(zip(rows, outs))
        else:
            zk = st.z[k]
            keys = list(self.A.keys())
            phis = [st.ZAX[src][key] for key in keys]
            cores = [self.A[key][k] for key in keys]
            tkeys = []
            for (i, j), (p, t) in self.transposes.items():
                tkeys.append((p, t))
                phis.append(st.ZAX[src][p, t])
                cores.append(self.A[i, j][k].permute(0, 2, 1, 3))
            outs = K.phi_update(phis, cores, zk, xk, not bck, rt=rt)
            st.ZAX[dst] = dict(zip(keys + tkeys, outs))
            outs = K.phi_rhs_update([st.Zb[src][i] for i in rows], [self.b[i][k] for i in rows], zk, not bck, rt=rt)
            st.Zb[dst] = dict(zip(rows, outs))

    # ------------------------------------------------------------------------------------------
    def _sweep(self, st, direction, swp, last):
        rt = self.rt
        d, bs, N = st.d, self.block_size, st.N
        x, z, rx, rz = st.x, st.z, st.rx, st.rz
        bck = direction > 0
        local_res = math.inf if swp == 0 else 0.0
        local_dx = math.inf if swp == 0 else 0.0
        solving = swp > 0 and not last
        for k in (range(d - 1, -1, -1) if bck else range(d)):
            inner = (k > 0) if bck else (k < d - 1)
            n = N[k]
            resz = None
            if solving:
                prev = x[k]
                sys_ = self._local_system(st, k)
                sol, r_old, r_new, rhs, norm_rhs, st.direct_solve_failure = solve_local(
                    sys_, prev, 3 * d, not st.direct_solve_failure, self.ineq, stats=self.stats)
                self.stats["local_solves"] += 1
                self.trace.append((swp, k, r_old, r_new, prev.shape[0] * prev.shape[3]))
                local_res = max(local_res, r_old)
                _, dnum = K.ewise(sol, 1.0, b=prev, beta=-1.0, want_sumsq=True, store=False, rt=rt)
                _, dden = K.ewise(sol, want_sumsq=True, store=False, rt=rt)
                dnum, dden = _host_sums(rt, dnum, dden)
                local_dx = max(local_dx, math.sqrt(dnum) / math.sqrt(dden))
                if st.amen:
                    zshape = (int(rz[k]), bs, n, int(rz[k + 1]))
                    rhsz = self._rhs_block(k, st.Zb[k], st.Zb[k + 1], zshape)
                    resz = K.block_matvec(self._mixed_terms(k, st.ZAX[k], st.ZAX[k + 1], True, True), sol, bs,
                                          (zshape[0], zshape[3]), sub=rhsz, y_scale=-1.0, sub_scale=1.0, rt=rt)
            else:
                sol = x[k]
                if st.amen and not last:
                    resz = z[k]
            scales = K.block_norms(sol, rt=rt)
            r_k, R_k = int(rx[k]), int(rx[k + 1])
            if bck:
                S = K.permute4(sol, (0, 1, 2, 3), scale=scales, scale_axis=1, rt=rt)           # (r, b, n, R)
                mat = S.reshape(r_k * bs, n * R_k).t()                                           # (n R, r b) view
                rzm = resz.reshape(int(rz[k]) * bs, n * int(rz[k + 1])).t() if resz is not None else None
            else:
                S = K.permute4(sol, (0, 2, 1, 3), scale=scales, scale_axis=2, rt=rt)           # (r, n, b, R)
                mat = S.reshape(r_k * n, bs * R_k)
                rzm = (K.permute4(resz, (0, 2, 1, 3), rt=rt).reshape(int(rz[k]) * n, bs * int(rz[k + 1]))
                       if resz is not None else None)
            if not inner:
                if bck:
                    x[k] = K.permute4(S, (0, 1, 2, 3), scale=scales, scale_axis=1, divide=True, rt=rt)
                    if st.amen and not last:
                        z[k] = K.permute4(resz.contiguous(), (0, 1, 2, 3), scale=scales, scale_axis=1, divide=True, rt=rt)
                else:
                    x[k] = K.permute4(S, (0, 2, 1, 3), scale=scales, scale_axis=1, divide=True, rt=rt)
                    if st.amen and not last:
                        z[k] = K.permute4(resz.contiguous(), (0, 1, 2, 3), scale=scales, scale_axis=1, divide=True, rt=rt)
                continue

            U, Sv, W = K.svd_left(mat, rt=rt)           # U (M, K), s (K), W = s * Vt (K, Nc)
            s_host = rt.to_host(Sv)

            if solving:
                trunc_lim = max(2 * st.trunc_tol, r_new)
                r0 = min(prune_singular_vals(s_host, st.eps), st.r_max)
                if bck:
                    # core layout (r b, n R) = W[:r0]^T U[:, :r0]^T
                    sol_r0 = K.gemm(W[:r0].t(), U[:, :r0].t(), rt=rt).reshape(r_k, bs, n, R_k)
                    layout = "rbnR"
                else:
                    sol_r0 = K.gemm(U[:, :r0], W[:r0], rt=rt).reshape(r_k, n, bs, R_k)
                    layout = "rnbR"
                full = sys_.full_terms()
                res = K.block_matvec(full, sol_r0, bs, (r_k, R_k), sub=rhs, x_layout=layout, rt=rt)
                r = r0
                if r0 > 1:
                    # all rank-one candidates q = 1..r0-1 in one batched launch
                    if bck:
                        terms = K.gemm(W[1:r0].unsqueeze(2), U[:, 1:r0].t().unsqueeze(1), rt=rt)
                        terms = terms.reshape(r0 - 1, r_k, bs, n, R_k)
                    else:
                        terms = K.gemm(U[:, 1:r0].t().unsqueeze(2), W[1:r0].unsqueeze(1), rt=rt)
                        terms = terms.reshape(r0 - 1, r_k, n, bs, R_k)
                    Y = K.block_matvec(full, terms, bs, (r_k, R_k), x_layout=layout, rt=rt)
                    parts = rt.to_host(K.trunc_resnorms(res, Y, rt=rt)).sum(axis=1)       # index q-1
                    r = 1
                    for q in range(r0 - 1, 0, -1):
                        if math.sqrt(parts[q - 1]) / norm_rhs > trunc_lim:
                            r = q
                            break
                    else:
                        r = 1
                r += 1
                r = min(r, U.shape[1])
                uk, vk = U[:, :r], W[:r]
                if st.amen:
                    if bck:
                        eshape = (int(rz[k]), bs, n, R_k)
                        rhsxz = self._rhs_block(k, st.Zb[k], st.Xb[k + 1], eshape)
                        resxz = K.block_matvec(self._mixed_terms(k, st.ZAX[k], st.XAX[k + 1], True, False), sol_r0, bs,
                                               (eshape[0], eshape[3]), sub=rhsxz, y_scale=-1.0, sub_scale=1.0,
                                               x_layout=layout, rt=rt)
                        kr = min(st.kick_rank, eshape[0] * bs, n * R_k)
                        Uz, _, _ = K.svd_left(resxz.reshape(eshape[0] * bs, n * R_k).t(), rt=rt)
                    else:
                        eshape = (r_k, bs, n, int(rz[k + 1]))
                        sol_r = K.gemm(uk, vk, rt=rt).reshape(r_k, n, bs, R_k)
                        rhsxz = self._rhs_block(k, st.Xb[k], st.Zb[k + 1], eshape)
                        resxz = K.block_matvec(self._mixed_terms(k, st.XAX[k], st.ZAX[k + 1], False, True), sol_r, bs,
                                               (eshape[0], eshape[3]), sub=rhsxz, y_scale=-1.0, sub_scale=1.0,
                                               x_layout=layout, rt=rt)
                        kr = min(st.kick_rank, r_k * n, bs * eshape[3])
                        Uz, _, _ = K.svd_left(K.permute4(resxz, (0, 2, 1, 3), rt=rt).reshape(r_k * n, bs * eshape[3]),
                                              rt=rt)
                    cat = rt.empty(U.shape[0], r + kr)
                    cat[:, :r].copy_(uk)
                    cat[:, r:].copy_(Uz[:, :kr])
                    Q, Rf = K.qr(cat, rt=rt)
                    vk = K.gemm(Rf[:, :r], vk, rt=rt)
                    uk = Q
                    r = uk.shape[1]
            else:
                r = min(prune_singular_vals(s_host, st.eps), st.r_max)
                uk, vk = U[:, :r], W[:r]

            if bck:
                x[k] = uk.t().contiguous().reshape(r, n, R_k)
                vT = vk.t().contiguous()                                   # (r_k * bs, r) = (c, b, Rnew)
                a, dd, c = x[k - 1].shape
                G = K.gemm(x[k - 1].reshape(a * dd, c), vT.reshape(c, bs * r), rt=rt).reshape(a, dd, bs, r)
                x[k - 1] = K.permute4(G, (0, 2, 1, 3), scale=scales, scale_axis=1, divide=True, rt=rt)
                rx[k] = r
            else:
                x[k] = uk.contiguous().reshape(r_k, n, r)
                Rn, dd, k2 = x[k + 1].shape
                G = K.gemm(vk.contiguous().reshape(r * bs, Rn), x[k + 1].reshape(Rn, dd * k2), rt=rt)
                x[k + 1] = K.permute4(G.reshape(r, bs, dd, k2), (0, 1, 2, 3), scale=scales, scale_axis=1, divide=True,
                                      rt=rt)
                rx[k + 1] = r
            self._update_interfaces(st, k, bck, "x")

            if st.amen and not last:
                kr = min(st.kick_rank, rzm.shape[0], rzm.shape[1])
                Uz, _, Wz = K.svd_left(rzm, rt=rt)
                if bck:
                    z[k] = Uz[:, :kr].t().contiguous().reshape(kr, n, int(rz[k + 1]))
                    vT = Wz[:kr].t().contiguous()                          # (rz_k * bs, kr)
                    a, dd, c = z[k - 1].shape
                    G = K.gemm(z[k - 1].reshape(a * dd, c), vT.reshape(c, bs * kr), rt=rt).reshape(a, dd, bs, kr)
                    z[k - 1] = K.permute4(G, (0, 2, 1, 3), scale=scales, scale_axis=1, divide=True, rt=rt)
                    rz[k] = kr
                else:
                    z[k] = Uz[:, :kr].contiguous().reshape(int(rz[k]), n, kr)
                    Rn, dd, k2 = z[k + 1].shape
                    G = K.gemm(Wz[:kr].contiguous().reshape(kr * bs, Rn), z[k + 1].reshape(Rn, dd * k2), rt=rt)
                    z[k + 1] = K.permute4(G.reshape(kr, bs, dd, k2), (0, 1, 2, 3), scale=scales, scale_axis=1,
                                          divide=True, rt=rt)
                    rz[k + 1] = kr
                self._update_interfaces(st, k, bck, "z")
        return local_res, local_dx

    # ------------------------------------------------------------------------------------------
    def prepare(self, x0=None, kick_rank=2, amen=True):
        """Host part of tt_block_amen's set-up (reference src/tt_als.py:527-585): validate the warm start,
        draw the random residual train in the reference's RNG order and upload everything once."""
        rt = self.rt
        bs = self.block_size
        model = next(iter(self.b.values()))
        x_shape = tuple(model[0].shape[1:-1])

        def fresh():
            from . import tt as T
            from .runtime import use_runtime
            with use_runtime(rt):
                cores = T.tt_normalise([np.random.randn(1, *x_shape, 1) for _ in model[:-1]])
            return cores + [np.random.randn(1, bs, *x_shape, 1)]

        direction = 1
        if x0 is None:
            xh = fresh()
        else:
            xh = x0
            where = [i for i, c in enumerate(xh) if c.ndim == 4 and c.shape[1] == bs]
            if len(where) != 1 or where[0] not in (0, len(xh) - 1):
                xh = fresh()
            elif where[0] == 0:
                direction = -1
        st = _State()
        st.direction = direction
        st.N = [c.shape[-2] for c in xh]
        st.d = d = len(st.N)
        st.rx = np.array([1] + [c.shape[0] for c in xh[1:]] + [1])
        st.amen = amen
        st.kick_rank = kick_rank
        one3 = lambda: rt.to_device(np.ones((1, 1, 1)))
        one2 = lambda: rt.to_device(np.ones((1, 1)))
        keys = list(self.A.keys())
        rows = list(self.b.keys())
        st.XAX = [{key: one3() for key in keys}] + [dict() for _ in range(d - 1)] + [{key: one3() for key in keys}]
        st.Xb = [{i: one2() for i in rows}] + [dict() for _ in range(d - 1)] + [{i: one2() for i in rows}]
        st.ZAX = st.Zb = st.z = st.rz = None
        if amen:
            tk = _tkeys(keys, self.transposes)
            st.ZAX = [{key: one3() for key in tk}] + [dict() for _ in range(d - 1)] + [{key: one3() for key in tk}]
            st.Zb = [{i: one2() for i in rows}] + [dict() for _ in range(d - 1)] + [{i: one2() for i in rows}]
            kr = kick_rank
            zh = [np.divide(1, np.prod(xh[0].shape[1:-1]) * kr ** 2) * np.random.randn(*xh[0].shape[:-1], kr)]
            zh += [np.divide(1, np.prod(c.shape[1:-1]) * kr ** 2) * np.random.randn(kr, *c.shape[1:-1], kr)
                   for c in xh[1:-1]]
            zh += [np.divide(1, np.prod(xh[-1].shape[1:-1]) * kr ** 2) * np.random.randn(kr, *xh[-1].shape[1:])]
            st.rz = np.array([1] + [c.shape[0] for c in zh[1:]] + [1])
            st.z = [rt.to_device(c) for c in zh]
        st.x = [rt.to_device(c) for c in xh]
        return st

    def run(self, st, term_tol, r_max=100, eps=1e-12, nswp=22):
        """The sweeps of tt_block_amen (reference src/tt_als.py:586-670) on a prepared, device-resident state.
        Returns (device cores, final local residual)."""
        d = st.d
        st.eps, st.r_max = eps, r_max
        st.trunc_tol = term_tol / math.sqrt(d)
        st.direct_solve_failure = False
        direction = st.direction
        last = False
        final_res = math.inf
        self.sweeps = 0
        for swp in range(nswp + 1):
            local_res, local_dx = self._sweep(st, direction, swp, last)
            self.sweeps = swp
            if last:
                break
            if local_res < term_tol or local_dx < eps or swp == nswp - 2:
                last = True
                final_res = local_res
            direction *= -1
        self.ranks = [int(v) for v in st.rx[1:-1]]
        return st.x, final_res

    def solve(self, term_tol, r_max=100, eps=1e-12, nswp=22, x0=None, kick_rank=2, amen=True):
        """tt_block_amen (reference src/tt_als.py:525-670); x0: list of numpy cores or None.
        Returns (x_cores as numpy list, final local residual)."""
        st = self.prepare(x0, kick_rank, amen)
        xd, final_res = self.run(st, term_tol, r_max, eps, nswp)
        return [self.rt.to_host(c) for c in xd], final_res


class NativeBlockAmen:
    """The same solve driven by the native C++ sweep driver (ttipm_amen_* in include/ttipm.h): Python uploads the
    operands once, makes the reference's NumPy RNG draws, calls run() and fetches the result; every sweep, local
    solve and rank decision in between happens inside libttipm_b200 without returning to the interpreter."""

    def __init__(self, block_A, aliases, transposes, block_b, ineq, rt=None, stats=None):
        import ctypes as C
        self._C = C
        self.rt = rt or get_runtime()
        self.lib = self.rt.lib
        self.block_size = max(k[0] for k in block_A.keys()) + 1
        self.ineq = bool(ineq)
        model = next(iter(block_b.values()))
        self.d = len(model)
        self.x_shape = tuple(model[0].shape[1:-1])
        self.h = C.c_void_p(self.lib.ttipm_amen_create(self.d, self.block_size, int(self.ineq), self.rt.stream()))
        if not self.h.value:
            raise RuntimeError("ttipm_amen_create failed: " + self.lib.ttipm_last_error().decode())
        self._keep = []
        for (i, j), cores in block_A.items():
            for k, c in enumerate(cores):
                a = np.ascontiguousarray(c, dtype=np.float64)
                self._chk(self.lib.ttipm_amen_set_block(self.h, i, j, k, a.ctypes.data, a.shape[0], a.shape[1], a.shape[3]))
        for (i, j), (p, t) in aliases.items():
            self._chk(self.lib.ttipm_amen_add_alias(self.h, i, j, p, t, 0))
        for (i, j), (p, t) in transposes.items():
            self._chk(self.lib.ttipm_amen_add_alias(self.h, i, j, p, t, 1))
        for i, cores in block_b.items():
            for k, c in enumerate(cores):
                a = np.ascontiguousarray(c, dtype=np.float64)
                self._chk(self.lib.ttipm_amen_set_rhs(self.h, i, k, a.ctypes.data, a.shape[0], a.shape[1], a.shape[2]))
        self.stats = stats if stats is not None else {}
        self.trace = []
        self.sweeps = 0
        self.ranks = []

    def _chk(self, code):
        if code != 0:
            from .runtime import TTIPMError
            raise TTIPMError(f"native AMEn driver failed ({code}): {self.lib.ttipm_last_error().decode()}")

    def __del__(self):
        try:
            if self.h:
                self.lib.ttipm_amen_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def _set_train(self, which, cores):
        bs = self.block_size
        for k, c in enumerate(cores):
            a = np.ascontiguousarray(c, dtype=np.float64)
            if a.ndim == 4:
                self._chk(self.lib.ttipm_amen_set_core(self.h, which, k, a.ctypes.data, a.shape[0], a.shape[1], a.shape[2], a.shape[3]))
            else:
                self._chk(self.lib.ttipm_amen_set_core(self.h, which, k, a.ctypes.data, a.shape[0], 0, a.shape[1], a.shape[2]))

    def prepare(self, x0=None, kick_rank=2, amen=True):
        """Host set-up of tt_block_amen (reference src/tt_als.py:527-585): warm-start validation and the random
        residual train, drawn in the reference's RNG order, then one upload."""
        bs = self.block_size

        def fresh():
            from . import tt as T
            from .runtime import use_runtime
            with use_runtime(self.rt):
                cores = T.tt_normalise([np.random.randn(1, *self.x_shape, 1) for _ in range(self.d - 1)])
            return cores + [np.random.randn(1, bs, *self.x_shape, 1)]

        direction = 1
        if x0 is None:
            xh = fresh()
        else:
            xh = x0
            where = [i for i, c in enumerate(xh) if c.ndim == 4 and c.shape[1] == bs]
            if len(where) != 1 or where[0] not in (0, len(xh) - 1):
                xh = fresh()
            elif where[0] == 0:
                direction = -1
        self._set_train(0, xh)
        if amen:
            kr = kick_rank
            zh = [np.divide(1, np.prod(xh[0].shape[1:-1]) * kr ** 2) * np.random.randn(*xh[0].shape[:-1], kr)]
            zh += [np.divide(1, np.prod(c.shape[1:-1]) * kr ** 2) * np.random.randn(kr, *c.shape[1:-1], kr)
                   for c in xh[1:-1]]
            zh += [np.divide(1, np.prod(xh[-1].shape[1:-1]) * kr ** 2) * np.random.randn(kr, *xh[-1].shape[1:])]
            self._set_train(1, zh)
        st = _State()
        st.direction, st.amen, st.kick_rank = direction, amen, kick_rank
        return st

    def run(self, st, term_tol, r_max=100, eps=1e-12, nswp=22):
        C = self._C
        if self.stats.get("profile") is not None:
            self.lib.ttipm_amen_set_profile(self.h, 1)
        res = C.c_double(0.0)
        sw = C.c_int(0)
        self._chk(self.lib.ttipm_amen_run(self.h, term_tol, int(r_max), eps, int(nswp), int(st.kick_rank), int(st.amen),
                                          int(st.direction), C.byref(res), C.byref(sw)))
        self.sweeps = sw.value
        return None, res.value

    def fetch(self):
        """Download the solution train (list of NumPy cores) and the run statistics."""
        C = self._C
        out = []
        dims = (_cabi_i32() * 4)()
        for k in range(self.d):
            self._chk(self.lib.ttipm_amen_core_shape(self.h, k, dims))
            shape = (dims[0], dims[1], dims[2], dims[3]) if dims[1] > 0 else (dims[0], dims[2], dims[3])
            a = np.empty(shape, dtype=np.float64)
            self._chk(self.lib.ttipm_amen_get_core(self.h, k, a.ctypes.data))
            out.append(a)
        self.ranks = [c.shape[0] for c in out[1:]]
        stats = np.zeros(12)
        self._chk(self.lib.ttipm_amen_stats(self.h, stats.ctypes.data, None, 0))
        nrows = int(stats[9])
        tr = np.zeros((max(nrows, 1), 5))
        self._chk(self.lib.ttipm_amen_stats(self.h, stats.ctypes.data, tr.ctypes.data, nrows))
        self.trace = [tuple(row) for row in tr[:nrows]]
        self.native_stats = dict(zip(["sweeps", "local_solves", "dense_solves", "krylov_solves", "krylov_its",
                                      "krylov_matvecs", "launches", "syncs", "peak_bytes", "trace_rows", "krylov_seconds",
                                      "krylov_flops"], stats.tolist()))
        self.stats.update(self.native_stats)
        prof = np.zeros(27)
        self._chk(self.lib.ttipm_amen_profile(self.h, prof.ctypes.data))
        # per kernel category of a profiled run: (seconds, algorithmic flops | bytes, launches)
        self.native_profile = {name: tuple(prof[3 * q:3 * q + 3]) for q, name in enumerate(PROFILE_CATEGORIES)}
        return out

    def solve(self, term_tol, r_max=100, eps=1e-12, nswp=22, x0=None, kick_rank=2, amen=True):
        st = self.prepare(x0, kick_rank, amen)
        _, res = self.run(st, term_tol, r_max, eps, nswp)
        return self.fetch(), res


PROFILE_CATEGORIES = ("block_matvec", "phi_update", "rhs_contract", "bond_gemm", "qr", "svd", "memory_bound",
                      "dense_schur", "krylov")


def _cabi_i32():
    from . import _cabi
    return _cabi.i32
