"""Tensor-level wrappers over the C ABI (include/ttipm.h).

Tensors are float64 torch tensors on the runtime's device; non-contiguous views are fine
wherever the ABI takes strides (that is how transposed / permuted operands are expressed).
"""
import ctypes as C

import torch

from . import _cabi
from .runtime import get_runtime


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class TermList:
    """A list of projected operator blocks; keeps the tensors alive while the C structs are in use."""

    def __init__(self):
        self.items = []

    def add(self, P1, A, P2, in_block, out_block, alpha=1.0):
        assert P1.dim() == 3 and A.dim() == 4 and P2.dim() == 3
        assert P1.shape[1] == A.shape[0] and P2.shape[1] == A.shape[3], (P1.shape, A.shape, P2.shape)
        self.items.append((P1, A, P2, int(in_block), int(out_block), float(alpha)))
        return self

    def __len__(self):
        return len(self.items)

    def carray(self):
        n = len(self.items)
        arr = (_cabi.Term * max(n, 1))()
        for q, (P1, A, P2, ib, ob, alpha) in enumerate(self.items):
            t = arr[q]
            t.P1, t.A, t.P2 = P1.data_ptr(), A.data_ptr(), P2.data_ptr()
            t.p1_strides[:] = list(P1.stride())
            t.a_strides[:] = list(A.stride())
            t.p2_strides[:] = list(P2.stride())
            t.s, t.S = A.shape[0], A.shape[3]
            t.in_block, t.out_block, t.alpha = ib, ob, alpha
        return arr, n


def block_matvec(terms, x, nb_out, out_ranks, sub=None, y_scale=1.0, sub_scale=-1.0, want_norm=False,
                 x_layout="rbnR", rt=None):
    """y[:, i] = y_scale * sum_terms alpha * (P1 A P2) x[:, in_block] (+ sub_scale * sub[:, i])
    (reference src/tt_als.py:190-238).

    x: (r, b_in, n, R) ["rbnR"] or the forward unfolding (r, n, b_in, R) ["rnbR"], optionally with a
    leading batch axis; returns y (l, nb_out, n, L) [batched likewise] and, if want_norm, a device
    tensor with the squared-norm partials (sum them on the host)."""
    rt = rt or get_runtime()
    batched = x.dim() == 5
    xb = x if batched else x.unsqueeze(0)
    if x_layout == "rbnR":
        B, r, b_in, n, R = xb.shape
        x_bs, x_rs, x_ns = xb.stride(2), xb.stride(1), xb.stride(3)
    else:
        B, r, n, b_in, R = xb.shape
        x_bs, x_rs, x_ns = xb.stride(3), xb.stride(1), xb.stride(2)
    assert xb.stride(4) == 1, "x needs a unit stride along R"
    l, L = out_ranks
    y = rt.empty(B, l, nb_out, n, L)
    arr, nt = terms.carray()
    for (P1, A, P2, ib, ob, _) in terms.items:
        assert P1.shape[0] == l and P1.shape[2] == r and P2.shape[0] == L and P2.shape[2] == R, \
            (P1.shape, P2.shape, (l, r, L, R))
        assert A.shape[1] == n and A.shape[2] == n and ib < b_in and ob < nb_out
    sumsq = rt.empty(B, nb_out * L) if want_norm else None
    if sub is not None:
        assert sub.shape == (l, nb_out, n, L) and sub.is_contiguous()
    code = rt.lib.ttipm_block_matvec(arr, nt, l, L, r, R, n, nb_out, _ptr(xb), x_bs, x_rs, x_ns, xb.stride(0),
                                     _ptr(y), y.stride(2), y.stride(1), y.stride(3), y.stride(0), y_scale,
                                     _ptr(sub), sub_scale, _ptr(sumsq), B, rt.stream())
    rt.check(code, "ttipm_block_matvec")
    if not batched:
        y = y[0]
    return (y, sumsq) if want_norm else y


def local_diag(P1, A, P2, invert=False, rt=None):
    """'lsr,smnS,LSR->lmL' (reference src/tt_ipm.py:191)."""
    rt = rt or get_runtime()
    tl = TermList().add(P1, A, P2, 0, 0)
    arr, _ = tl.carray()
    l, L, n = P1.shape[0], P2.shape[0], A.shape[1]
    out = rt.empty(l, n, L)
    rt.check(rt.lib.ttipm_local_diag(arr, l, L, n, int(invert), _ptr(out), rt.stream()), "ttipm_local_diag")
    return out


def local_dense(P1, A, P2, rt=None):
    """'lsr,smnS,LSR->lmLrnR' as an (l n L) x (r n R) matrix (reference src/tt_ipm.py:201-212)."""
    rt = rt or get_runtime()
    tl = TermList().add(P1, A, P2, 0, 0)
    arr, _ = tl.carray()
    l, r, L, R, n = P1.shape[0], P1.shape[2], P2.shape[0], P2.shape[2], A.shape[1]
    out = rt.empty(l * n * L, r * n * R)
    rt.check(rt.lib.ttipm_local_dense(arr, l, L, r, R, n, _ptr(out), rt.stream()), "ttipm_local_dense")
    return out


def phi_update(phis, cores_A, U, V, forward, rt=None):
    """Interface updates for several stored blocks in one launch (reference src/tt_als.py:252-257).

    phis: list of Phi tensors (contiguous), cores_A: list of operator cores (any strides),
    U: left core (ul, n, uL), V: right core (vr, n, vR).  Returns the list of new interfaces."""
    rt = rt or get_runtime()
    n = len(phis)
    assert U.is_contiguous() and V.is_contiguous()
    ul, nm, uL = U.shape
    vr, _, vR = V.shape
    arr = (_cabi.PhiTerm * n)()
    outs = []
    for q, (Phi, A) in enumerate(zip(phis, cores_A)):
        assert Phi.is_contiguous()
        s, S = A.shape[0], A.shape[3]
        if forward:
            assert Phi.shape == (ul, s, vr), (Phi.shape, (ul, s, vr))
            out = rt.empty(uL, S, vR)
        else:
            assert Phi.shape == (uL, S, vR), (Phi.shape, (uL, S, vR))
            out = rt.empty(ul, s, vr)
        outs.append(out)
        t = arr[q]
        t.Phi, t.A, t.out = Phi.data_ptr(), A.data_ptr(), out.data_ptr()
        t.a_strides[:] = list(A.stride())
        t.s, t.S = s, S
    code = rt.lib.ttipm_phi_update(arr, n, int(forward), _ptr(U), ul, uL, _ptr(V), vr, vR, nm, rt.stream())
    rt.check(code, "ttipm_phi_update")
    return outs


def rhs_project(Xb1s, Bs, Xb2s, out, blocks, rt=None):
    """out[:, blocks[q]] = 'br,bnB,BR->rnR' (reference src/tt_als.py:82); out: (r, nb, n, R) contiguous."""
    rt = rt or get_runtime()
    n = len(Bs)
    r, nb, nm, R = out.shape
    arr = (_cabi.RhsTerm * n)()
    for q in range(n):
        X1, Bc, X2 = Xb1s[q], Bs[q], Xb2s[q]
        assert X1.is_contiguous() and Bc.is_contiguous() and X2.is_contiguous()
        assert X1.shape == (Bc.shape[0], r) and X2.shape == (Bc.shape[2], R), (X1.shape, Bc.shape, X2.shape, out.shape)
        t = arr[q]
        t.Xb1, t.B, t.Xb2 = X1.data_ptr(), Bc.data_ptr(), X2.data_ptr()
        t.out = out.data_ptr() + 8 * blocks[q] * out.stride(1)
        t.b, t.Bp = Bc.shape[0], Bc.shape[2]
    code = rt.lib.ttipm_rhs_contract(arr, n, 0, C.c_void_p(0), r, R, nm, out.stride(0), rt.stream())
    rt.check(code, "ttipm_rhs_contract")
    return out


def phi_rhs_update(Xbs, Bs, core, forward, rt=None):
    """'br,bnB,rnR->BR' (forward) / 'BR,bnB,rnR->br' (backward), reference src/tt_als.py:260-265."""
    rt = rt or get_runtime()
    n = len(Bs)
    assert core.is_contiguous()
    r, nm, R = core.shape
    arr = (_cabi.RhsTerm * n)()
    outs = []
    for q in range(n):
        Xb, Bc = Xbs[q], Bs[q]
        assert Xb.is_contiguous() and Bc.is_contiguous()
        t = arr[q]
        t.B = Bc.data_ptr()
        t.b, t.Bp = Bc.shape[0], Bc.shape[2]
        if forward:
            assert Xb.shape == (Bc.shape[0], r)
            out = rt.empty(Bc.shape[2], R)
            t.Xb1, t.Xb2 = Xb.data_ptr(), 0
        else:
            assert Xb.shape == (Bc.shape[2], R)
            out = rt.empty(Bc.shape[0], r)
            t.Xb1, t.Xb2 = 0, Xb.data_ptr()
        t.out = out.data_ptr()
        outs.append(out)
    code = rt.lib.ttipm_rhs_contract(arr, n, 1 if forward else 2, _ptr(core), r, R, nm, 0, rt.stream())
    rt.check(code, "ttipm_rhs_contract")
    return outs


def gemm(A, B, out=None, alpha=1.0, beta=0.0, rt=None):
    """C = alpha * A @ B + beta * C for 2-D (or batched 3-D) strided views."""
    rt = rt or get_runtime()
    batched = A.dim() == 3
    A3 = A if batched else A.unsqueeze(0)
    B3 = B if B.dim() == 3 else B.unsqueeze(0)
    nb, M, K = A3.shape
    N = B3.shape[2]
    assert B3.shape[1] == K
    if out is None:
        out = rt.empty(nb, M, N) if batched else rt.empty(M, N)
        beta = 0.0
    C3 = out if out.dim() == 3 else out.unsqueeze(0)
    bs = lambda t: t.stride(0) if t.shape[0] > 1 else 0
    code = rt.lib.ttipm_gemm(M, N, K, alpha, _ptr(A3), A3.stride(1), A3.stride(2), bs(A3), _ptr(B3), B3.stride(1),
                             B3.stride(2), bs(B3) if B3.shape[0] == nb else 0, beta, _ptr(C3), C3.stride(1),
                             C3.stride(2), bs(C3), nb, rt.stream())
    rt.check(code, "ttipm_gemm")
    return out


class ReducedOperator:
    """Schur-reduced local KKT operator + device LGMRES (reference cy_src/lgmres_cy.pyx:203-510 and
    src/tt_ipm.py:101-162).  P1/A/P2: dicts keyed by block (i, j) of device tensors."""

    KEYS_EQ = [(0, 0), (0, 1), (2, 1), (2, 2)]
    KEYS_INEQ = KEYS_EQ + [(3, 1), (3, 3)]

    def __init__(self, P1, A, P2, inv_I, ineq, rt=None):
        self.rt = rt or get_runtime()
        self.ineq = bool(ineq)
        self.inv_I = inv_I
        assert inv_I.is_contiguous()
        self.r, self.n, self.R = inv_I.shape
        self.nblk = 3 if ineq else 2
        self._keep = []
        self._terms = []
        for key in (self.KEYS_INEQ if ineq else self.KEYS_EQ):
            tl = TermList().add(P1[key], A[key], P2[key], 0, 0)
            arr, _ = tl.carray()
            self._keep.append((tl, arr))
            self._terms.append(arr)
        while len(self._terms) < 6:
            self._terms.append(None)
        self._ws = None
        self._ws_key = None

    def _workspace(self, restart, augment):
        key = (restart, augment)
        if self._ws_key != key:
            n = self.rt.lib.ttipm_lgmres_workspace(int(self.ineq), self.r, self.R, self.n, restart, augment)
            self._ws = self.rt.empty(int(n))
            self._ws_key = key
        return self._ws

    def _call(self, rhs, restart, augment, max_it, rtol, apply_only, grid_hint):
        rt = self.rt
        assert rhs.is_contiguous() and rhs.numel() == self.nblk * self.r * self.n * self.R
        ws = self._workspace(restart, augment)
        x = rt.empty(self.nblk, self.r, self.n, self.R)
        info = rt.empty(6)
        t = self._terms
        code = rt.lib.ttipm_local_lgmres(int(self.ineq), t[0], t[1], t[2], t[3], t[4], t[5], _ptr(self.inv_I), self.r,
                                         self.R, self.n, _ptr(rhs), _ptr(x), _ptr(ws), ws.numel(), restart, augment,
                                         max_it, rtol, int(apply_only), grid_hint, _ptr(info), rt.stream())
        rt.check(code, "ttipm_local_lgmres")
        return x, info

    def matvec(self, v, grid_hint=0):
        """MatVecWrapper.matvec / IneqMatVecWrapper.matvec on a block-major device vector."""
        return self._call(v, 4, 3, 1, 1e-5, True, grid_hint)[0]

    def solve(self, rhs, restart, augment, max_it=300, rtol=1e-5, grid_hint=0):
        """-> (x, info) ; info = device tensor [its, matvecs, reason, cycles, residual estimate, grid]."""
        return self._call(rhs, restart, augment, max_it, rtol, False, grid_hint)


# ---- dense factorisations -------------------------------------------------------------------------
def rt_error(rt, what):
    from .runtime import TTIPMError
    return TTIPMError(f"{what} failed: {rt.lib.ttipm_last_error().decode()}")


def _mat3(A):
    return A if A.dim() == 3 else A.unsqueeze(0)


def qr(A, rt=None):
    """Economic QR of a strided 2-D view (or a batch): returns Q (M, K), R (K, N)."""
    rt = rt or get_runtime()
    A3 = _mat3(A)
    nb, M, N = A3.shape
    Kk = min(M, N)
    Q, R = rt.empty(nb, M, Kk), rt.empty(nb, Kk, N)
    nws = int(rt.lib.ttipm_qr_workspace(M, N, nb))
    if nws <= 0:
        raise rt_error(rt, f"ttipm_qr_workspace({M}, {N}, {nb})")
    ws = rt.empty(nws)
    code = rt.lib.ttipm_qr(_ptr(A3), A3.stride(1), A3.stride(2), A3.stride(0) if nb > 1 else 0, M, N, _ptr(Q), _ptr(R),
                           _ptr(ws), nb, rt.stream())
    rt.check(code, "ttipm_qr")
    return (Q, R) if A.dim() == 3 else (Q[0], R[0])


def svd_left(A, rt=None):
    """U (M, K), s (K) descending, W = diag(s) V^T (K, N) of a strided 2-D view (or a batch)."""
    rt = rt or get_runtime()
    A3 = _mat3(A)
    nb, M, N = A3.shape
    Kk = min(M, N)
    U, S, W = rt.empty(nb, M, Kk), rt.empty(nb, Kk), rt.empty(nb, Kk, N)
    nws = int(rt.lib.ttipm_svd_workspace(M, N, nb))
    if nws <= 0:
        raise rt_error(rt, f"ttipm_svd_workspace({M}, {N}, {nb})")
    ws = rt.empty(nws)
    code = rt.lib.ttipm_svd_left(_ptr(A3), A3.stride(1), A3.stride(2), A3.stride(0) if nb > 1 else 0, M, N, _ptr(U),
                                 _ptr(S), _ptr(W), _ptr(ws), C.c_void_p(0), nb, rt.stream())
    rt.check(code, "ttipm_svd_left")
    return (U, S, W) if A.dim() == 3 else (U[0], S[0], W[0])


# ---- memory-bound helpers ---------------------------------------------------------------------------
def permute4(x, perm, scale=None, scale_axis=0, divide=False, rt=None):
    """out = x.permute(perm) materialised (x 4-D contiguous), optionally scaled along an OUTPUT axis."""
    rt = rt or get_runtime()
    assert x.dim() == 4 and x.is_contiguous()
    dims = (_cabi.i32 * 4)(*x.shape)
    pm = (_cabi.i32 * 4)(*perm)
    out = rt.empty(*[x.shape[q] for q in perm])
    if scale is not None:
        assert scale.is_contiguous() and scale.numel() == out.shape[scale_axis]
    code = rt.lib.ttipm_permute4(_ptr(x), dims, pm, _ptr(out), _ptr(scale), scale_axis, 2 if divide else 1,
                                 rt.stream())
    rt.check(code, "ttipm_permute4")
    return out


def block_norms(x, floor=1e-10, rt=None):
    """scales[j] = max(||x[:, j]||, floor) for x (r, b, n, R) contiguous (reference src/tt_als.py:321)."""
    rt = rt or get_runtime()
    assert x.is_contiguous()
    r, b = x.shape[0], x.shape[1]
    out = rt.empty(b)
    rt.check(rt.lib.ttipm_block_norms(_ptr(x), r, b, x.numel() // (r * b), floor, _ptr(out), rt.stream()),
             "ttipm_block_norms")
    return out


def _panels(tensors):
    """Common (rows, inner) panel view of equally shaped tensors that are contiguous or x[:, j]-style slices;
    returns rows, inner and the row stride of every tensor (0 for None)."""
    live = [t for t in tensors if t is not None]
    shape = live[0].shape
    for t in live:
        assert t.shape == shape, (t.shape, shape)
    if all(t.is_contiguous() for t in live):
        n = live[0].numel()
        return 1, n, [n if t is not None else 0 for t in tensors]
    rows = shape[0]
    inner = 1
    for q in shape[1:]:
        inner *= q
    strides = []
    for t in tensors:
        if t is None:
            strides.append(0)
            continue
        acc = 1
        for q in range(t.dim() - 1, 0, -1):
            assert t.shape[q] == 1 or t.stride(q) == acc, "unsupported view"
            acc *= t.shape[q]
        strides.append(t.stride(0))
    return rows, inner, strides


def ewise(a, alpha=1.0, b=None, beta=0.0, c=None, gamma=0.0, w=None, out=None, want_sumsq=False, store=True, rt=None):
    """out = w .* (alpha a + beta b) + gamma c on equally shaped tensors (contiguous or [:, j]-style slices)."""
    rt = rt or get_runtime()
    if out is None and store:
        out = rt.empty(*a.shape)
    rows, inner, (a_rs, b_rs, c_rs, w_rs, o_rs) = _panels([a, b, c, w, out])
    ss = rt.empty(256) if want_sumsq else None
    code = rt.lib.ttipm_ewise(rows, inner, alpha, _ptr(a), a_rs, beta, _ptr(b), b_rs, gamma, _ptr(c), c_rs, _ptr(w),
                              w_rs, _ptr(out), o_rs, _ptr(ss), rt.stream())
    rt.check(code, "ttipm_ewise")
    return (out, ss) if want_sumsq else out


def trunc_resnorms(base, Y, rt=None):
    """partials (q, 256) of || base - sum_{i>=j} Y_i ||^2 (reference src/tt_als.py:338-345)."""
    rt = rt or get_runtime()
    assert base.is_contiguous() and Y.is_contiguous()
    q = Y.shape[0]
    ln = base.numel()
    assert Y.numel() == q * ln
    out = rt.empty(q, 256)
    rt.check(rt.lib.ttipm_trunc_resnorms(_ptr(base), _ptr(Y), q, ln, _ptr(out), rt.stream()), "ttipm_trunc_resnorms")
    return out


# ---- TT primitives --------------------------------------------------------------------------------
def block_diag(a, b, where, rt=None):
    """One core of tt_add (reference cy_src/tt_ops_cy.pyx:229-258); where in {'first','mid','last'}."""
    rt = rt or get_runtime()
    a, b = a.contiguous(), b.contiguous()
    mode = {"first": 0, "mid": 1, "last": 2}[where]
    ra, Ra, rb, Rb = a.shape[0], a.shape[-1], b.shape[0], b.shape[-1]
    mid = tuple(a.shape[1:-1])
    assert mid == tuple(b.shape[1:-1])
    n = 1
    for q in mid:
        n *= q
    ro = ra if mode == 0 else ra + rb
    Ro = Ra if mode == 2 else Ra + Rb
    out = rt.empty(ro, *mid, Ro)
    rt.check(rt.lib.ttipm_block_diag(_ptr(a), _ptr(b), _ptr(out), ra, Ra, rb, Rb, n, mode, rt.stream()),
             "ttipm_block_diag")
    return out


def embed(core, kind, rt=None):
    """'IkronM' / 'MkronI' on a (r,2,2,R) core -> (r,4,4,R); 'diag' on a (r,q,R) core -> (r,q,q,R)
    (reference src/tt_ops.py:360-375, :312-316)."""
    rt = rt or get_runtime()
    core = core.contiguous()
    mode = {"IkronM": 0, "MkronI": 1, "diag": 2}[kind]
    r, R = core.shape[0], core.shape[-1]
    if mode == 2:
        q = core.shape[1]
        out = rt.empty(r, q, q, R)
    else:
        assert tuple(core.shape[1:3]) == (2, 2)
        q = 4
        out = rt.empty(r, 4, 4, R)
    rt.check(rt.lib.ttipm_embed(_ptr(core), _ptr(out), r, R, q, mode, rt.stream()), "ttipm_embed")
    return out


def _scale2d(M, s, axis, divide, rt):
    rt = rt or get_runtime()
    rows, cols = M.shape
    out = rt.empty(rows, cols)
    s = s.contiguous()
    rt.check(rt.lib.ttipm_scale2d(_ptr(M), M.stride(0), M.stride(1), rows, cols, _ptr(s), axis, int(divide),
                                  _ptr(out), rt.stream()), "ttipm_scale2d")
    return out


def scale_cols(M, s, divide=False, rt=None):
    return _scale2d(M, s, 1, divide, rt)


def scale_rows(M, s, divide=False, rt=None):
    return _scale2d(M, s, 0, divide, rt)


# ---- local eigenvalue problems of the step-size sweeps ----------------------------------------------
def eig_assemble(P1, A1, A2, P2, symmetrise=True, rt=None):
    """Dense projection 'lsr,smnk,kptS,LSR->lmpLrntR' (A2 None: 'lsr,smnS,LSR->lmLrnR') as an (m, m) matrix,
    0.5 (M + M^T) if symmetrise (reference src/tt_als.py:952-959, :1037-1041, :1305, :1346)."""
    rt = rt or get_runtime()
    op = _cabi.EigOp()
    op.P1, op.A1, op.P2 = P1.data_ptr(), A1.data_ptr(), P2.data_ptr()
    op.A2 = A2.data_ptr() if A2 is not None else None
    op.p1_strides[:] = list(P1.stride())
    op.a1_strides[:] = list(A1.stride())
    op.a2_strides[:] = list(A2.stride()) if A2 is not None else [0, 0, 0, 0]
    op.p2_strides[:] = list(P2.stride())
    l, s, L = P1.shape[0], A1.shape[0], P2.shape[0]
    n1 = A1.shape[1]
    if A2 is not None:
        k, n2, S = A1.shape[3], A2.shape[1], A2.shape[3]
    else:
        k, n2, S = A1.shape[3], 1, A1.shape[3]
    assert P1.shape == (l, s, l) and P2.shape == (L, S, L), (P1.shape, A1.shape, P2.shape)
    op.l, op.s, op.k, op.S, op.L, op.n1, op.n2 = l, s, k, S, L, n1, n2
    m = l * n1 * n2 * L
    out = rt.empty(m, m)
    rt.check(rt.lib.ttipm_eig_assemble(C.byref(op), int(symmetrise), _ptr(out), rt.stream()), "ttipm_eig_assemble")
    return out


def _eig_ws(rt, m, K):
    n = int(rt.lib.ttipm_eig_workspace(m, K))
    if n <= 0:
        raise rt_error(rt, f"ttipm_eig_workspace({m}, {K})")
    return rt.empty(n)


def eig_lanczos(A, cA=1.0, D=None, cD=0.0, v0=None, largest=False, K=64, max_cycles=30, tol=1e-9, rt=None):
    """Extreme eigenpair of cA A + cD D (dense symmetric).  Returns (x device vector, out device[10]):
    out = eigenvalue, residual, matvecs, converged, v0^T M v0, ||M v0 - (v0^T M v0) v0||, cycles, ||v0||,
    ||M v0 - eigenvalue v0||, reserved."""
    rt = rt or get_runtime()
    m = A.shape[0]
    K = max(2, min(K, m))
    x, out = rt.empty(m), rt.zeros(10)
    ws = _eig_ws(rt, m, K)
    rt.check(rt.lib.ttipm_eig_lanczos(_ptr(A), float(cA), _ptr(D), float(cD), m, _ptr(v0), int(largest), K,
                                      int(max_cycles), float(tol), _ptr(x), _ptr(out), _ptr(ws), rt.stream()),
             "ttipm_eig_lanczos")
    return x, out


def eig_gen_largest(A, D, v0=None, K=64, max_cycles=30, tol=1e-9, rt=None):
    """Largest eigenpair of (-D) x = lambda A x (A positive definite); out[3] == 0 if A is not positive definite."""
    rt = rt or get_runtime()
    m = A.shape[0]
    K = max(2, min(K, m))
    x, out = rt.empty(m), rt.zeros(10)
    ws = _eig_ws(rt, m, K)
    rt.check(rt.lib.ttipm_eig_gen_largest(_ptr(A), _ptr(D), m, _ptr(v0), K, int(max_cycles), float(tol), _ptr(x),
                                          _ptr(out), _ptr(ws), rt.stream()), "ttipm_eig_gen_largest")
    return x, out
