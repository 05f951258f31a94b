"""TT algebra of the IPM driver on the device.

Same functions, argument meaning and mutation behaviour as the reference's cy_src/tt_ops_cy.pyx and
the IPM-used subset of src/tt_ops.py.  Inputs are lists of NumPy cores (the reference's boundary
type) or the lazy device-resident lists (`devtt.TTList`) these functions return; the arithmetic runs in
the native TT-algebra driver of libttipm_b200 (`ttipm_tt_*`: gemm / QR / Jacobi SVD / block-diagonal
assembly / embeddings launched back to back from C++), so a chain of TT operations keeps its
intermediates in HBM and NumPy cores appear only when the caller looks at them.  Rank decisions use the
reference's exact rule (prune_singular_vals) on the device-computed singular values.
"""
from ctypes import c_double as C_double

import numpy as np

from . import kernels as K
from .devtt import Handle, TTList, as_handle, shapes_of
from .runtime import get_runtime


def _native(ptr, rt):
    """TTList around a freshly created ttipm_tt handle (raises on NULL)."""
    return TTList(Handle(ptr, rt))


# ---- pure bookkeeping (no arithmetic): kept on the host exactly as in the reference ---------------
def tt_ranks(tt):
    """cy_src/tt_ops_cy.pyx:82-92 (shapes only: a device-resident train is not downloaded)."""
    return [sh[0] for sh in shapes_of(tt)[1:]]


def tt_identity(dim):
    """cy_src/tt_ops_cy.pyx:21-29 (d references to ONE array, like the reference)."""
    I = np.eye(2).reshape(1, 2, 2, 1)
    return [I] * dim


def tt_zero_matrix(dim):
    """cy_src/tt_ops_cy.pyx:33-41."""
    Z = np.zeros((1, 2, 2, 1))
    return [Z] * dim


def tt_one_matrix(dim):
    """cy_src/tt_ops_cy.pyx:45-53."""
    W = np.ones((1, 2, 2, 1))
    return [W] * dim


def tt_transpose(tt):
    """cy_src/tt_ops_cy.pyx:57-78 (views; axes 1,2 swapped from the first 4-D core onward)."""
    if isinstance(tt, TTList) and all(len(sh) == 4 for sh in tt.shapes()):
        h = tt.handle()
        if h is not None:
            return _native(h.rt.lib.ttipm_tt_transpose(h.ptr), h.rt)
    split = int(np.argmax([np.ndim(c) for c in tt]))
    return list(tt[:split]) + [np.swapaxes(c, 1, 2) for c in tt[split:]]


def tt_swap_all(tt):
    """cy_src/tt_ops_cy.pyx:118-128."""
    return [np.swapaxes(c, 0, -1) for c in reversed(tt)]


def tt_merge_cores(tt):
    """src/tt_ops.py:335-339: neighbouring cores pairwise into one core whose modes interleave as (i s) resp.
    (i s, j d).  Host NumPy: only the problem generators reach it (psd_system/max_stable_set, graphm), on tiny cores."""
    if tt[0].ndim == 3:
        return [np.einsum("kir,rsK->kisK", a, b) for a, b in zip(tt[:-1:2], tt[1::2])]
    return [np.einsum("kijr,rsdK->kisjdK", a, b) for a, b in zip(tt[:-1:2], tt[1::2])]


def tt_reshape(tt, shape):
    """src/tt_ops.py:330-333."""
    shapes = shapes_of(tt)
    merge = np.prod(shape) > np.prod(shapes[0][1:-1])
    if isinstance(tt, TTList) and not merge and len(shape) in (1, 2):
        h = tt.handle()
        if h is not None:
            return _native(h.rt.lib.ttipm_tt_reshape(h.ptr, int(shape[0]), int(shape[1]) if len(shape) == 2 else 0), h.rt)
    if merge:
        tt = tt_merge_cores(tt)
    return [c.reshape(c.shape[0], *shape, c.shape[-1]) for c in tt]


def tt_scale(alpha, tt):
    """cy_src/tt_ops_cy.pyx:96-114: the C signature rounds alpha to float32 and ONE core, drawn with
    np.random.randint (advancing the global RNG), is scaled; all other cores are shared."""
    idx = np.random.randint(0, len(tt))
    a32 = float(np.float32(alpha))
    rt = get_runtime()
    h = as_handle(tt, rt).clone()
    rt.check(rt.lib.ttipm_tt_scale_core(h.ptr, int(idx), a32), "ttipm_tt_scale_core")
    return TTList(h)


def prune_singular_vals(s, eps):
    """cy_src/tt_ops_cy.pyx:162-177."""
    s = np.asarray(s, dtype=np.float64)
    if np.linalg.norm(s) == 0.0:
        return 1
    sc = np.cumsum(np.abs(s[::-1]) ** 2)[::-1]
    R = max(int(np.argmax(sc < eps ** 2)), 1)
    if sc[-1] > eps ** 2:
        R = s.size
    return R


# ---- device-side building blocks -------------------------------------------------------------------
def _up(tt, rt):
    return [rt.to_device(c) for c in tt]


def _down(tt, rt):
    return [rt.to_host(c) for c in tt]


def _all_rank_one(tt):
    return len(tt) == 1 or all(r == 1 for r in tt_ranks(tt))


# ---- NumPy-boundary API (reference signatures) -----------------------------------------------------
def _assign(train_tt, h):
    """The reference's in-place semantics: the caller's list object now holds the train behind `h`."""
    h._shapes = None
    if isinstance(train_tt, TTList):
        train_tt.rebind(h)
    else:
        train_tt[:] = h.to_numpy()
    return train_tt


def _own_handle(train_tt, rt):
    """A handle the call may modify in place: the train's own handle, or a fresh upload."""
    h = as_handle(train_tt, rt)
    return h if not isinstance(train_tt, TTList) or not train_tt._live else h.clone()


def tt_rl_orthogonalise(train_tt):
    """cy_src/tt_ops_cy.pyx:132-159; mutates and returns the input list."""
    if len(train_tt) == 1:
        return train_tt
    rt = get_runtime()
    h = _own_handle(train_tt, rt)
    rt.check(rt.lib.ttipm_tt_rl_orthogonalise(h.ptr), "ttipm_tt_rl_orthogonalise")
    return _assign(train_tt, h)


def _round_native(h, eps, rt, collect=False):
    """rl-orthogonalise + truncation sweep on a handle with the per-train eps of the reference (eps / sqrt(d - 1) per
    bond); returns the discarded energy of the collecting variant."""
    d = len(h.shapes())
    dropped = C_double(0.0)
    rt.check(rt.lib.ttipm_tt_round(h.ptr, float(eps / np.sqrt(d - 1)) if d > 1 else float(eps), int(collect), dropped),
             "ttipm_tt_round")
    h._shapes = None
    return float(dropped.value)


def tt_rank_reduce(train_tt, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:180-226; mutates and returns the input list."""
    if _all_rank_one(train_tt):
        return train_tt
    rt = get_runtime()
    h = _own_handle(train_tt, rt)
    _round_native(h, eps, rt)
    return _assign(train_tt, h)


def tt_add(train_1_tt, train_2_tt):
    """cy_src/tt_ops_cy.pyx:244-258."""
    rt = get_runtime()
    a, b = as_handle(train_1_tt, rt), as_handle(train_2_tt, rt)
    return _native(rt.lib.ttipm_tt_add(a.ptr, b.ptr), rt)


def tt_sub(train_1_tt, train_2_tt):
    """src/tt_ops.py:189-190."""
    return tt_add(train_1_tt, tt_scale(-1, train_2_tt))


def _psd_like(train_tt, extra_fn, eps):
    d = len(train_tt)
    eps = eps / 2.0
    if _all_rank_one(train_tt):
        return train_tt
    rt = get_runtime()
    h = _own_handle(train_tt, rt)
    dropped = _round_native(h, eps, rt, collect=True)
    _assign(train_tt, h)                   # the reference mutates its input before adding the correction
    factor = pow(dropped, 1.0 / (2 * d))
    extra = Handle.from_numpy(extra_fn(factor), rt)
    return _native(rt.lib.ttipm_tt_add(h.ptr, extra.ptr), rt)


def tt_psd_rank_reduce(train_tt, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:262-325: rounding that adds the discarded energy back as factor * I."""
    shape = train_tt[0].shape
    return _psd_like(train_tt, lambda f: [f * np.eye(shape[1]).reshape(1, *shape[1:-1], 1)] * len(train_tt), eps)


def tt_mask_rank_reduce(train_tt, mask_tt, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:329-388."""
    return _psd_like(train_tt, lambda f: [f * c for c in mask_tt], eps)


def tt_inner_prod(train_1_tt, train_2_tt):
    """cy_src/tt_ops_cy.pyx:506-520."""
    rt = get_runtime()
    a, b = as_handle(train_1_tt, rt), as_handle(train_2_tt, rt)
    out = C_double(0.0)
    rt.check(rt.lib.ttipm_tt_inner(a.ptr, b.ptr, out), "ttipm_tt_inner")
    return float(out.value)


def tt_norm(train_tt):
    """src/tt_ops.py:306-310."""
    v = tt_inner_prod(train_tt, train_tt)
    return float(np.sqrt(v)) if v > 0 else 0.0


def tt_normalise(train_tt, radius=1):
    """cy_src/tt_ops_cy.pyx:524-526 (`radius` is a C int there, i.e. truncated)."""
    factor = np.divide(int(radius), np.sqrt(tt_inner_prod(train_tt, train_tt)))
    return tt_scale(factor, train_tt)


def tt_fast_matrix_vec_mul(matrix_tt, vec_tt, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:430-447."""
    rt = get_runtime()
    A, B = as_handle(matrix_tt, rt), as_handle(vec_tt, rt)
    return _native(rt.lib.ttipm_tt_zipup(0, A.ptr, B.ptr, float(eps)), rt)


def tt_fast_mat_mat_mul(matrix_tt_1, matrix_tt_2, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:451-464."""
    rt = get_runtime()
    A, B = as_handle(matrix_tt_1, rt), as_handle(matrix_tt_2, rt)
    return _native(rt.lib.ttipm_tt_zipup(1, A.ptr, B.ptr, float(eps)), rt)


def tt_fast_hadamard(train_tt_1, train_tt_2, eps=1e-18):
    """cy_src/tt_ops_cy.pyx:468-502."""
    rt = get_runtime()
    A, B = as_handle(train_tt_1, rt), as_handle(train_tt_2, rt)
    return _native(rt.lib.ttipm_tt_zipup(2, A.ptr, B.ptr, float(eps)), rt)


def tt_random_gaussian(target_ranks, shape=(2,)):
    """cy_src/tt_ops_cy.pyx:529-533."""
    rr = [1] + list(target_ranks) + [1]
    return tt_normalise([np.divide(1, a * np.prod(shape) * b) * np.random.randn(a, *shape, b)
                         for a, b in zip(rr[:-1], rr[1:])])


def tt_mat_vec_mul(mat, vec, op_tol, eps, verbose=False):
    """src/tt_als.py:1765-1768: exact zip-up product rounded to op_tol while every rank product is <= 80, otherwise the
    ALS fit of the product (als_product.tt_approx_mat_vec_mul, which draws from the global NumPy RNG like the
    reference's)."""
    if np.max(np.array(tt_ranks(mat)) * np.array(tt_ranks(vec))) <= 80:
        out = tt_fast_matrix_vec_mul(mat, vec, eps)
        _round_native(out._h, op_tol, get_runtime())
        return out
    from .als_product import tt_approx_mat_vec_mul
    return tt_approx_mat_vec_mul(mat, vec, tol=op_tol, verbose=verbose)


def tt_mat_mat_mul(mat1, mat2, op_tol, eps, verbose=False):
    """src/tt_als.py:1631-1634: exact zip-up product rounded to op_tol up to a rank product of 40, ALS fit above."""
    if np.max(np.array(tt_ranks(mat1)) * np.array(tt_ranks(mat2))) <= 40:
        out = tt_fast_mat_mat_mul(mat1, mat2, eps)
        _round_native(out._h, op_tol, get_runtime())
        return out
    from .als_product import tt_approx_mat_mat_mul
    return tt_approx_mat_mat_mul(mat1, mat2, tol=op_tol, verbose=verbose)


def _embed(tt, kind):
    rt = get_runtime()
    h = as_handle(tt, rt)                  # keep the (possibly temporary) handle alive across the call
    return _native(rt.lib.ttipm_tt_embed(h.ptr, kind), rt)


def tt_IkronM(matrix_tt):
    """src/tt_ops.py:360-363: I (x) M per core."""
    return _embed(matrix_tt, 0)


def tt_MkronI(matrix_tt):
    """src/tt_ops.py:365-368."""
    return _embed(matrix_tt, 1)


def _diag_embed_tt(tt, eps):
    out = _embed(tt, 2)
    _round_native(out._h, eps, get_runtime())
    return out


def tt_diag(vec_tt, eps=1e-18):
    """src/tt_ops.py:312-316."""
    return _diag_embed_tt(vec_tt, eps)


def tt_diag_op(matrix_tt, eps=1e-18):
    """src/tt_ops.py:371-375."""
    return _diag_embed_tt(matrix_tt, eps)


def tt_diagonal(matrix_tt):
    """src/tt_ops.py:318-319 (pure view bookkeeping)."""
    return [np.transpose(np.diagonal(c, axis1=1, axis2=2), (0, 2, 1)) for c in matrix_tt]


def tt_entrywise_sum(train_tt):
    """src/tt_ops.py:342-352: inner product with the all-ones train."""
    ones = [np.ones((1, *sh[1:-1], 1)) for sh in shapes_of(train_tt)]
    return tt_inner_prod(train_tt, ones)


def tt_sum(*args, op_tol=1e-18, rank_reduce=True):
    """src/tt_ops.py:321-328."""
    out = args[0]
    for a in args[1:]:
        out = tt_rank_reduce(tt_add(out, a), op_tol) if rank_reduce else tt_add(out, a)
    return out


def tt_rl_orthogonalise_py(train_tt):
    """src/tt_ops.py:30-42: loops down to i = 0, so core 0 is normalised and its 1x1 R factor wraps
    around into the LAST core."""
    d = len(train_tt)
    if d == 1:
        return train_tt
    rt = get_runtime()
    dev = _up(train_tt, rt)
    for i in range(d - 1, -1, -1):
        sh = tuple(dev[i].shape)
        prev = (i - 1) % d
        shm = tuple(dev[prev].shape)
        Q, R = K.qr(dev[i].reshape(sh[0], -1).t(), rt=rt)
        dev[i] = Q.t().contiguous().reshape(-1, *sh[1:-1], sh[-1])
        dev[prev] = K.gemm(dev[prev].reshape(-1, R.shape[1]), R.t(), rt=rt).reshape(-1, *shm[1:-1], dev[i].shape[0])
    train_tt[:] = _down(dev, rt)
    return train_tt


def tt_rank_retraction(train_tt, upper_ranks):
    """src/tt_ops.py:132-152: cap the ranks by keeping the top-k singular triplets."""
    train_tt = tt_rl_orthogonalise_py(train_tt)
    rt = get_runtime()
    dev = _up(train_tt, rt)
    rank = 1
    for idx, ur in enumerate(upper_ranks):
        sh = tuple(dev[idx].shape)
        shn = tuple(dev[idx + 1].shape)
        U, S, W = K.svd_left(dev[idx].reshape(rank * int(np.prod(sh[1:-1])), -1), rt=rt)
        nr = min(int(ur), S.shape[0])
        # the reference selects with argpartition (unordered); the kept subspace is the same top-nr set
        dev[idx] = U[:, :nr].contiguous().reshape(rank, *sh[1:-1], nr)
        dev[idx + 1] = K.gemm(W[:nr], dev[idx + 1].reshape(shn[0], -1), rt=rt).reshape(nr, *shn[1:-1], shn[-1])
        rank = nr
    train_tt[:] = _down(dev, rt)
    return train_tt
