"""Host mirror of the reference's src/tt_als.py for the Newton-system path.

Same names, argument meaning, return conventions and error behaviour as the reference
(TTBlockMatrix / TTBlockVector containers, tt_block_amen, tt_restarted_block_amen, the
tt_mat_vec_mul / tt_mat_mat_mul dispatchers), so the unmodified IPM driver (src/tt_ipm.py) runs on
top of it; the numerical work happens in the CUDA kernels of libttipm_b200 (ttipm_b200.amen,
ttipm_b200.local_solve, ttipm_b200.tt).  Containers are plain bookkeeping and stay NumPy-typed at
the boundary, exactly like the reference's.
"""
import numpy as np

from . import kernels as K
from . import tt as T
from .amen import DeviceBlockAmen, NativeBlockAmen
from .runtime import get_runtime
from .als_product import tt_approx_mat_mat_mul, tt_approx_mat_vec_mul  # noqa: F401  (reference src/tt_als.py:1502,1637)
from .tt import tt_mat_mat_mul, tt_mat_vec_mul  # noqa: F401  (re-exported, reference src/tt_als.py:1631,1765)
from .tt_ops import cached_einsum  # noqa: F401  (reference src/tt_als.py imports it from src.tt_ops)
from .eigen import tt_max_generalised_eigen, tt_min_eig  # noqa: E402,F401  (reference src/tt_als.py:1132-1499)


def _tt_get_block(i, block_matrix_tt):
    """reference src/tt_als.py:12-14."""
    b = int(np.argmax([len(c.shape) for c in block_matrix_tt]))
    return block_matrix_tt[:b] + [block_matrix_tt[b][:, i]] + block_matrix_tt[b + 1:]


# ---- containers (reference src/tt_als.py:16-250) ---------------------------------------------------
class TTBlockVector:
    def __init__(self):
        self._data = {}

    def __setitem__(self, index, value):
        if not isinstance(value, list):
            raise ValueError("Each entry must be a list")
        self._data[index] = value

    def get_row(self, index):
        return self._data.get(index, None)

    def __getitem__(self, list_index):
        return TTBlockVectorView(self._data, list_index)

    def __iter__(self):
        return iter(self._data)

    def keys(self):
        return self._data.keys()

    def values(self):
        return self._data.values()

    def __repr__(self):
        return repr(self._data)

    @property
    def norm(self):
        return np.sqrt(sum(T.tt_inner_prod(v, v) for v in self._data.values()))

    def __sub__(self, other):
        out = TTBlockVector()
        for i in self._data.keys():
            out[i] = T.tt_rank_reduce(T.tt_sub(self.get_row(i), other.get_row(i)), 1e-12)
        return out

    def scale(self, s):
        self._data = {key: T.tt_rank_reduce(T.tt_scale(s, value), 1e-12) for key, value in self._data.items()}


class TTBlockVectorView:
    def __init__(self, data, list_index):
        self._data = data
        self._list_index = list_index

    def __getitem__(self, row_index):
        return self._data[row_index][self._list_index]

    def items(self):
        for i, row in self._data.items():
            if self._list_index < len(row):
                yield (i, row[self._list_index])

    def __iter__(self):
        return iter(self._data)

    def __repr__(self):
        return repr(dict(self.items()))

    def block_local_product(self, Xb_k, Xb_kp1, nrmsc, shape):
        """'br,bnB,BR->rnR' per row (reference src/tt_als.py:79-83) on the device."""
        rt = get_runtime()
        out = rt.zeros(*shape)
        rows = sorted(self._data.keys())
        up = rt.to_device
        K.rhs_project([up(Xb_k[i]) for i in rows], [up(nrmsc * self._data[i][self._list_index]) for i in rows],
                      [up(Xb_kp1[i]) for i in rows], out, rows, rt=rt)
        return rt.to_host(out)


class _BlockKeys:
    """Key sets of a block operator: stored blocks plus the positions their (transpose) aliases fill
    (reference src/tt_als.py:113-129, :176-188).  Shared by the matrix and its per-core view."""

    def __iter__(self):
        return iter(self._data)

    def keys(self):
        return self._data.keys()

    def _with(self, *alias_maps):
        out = set(self._data.keys())
        for m in alias_maps:
            out |= set(m.values())
        return out

    def tkeys(self):
        return self._with(self._transposes)

    def akeys(self):
        return self._with(self._aliases)

    def all_keys(self):
        return self._with(self._aliases, self._transposes)


class TTBlockMatrix(_BlockKeys):
    def __init__(self):
        self._data = {}
        self._aliases = {}
        self._transposes = {}

    def add_alias(self, key1, key2, is_transpose=False):
        (self._transposes if is_transpose else self._aliases)[key1] = key2

    def __getitem__(self, key):
        if isinstance(key, tuple) and len(key) == 2:
            return self._data.setdefault(key, [])
        if isinstance(key, int):
            return TTBlockMatrixView(self._data, self._aliases, self._transposes, key)
        raise KeyError(f"Invalid key format: {key}")

    def __setitem__(self, key, value):
        if not (isinstance(key, tuple) and len(key) == 2):
            raise KeyError(f"Invalid key format: {key}")
        self._data[key] = value

    def __repr__(self):
        return f"{self.__class__.__name__}({self._data})"

    def block_product(self, x_cores, op_tol, eps=1e-12):
        """A x in TT format, block row by block row (reference src/tt_als.py:132-155)."""
        result = TTBlockVector()

        def accumulate(row, term):
            if row in result.keys():
                result[row] = T.tt_rank_reduce(T.tt_add(result.get_row(row), term), eps)
            else:
                result[row] = term

        for (i, j), blk in self._data.items():
            accumulate(i, tt_mat_vec_mul(blk, _tt_get_block(j, x_cores), op_tol, eps))
            if (i, j) in self._transposes:
                p, t = self._transposes[i, j]
                accumulate(p, tt_mat_vec_mul(T.tt_transpose(blk), _tt_get_block(t, x_cores), op_tol, eps))
            if (i, j) in self._aliases:
                p, t = self._aliases[i, j]
                accumulate(p, tt_mat_vec_mul(blk, _tt_get_block(t, x_cores), op_tol, eps))
        return result

    def get_submatrix(self, row_index, col_index):
        sub = TTBlockMatrix()
        sub._data = {(i, j): v for (i, j), v in self._data.items() if i <= row_index and j <= col_index}
        sub._aliases = {k: (p, t) for k, (p, t) in self._aliases.items() if p <= row_index and t <= col_index}
        sub._transposes = {k: (p, t) for k, (p, t) in self._transposes.items() if p <= row_index and t <= col_index}
        return sub


class TTBlockMatrixView(_BlockKeys):
    """Per-core view; the four local products run on the device (reference src/tt_als.py:190-238)."""

    def __init__(self, data, aliases, transposes, list_index):
        self._data = data
        self._aliases = aliases
        self._transposes = transposes
        self._idx = list_index

    def __getitem__(self, key):
        if not isinstance(key, tuple) or len(key) != 2:
            raise KeyError("Key must be (row, col)")
        return self._data[key][self._idx]

    def items(self):
        for coord, values in self._data.items():
            if len(values) > self._idx:
                yield coord, values[self._idx]

    def __repr__(self):
        return f"IndexView({self._idx})"

    def _product(self, left, right, x_core, shape, left_is_z, right_is_z):
        rt = get_runtime()
        up = rt.to_device
        tl = K.TermList()
        for (i, j), cores in self._data.items():
            A = up(cores[self._idx])
            L, R = up(left[i, j]), up(right[i, j])
            tl.add(L, A, R, j, i)
            if (i, j) in self._transposes:
                p, t = self._transposes[i, j]
                Pl = up(left[p, t]) if left_is_z else L.permute(2, 1, 0)
                Pr = up(right[p, t]) if right_is_z else R.permute(2, 1, 0)
                tl.add(Pl, A.permute(0, 2, 1, 3), Pr, t, p)
            if (i, j) in self._aliases:
                p, t = self._aliases[i, j]
                tl.add(L, A, R, t, p)
        y = K.block_matvec(tl, up(x_core), shape[1], (shape[0], shape[3]), rt=rt)
        return rt.to_host(y)

    def block_local_product(self, XAX_k, XAX_kp1, x_core):
        return self._product(XAX_k, XAX_kp1, x_core, x_core.shape, False, False)

    def compressed_block_local_product(self, ZAX_k, ZAX_kp1, x_core, shape):
        return self._product(ZAX_k, ZAX_kp1, x_core, shape, True, True)

    def lcompressed_block_local_product(self, ZAX_k, XAX_kp1, x_core, shape):
        return self._product(ZAX_k, XAX_kp1, x_core, shape, True, False)

    def rcompressed_block_local_product(self, XAX_k, ZAX_kp1, x_core, shape):
        return self._product(XAX_k, ZAX_kp1, x_core, shape, False, True)


def truncated_svd(matrix, trunc_rank):
    """reference src/tt_als.py:269-274: leading trunc_rank left singular vectors and the matching rows of S V^T."""
    rt = get_runtime()
    U, _, W = K.svd_left(rt.to_device(matrix), rt=rt)
    return rt.to_host(U[:, :trunc_rank]), rt.to_host(W[:trunc_rank])


# ---- interface updates on NumPy operands (reference src/tt_als.py:252-265) --------------------------
def compute_phi_bck_A(Phi_now, core_left, core_A, core_right):
    rt = get_runtime()
    up = rt.to_device
    return rt.to_host(K.phi_update([up(Phi_now)], [up(core_A)], up(core_left), up(core_right), False, rt=rt)[0])


def compute_phi_fwd_A(Phi_now, core_left, core_A, core_right):
    rt = get_runtime()
    up = rt.to_device
    return rt.to_host(K.phi_update([up(Phi_now)], [up(core_A)], up(core_left), up(core_right), True, rt=rt)[0])


def compute_phi_bck_rhs(Phi_now, core_b, core):
    rt = get_runtime()
    up = rt.to_device
    return rt.to_host(K.phi_rhs_update([up(Phi_now)], [up(core_b)], up(core), False, rt=rt)[0])


def compute_phi_fwd_rhs(Phi_now, core_rhs, core):
    rt = get_runtime()
    up = rt.to_device
    return rt.to_host(K.phi_rhs_update([up(Phi_now)], [up(core_rhs)], up(core), True, rt=rt)[0])


# ---- solvers -------------------------------------------------------------------------------------
_INEQ_NAMES = ("_ipm_local_solver_ineq",)
_EQ_NAMES = ("_ipm_local_solver",)


def _solver_kind(local_solver, block_A):
    """The reference passes its local solver as a Python callback that works on host arrays
    (src/tt_ipm.py:183, :284).  The device sweep cannot call it; it dispatches on the callback's
    identity instead and runs the device implementation of the same solver."""
    name = getattr(local_solver, "__name__", None)
    if name in _INEQ_NAMES:
        return True
    if name in _EQ_NAMES:
        return False
    if local_solver is None or name == "_default_local_solver":
        raise NotImplementedError("the generic sparse local solver (reference src/tt_als.py:672-741) is not part of "
                                  "the IPM Newton path; pass _ipm_local_solver or _ipm_local_solver_ineq")
    if isinstance(local_solver, str):
        return local_solver == "ineq"
    raise NotImplementedError(f"unknown local solver {local_solver!r}: the device sweep implements the reference's "
                              "_ipm_local_solver and _ipm_local_solver_ineq")


def tt_block_amen(block_A, block_b, term_tol, r_max=100, eps=1e-12, nswp=22, x0=None, local_solver=None, kick_rank=2,
                  amen=False, verbose=False, _stats=None):
    """reference src/tt_als.py:525-670.  Returns (x_cores, final_local_residual)."""
    ineq = _solver_kind(local_solver, block_A)
    driver = DeviceBlockAmen if (_stats or {}).get("driver") == "python" else NativeBlockAmen
    dev = driver(block_A._data, block_A._aliases, block_A._transposes, block_b._data, ineq, stats=_stats)
    x, res = dev.solve(term_tol, r_max=r_max, eps=eps, nswp=nswp, x0=x0, kick_rank=kick_rank, amen=amen)
    if _stats is not None:            # test / harness hook: one record per solve, the local-solve trace of all of them
        _stats.setdefault("solves", []).append(dict(r_max=r_max, kick_rank=kick_rank, sweeps=dev.sweeps, res=float(res)))
        _stats.setdefault("local_trace", []).extend([list(t) for t in dev.trace])
    if verbose:
        print(f"\tSolution rank is {dev.ranks}\n\tResidual {res:.3e}\n\tNumber of sweeps {dev.sweeps}", flush=True)
    return x, res


def tt_restarted_block_amen(block_A, block_b, rank_restriction, op_tol, termination_tol=1e-3, eps=1e-11,
                            num_restarts=3, inner_m=10, x0=None, local_solver=None, verbose=False, _stats=None):
    """reference src/tt_als.py:744-825: warm-start retraction, first solve, global-residual check with
    leniency, rank-growing restarts; raises RuntimeError like the reference."""
    if x0 is not None:
        dim = len(x0)
        x0 = T.tt_rank_retraction(x0, [dim] * (dim - 1))

    def solve_als(rank, start, kick_rank):
        return tt_block_amen(block_A, block_b, termination_tol, r_max=rank, eps=eps, nswp=inner_m, x0=start,
                             local_solver=local_solver, kick_rank=kick_rank, amen=True, verbose=verbose, _stats=_stats)

    def residual_norm(x_cores):
        return (block_b - block_A.block_product(x_cores, 0.1 * op_tol)).norm

    orig_rhs_norm = block_b.norm
    if orig_rhs_norm < 0.5 * op_tol:
        raise RuntimeError(f"\n\tAbsolute tolerance already reached: {orig_rhs_norm:4f} < {op_tol:4f}")
    x_cores, res = solve_als(rank_restriction, x0, 2)
    if res < termination_tol:
        return x_cores, res
    rhs_norm = residual_norm(x_cores)
    if rhs_norm < termination_tol * orig_rhs_norm or rhs_norm < orig_rhs_norm:
        return x_cores, res
    for _ in range(1, num_restarts):
        dim = len(x_cores)
        x_cores = T.tt_rank_retraction(x_cores, [2 * dim] * (dim - 1))
        x_cores, res = solve_als(rank_restriction + 4, x_cores, 4)
        rhs_norm = residual_norm(x_cores)
        if rhs_norm < termination_tol * orig_rhs_norm or rhs_norm < orig_rhs_norm:
            return x_cores, res
    raise RuntimeError(f"\n\tNumber of restarts exhausted, Relative Error = {rhs_norm / orig_rhs_norm:3e}. "
                       "Consider increasing rank ceiling.")
