"""Host-side problem-generator helpers of the reference's src/tt_ops.py that psd_system/*/create_problem imports
next to the hot-path names (SURVEY 8b "names the drivers actually use", 8f-4): graph sampler, bond splitting,
triangular all-one matrices, TT -> dense conversion.  These run once per problem instance on KB-sized cores, so they
stay NumPy on the host; everything they call that IS on the hot path (tt_rank_reduce, tt_norm, tt_reshape) is the
device implementation of ttipm_b200.tt.

The sampler must consume the global NumPy RNG in exactly the reference's order (the problem instance is a function
of `np.random.seed(seed)`, reference src/utils.py:260), which fixes the order of the draws below.
"""
import numpy as np
import scipy.linalg

from . import tt as T


def _unit(i, j):
    e = np.zeros((1, 2, 2, 1))
    e[0, i, j, 0] = 1.0
    return e


# ---------------------------------------------------------------------------------------------------------------
# dense conversions (reference src/tt_ops.py:192-216, 227-229)
# ---------------------------------------------------------------------------------------------------------------
def tt_to_tensor(tt_train):
    """Full tensor of a train, boundary ranks summed out (reference src/tt_ops.py:192-196)."""
    full = tt_train[0]
    for core in tt_train[1:]:
        full = np.tensordot(full, core, axes=(-1, 0))
    return full.sum(axis=(0, -1))


def tt_matrix_to_matrix(matrix_tt):
    """Dense matrix of a TT matrix with cores (r, m, n, R): row modes first (reference src/tt_ops.py:211-217)."""
    if len(matrix_tt) == 1:
        return np.squeeze(matrix_tt)
    full = tt_to_tensor(matrix_tt)
    nd = full.ndim
    order = list(range(0, nd - 1, 2)) + list(range(1, nd, 2))
    rows = int(np.prod(full.shape[:nd // 2]))
    return full.transpose(order).reshape(rows, -1)


def tt_vec_to_vec(vec_tt):
    return tt_to_tensor(vec_tt).reshape(-1, 1)


def tt_trace(matrix_tt):
    """reference src/tt_ops.py:206-208."""
    return T.tt_inner_prod(matrix_tt, T.tt_identity(len(matrix_tt)))


def tt_kron(matrix_tt_1, matrix_tt_2):
    """Core-wise Kronecker product (reference src/tt_ops.py:199-203)."""
    out = []
    for a, b in zip(matrix_tt_1, matrix_tt_2):
        c = np.einsum("rmnR,lijL->rlminjRL", a, b)
        out.append(c.reshape(a.shape[0] * b.shape[0], a.shape[1] * b.shape[1], a.shape[2] * b.shape[2],
                             a.shape[3] * b.shape[3]))
    return out


# ---------------------------------------------------------------------------------------------------------------
# bond splitting / merging (reference src/tt_ops.py:247-270)
# ---------------------------------------------------------------------------------------------------------------
def _split_core(core, err_bound=1e-18):
    """One core (r, m, n, R) -> (r, m, q), (q, n, R) by an SVD of the (r m) x (n R) unfolding; singular values
    <= err_bound dropped, at least one kept (reference src/tt_ops.py:247-262)."""
    shp = core.shape
    half = len(shp) // 2
    unfolding = core.reshape(int(np.prod(shp[:half])), -1)
    u, s, vt = scipy.linalg.svd(unfolding, full_matrices=False, check_finite=False)
    keep = np.flatnonzero(np.abs(s) > err_bound)
    if keep.size == 0:
        keep = np.array([0])
    q = keep.size
    left = u[:, keep].reshape(*shp[:half], q)
    right = (np.diag(s[keep]) @ vt[keep, :]).reshape(q, *shp[half:])
    return [left, right]


def tt_split_bonds(matrix_tt):
    out = []
    for core in matrix_tt:
        out.extend(_split_core(core))
    return out


def tt_merge_bonds(vec_tt):
    """Inverse of tt_split_bonds: neighbouring 3-D cores pairwise into 4-D cores (reference src/tt_ops.py:268-270)."""
    return [np.einsum("abc,cde->abde", a, b) for a, b in zip(vec_tt[:-1:2], vec_tt[1::2])]


# ---------------------------------------------------------------------------------------------------------------
# lower / upper triangular all-one matrices of size 2^dim (reference src/tt_ops.py:377-395)
# ---------------------------------------------------------------------------------------------------------------
def _tri_one(dim, i, j):
    """Rank-2 TT of the triangular all-one matrix whose strict part lies at 2x2 position (i, j)."""
    if dim == 1:
        m = np.eye(2)
        m[i, j] = 1.0
        return [m.reshape(1, 2, 2, 1)]
    ones, zeros = np.ones((1, 2, 2, 1)), np.zeros((1, 2, 2, 1))
    strict, eye = _unit(i, j), _unit(0, 0) + _unit(1, 1)
    first = np.concatenate((strict, eye), axis=-1)
    mid_top = np.concatenate((ones, strict), axis=0)
    mid_bot = np.concatenate((zeros, eye), axis=0)
    last = np.concatenate((ones, strict + eye), axis=0)
    return [first] + [np.concatenate((mid_top, mid_bot), axis=-1) for _ in range(dim - 2)] + [last]


def tt_tril_one_matrix(dim):
    return _tri_one(dim, 1, 0)


def tt_triu_one_matrix(dim):
    return _tri_one(dim, 0, 1)


# ---------------------------------------------------------------------------------------------------------------
# random graph sampler (reference src/tt_ops.py:398-520); RNG draw order is part of the contract
# ---------------------------------------------------------------------------------------------------------------
def skewed_probabilities(n, skew=0.0):
    w = np.exp(-skew * np.linspace(0, 1, n))
    return w / w.sum()


class _BasisSampler:
    """Orthonormal basis rows 1..rank (row 0 = zero vector) with skewed selection probabilities."""

    def __init__(self, rank, skew):
        q, _ = np.linalg.qr(np.random.randn(rank, rank), mode="reduced")
        self.vec = np.vstack((np.zeros(rank), q.T))
        self.size = rank + 1
        self.prob = skewed_probabilities(self.size, skew)

    def _redirect(self, proj, i, j):
        proj += np.outer(self.vec[i], self.vec[j] - self.vec[i])

    def off_diagonal(self):
        """reference _random_projector (src/tt_ops.py:438-452): randint, choice without replacement, skewed choice."""
        n = self.size
        count = np.random.randint(n)
        src = np.random.choice(n, size=count, replace=False)
        dst = np.random.choice(n, size=count, replace=True, p=self.prob)
        proj = np.eye(n - 1)
        for i, j in zip(src, dst):
            self._redirect(proj, i, j)
        return proj

    def diagonal_pair(self, discarded, limit):
        """reference _diag_projector (src/tt_ops.py:405-436): two projectors sharing their sources; a discarded source
        is only redirected (to non-zero targets) while the discarded set is small or a target is already discarded."""
        n = self.size
        count = np.random.randint(n) if n > 0 else 0
        src = np.random.choice(n, size=count, replace=False)
        dst1 = np.random.choice(n, size=count, replace=True, p=self.prob)
        dst2 = np.random.choice(n, size=count, replace=True, p=self.prob)
        p1, p2 = np.eye(n - 1), np.eye(n - 1)
        new_discarded = set(discarded)
        for i, j1, j2 in zip(src, dst1, dst2):
            guarded = i in discarded and j1 != 0 and j2 != 0
            if guarded:
                if not (len(new_discarded) <= limit or j1 in discarded or j2 in discarded):
                    continue
                new_discarded.discard(i)
                new_discarded.update((j1, j2))
            self._redirect(p1, i, j1)
            self._redirect(p2, i, j2)
        return p1, p2, new_discarded


def tt_random_binary_sym(dim, rank, skew=5.0):
    """Random symmetric 0/1-structured TT with vectorised 2x2 modes (cores (r, 4, R)); reference src/tt_ops.py:455-502."""
    if rank <= 0:
        return []
    bs = _BasisSampler(rank, skew)
    pick = np.random.choice(bs.size, size=3, replace=True, p=bs.prob)
    first = np.zeros((1, 4, rank))
    first[0] = bs.vec[[pick[0], pick[1], pick[1], pick[2]]]
    discarded = {int(p) for p in (pick[0], pick[2]) if p != 0}
    cores = [first]
    if dim <= 1:
        return cores
    for _ in range(dim - 2):
        core = np.empty((rank, 4, rank))
        off = bs.off_diagonal()
        core[:, 1, :] = off
        core[:, 0, :], core[:, 3, :], discarded = bs.diagonal_pair(discarded, limit=rank)
        core[:, 2, :] = off
        cores.append(core)
    allowed = sorted(set(range(bs.size)) - discarded)
    p_allowed = bs.prob[allowed] / sum(bs.prob[allowed])
    ends = np.random.choice(allowed, size=2, replace=True, p=p_allowed)
    mid = np.random.choice(bs.size, size=1, replace=True, p=bs.prob)
    last = np.zeros((rank, 4, 1))
    last[:, :, 0] = bs.vec[[ends[0], mid[0], mid[0], ends[1]]].T
    cores.append(last)
    return cores


def _host_norm(train):
    """Norm of a small train on the host with the reference's contraction order (cy_src/tt_ops_cy.pyx:506-520,
    src/tt_ops.py:306-310).  The sampler's rejection test `norm > 1e-12` has to see an exactly cancelling sample as 0,
    which only the same summation order guarantees (a device inner product returns ~1e-8 of rounding noise there)."""
    acc = np.ones((1, 1))
    for core in train:
        lead = list(range(core.ndim - 1))
        acc = np.tensordot(np.tensordot(acc, core, axes=([0], [0])), core, axes=(lead, lead))
    v = acc[0, 0]
    return float(np.sqrt(v)) if v > 0 else 0.0


def tt_random_graph(dim, r, skew=-1.0, eps=1e-12):
    """Up to 999 rejection draws of tt_random_binary_sym(dim, 2 r), keeping the rounded sample whose maximal TT rank is
    the largest one <= r; stops at rank r (reference src/tt_ops.py:505-520; prints like the reference)."""
    best_rank, best = 0, None
    for _ in range(1, 1000):
        sample = tt_random_binary_sym(dim, 2 * r, skew=skew)
        if _host_norm(sample) > 1e-12:
            sample = T.tt_rank_reduce(T.tt_reshape(sample, (2, 2)), 1e-12)
            top = np.max(T.tt_ranks(sample))
            if best_rank <= top <= r:
                best_rank, best = top, sample
            if best_rank == r:
                break
    else:
        best = [np.array([[0.0, 1.0], [1.0, 0.0]]).reshape(1, 2, 2, 1) for _ in range(dim)]
    print("===Terminated Graph Sampling=== rank: ", T.tt_ranks(best), flush=True)
    return best
