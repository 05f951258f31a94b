"""Host mirror of the reference's src/tt_ops.py (IPM-used subset) and cy_src/tt_ops_cy.pyx.

`from ttipm_b200.tt_ops import *` gives the driver the same names the reference's star import does,
including the implicit ones (np, scp, List, E, cached_einsum; reference src/tt_ipm.py uses scp.linalg
without importing it)."""
from typing import *  # noqa: F401,F403

import numpy as np
import scipy as scp  # noqa: F401
import scipy.linalg  # noqa: F401
import scipy.sparse  # noqa: F401
import scipy.sparse.linalg  # noqa: F401

from .als_product import add_kick_rank, symmetric_powers_of_two  # noqa: F401  (cy_src/tt_ops_cy.pyx:538-579)
from .generators import (skewed_probabilities, tt_kron, tt_matrix_to_matrix, tt_merge_bonds,  # noqa: F401
                         tt_random_binary_sym, tt_random_graph, tt_split_bonds, tt_to_tensor, tt_trace,
                         tt_tril_one_matrix, tt_triu_one_matrix, tt_vec_to_vec)
from .lgmres import IneqMatVecWrapper, MatVecWrapper  # noqa: F401
from .tt import (prune_singular_vals, tt_add, tt_diag, tt_diag_op, tt_diagonal, tt_entrywise_sum,  # noqa: F401
                 tt_fast_hadamard, tt_fast_mat_mat_mul, tt_fast_matrix_vec_mul, tt_identity, tt_IkronM,
                 tt_inner_prod, tt_mask_rank_reduce, tt_merge_cores, tt_MkronI, tt_norm, tt_normalise, tt_one_matrix,
                 tt_psd_rank_reduce, tt_random_gaussian, tt_rank_reduce, tt_rank_retraction, tt_ranks, tt_reshape,
                 tt_rl_orthogonalise, tt_rl_orthogonalise_py, tt_scale, tt_sub, tt_sum, tt_swap_all, tt_transpose,
                 tt_zero_matrix)


def E(i, j):
    """reference src/tt_ops.py:16-19."""
    out = np.zeros((1, 2, 2, 1))
    out[:, i, j] += 1
    return out


def cached_einsum(equation, *operands):
    """reference src/tt_ops.py:22-28.  Kept for the driver code that still contracts tiny host arrays
    with it; the hot-path contractions do not go through here."""
    return np.einsum(equation, *operands, optimize="greedy")
