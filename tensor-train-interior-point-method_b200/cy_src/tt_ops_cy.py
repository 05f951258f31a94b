"""Same module name as reference cy_src/tt_ops_cy.pyx."""
from ttipm_b200.als_product import add_kick_rank, symmetric_powers_of_two  # noqa: F401
from ttipm_b200.tt import (prune_singular_vals, tt_add, tt_fast_hadamard, tt_fast_mat_mat_mul,  # noqa: F401
                           tt_fast_matrix_vec_mul, tt_identity, tt_inner_prod, tt_mask_rank_reduce, tt_normalise,
                           tt_one_matrix, tt_psd_rank_reduce, tt_random_gaussian, tt_rank_reduce, tt_ranks,
                           tt_rl_orthogonalise, tt_scale, tt_swap_all, tt_transpose, tt_zero_matrix)
