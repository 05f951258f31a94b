"""Same module name as reference src/tt_als.py for the Newton-system path (a1-a11, a25; 8f-1, 8f-2)."""
from ttipm_b200.tt_ops import *  # noqa: F401,F403
from ttipm_b200.tt_als import (TTBlockMatrix, TTBlockMatrixView, TTBlockVector, TTBlockVectorView,  # noqa: F401
                               _tt_get_block, compute_phi_bck_A, compute_phi_bck_rhs, compute_phi_fwd_A,
                               compute_phi_fwd_rhs, truncated_svd, tt_approx_mat_mat_mul,
                               tt_approx_mat_vec_mul, tt_block_amen, tt_mat_mat_mul, tt_mat_vec_mul,
                               tt_max_generalised_eigen, tt_min_eig, tt_restarted_block_amen)
from ttipm_b200.tt_ops import cached_einsum  # noqa: F401
