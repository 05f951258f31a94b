"""Swap the Newton-system hot path of an already imported reference tree for the B200 path.

    import src.tt_ipm                      # the UNMODIFIED reference driver
    import ttipm_b200.dropin as dropin
    dropin.install()                       # rebinds the hot-path names in every reference module

After install(), src.tt_ipm.tt_ipm and psd_system/*/create_problem drive the CUDA path: every module
whose namespace holds one of the reference's hot-path callables (the reference spreads them with
`from src.tt_ops import *`) gets that name rebound to the ttipm_b200 implementation of the same
signature (the step-size eigen sweeps of SURVEY 8f-1 included).  Everything else stays the reference's own code.
"""
import sys

from . import lgmres, tt, tt_als

TT_OPS = ["tt_add", "tt_sub", "tt_scale", "tt_inner_prod", "tt_norm", "tt_normalise", "tt_rank_reduce",
          "tt_psd_rank_reduce", "tt_mask_rank_reduce", "tt_rl_orthogonalise", "tt_fast_matrix_vec_mul",
          "tt_fast_mat_mat_mul", "tt_fast_hadamard", "tt_IkronM", "tt_MkronI", "tt_diag", "tt_diag_op",
          "tt_entrywise_sum", "tt_rank_retraction", "tt_rl_orthogonalise_py", "tt_sum", "prune_singular_vals",
          # shape / view helpers: the reference's Cython versions are typed `list` and reject the lazy TTList the
          # functions above return; these also keep device-resident trains on the device (tt_ranks, tt_reshape, tt_transpose)
          "tt_ranks", "tt_reshape", "tt_transpose", "tt_swap_all", "tt_diagonal", "tt_random_gaussian", "tt_merge_cores"]
TT_ALS = ["tt_restarted_block_amen", "tt_block_amen", "tt_mat_vec_mul", "tt_mat_mat_mul", "tt_approx_mat_vec_mul",
          "tt_approx_mat_mat_mul", "tt_max_generalised_eigen", "tt_min_eig",
          "compute_phi_bck_A", "compute_phi_fwd_A", "compute_phi_bck_rhs", "compute_phi_fwd_rhs"]
# The containers TTBlockMatrix / TTBlockVector are NOT rebound: in drop-in mode the reference's own classes stay (the IPM
# driver reaches into their _data / _aliases / _transposes dicts, src/tt_ipm.py:534-556); the device sweep only reads
# those three dicts, and the containers' own TT arithmetic (block_product, __sub__, norm) calls module-level functions
# that ARE rebound.  ttipm_b200.tt_als carries equivalent classes for the stand-alone layout (Option B).
LGMRES = ["MatVecWrapper", "IneqMatVecWrapper"]


def install(prefixes=("src", "cy_src", "psd_system", "refproblem_")):
    """Rebind the hot-path names in every loaded reference module; returns {module: [names]}."""
    table = {}
    for name in TT_OPS:
        table[name] = getattr(tt, name)
    for name in TT_ALS:
        table[name] = getattr(tt_als, name)
    for name in LGMRES:
        table[name] = getattr(lgmres, name)
    done = {}
    for modname, mod in list(sys.modules.items()):
        if mod is None or not any(modname == p or modname.startswith(p + ".") or
                                  (p.endswith("_") and modname.startswith(p)) for p in prefixes):
            continue
        if modname.startswith("ttipm_b200"):
            continue
        for name, impl in table.items():
            if hasattr(mod, name) and getattr(mod, name) is not impl:
                setattr(mod, name, impl)
                done.setdefault(modname, []).append(name)
    return done
