"""ctypes declarations for include/ttipm.h."""
import ctypes as C
import os

c_double_p = C.c_void_p   # device pointers are passed as integers
i64 = C.c_int64
i32 = C.c_int32

MAX_TERMS = 16


class Term(C.Structure):
    _fields_ = [("P1", C.c_void_p), ("A", C.c_void_p), ("P2", C.c_void_p),
                ("p1_strides", i64 * 3), ("a_strides", i64 * 4), ("p2_strides", i64 * 3),
                ("s", i32), ("S", i32), ("in_block", i32), ("out_block", i32), ("alpha", C.c_double)]


class PhiTerm(C.Structure):
    _fields_ = [("Phi", C.c_void_p), ("A", C.c_void_p), ("out", C.c_void_p), ("a_strides", i64 * 4),
                ("s", i32), ("S", i32)]


class RhsTerm(C.Structure):
    _fields_ = [("Xb1", C.c_void_p), ("B", C.c_void_p), ("Xb2", C.c_void_p), ("out", C.c_void_p),
                ("b", i32), ("Bp", i32)]


class EigOp(C.Structure):
    _fields_ = [("P1", C.c_void_p), ("A1", C.c_void_p), ("A2", C.c_void_p), ("P2", C.c_void_p),
                ("p1_strides", i64 * 3), ("a1_strides", i64 * 4), ("a2_strides", i64 * 4), ("p2_strides", i64 * 3),
                ("l", i32), ("s", i32), ("k", i32), ("S", i32), ("L", i32), ("n1", i32), ("n2", i32)]


SIGNATURES = {
    "ttipm_abi_version": (C.c_int, []),
    "ttipm_last_error": (C.c_char_p, []),
    "ttipm_use_pdl": (C.c_int, [C.c_int]),
    "ttipm_device_info": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ttipm_block_matvec": (C.c_int, [C.POINTER(Term), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, i64, i64, i64, i64, C.c_void_p, i64, i64, i64, i64, C.c_double,
                                     C.c_void_p, C.c_double, C.c_void_p, C.c_int, C.c_void_p]),
    "ttipm_local_diag": (C.c_int, [C.POINTER(Term), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ttipm_local_dense": (C.c_int, [C.POINTER(Term), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p]),
    "ttipm_phi_update": (C.c_int, [C.POINTER(PhiTerm), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ttipm_rhs_contract": (C.c_int, [C.POINTER(RhsTerm), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, i64,
                                     C.c_void_p]),
    "ttipm_lgmres_workspace": (i64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ttipm_local_lgmres": (C.c_int, [C.c_int] + [C.POINTER(Term)] * 6 + [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.c_void_p, i64, C.c_int, C.c_int, C.c_int, C.c_double,
                                     C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "ttipm_linalg_coop_min_dim": (C.c_int, [C.c_int]),
    "ttipm_qr_workspace": (i64, [C.c_int, C.c_int, C.c_int]),
    "ttipm_qr": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                           C.c_void_p]),
    "ttipm_svd_workspace": (i64, [C.c_int, C.c_int, C.c_int]),
    "ttipm_svd_left": (C.c_int, [C.c_void_p, i64, i64, i64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ttipm_permute4": (C.c_int, [C.c_void_p, C.POINTER(i32), C.POINTER(i32), C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.c_void_p]),
    "ttipm_block_norms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]),
    "ttipm_ewise": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_void_p, i64, C.c_double, C.c_void_p, i64, C.c_double,
                              C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, i64, C.c_void_p, C.c_void_p]),
    "ttipm_trunc_resnorms": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, i64, C.c_void_p, C.c_void_p]),
    "ttipm_block_diag": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p]),
    "ttipm_embed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ttipm_scale2d": (C.c_int, [C.c_void_p, i64, i64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                C.c_void_p]),
    "ttipm_amen_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ttipm_amen_destroy": (None, [C.c_void_p]),
    "ttipm_amen_set_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ttipm_amen_add_alias": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ttipm_amen_set_rhs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    "ttipm_amen_set_core": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ttipm_amen_run": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                 C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "ttipm_amen_set_profile": (C.c_int, [C.c_void_p, C.c_int]),
    "ttipm_amen_core_shape": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(i32)]),
    "ttipm_amen_get_core": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ttipm_amen_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "ttipm_amen_profile": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ttipm_amen_host_krylov": (C.c_int, [C.c_int]),
    "ttipm_lgmres_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(C.c_int)]),
    "ttipm_cgs_parts": (C.c_int, [C.c_int64]),
    "ttipm_cgs_project": (C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "ttipm_lincomb": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_int64,
                                C.c_void_p]),
    "ttipm_matvec_big_min_flops": (C.c_double, [C.c_double]),
    "ttipm_cgemm_force_ksplit": (C.c_int, [C.c_int]),
    "ttipm_cgemm_force_cfg": (C.c_int, [C.c_int]),
    "ttipm_cgemm_vector_loads": (C.c_int, [C.c_int]),
    "ttipm_linalg_noise_floor": (C.c_double, [C.c_double]),
    "ttipm_linalg_early_exit": (C.c_int, [C.c_int]),
    "ttipm_linalg_tall_triple_qr": (C.c_int, [C.c_int]),
    "ttipm_linalg_block_rows": (C.c_int, [C.c_int]),
    "ttipm_linalg_threads": (C.c_int, [C.c_int]),
    "ttipm_linalg_use_cluster": (C.c_int, [C.c_int]),
    "ttipm_tt_create": (C.c_void_p, [C.c_int, C.c_void_p]),
    "ttipm_tt_destroy": (None, [C.c_void_p]),
    "ttipm_tt_length": (C.c_int, [C.c_void_p]),
    "ttipm_tt_set_cores": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ttipm_tt_shapes": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ttipm_tt_get_cores": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ttipm_tt_clone": (C.c_void_p, [C.c_void_p]),
    "ttipm_tt_scale_core": (C.c_int, [C.c_void_p, C.c_int, C.c_double]),
    "ttipm_tt_rl_orthogonalise": (C.c_int, [C.c_void_p]),
    "ttipm_tt_round": (C.c_int, [C.c_void_p, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "ttipm_tt_add": (C.c_void_p, [C.c_void_p, C.c_void_p]),
    "ttipm_tt_inner": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]),
    "ttipm_tt_inner_chain": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "ttipm_tt_zipup": (C.c_void_p, [C.c_int, C.c_void_p, C.c_void_p, C.c_double]),
    "ttipm_tt_reshape": (C.c_void_p, [C.c_void_p, C.c_int, C.c_int]),
    "ttipm_tt_transpose": (C.c_void_p, [C.c_void_p]),
    "ttipm_tt_embed": (C.c_void_p, [C.c_void_p, C.c_int]),
    "ttipm_tt_counters": (C.c_int, [C.c_void_p, C.POINTER(i64), C.POINTER(i64)]),
    "ttipm_eig_assemble": (C.c_int, [C.POINTER(EigOp), C.c_int, C.c_void_p, C.c_void_p]),
    "ttipm_eig_workspace": (i64, [C.c_int, C.c_int]),
    "ttipm_eig_lanczos": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ttipm_eig_gen_largest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ttipm_eig_force_cluster": (C.c_int, [C.c_int]),
    "ttipm_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, i64, i64, i64, C.c_void_p, i64, i64,
                             i64, C.c_double, C.c_void_p, i64, i64, i64, C.c_int, C.c_void_p]),
}

DEFAULT_LIB = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "libttipm_b200.so"))


def load(path=None):
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build it with `python tensor-train-interior-point-method_b200/build.py` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ttipm_abi_version() != 1:
        raise RuntimeError("libttipm ABI version mismatch")
    return lib
