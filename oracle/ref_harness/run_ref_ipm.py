"""Harness (this container only): run the unmodified reference tt_ipm on a config
exactly as src/utils.py:245-309 (run_and_record) does, without its YAML rewriting.

  python oracle/ref_harness/run_ref_ipm.py maxcut 5 1 319 [--quiet]
"""
import os
import sys
import time

import numpy as np
import yaml

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_env  # noqa: E402


def load_config(problem, dim):
    with open(os.path.join(ref_env.REF, "configs", f"{problem}_{dim}.yaml")) as f:
        return yaml.safe_load(f)


def build_problem(ref, problem, dim, rank, seed):
    """np.random.seed(seed); create_problem(dim, rank) + the reshapes of src/utils.py:260-271."""
    np.random.seed(seed)
    prob = ref_env.create_problem(problem, dim, rank)
    T = ref.tt_ops
    if len(prob) == 5:
        obj_tt, L_op_tt, bias_tt, ineq_mask, lag_maps = prob
    else:
        obj_tt, L_op_tt, bias_tt, lag_y = prob
        ineq_mask = None
        lag_maps = {"y": lag_y}
    lag_maps = {k: T.tt_reshape(v, (4, 4)) for k, v in lag_maps.items()}
    obj_tt = T.tt_reshape(obj_tt, (4,))
    bias_tt = T.tt_reshape(bias_tt, (4,))
    return obj_tt, L_op_tt, bias_tt, ineq_mask, lag_maps


def run(problem, dim, rank, seed, verbose=True):
    ref = ref_env.load()
    cfg = load_config(problem, dim)
    obj_tt, L_op_tt, bias_tt, ineq_mask, lag_maps = build_problem(ref, problem, dim, rank, seed)
    T = ref.tt_ops
    t0 = time.time()
    X, Y, Tt, Z, info = ref.tt_ipm.tt_ipm(
        lag_maps, obj_tt, L_op_tt, bias_tt, ineq_mask=ineq_mask,
        max_iter=int(os.environ.get("TTIPM_MAX_ITER", cfg["max_iter"])),       # env: profiling runs stop early
        verbose=verbose, gap_tol=float(cfg["gap_tol"]),
        op_tol=float(cfg["op_tol"]), warm_up=cfg["warm_up"], abs_tol=float(cfg["abs_tol"]),
        aho_direction=False, mals_restarts=cfg["mals_restarts"],
        max_refinement=cfg["max_refinement"], lambdaStar=float(cfg.get("lambdaStar", 1)),
        lambdaStarIneq=float(cfg.get("lambdaStarIneq", 1)))
    wall = time.time() - t0
    gap = abs(T.tt_inner_prod(X, Z))
    pres = T.tt_rank_reduce(T.tt_sub(T.tt_fast_matrix_vec_mul(L_op_tt, T.tt_reshape(X, (4,))), bias_tt), eps=1e-12)
    dres = T.tt_rank_reduce(T.tt_sub(T.tt_fast_matrix_vec_mul(T.tt_transpose(L_op_tt), T.tt_reshape(Y, (4,)), eps=1e-12),
                                     T.tt_rank_reduce(T.tt_add(T.tt_reshape(Z, (4,)), obj_tt), eps=1e-12)), eps=1e-12)
    if info["status"].ineq_status is ref.tt_ipm.IneqStatus.ACTIVE:
        dres = T.tt_rank_reduce(T.tt_sub(dres, T.tt_reshape(Tt, (4,))), eps=1e-12)
    out = dict(problem=problem, dim=dim, rank=rank, seed=seed, iters=int(info["num_iters"]), wall_s=wall,
               gap=float(gap), primal_sq=float(T.tt_inner_prod(pres, pres)), dual_sq=float(T.tt_inner_prod(dres, dres)),
               ranksX=list(map(int, info["ranksX"])))
    return out


if __name__ == "__main__":
    problem, dim, rank, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    res = run(problem, dim, rank, seed, verbose="--quiet" not in sys.argv)
    from petsc4py import PETSc
    print("RESULT", res, PETSc.STATS)
