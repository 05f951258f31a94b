"""CPU tier: the whole hot path (native C++ sweep driver + the real kernel sources under -DTTIPM_EMU) on the
smallest traced KKT systems vs the oracle, and the host mirror's containers / dispatch logic."""
import os
import numpy as np
import pytest

import golden_io as G
import rt_util
import amen_cases as AC
import tt_oracle as O


@pytest.fixture(scope="module")
def rt():
    return rt_util.emu_runtime()


@pytest.mark.parametrize("native", [True, False], ids=["native", "python"])
def test_block_amen_maxcut5(rt, native):
    out = AC.run_block_amen(rt, G.amen_files("amen_maxcut_5_r1_s319_1*")[0], native=native)
    assert out["sweeps_dev"] == out["sweeps_oracle"] and out["solves_dev"] == out["solves_oracle"], out
    assert out["ranks_dev"] == out["ranks_oracle"], out
    assert out["sol_rel_diff"] < 1e-7 and out["trace_absdiff"] < 1e-9, out   # norm via inner products: sqrt(eps) floor


def test_restarted_block_amen_through_mirror(rt):
    """ttipm_b200.tt_als.tt_restarted_block_amen with the reference's call convention (local_solver callback
    dispatched by name, NumPy in / NumPy out, warm-start retraction) vs the reference's own output."""
    from ttipm_b200 import tt_als, use_runtime
    g = G.load_amen(G.amen_files("amen_maxcut_5_r1_s319_6*")[0])
    A = tt_als.TTBlockMatrix()
    for key, cores in g["A"].items():
        A[key] = [c.copy() for c in cores]
    for k1, k2 in g["aliases"].items():
        A.add_alias(k1, k2)
    for k1, k2 in g["transposes"].items():
        A.add_alias(k1, k2, is_transpose=True)
    b = tt_als.TTBlockVector()
    for i, cores in g["b"].items():
        b[i] = [c.copy() for c in cores]

    def _ipm_local_solver(*a, **k):     # never called: the device sweep dispatches on the name
        raise AssertionError("host callback must not run")

    np.random.set_state(g["rng_state"])
    with use_runtime(rt):
        x, res = tt_als.tt_restarted_block_amen(A, b, g["rank_restriction"], g["op_tol"], termination_tol=g["termination_tol"],
                                                eps=g["eps"], num_restarts=g["num_restarts"], inner_m=g["inner_m"],
                                                x0=[c.copy() for c in g["x0"]], local_solver=_ipm_local_solver)
    assert O.tt_ranks(x) == O.tt_ranks(g["out_x"])
    for i in range(3):
        assert AC.rel(AC.dense_block(x, i), AC.dense_block(g["out_x"], i)) < 1e-6
    assert res <= max(10 * g["out_res"], 1e-10)


def test_container_semantics():
    from ttipm_b200 import tt_als
    A = tt_als.TTBlockMatrix()
    assert A[0, 1] == [] and (0, 1) in A.keys()              # __getitem__ has setdefault semantics (src/tt_als.py:100-102)
    A[2, 1] = [np.zeros((1, 4, 4, 1))]
    A[3, 3] = [np.zeros((1, 4, 4, 1))]
    A.add_alias((0, 1), (1, 0), is_transpose=True)
    A.add_alias((1, 2), (1, 3))
    assert A.tkeys() == {(0, 1), (2, 1), (3, 3), (1, 0)}
    assert A.akeys() == {(0, 1), (2, 1), (3, 3), (1, 3)}
    sub = A.get_submatrix(2, 2)
    assert (3, 3) not in sub.keys() and sub._aliases == {} and sub._transposes == {(0, 1): (1, 0)}
    with pytest.raises(KeyError):
        A["x"]
    b = tt_als.TTBlockVector()
    with pytest.raises(ValueError):
        b[0] = np.zeros(3)
    b[1] = [np.ones((1, 4, 1))]
    assert b.get_row(0) is None and list(b.keys()) == [1]
    def _ipm_local_solver_ineq():
        pass

    def _ipm_local_solver():
        pass
    assert tt_als._solver_kind(_ipm_local_solver_ineq, A) is True
    assert tt_als._solver_kind(_ipm_local_solver, A) is False
    with pytest.raises(NotImplementedError):
        tt_als._solver_kind(None, A)


def test_restarted_raises_on_tiny_rhs(rt):
    from ttipm_b200 import tt_als, use_runtime
    A = tt_als.TTBlockMatrix()
    A[0, 0] = [np.eye(4).reshape(1, 4, 4, 1)] * 2
    b = tt_als.TTBlockVector()
    b[0] = [1e-9 * np.ones((1, 4, 1))] * 2

    def _ipm_local_solver():
        pass
    with use_runtime(rt), pytest.raises(RuntimeError):
        tt_als.tt_restarted_block_amen(A, b, 10, 1e-4, local_solver=_ipm_local_solver)


def test_dropin_rebinds_only_hot_path_names():
    """dropin.install() swaps the hot-path callables (SURVEY 8b names, incl. the ALS product pair) inside modules of a
    reference-like tree (the step-size eigen sweeps of 8f-1 included) and leaves everything else alone"""
    import sys
    import types
    from ttipm_b200 import dropin, tt, tt_als, lgmres
    mod = types.ModuleType("refproblem_fake")
    sentinel = lambda *a, **k: None
    for name in ("tt_add", "tt_rank_reduce", "tt_mat_vec_mul", "tt_approx_mat_vec_mul", "tt_approx_mat_mat_mul",
                 "tt_restarted_block_amen", "TTBlockMatrix", "MatVecWrapper", "tt_min_eig", "tt_max_generalised_eigen",
                 "create_problem"):
        setattr(mod, name, sentinel)
    sys.modules["refproblem_fake"] = mod
    try:
        done = dropin.install(prefixes=("refproblem_",))
        assert sorted(done["refproblem_fake"]) == sorted(
            ["tt_add", "tt_rank_reduce", "tt_mat_vec_mul", "tt_approx_mat_vec_mul", "tt_approx_mat_mat_mul",
             "tt_restarted_block_amen", "MatVecWrapper", "tt_min_eig", "tt_max_generalised_eigen"])
        assert mod.tt_add is tt.tt_add and mod.tt_approx_mat_vec_mul is tt_als.tt_approx_mat_vec_mul
        assert mod.TTBlockMatrix is sentinel and mod.MatVecWrapper is lgmres.MatVecWrapper     # containers stay the reference's
        assert mod.tt_min_eig is tt_als.tt_min_eig and mod.tt_max_generalised_eigen is tt_als.tt_max_generalised_eigen
        assert mod.create_problem is sentinel
        assert dropin.install(prefixes=("refproblem_",)).get("refproblem_fake") is None       # idempotent
    finally:
        del sys.modules["refproblem_fake"]


def test_krylov_setup_grid_search_terminates():
    """lg_setup's grid search on a 148-SM device (TTIPM_EMU_SMS=148 in a fresh process; planning needs no device): before the
    fix of round 2 / session 3 a matvec plan that filled the shared memory to the last KB made the search alternate between
    G = items and 2 G forever on the HOST -- graphm_3, IPM iteration 4, local block (r, R) = (150, 16), operator ranks 30,
    restart 100.  Every shape of that solve (and a few around it) must come back at once with a plan inside the 227 KB or
    with error 4."""
    import subprocess
    import sys
    code = r'''
import ctypes as C, os, sys
sys.path[:0] = [os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
import build_emu
lib = C.CDLL(build_emu.build())
lib.ttipm_lgmres_plan.argtypes = [C.c_int] * 6 + [C.POINTER(C.c_int)] * 2
for ineq in (0, 1):
    for (r, R, s) in ((150, 16, 30), (130, 16, 30), (188, 16, 30), (61, 64, 36), (16, 125, 30), (100, 108, 5), (8, 8, 3),
                      (150, 16, 36), (400, 4, 30)):
        for restart in (4, 100):
            g, b = C.c_int(0), C.c_int(0)
            rc = lib.ttipm_lgmres_plan(ineq, r, R, 4, s, restart, C.byref(g), C.byref(b))
            assert rc in (0, 4), (r, R, s, restart, rc)
            if rc == 0:
                assert 1 <= g.value <= 148 and 0 < b.value <= 227 * 1024, (r, R, s, restart, g.value, b.value)
print("ok")
'''.replace("ROOT", repr(os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))))
    env = dict(os.environ, TTIPM_EMU_SMS="148")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip().endswith("ok"), out.stdout + out.stderr
