"""Path-level parity (VERDICT r1 "close path-level parity"):

 * `ttipm_b200.tt_als.tt_restarted_block_amen` driven through the TTBlockMatrix / TTBlockVector containers with the
   reference's call convention on every traced KKT system that holds the REFERENCE's own output (`out_x`, `out_res`,
   `trace` written by oracle/ref_harness/make_golden.py amen from the unmodified src/tt_als.py), not the oracle re-run;
 * the restart ladder (reference src/tt_als.py:806-825) and the tiny right-hand-side RuntimeError (:781);
 * the Schur-reduced operators through `lgmres.MatVecWrapper` / `IneqMatVecWrapper`'s reference constructor signatures;
 * oracle/lgmres_ref.py against scipy.sparse.linalg.lgmres (an independent third implementation) on traced reduced
   operators (CPU tier).
"""
import glob
import os

import numpy as np
import pytest

import amen_cases as AC
import golden_io as G
import kernel_cases as KC
import rt_util
import tt_oracle as O

# the extra maxcut_13 seeds are bench inputs only (see tests/test_oracle_vs_golden.py)
WITH_REF_OUTPUT = [f for f in G.amen_files("amen_*.npz") if "out/x/0" in np.load(f).files and
                   not any(f"maxcut_13_r2_s{s}_" in f for s in (45, 23, 53, 12))]


def _containers(g):
    from ttipm_b200 import tt_als
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
    return A, b


def _solver_stub(ineq):
    # the device sweep dispatches on the callback's NAME (reference src/tt_ipm.py:958-981 passes these two functions)
    def _ipm_local_solver(*a, **k):
        raise AssertionError("host callback must not run")

    def _ipm_local_solver_ineq(*a, **k):
        raise AssertionError("host callback must not run")
    return _ipm_local_solver_ineq if ineq else _ipm_local_solver


def _restarted_vs_reference(rt, path, sol_tol=1e-5):
    from ttipm_b200 import tt_als, use_runtime
    g = G.load_amen(path)
    A, b = _containers(g)
    stats = {}
    np.random.set_state(g["rng_state"])
    with use_runtime(rt):
        x, res = tt_als.tt_restarted_block_amen(A, b, g["rank_restriction"], g["op_tol"], termination_tol=g["termination_tol"],
                                                eps=g["eps"], num_restarts=g["num_restarts"], inner_m=g["inner_m"],
                                                x0=[c.copy() for c in g["x0"]] if g["x0"] is not None else None,
                                                local_solver=_solver_stub(g["ineq"]), _stats=stats)
    bm = O.BlockMatrix(g["A"], g["aliases"], g["transposes"])
    ref = g["out_x"]
    if "amen_maxcut_13_r2_s83_2" in path:       # ranks inside the rounding-noise plateau (tests/test_oracle_vs_golden.py)
        a, b = np.array(O.tt_ranks(x), dtype=float), np.array(O.tt_ranks(ref), dtype=float)
        # (LAPACK vs LAPACK differs by up to 10 % there; the Jacobi kernel vs the reference's LAPACK is given 30 %)
        assert np.all(np.abs(a - b) <= np.maximum(3.0, 0.30 * b)), (a, b)
    else:
        assert O.tt_ranks(x) == O.tt_ranks(ref), (O.tt_ranks(x), O.tt_ranks(ref))
    nn = AC.block_inner(bm, ref)
    diff2 = nn - 2 * AC.block_inner(bm, x, ref) + AC.block_inner(bm, x)
    # norm through inner products: sqrt(eps) floor ~1e-8
    assert np.sqrt(max(diff2, 0.0) / nn) < sol_tol, (np.sqrt(max(diff2, 0.0) / nn), path)
    assert res <= max(10 * g["out_res"], 1e-9), (res, g["out_res"])
    # the reference's local-solve trace: same number of local solves, same residuals before / after every solve
    tr = np.asarray(stats.get("local_trace", []))
    if tr.size:
        want = g["trace"]
        assert tr.shape[0] == want.shape[0], (tr.shape, want.shape)
        assert np.max(np.abs(tr[:, 2:4] - want[:, 1:3])) < 1e-6 * max(1.0, np.max(np.abs(want[:, 1:3])))
    if g["d"] <= 8:                                    # dense blocks where they are small enough
        nb = O.block_size_of(bm)
        for i in range(nb):
            assert AC.rel(AC.dense_block(x, i), AC.dense_block(ref, i)) < 1e-6


def _restart_ladder(rt):
    """A doctored call (2 sweeps per solve, rank ceiling 2, unreachable tolerance) makes every global-residual check fail:
    first solve, then num_restarts - 1 restarts with rank_restriction + 4 / kick_rank 4 from the retracted iterate, then
    RuntimeError -- exactly as the oracle's restatement of src/tt_als.py:798-825 does on the same inputs."""
    from ttipm_b200 import tt_als, use_runtime
    g = G.load_amen(G.amen_files("amen_maxcut_5_r1_s319_1*")[0])
    A, b = _containers(g)
    bm, bv = O.BlockMatrix(g["A"], g["aliases"], g["transposes"]), O.BlockVector(g["b"])
    orig = O.block_norm(bv)
    calls = []
    real = O.tt_block_amen

    def counted(*a, **k):
        calls.append(k.get("r_max"))
        return real(*a, **k)
    O.tt_block_amen = counted
    try:
        np.random.seed(5)
        with pytest.raises(RuntimeError, match="restarts exhausted"):
            O.tt_restarted_block_amen(bm, bv, 2, 1.9 * orig, termination_tol=1e-30, eps=g["eps"], num_restarts=3,
                                      inner_m=2, x0=None, local_solver=O.local_solver_eq)
    finally:
        O.tt_block_amen = real
    assert calls == [2, 6, 6]
    stats = {"solves": []}
    np.random.seed(5)
    with use_runtime(rt):
        with pytest.raises(RuntimeError, match="restarts exhausted"):
            tt_als.tt_restarted_block_amen(A, b, 2, 1.9 * orig, termination_tol=1e-30, eps=g["eps"], num_restarts=3,
                                           inner_m=2, x0=None, local_solver=_solver_stub(False), _stats=stats)
        assert [s["r_max"] for s in stats["solves"]] == [2, 6, 6], stats["solves"]
        assert [s["kick_rank"] for s in stats["solves"]] == [2, 4, 4]
        # tiny right-hand side (reference :781)
        with pytest.raises(RuntimeError, match="Absolute tolerance already reached"):
            tt_als.tt_restarted_block_amen(A, b, 2, 2.5 * orig, local_solver=_solver_stub(False))
        # a call whose first solve already passes the lenient check returns without a restart
        stats = {"solves": []}
        np.random.seed(5)
        x, res = tt_als.tt_restarted_block_amen(A, b, 2, 1e-4, termination_tol=1e-30, eps=g["eps"], num_restarts=3,
                                                inner_m=4, x0=None, local_solver=_solver_stub(False), _stats=stats)
        assert len(stats["solves"]) == 1
    np.random.seed(5)
    xo, reso = O.tt_restarted_block_amen(bm, bv, 2, 1e-4, termination_tol=1e-30, eps=g["eps"], num_restarts=3, inner_m=4,
                                         x0=None, local_solver=O.local_solver_eq)
    assert O.tt_ranks(x) == O.tt_ranks(xo) and abs(res - reso) <= 1e-6 * max(reso, 1e-12)


def _wrappers(rt):
    """MatVecWrapper(XAX_k_00, ..., inv_I, r, n, R).matvec(flat) with the reference's positional constructor
    (cy_src/lgmres_cy.pyx:216-232, :392-414) against the output of the compiled reference class (kernels.npz red_y)."""
    from ttipm_b200 import lgmres, use_runtime
    with use_runtime(rt):
        for case in ("eq_small", "eq_mid", "ineq_small"):
            c = KC.load_blp_case(case)
            r, R = c["x"].shape[0], c["x"].shape[3]
            keys = [(0, 0), (0, 1), (2, 1), (2, 2)] + ([(3, 1), (3, 3)] if c["nb"] == 4 else [])
            args = [c["P1"][k] for k in keys] + [c["A"][k] for k in keys] + [c["P2"][k] for k in keys] + [c["inv_I"], r, 4, R]
            cls = lgmres.IneqMatVecWrapper if c["nb"] == 4 else lgmres.MatVecWrapper
            w = cls(*args)
            assert isinstance(w, lgmres.BaseMatVec)
            y = w.matvec(c["red_x"].copy())
            assert y.shape == c["red_y"].shape and KC.rel(y, c["red_y"]) < 1e-10, (case, KC.rel(y, c["red_y"]))
            # the device LGMRES behind the same object takes the same steps as the oracle LGMRES driven through
            # w.matvec (random, indefinite operands: 40 steps, compared iterate for iterate, converged or not)
            import lgmres_ref
            restart = min(c["red_y"].size, 100)
            sol = w.solve(c["red_y"], rtol=1e-10, restart=restart, outer_k=10, max_iter=40)
            want = lgmres_ref.lgmres(w.matvec, c["red_y"], rtol=1e-10, max_it=40, restart=restart, augment=10)
            assert KC.rel(sol, want.x) < 1e-7, (case, KC.rel(sol, want.x))


# ---- CPU tier --------------------------------------------------------------------------------------------------
def test_restart_ladder_emu():
    _restart_ladder(rt_util.emu_runtime())


@pytest.mark.parametrize("path", [f for f in WITH_REF_OUTPUT if "maxcut_5" in f], ids=os.path.basename)
def test_restarted_amen_vs_reference_output_emu(path):
    _restarted_vs_reference(rt_util.emu_runtime(), path)


def test_lgmres_oracle_vs_scipy():
    """oracle/lgmres_ref.py (restated from PETSc's published LGMRES; parity with PETSc itself is unpinned) against
    scipy.sparse.linalg.lgmres on the traced reduced local operators: both reach the requested relative residual and
    agree on the solution to the accuracy that residual allows; at a tight tolerance the solutions coincide."""
    import scipy.sparse.linalg as spla
    import lgmres_ref
    files = sorted(glob.glob(os.path.join(G.GOLD, "local_*.npz")))
    assert files
    checked = 0
    for f in files:
        z = np.load(f)
        ineq = bool(int(z["ineq"]))
        P1, A, P2 = G.keyed(z, "P1"), G.keyed(z, "A"), G.keyed(z, "P2")
        rhs = z["rhs"]
        r, nb, n, R = rhs.shape
        inv_I = 1.0 / O.local_diag(P1[1, 2], A[1, 2], P2[1, 2])
        op = (O.ReducedOperatorIneq if ineq else O.ReducedOperatorEq)(P1, A, P2, inv_I)
        nred = 3 if ineq else 2
        lrhs = np.empty((nred, r, n, R))
        lrhs[0] = rhs[:, 0]
        lrhs[1] = rhs[:, 2] - O.local_matvec(P1[2, 2], A[2, 2], P2[2, 2], inv_I * rhs[:, 1])
        if ineq:
            lrhs[2] = rhs[:, 3]
        bvec = lrhs.reshape(-1)
        N = bvec.size
        restart = min(N // nred, 100)
        aug = max(restart // 10, 3)
        lo = spla.LinearOperator((N, N), matvec=lambda v: op.matvec(np.asarray(v).reshape(-1)), dtype=np.float64)
        for rtol, agree, slack in ((1e-5, 5e-3, 1.05), (1e-11, 1e-6, 20.0)):       # slack: true vs recurrence residual
            mine = lgmres_ref.lgmres(op.matvec, bvec, rtol=rtol, max_it=5000, restart=restart, augment=aug)
            theirs, info = spla.lgmres(lo, bvec, rtol=rtol, atol=0.0, inner_m=restart - aug, outer_k=aug, maxiter=2000)
            assert info == 0
            nb_ = np.linalg.norm(bvec)
            assert np.linalg.norm(op.matvec(mine.x) - bvec) <= slack * rtol * nb_, \
                (f, rtol, mine.reason, np.linalg.norm(op.matvec(mine.x) - bvec) / nb_)
            assert np.linalg.norm(op.matvec(theirs) - bvec) <= slack * rtol * nb_
            assert np.linalg.norm(mine.x - theirs) <= agree * np.linalg.norm(theirs), \
                (f, rtol, np.linalg.norm(mine.x - theirs) / np.linalg.norm(theirs))
        checked += 1
    assert checked >= 10


# ---- GPU tier --------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("path", WITH_REF_OUTPUT, ids=os.path.basename)
def test_restarted_amen_vs_reference_output_gpu(path):
    _restarted_vs_reference(rt_util.cuda_runtime(), path)


def _with_host_krylov(rt, fn):
    """run fn() with EVERY Krylov solve of the native driver host-driven (ttipm_amen_host_krylov(2)): block matvec through
    ttipm_block_matvec, Gram-Schmidt / combinations through csrc/krylov_ops.cu, Hessenberg matrix on the host"""
    old = rt.lib.ttipm_amen_host_krylov(2)
    try:
        return fn()
    finally:
        rt.lib.ttipm_amen_host_krylov(old)


@pytest.mark.gpu
@pytest.mark.parametrize("path", WITH_REF_OUTPUT, ids=os.path.basename)
def test_restarted_amen_host_krylov_vs_reference_output_gpu(path):
    """The host-driven LGMRES (large local blocks; forced for every block here) against the REFERENCE's stored outputs:
    same ranks, same solution, same residual before / after every local solve as the PETSc-backed reference run."""
    rt = rt_util.cuda_runtime()
    _with_host_krylov(rt, lambda: _restarted_vs_reference(rt, path))


def test_block_amen_host_krylov_vs_oracle_emu():
    """CPU tier of the same: an inequality system whose local solves are Krylov solves, kernels on the emulator."""
    rt = rt_util.emu_runtime()
    out = _with_host_krylov(rt, lambda: AC.run_block_amen(rt, G.amen_files("amen_corr_clust_8_r1_s208_7*")[0],
                                                          use_oracle=True, native=True))
    assert out["sweeps_dev"] == out["sweeps_oracle"] and out["solves_dev"] == out["solves_oracle"], out
    assert out["ranks_dev"] == out["ranks_oracle"], out
    assert out["trace_absdiff"] < 1e-9 and out["sol_rel_diff"] < 1e-6, out


@pytest.mark.gpu
def test_restart_ladder_gpu():
    _restart_ladder(rt_util.cuda_runtime())


@pytest.mark.gpu
def test_matvec_wrappers_reference_signature_gpu():
    _wrappers(rt_util.cuda_runtime())


def test_matvec_wrappers_reference_signature_emu():
    _wrappers(rt_util.emu_runtime())


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_graphm3_iteration4_predictor_system_returns():
    """The KKT system of graphm_3 rank 2 (seed 256) at IPM iteration 4, dumped from the end-to-end run of the unmodified
    reference driver on the B200 (oracle/ref_harness/run_dropin_ipm.py --dump-latest): solution ranks up to 129, KKT
    operator ranks up to 36, local blocks of (r, R) = (150, 16).  There the matvec plan of the Krylov kernel filled the
    shared memory to the last KB and lg_setup() looped forever on the host (the end-to-end run never came back from this
    call).  No reference output exists for this system (the CPU reference needs hours to get here): the test pins that
    the call returns, in the sweep budget, with a solution whose GLOBAL residual -- evaluated independently through the
    containers' block_product -- is at the tolerance."""
    from ttipm_b200 import tt_als, use_runtime
    rt = rt_util.cuda_runtime()
    g = G.load_amen(os.path.join(os.path.dirname(G.amen_files("amen_*")[0]), "big_graphm_3_r2_s256_it4_predictor.npz"))
    A, b = _containers(g)
    np.random.set_state(g["rng_state"])
    stats = {}
    with use_runtime(rt):
        x, res = tt_als.tt_restarted_block_amen(A, b, g["rank_restriction"], g["op_tol"], termination_tol=g["termination_tol"],
                                                eps=g["eps"], num_restarts=g["num_restarts"], inner_m=g["inner_m"],
                                                x0=[c.copy() for c in g["x0"]], local_solver=_solver_stub(g["ineq"]),
                                                _stats=stats)
    assert res < g["termination_tol"], res
    assert max(c.shape[-1] for c in x) <= 200, [c.shape for c in x]
