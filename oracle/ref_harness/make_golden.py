"""Harness (this container only): generate tests/golden/*.npz from the UNMODIFIED
reference (imported from /root/reference through ref_env, with the stand-ins of
oracle/ref_harness/standins).  The fixtures travel to the GPU box; the reference
does not.

  python oracle/ref_harness/make_golden.py kernels
  python oracle/ref_harness/make_golden.py amen maxcut 5 1 319 0,1,5 [--budget 600]
  python oracle/ref_harness/make_golden.py e2e maxcut 5 1 319
  python oracle/ref_harness/make_golden.py als
  python oracle/ref_harness/make_golden.py eigen maxcut 5 1 319 0,1,4,9

File formats (all float64 unless noted):
  kernels_*.npz : flat dict "case/<name>/in_*", ".../out_*"
  amen_<cfg>_<n>.npz : one traced call of tt_restarted_block_amen
      A/<i><j>/<k>        operator cores, keys of lhs._data
      aliases, transposes int arrays (n,4): i,j,k,t
      b/<i>/<k>           rhs cores
      x0/<k>              warm start (absent if None)
      rng_keys, rng_pos   NumPy global RNG state at entry (MT19937)
      args                [rank_restriction, op_tol, termination_tol, eps, num_restarts, inner_m, ineq]
      out/x/<k>, out_res  reference result
      trace               (n,5): swp,k,res_old,res_new,r*R per local solve
"""
import copy
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_env  # noqa: E402
from run_ref_ipm import build_problem, load_config  # noqa: E402

GOLD = os.environ.get("TTIPM_GOLD_DIR") or os.path.abspath(os.path.join(HERE, "..", "..", "tests", "golden"))


def _rand_tt(rng, ranks, mode):
    rr = [1] + list(ranks) + [1]
    return [rng.standard_normal((a, *mode, b)) for a, b in zip(rr[:-1], rr[1:])]


def _cp(tt):
    return [np.array(c, copy=True) for c in tt]


def _put_tt(out, prefix, tt):
    out[prefix + "/n"] = np.array(len(tt))
    for k, c in enumerate(tt):
        out[f"{prefix}/{k}"] = np.ascontiguousarray(c)


def kernels():
    ref = ref_env.load()
    T, ALS, CY, LG = ref.tt_ops, ref.tt_als, ref.tt_ops_cy, ref.lgmres_cy
    rng = np.random.default_rng(20251018)
    out = {}

    # ---- K1/K2/K3 contractions through the reference containers -----------------
    for name, (r, R, nb, ranks) in {
        "eq_small": (4, 5, 3, {(0, 0): (2, 2), (0, 1): (1, 1), (1, 2): (1, 1), (2, 1): (4, 3), (2, 2): (3, 4)}),
        "ineq_small": (4, 6, 4, {(0, 0): (2, 3), (0, 1): (2, 1), (1, 2): (1, 1), (2, 1): (3, 3), (2, 2): (4, 4),
                                 (3, 1): (1, 1), (3, 3): (2, 2)}),
        "eq_mid": (13, 9, 3, {(0, 0): (3, 3), (0, 1): (2, 2), (1, 2): (1, 1), (2, 1): (5, 6), (2, 2): (6, 4)}),
    }.items():
        bm = ALS.TTBlockMatrix()
        XAX_k, XAX_k1 = {}, {}
        for key, (s, S) in ranks.items():
            bm[key] = [rng.standard_normal((s, 4, 4, S))]
            XAX_k[key] = rng.standard_normal((r, s, r))
            XAX_k1[key] = rng.standard_normal((R, S, R))
        bm.add_alias((0, 1), (1, 0), is_transpose=True)
        if nb == 4:
            bm.add_alias((1, 2), (1, 3))
        x = rng.standard_normal((r, nb, 4, R))
        view = bm[0]
        p = f"blp/{name}"
        out[p + "/x"] = x
        out[p + "/nb"] = np.array(nb)
        for key in ranks:
            tag = f"{key[0]}{key[1]}"
            out[f"{p}/A/{tag}"] = bm[key][0]
            out[f"{p}/P1/{tag}"] = XAX_k[key]
            out[f"{p}/P2/{tag}"] = XAX_k1[key]
        out[p + "/y"] = view.block_local_product(XAX_k, XAX_k1, x)
        # compressed variants with residual interfaces of rank rz
        rz, rz1 = 2, 3
        tkeys = list(ranks.keys()) + [(1, 0)]
        ZAX_k = {key: rng.standard_normal((rz, (ranks[key] if key in ranks else ranks[0, 1])[0], r)) for key in tkeys}
        ZAX_k1 = {key: rng.standard_normal((rz1, (ranks[key] if key in ranks else ranks[0, 1])[1], R)) for key in tkeys}
        for key in tkeys:
            tag = f"{key[0]}{key[1]}"
            out[f"{p}/Z1/{tag}"] = ZAX_k[key]
            out[f"{p}/Z2/{tag}"] = ZAX_k1[key]
        out[p + "/y_cc"] = view.compressed_block_local_product(ZAX_k, ZAX_k1, x, (rz, nb, 4, rz1))
        out[p + "/y_lc"] = view.lcompressed_block_local_product(ZAX_k, XAX_k1, x, (rz, nb, 4, R))
        out[p + "/y_rc"] = view.rcompressed_block_local_product(XAX_k, ZAX_k1, x, (r, nb, 4, rz1))
        # interface updates
        left = rng.standard_normal((r, 4, R))
        zleft = rng.standard_normal((rz, 4, rz1))
        out[p + "/left"] = left
        out[p + "/zleft"] = zleft
        for key in ranks:
            tag = f"{key[0]}{key[1]}"
            out[f"{p}/phi_bck/{tag}"] = ALS.compute_phi_bck_A(XAX_k1[key], left, bm[key][0], left)
            out[f"{p}/phi_fwd/{tag}"] = ALS.compute_phi_fwd_A(XAX_k[key], left, bm[key][0], left)
            out[f"{p}/zphi_bck/{tag}"] = ALS.compute_phi_bck_A(ZAX_k1[key], zleft, bm[key][0], left)
            out[f"{p}/zphi_fwd/{tag}"] = ALS.compute_phi_fwd_A(ZAX_k[key], zleft, bm[key][0], left)
        # rhs projection / rhs interfaces
        bv = ALS.TTBlockVector()
        Xb_k, Xb_k1 = {}, {}
        for i in range(nb):
            rb, rb1 = 2 + i % 2, 3
            bv[i] = [rng.standard_normal((rb, 4, rb1))]
            Xb_k[i] = rng.standard_normal((rb, r))
            Xb_k1[i] = rng.standard_normal((rb1, R))
            out[f"{p}/b/{i}"] = bv.get_row(i)[0]
            out[f"{p}/Xb1/{i}"] = Xb_k[i]
            out[f"{p}/Xb2/{i}"] = Xb_k1[i]
            out[f"{p}/phib_bck/{i}"] = ALS.compute_phi_bck_rhs(Xb_k1[i], bv.get_row(i)[0], left)
            out[f"{p}/phib_fwd/{i}"] = ALS.compute_phi_fwd_rhs(Xb_k[i], bv.get_row(i)[0], left)
        out[p + "/rhs"] = bv[0].block_local_product(Xb_k, Xb_k1, 1, (r, nb, 4, R))
        # reduced operators (cy_src/lgmres_cy.pyx)
        inv_I = 1.0 / (1.0 + rng.random((r, 4, R)))
        out[p + "/inv_I"] = inv_I
        if nb == 3:
            w = LG.MatVecWrapper(XAX_k[0, 0], XAX_k[0, 1], XAX_k[2, 1], XAX_k[2, 2],
                                 bm[0, 0][0], bm[0, 1][0], bm[2, 1][0], bm[2, 2][0],
                                 XAX_k1[0, 0], XAX_k1[0, 1], XAX_k1[2, 1], XAX_k1[2, 2], inv_I, r, 4, R)
            v = rng.standard_normal(2 * r * 4 * R)
        else:
            w = LG.IneqMatVecWrapper(XAX_k[0, 0], XAX_k[0, 1], XAX_k[2, 1], XAX_k[2, 2], XAX_k[3, 1], XAX_k[3, 3],
                                     bm[0, 0][0], bm[0, 1][0], bm[2, 1][0], bm[2, 2][0], bm[3, 1][0], bm[3, 3][0],
                                     XAX_k1[0, 0], XAX_k1[0, 1], XAX_k1[2, 1], XAX_k1[2, 2], XAX_k1[3, 1],
                                     XAX_k1[3, 3], inv_I, r, 4, R)
            v = rng.standard_normal(3 * r * 4 * R)
        out[p + "/red_x"] = v
        out[p + "/red_y"] = np.array(w.matvec(v.copy()), copy=True)

    # ---- TT algebra ------------------------------------------------------------
    a4 = _rand_tt(rng, [3, 4, 5, 2], (2, 2))
    b4 = _rand_tt(rng, [2, 3, 3, 4], (2, 2))
    a3 = _rand_tt(rng, [3, 5, 4, 2], (4,))
    b3 = _rand_tt(rng, [2, 2, 3, 3], (4,))
    op = _rand_tt(rng, [2, 3, 2, 2], (4, 4))
    for nm, tt in (("a4", a4), ("b4", b4), ("a3", a3), ("b3", b3), ("op", op)):
        _put_tt(out, "tt/in/" + nm, tt)
    _put_tt(out, "tt/add4", T.tt_add(a4, b4))
    _put_tt(out, "tt/add3", T.tt_add(a3, b3))
    out["tt/inner4"] = np.array(T.tt_inner_prod(a4, b4))
    out["tt/inner3"] = np.array(T.tt_inner_prod(a3, b3))
    out["tt/norm3"] = np.array(T.tt_norm(a3))
    out["tt/esum4"] = np.array(T.tt_entrywise_sum(a4))
    _put_tt(out, "tt/hadamard4", T.tt_fast_hadamard(_cp(a4), _cp(b4), 1e-12))
    _put_tt(out, "tt/hadamard3", T.tt_fast_hadamard(_cp(a3), _cp(b3), 1e-12))
    _put_tt(out, "tt/matvec", T.tt_fast_matrix_vec_mul(_cp(op), _cp(a3), 1e-12))
    _put_tt(out, "tt/matmat", T.tt_fast_mat_mat_mul(_cp(a4), _cp(b4), 1e-12))
    _put_tt(out, "tt/IkronM", T.tt_IkronM(a4))
    _put_tt(out, "tt/MkronI", T.tt_MkronI(a4))
    _put_tt(out, "tt/diag_op", T.tt_diag_op(_cp(a4), 1e-12))
    _put_tt(out, "tt/diag", T.tt_diag([c[:, :2] for c in _cp(a3)], 1e-12))
    _put_tt(out, "tt/transpose", T.tt_transpose(a4))
    _put_tt(out, "tt/rl_orth", CY.tt_rl_orthogonalise(_cp(a3)))
    # rounding: sum with a tiny perturbation so truncation really happens
    np.random.seed(7)
    big = T.tt_add(T.tt_add(a4, b4), T.tt_scale(1e-7, _rand_tt(rng, [2, 2, 2, 2], (2, 2))))
    _put_tt(out, "tt/in/big", big)
    for eps in (1e-12, 1e-5, 1e-1):
        red = T.tt_rank_reduce(_cp(big), eps)
        _put_tt(out, f"tt/round/{eps:g}", red)
    sym = T.tt_add(a4, T.tt_transpose(a4))
    sym = T.tt_add(sym, [5.0 ** (1 / 5) * np.eye(2).reshape(1, 2, 2, 1)] * 5)
    _put_tt(out, "tt/in/sym", sym)
    _put_tt(out, "tt/psd_round", T.tt_psd_rank_reduce(_cp(sym), 1e-2))
    mask = [np.array([[0.0, 1.0], [1.0, 0.0]]).reshape(1, 2, 2, 1) for _ in range(5)]
    _put_tt(out, "tt/in/mask", mask)
    _put_tt(out, "tt/mask_round", T.tt_mask_rank_reduce(_cp(sym), mask, 1e-2))
    _put_tt(out, "tt/retract", T.tt_rank_retraction(_cp(big), [2, 3, 3, 2]))
    np.random.seed(11)
    _put_tt(out, "tt/scale", T.tt_scale(0.1, a3))
    np.random.seed(11)
    _put_tt(out, "tt/normalise", T.tt_normalise(a3, radius=np.sqrt(5)))
    s = np.array([3.0, 1.0, 1e-3, 1e-7, 1e-9, 0.0])
    out["tt/prune_s"] = s
    out["tt/prune_eps"] = np.array([1e-12, 1e-8, 1e-6, 1e-2, 2.0, 10.0])
    out["tt/prune_out"] = np.array([CY.prune_singular_vals(s, e) for e in out["tt/prune_eps"]])

    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "kernels.npz"), **out)
    print("wrote kernels.npz with", len(out), "arrays")


def als():
    """ALS-fitted TT products (reference src/tt_als.py:1502-1762): inputs, the NumPy seed set
    before the call, the reference's result cores.  als_products.npz: "<case>/A|D|out/<k>", "<case>/seed|tol"."""
    ref = ref_env.load()
    ALS = ref.tt_als
    rng = np.random.default_rng(20251019)
    out = {}
    cases = {
        "matvec_d5": ("vec", [3, 5, 5, 3], [4, 6, 6, 4], 2, 1e-6),
        "matvec_d6": ("vec", [2, 4, 6, 4, 2], [3, 8, 9, 8, 3], 4, 1e-4),
        "matmat_d4": ("mat", [3, 4, 3], [2, 5, 2], 2, 1e-6),
        "matmat_d5": ("mat", [2, 3, 3, 2], [3, 4, 4, 3], 4, 1e-4),
    }
    for seed, (name, (kind, ra, rd, n, tol)) in enumerate(cases.items(), start=101):
        A = _rand_tt(rng, ra, (n, n))
        D = _rand_tt(rng, rd, (n,) if kind == "vec" else (n, n))
        fn = ALS.tt_approx_mat_vec_mul if kind == "vec" else ALS.tt_approx_mat_mat_mul
        np.random.seed(seed)
        res = fn(_cp(A), _cp(D), tol=tol)
        _put_tt(out, name + "/A", A)
        _put_tt(out, name + "/D", D)
        _put_tt(out, name + "/out", res)
        out[name + "/seed"] = np.array(seed)
        out[name + "/tol"] = np.array(tol)
        print(name, "ranks", [c.shape[-1] for c in res[:-1]])
    np.savez_compressed(os.path.join(GOLD, "als_products.npz"), **out)
    print("wrote als_products.npz with", len(out), "arrays")


def generators():
    """Problem-generator helpers of src/tt_ops.py: seeded sampler outputs, triangular matrices,
    bond splitting, and whole create_problem outputs as dense matrices.  generators.npz."""
    ref = ref_env.load()
    T = ref.tt_ops
    out = {}
    for seed, dim, rank in ((319, 5, 2), (83, 7, 4), (7, 3, 1), (11, 1, 2)):
        np.random.seed(seed)
        _put_tt(out, f"binary_sym/{seed}_{dim}_{rank}", T.tt_random_binary_sym(dim, rank, skew=-1.0))
    for seed, dim, r in ((319, 5, 1), (41, 6, 1), (83, 6, 2)):
        np.random.seed(seed)
        g = T.tt_random_graph(dim, r)
        out[f"graph/{seed}_{dim}_{r}/dense"] = T.tt_matrix_to_matrix(g)
        out[f"graph/{seed}_{dim}_{r}/ranks"] = np.array(T.tt_ranks(g))
        out[f"graph/{seed}_{dim}_{r}/rng_pos"] = np.array(np.random.get_state()[2])
    for dim in (1, 2, 4):
        _put_tt(out, f"tril/{dim}", T.tt_tril_one_matrix(dim))
        _put_tt(out, f"triu/{dim}", T.tt_triu_one_matrix(dim))
    rng = np.random.default_rng(5)
    m = _rand_tt(rng, [3, 2, 3], (2, 2))
    _put_tt(out, "split/in", m)
    _put_tt(out, "split/out", T.tt_split_bonds(_cp(m)))
    out["to_matrix"] = T.tt_matrix_to_matrix(m)
    for name, dim, rank, seed in (("maxcut", 4, 1, 319), ("corr_clust", 4, 1, 208), ("max_stable_set", 3, 1, 876),
                                  ("graphm", 2, 1, 5)):
        np.random.seed(seed)
        res = ref_env.create_problem(name, dim, rank)
        out[f"problem/{name}/n"] = np.array(len(res))
        for q, item in enumerate(res):
            if item is None:
                out[f"problem/{name}/{q}/none"] = np.array(1)
            elif isinstance(item, dict):
                for key, tt in item.items():
                    _put_tt(out, f"problem/{name}/{q}/dict/{key}", tt)
            else:
                _put_tt(out, f"problem/{name}/{q}/tt", item)
        out[f"problem/{name}/args"] = np.array([dim, rank, seed])
    np.savez_compressed(os.path.join(GOLD, "generators.npz"), **out)
    print("wrote generators.npz with", len(out), "arrays")


class _Budget(BaseException):
    pass


def amen(problem, dim, rank, seed, which, budget_s, local=False, inputs_only=False):
    """Trace calls of tt_restarted_block_amen made by the reference IPM."""
    ref = ref_env.load()
    ipm, ALS = ref.tt_ipm, ref.tt_als
    cfg = load_config(problem, dim)
    obj_tt, L_op_tt, bias_tt, ineq_mask, lag_maps = build_problem(ref, problem, dim, rank, seed)
    orig = ipm.tt_restarted_block_amen
    counter = [0]
    t_start = time.time()
    tag = f"{problem}_{dim}_r{rank}_s{seed}"
    solver_trace = []
    orig_eq, orig_ineq = ipm._ipm_local_solver, ipm._ipm_local_solver_ineq
    local_dump = []

    def wrap_solver(fn, is_ineq):
        def inner(XAX_k, A_k, XAX_k1, Xb_k, b_k, Xb_k1, prev, size_limit, dense_solve=True, rtol=1e-5):
            prev_c = np.array(prev, copy=True)
            res = fn(XAX_k, A_k, XAX_k1, Xb_k, b_k, Xb_k1, prev, size_limit, dense_solve, rtol)
            solver_trace.append((A_k._idx, res[1], res[2], prev.shape[0] * prev.shape[3], float(res[5])))
            if local and counter[0] - 1 in which and len(local_dump) < 6 and (len(solver_trace) % 7 == 3):
                d = {"ineq": np.array(int(is_ineq)), "prev": prev_c, "size_limit": np.array(size_limit),
                     "dense_solve": np.array(int(dense_solve)), "sol": np.array(res[0], copy=True),
                     "res_old": np.array(res[1]), "res_new": np.array(res[2]), "rhs": np.array(res[3], copy=True),
                     "norm_rhs": np.array(res[4]), "direct_fail": np.array(int(res[5]))}
                for key, v in A_k.items():
                    d[f"A/{key[0]}{key[1]}"] = np.array(v, copy=True)
                for key, v in XAX_k.items():
                    d[f"P1/{key[0]}{key[1]}"] = np.array(v, copy=True)
                for key, v in XAX_k1.items():
                    d[f"P2/{key[0]}{key[1]}"] = np.array(v, copy=True)
                for i, v in b_k.items():
                    d[f"b/{i}"] = np.array(v, copy=True)
                for i, v in Xb_k.items():
                    d[f"Xb1/{i}"] = np.array(v, copy=True)
                for i, v in Xb_k1.items():
                    d[f"Xb2/{i}"] = np.array(v, copy=True)
                local_dump.append(d)
            return res
        return inner

    def traced(block_A, block_b, rank_restriction, op_tol, termination_tol=1e-3, eps=1e-11, num_restarts=3,
               inner_m=10, x0=None, local_solver=None, verbose=False):
        idx = counter[0]
        counter[0] += 1
        rec = idx in which
        is_ineq = local_solver is orig_ineq
        ls = wrap_solver(local_solver, is_ineq)
        if rec:
            out = {}
            for (i, j), cores in block_A._data.items():
                for k, c in enumerate(cores):
                    out[f"A/{i}{j}/{k}"] = np.array(c, copy=True)
            out["aliases"] = np.array([[*a, *b] for a, b in block_A._aliases.items()], dtype=np.int64).reshape(-1, 4)
            out["transposes"] = np.array([[*a, *b] for a, b in block_A._transposes.items()], dtype=np.int64).reshape(-1, 4)
            for i, cores in block_b._data.items():
                for k, c in enumerate(cores):
                    out[f"b/{i}/{k}"] = np.array(c, copy=True)
            if x0 is not None:
                for k, c in enumerate(x0):
                    out[f"x0/{k}"] = np.array(c, copy=True)
            st = np.random.get_state()
            out["rng_keys"] = st[1].copy()
            out["rng_pos"] = np.array([st[2], st[3]], dtype=np.int64)
            out["rng_gauss"] = np.array(st[4])
            out["args"] = np.array([rank_restriction, op_tol, termination_tol, eps, num_restarts, inner_m,
                                    float(is_ineq)])
            out["d"] = np.array(len(next(iter(block_b._data.values()))))
            n0 = len(solver_trace)
            t0 = time.time()
        if time.time() - t_start > budget_s and not rec:
            raise _Budget()
        if rec and inputs_only:
            out["wall_s"] = np.array(np.nan)
            out["raised"] = np.array(0)
            out["trace"] = np.zeros((0, 5))
            path = os.path.join(GOLD, f"amen_{tag}_{idx}.npz")
            np.savez_compressed(path, **out)
            print(f"wrote {path} (inputs only, {os.path.getsize(path) / 1e3:.0f} kB)", flush=True)
            if idx >= max(which):
                raise _Budget()
        err = None
        try:
            x, res = orig(block_A, block_b, rank_restriction, op_tol, termination_tol=termination_tol, eps=eps,
                          num_restarts=num_restarts, inner_m=inner_m, x0=x0, local_solver=ls, verbose=verbose)
        except RuntimeError as e:
            err = e
        if rec:
            out["wall_s"] = np.array(time.time() - t0)
            out["raised"] = np.array(int(err is not None))
            if err is None:
                for k, c in enumerate(x):
                    out[f"out/x/{k}"] = np.array(c, copy=True)
                out["out_res"] = np.array(res)
            out["trace"] = np.array(solver_trace[n0:], dtype=np.float64).reshape(-1, 5)
            path = os.path.join(GOLD, f"amen_{tag}_{idx}.npz")
            np.savez_compressed(path, **out)
            print(f"wrote {path}  ({os.path.getsize(path) / 1e3:.0f} kB, {out['wall_s']:.2f}s, "
                  f"{len(solver_trace) - n0} local solves)", flush=True)
            for q, dmp in enumerate(local_dump):
                lp = os.path.join(GOLD, f"local_{tag}_{idx}_{q}.npz")
                np.savez_compressed(lp, **dmp)
            local_dump.clear()
            if idx >= max(which):
                raise _Budget()
        if err is not None:
            raise err
        return x, res

    ipm.tt_restarted_block_amen = traced
    try:
        ipm.tt_ipm(lag_maps, obj_tt, L_op_tt, bias_tt, ineq_mask=ineq_mask, max_iter=cfg["max_iter"], verbose=False,
                   gap_tol=float(cfg["gap_tol"]), op_tol=float(cfg["op_tol"]), warm_up=cfg["warm_up"],
                   abs_tol=float(cfg["abs_tol"]), aho_direction=False, mals_restarts=cfg["mals_restarts"],
                   max_refinement=cfg["max_refinement"], lambdaStar=float(cfg.get("lambdaStar", 1)),
                   lambdaStarIneq=float(cfg.get("lambdaStarIneq", 1)))
    except _Budget:
        pass
    finally:
        ipm.tt_restarted_block_amen = orig
    print("AMEn calls seen:", counter[0])


def eigen(problem, dim, rank, seed, which):
    """Trace calls of the step-size eigen sweeps (reference src/tt_als.py:1132-1283 tt_max_generalised_eigen,
    :1392-1499 tt_min_eig) made by the reference IPM.  eigen_<cfg>_<n>.npz: "kind" (0 = generalised,
    1 = min_eig), "A/<k>", "Delta/<k>" (generalised only), "x0/<k>" (absent if None), "tol", rng_keys / rng_pos at
    entry, "out_scalar" (step size / eigenvalue), "out/x/<k>"."""
    ref = ref_env.load()
    ipm = ref.tt_ipm
    cfg = load_config(problem, dim)
    obj_tt, L_op_tt, bias_tt, ineq_mask, lag_maps = build_problem(ref, problem, dim, rank, seed)
    tag = f"{problem}_{dim}_r{rank}_s{seed}"
    counter = [0]
    orig_gen, orig_min = ipm.tt_max_generalised_eigen, ipm.tt_min_eig

    def record(kind, A, Delta, x0, tol, call):
        idx = counter[0]
        counter[0] += 1
        if idx not in which:
            res = call()
            if idx > max(which):
                raise _Budget()
            return res
        out = {"kind": np.array(kind), "tol": np.array(tol)}
        _put_tt(out, "A", _cp(A))
        if Delta is not None:
            _put_tt(out, "Delta", _cp(Delta))
        if x0 is not None:
            _put_tt(out, "x0", _cp(x0))
        st = np.random.get_state()
        out["rng_keys"], out["rng_pos"] = np.array(st[1]), np.array(st[2])
        res = call()
        scalar, xs = (res[0], res[1]) if kind == 0 else (res[1], res[0])
        out["out_scalar"] = np.array(np.nan if scalar is None else scalar, dtype=np.float64)
        _put_tt(out, "out/x", _cp(xs))
        path = os.path.join(GOLD, f"eigen_{tag}_{idx}.npz")
        np.savez_compressed(path, **out)
        print(f"wrote {path} kind={kind} scalar={out['out_scalar']} ({os.path.getsize(path) / 1e3:.0f} kB)", flush=True)
        return res

    def traced_gen(A, Delta, x0=None, nswp=10, tol=1e-8, size_limit=256, verbose=False):
        return record(0, A, Delta, x0, tol, lambda: orig_gen(A, Delta, x0=x0, nswp=nswp, tol=tol,
                                                              size_limit=size_limit, verbose=verbose))

    def traced_min(A, x0=None, nswp=10, tol=1e-8, size_limit=64, return_eig_val=False, verbose=False):
        return record(1, A, None, x0, tol, lambda: orig_min(A, x0=x0, nswp=nswp, tol=tol, size_limit=size_limit,
                                                            return_eig_val=return_eig_val, verbose=verbose))

    ipm.tt_max_generalised_eigen, ipm.tt_min_eig = traced_gen, traced_min
    try:
        ipm.tt_ipm(lag_maps, obj_tt, L_op_tt, bias_tt, ineq_mask=ineq_mask, max_iter=cfg["max_iter"], verbose=False,
                   gap_tol=float(cfg["gap_tol"]), op_tol=float(cfg["op_tol"]), warm_up=cfg["warm_up"],
                   abs_tol=float(cfg["abs_tol"]), aho_direction=False, mals_restarts=cfg["mals_restarts"],
                   max_refinement=cfg["max_refinement"], lambdaStar=float(cfg.get("lambdaStar", 1)),
                   lambdaStarIneq=float(cfg.get("lambdaStarIneq", 1)))
    except _Budget:
        pass
    finally:
        ipm.tt_max_generalised_eigen, ipm.tt_min_eig = orig_gen, orig_min


def e2e(problem, dim, rank, seed):
    from run_ref_ipm import run
    res = run(problem, dim, rank, seed, verbose=False)
    path = os.path.join(GOLD, "e2e.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[f"{problem}_{dim}_r{rank}_s{seed}"] = res
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    print(res)


if __name__ == "__main__":
    cmd = sys.argv[1]
    if cmd == "kernels":
        kernels()
    elif cmd == "amen":
        budget = 600.0
        if "--budget" in sys.argv:
            budget = float(sys.argv[sys.argv.index("--budget") + 1])
        amen(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]),
             set(int(v) for v in sys.argv[6].split(",")), budget, local="--local" in sys.argv,
             inputs_only="--inputs-only" in sys.argv)
    elif cmd == "als":
        als()
    elif cmd == "generators":
        generators()
    elif cmd == "eigen":
        eigen(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]),
              set(int(v) for v in sys.argv[6].split(",")))
    elif cmd == "e2e":
        e2e(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]))
