"""Harness (this container only): run the UNMODIFIED reference IPM driver (src/tt_ipm.py::tt_ipm) with the
Newton-system hot path swapped for ttipm_b200 by dropin.install(), next to the pure reference run, and report
iteration counts / gap / feasibility of both (north-star end-to-end criterion: same IPM iteration count +-1).

There is no GPU in this container, so the kernels run through the -DTTIPM_EMU build of the same sources
(tests/emu); on a machine that has both the reference tree and a B200, pass --cuda.

  python oracle/ref_harness/run_dropin_ipm.py maxcut 5 1 319 [--cuda] [--skip-ref]
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path[:0] = [HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"),
                os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
import ref_env  # noqa: E402
import run_ref_ipm  # noqa: E402


def _dump_latest_amen_inputs():
    """Before every tt_restarted_block_amen call of the run write its inputs (the amen_*.npz fixture layout of
    make_golden.amen, inputs only) to gpurun_out/amen_dump_<index>.npz and print the call's wall time afterwards: a
    call that does not return inside the time limit of the run leaves its system behind for a stand-alone look."""
    import time
    import numpy as np
    ref = ref_env.load()
    ipm = ref.tt_ipm
    orig = ipm.tt_restarted_block_amen
    counter = [0]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)

    def traced(block_A, block_b, rank_restriction, op_tol, termination_tol=1e-3, eps=1e-11, num_restarts=3, inner_m=10,
               x0=None, local_solver=None, verbose=False):
        idx = counter[0]
        counter[0] += 1
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
        is_ineq = local_solver is not None and "ineq" in getattr(local_solver, "__name__", "")
        out["args"] = np.array([rank_restriction, op_tol, termination_tol, eps, num_restarts, inner_m, float(is_ineq)])
        out["d"] = np.array(len(next(iter(block_b._data.values()))))
        out["wall_s"] = np.array(np.nan)
        out["raised"] = np.array(0)
        out["trace"] = np.zeros((0, 5))
        path = os.path.join(ROOT, "gpurun_out", f"amen_dump_{idx % 2}.npz")     # the last two calls survive
        np.savez(path, **out)
        ranks = {key: max(c.shape[0] for c in cores) for key, cores in block_A._data.items()}
        print(f"AMEN-CALL {idx} -> {os.path.basename(path)} ineq={is_ineq} operator ranks {ranks} "
              f"x0 ranks {[c.shape[-1] for c in x0] if x0 is not None else None}", flush=True)
        t0 = time.time()
        res = orig(block_A, block_b, rank_restriction, op_tol, termination_tol=termination_tol, eps=eps,
                   num_restarts=num_restarts, inner_m=inner_m, x0=x0, local_solver=local_solver, verbose=verbose)
        print(f"AMEN-CALL {idx} done in {time.time() - t0:.3f} s, solution ranks {[c.shape[-1] for c in res[0]]}", flush=True)
        return res

    ipm.tt_restarted_block_amen = traced


def main():
    problem, dim, rank, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    out = {}
    if "--skip-ref" not in sys.argv:
        out["reference"] = run_ref_ipm.run(problem, dim, rank, seed, verbose=False)
        print("REFERENCE", json.dumps(out["reference"]), flush=True)
    import rt_util
    rt = rt_util.cuda_runtime() if "--cuda" in sys.argv else rt_util.emu_runtime()
    from ttipm_b200 import dropin, use_runtime
    ref_env.load()
    with use_runtime(rt):
        done = dropin.install()
        print("rebound:", {k: len(v) for k, v in done.items()}, flush=True)
        if "--dump-latest" in sys.argv:
            _dump_latest_amen_inputs()
        if "--profile" in sys.argv:
            import cProfile
            import pstats
            pr = cProfile.Profile()
            pr.enable()
        out["dropin"] = run_ref_ipm.run(problem, dim, rank, seed, verbose="--verbose" in sys.argv)
        if "--profile" in sys.argv:
            pr.disable()
            pstats.Stats(pr).sort_stats("cumulative").print_stats(70)
    from ttipm_b200 import als_product
    out["dropin"]["als_product_fits"] = dict(als_product.STATS)
    print("DROPIN", json.dumps(out["dropin"]), flush=True)
    for i, a in enumerate(sys.argv):
        if a == "--out":
            out["generator"] = "oracle/ref_harness/run_dropin_ipm.py " + " ".join(sys.argv[1:5])
            with open(sys.argv[i + 1], "w") as f:
                json.dump(out, f, indent=1)
    if "reference" in out:
        a, b = out["reference"], out["dropin"]
        print("COMPARE iters %d vs %d | gap %.3e vs %.3e | primal^2 %.3e vs %.3e | dual^2 %.3e vs %.3e" % (
            a["iters"], b["iters"], a["gap"], b["gap"], a["primal_sq"], b["primal_sq"], a["dual_sq"], b["dual_sq"]))


if __name__ == "__main__":
    main()
