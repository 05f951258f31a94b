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
