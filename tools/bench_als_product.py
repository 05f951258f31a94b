"""ALS-fitted TT product on the B200 next to the oracle port on the host cores, same inputs and seed.

  python tools/bench_als_product.py > gpurun_out/als_product.jsonl

One JSON line per case: wall seconds of the device fit (NumPy in -> NumPy out, launches + host syncs included), of the
oracle (NumPy/SciPy, BLAS threads = cores), half sweeps, final ranks, relative error against the exact product."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "tensor-train-interior-point-method_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def rand_tt(rng, ranks, mode):
    rr = [1] + list(ranks) + [1]
    return [rng.standard_normal((a, *mode, b)) / np.sqrt(a * b) for a, b in zip(rr[:-1], rr[1:])]


def dense(tt):
    t = tt[0]
    for c in tt[1:]:
        t = np.tensordot(t, c, axes=(-1, 0))
    return t


def main():
    import tt_oracle as O
    from ttipm_b200 import als_product as AP
    from ttipm_b200.runtime import get_runtime
    rt = get_runtime()
    rng = np.random.default_rng(1)
    cases = [("matvec", 7, [4, 6, 6, 6, 6, 4], [8, 16, 16, 16, 16, 8], 4, 1e-6),
             ("matmat", 6, [3, 5, 5, 5, 3], [4, 9, 9, 9, 4], 4, 1e-6),
             ("matvec", 9, [4, 6, 6, 6, 6, 6, 6, 4], [8, 20, 20, 20, 20, 20, 20, 8], 4, 1e-4)]
    for kind, d, ra, rd, n, tol in cases:
        A = rand_tt(rng, ra, (n, n))
        D = rand_tt(rng, rd, (n,) if kind == "matvec" else (n, n))
        cp = lambda tt: [c.copy() for c in tt]
        rec = {"kind": kind, "d": d, "ranks_A": ra, "ranks_D": rd, "mode": n, "tol": tol}
        for rep in range(2):                      # first pass warms the allocator / module load
            np.random.seed(3)
            tr = []
            l0 = rt.launches
            t0 = time.perf_counter()
            out = AP.als_fit_product(cp(A), cp(D), tol=tol, trace=tr)
            rt.sync()
            rec["device_s"] = time.perf_counter() - t0
            rec["launches"] = rt.launches - l0
        np.random.seed(3)
        tro = []
        t0 = time.perf_counter()
        ref = (O.tt_approx_mat_vec_mul if kind == "matvec" else O.tt_approx_mat_mat_mul)(cp(A), cp(D), tol=tol, trace=tro)
        rec["oracle_s"] = time.perf_counter() - t0
        rec["half_sweeps"] = [len(tr), len(tro)]
        rec["ranks"] = [c.shape[-1] for c in out[:-1]]
        rec["ranks_oracle"] = [c.shape[-1] for c in ref[:-1]]
        if d <= 7:
            ex = dense(O.tt_fast_matrix_vec_mul(A, D, 1e-14) if kind == "matvec" else O.tt_fast_mat_mat_mul(A, D, 1e-14))
            rec["rel_err_vs_exact"] = float(np.linalg.norm(dense(out) - ex) / np.linalg.norm(ex))
            rec["rel_err_oracle_vs_exact"] = float(np.linalg.norm(dense(ref) - ex) / np.linalg.norm(ex))
        rec["cores"] = os.cpu_count()
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
