"""ALS-fitted TT product on the B200 next to the oracle port on the host cores, same inputs and seed.

  python tools/bench_als_product.py > gpurun_out/als_product.jsonl
  python tools/bench_als_product.py --large     # device only: d = 8 mat-vec, rank products 96 (the oracle needs minutes)

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


def dense(tt):
    t = tt[0]
    for c in tt[1:]:
        t = np.tensordot(t, c, axes=(-1, 0))
    return t


def large():
    from ttipm_b200 import als_product as AP
    from ttipm_b200.runtime import get_runtime
    rt = get_runtime()
    rng = np.random.default_rng(12)
    d, n = 8, 4
    ra, rd = [1, 4, 6, 6, 6, 6, 6, 4, 1], [1, 8, 16, 16, 16, 16, 16, 8, 1]
    A = [rng.standard_normal((ra[k], n, n, ra[k + 1])) / np.sqrt(ra[k] * ra[k + 1]) for k in range(d)]
    v = [rng.standard_normal((rd[k], n, rd[k + 1])) / np.sqrt(rd[k] * rd[k + 1]) for k in range(d)]
    exact = dense([np.einsum("amkA,bkB->abmAB", a, b).reshape(a.shape[0] * b.shape[0], n, -1) for a, b in zip(A, v)])
    rec = {"case": "matvec_d8_large", "d": d, "ranks_A": ra[1:-1], "ranks_D": rd[1:-1], "mode": n, "tol": 1e-6}
    for rep in range(2):
        np.random.seed(21)
        tr = []
        l0 = rt.launches
        t0 = time.perf_counter()
        out = AP.als_fit_product([c.copy() for c in A], [c.copy() for c in v], tol=1e-6, trace=tr)
        rt.sync()
        rec["device_s"] = time.perf_counter() - t0
        rec["launches"] = rt.launches - l0
    rec["half_sweeps"] = len(tr)
    rec["ranks"] = [c.shape[-1] for c in out[:-1]]
    rec["rel_err_vs_exact"] = float(np.linalg.norm(dense(out) - exact) / np.linalg.norm(exact))
    print(json.dumps(rec), flush=True)


def main():
    if "--large" in sys.argv:
        return large()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import golden_io as G
    import tt_oracle as O
    from ttipm_b200 import als_product as AP
    from ttipm_b200.runtime import get_runtime
    rt = get_runtime()
    z = G.load("als_products.npz")          # the reference-generated cases (oracle/ref_harness/make_golden.py als)
    cp = lambda tt: [c.copy() for c in tt]
    for name in ("matmat_d4", "matmat_d5", "matvec_d6"):
        A, D, ref_out = (G.get_tt(z, f"{name}/{q}") for q in ("A", "D", "out"))
        tol, seed = float(z[name + "/tol"]), int(z[name + "/seed"])
        rec = {"case": name, "d": len(A), "ranks_A": [c.shape[-1] for c in A[:-1]], "ranks_D": [c.shape[-1] for c in D[:-1]],
               "mode": int(A[0].shape[1]), "tol": tol}
        for rep in range(2):                      # first pass warms the allocator / module load
            np.random.seed(seed)
            tr = []
            l0 = rt.launches
            t0 = time.perf_counter()
            out = AP.als_fit_product(cp(A), cp(D), tol=tol, trace=tr)
            rt.sync()
            rec["device_s"] = time.perf_counter() - t0
            rec["launches"] = rt.launches - l0
        np.random.seed(seed)
        tro = []
        t0 = time.perf_counter()
        (O.tt_approx_mat_vec_mul if D[0].ndim == 3 else O.tt_approx_mat_mat_mul)(cp(A), cp(D), tol=tol, trace=tro)
        rec["oracle_s"] = time.perf_counter() - t0
        rec["half_sweeps"] = [len(tr), len(tro)]
        rec["ranks"] = [c.shape[-1] for c in out[:-1]]
        ex = dense(ref_out)
        rec["rel_err_vs_reference_output"] = float(np.linalg.norm(dense(out) - ex) / np.linalg.norm(ex))
        rec["cores"] = os.cpu_count()
        rec["note"] = "oracle = NumPy port with pairwise tensordot (BLAS) contractions, BLAS threads = cores"
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
